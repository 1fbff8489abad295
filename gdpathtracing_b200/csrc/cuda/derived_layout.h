// derived_layout.h -- host-side construction of the WideNode / LeafRec / InstRec
// tables (pt_scene.cuh) from the uploaded reference arrays.  Pure C++ (no CUDA)
// so the CPU unit check of the device functions can reuse it.
#ifndef GDPT_DERIVED_LAYOUT_H
#define GDPT_DERIVED_LAYOUT_H

#include "pt_scene.cuh"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace gdpt {

struct DerivedLayout {
    std::vector<WideNode> wide_nodes;
    std::vector<LeafRec> leaf_recs;
    std::vector<WideNode> wide_tlas;
    std::vector<InstRec> inst_recs;
    uint32_t tlas_root_link = LINK_NONE;
};

// Returns "" on success, otherwise what is wrong with the input arrays.
inline std::string derive_layout(const gdpt_bvh_node *bvh, uint32_t n_nodes, const gdpt_blas_instance *blas, uint32_t n_blas,
                                 const gdpt_tlas_node *tlas, uint32_t n_tlas, DerivedLayout &out)
{
    char msg[160];
    if (n_tlas == 0 || n_blas == 0) return "empty TLAS / instance buffer";
    if (n_nodes >= LINK_INDEX_MASK) return "too many BVH nodes";

    // BLAS: internal nodes and leaves get their own dense numbering
    std::vector<uint32_t> link(n_nodes);
    uint32_t n_internal = 0, n_leaf = 0;
    for (uint32_t i = 0; i < n_nodes; i++) link[i] = bvh[i].tri_count > 0 ? (LINK_LEAF | n_leaf++) : n_internal++;
    out.wide_nodes.assign(n_internal, WideNode());
    out.leaf_recs.assign(n_leaf, LeafRec());
    for (uint32_t i = 0; i < n_nodes; i++) {
        const gdpt_bvh_node &n = bvh[i];
        if (n.tri_count > 0) {
            LeafRec &l = out.leaf_recs[link[i] & LINK_INDEX_MASK];
            l.first_tri = n.first_tri_index; l.tri_count = n.tri_count; l.orig = i; l.pad = 0;
            continue;
        }
        if (n.left_child >= n_nodes || n.right_child >= n_nodes) {
            std::snprintf(msg, sizeof(msg), "BVH node %u has a child index out of range", i);
            return msg;
        }
        WideNode &w = out.wide_nodes[link[i]];
        const gdpt_bvh_node &L = bvh[n.left_child], &R = bvh[n.right_child];
        for (int k = 0; k < 3; k++) {
            w.lmin[k] = L.aabb_min[k]; w.lmax[k] = L.aabb_max[k];
            w.rmin[k] = R.aabb_min[k]; w.rmax[k] = R.aabb_max[k];
        }
        w.left = link[n.left_child]; w.right = link[n.right_child]; w.orig = i; w.pad = 0;
    }

    // TLAS: same split; leaves index the instance table directly
    std::vector<uint32_t> tlink(n_tlas);
    uint32_t nt_internal = 0;
    for (uint32_t i = 0; i < n_tlas; i++) {
        if (tlas[i].left_right == 0) {
            if (tlas[i].blas >= n_blas) {
                std::snprintf(msg, sizeof(msg), "TLAS leaf %u names instance %u of %u", i, tlas[i].blas, n_blas);
                return msg;
            }
            tlink[i] = LINK_TLAS | LINK_LEAF | tlas[i].blas;
        } else tlink[i] = LINK_TLAS | nt_internal++;
    }
    out.wide_tlas.assign(nt_internal, WideNode());
    out.inst_recs.assign(n_blas, InstRec());
    for (uint32_t b = 0; b < n_blas; b++) {
        if (blas[b].root >= n_nodes) {
            std::snprintf(msg, sizeof(msg), "instance %u root out of range", b);
            return msg;
        }
        InstRec &r = out.inst_recs[b];
        std::memcpy(r.inv, blas[b].inverse_transform, sizeof(r.inv));
        r.root_link = link[blas[b].root];
        r.root_orig = blas[b].root;
        r.tlas_orig = 0; r.pad = 0;
    }
    for (uint32_t i = 0; i < n_tlas; i++) {
        const gdpt_tlas_node &n = tlas[i];
        if (n.left_right == 0) {
            if (i > 0) out.inst_recs[n.blas].tlas_orig = i; // leaves sit at 1..I in instance order (bvh.cpp:278-287)
            continue;
        }
        const uint32_t l = n.left_right & 0xFFFFu, r = n.left_right >> 16;
        if (l >= n_tlas || r >= n_tlas) {
            std::snprintf(msg, sizeof(msg), "TLAS node %u has a child out of range", i);
            return msg;
        }
        WideNode &w = out.wide_tlas[tlink[i] & LINK_INDEX_MASK];
        for (int k = 0; k < 3; k++) {
            w.lmin[k] = tlas[l].aabb_min[k]; w.lmax[k] = tlas[l].aabb_max[k];
            w.rmin[k] = tlas[r].aabb_min[k]; w.rmax[k] = tlas[r].aabb_max[k];
        }
        w.left = tlink[l]; w.right = tlink[r]; w.orig = i; w.pad = 0;
    }
    // node 0 is a copy of the final root (bvh.cpp:316); with a single instance that root is
    // itself the leaf, and the traversal pops it as node 0
    if (tlas[0].left_right == 0) out.inst_recs[tlas[0].blas].tlas_orig = 0;
    out.tlas_root_link = tlink[0];
    return "";
}

} // namespace gdpt
#endif
