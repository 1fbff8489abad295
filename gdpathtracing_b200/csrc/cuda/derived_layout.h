// derived_layout.h -- host-side construction of the WideNode / LeafRec / InstRec
// tables (pt_scene.cuh) from the uploaded reference arrays.  Pure C++ (no CUDA)
// so the CPU unit check of the device functions can reuse it.
#ifndef GDPT_DERIVED_LAYOUT_H
#define GDPT_DERIVED_LAYOUT_H

#include "pt_scene.cuh"

#include <cfloat>
#include <cmath>
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
    float world_reach = 0.0f; // SceneView::fast_world_reach
    std::vector<float> root_extent; // per BVH node index of a BLAS root: the effective extent its culling margins are built on
};

struct TightBox { float lo[3], hi[3]; };

inline TightBox tight_empty()
{
    TightBox b;
    for (int k = 0; k < 3; k++) { b.lo[k] = FLT_MAX; b.hi[k] = -FLT_MAX; }
    return b;
}
inline void tight_grow(TightBox &b, const float *p)
{
    for (int k = 0; k < 3; k++) { if (p[k] < b.lo[k]) b.lo[k] = p[k]; if (p[k] > b.hi[k]) b.hi[k] = p[k]; }
}
inline void tight_merge(TightBox &b, const TightBox &o)
{
    for (int k = 0; k < 3; k++) { if (o.lo[k] < b.lo[k]) b.lo[k] = o.lo[k]; if (o.hi[k] > b.hi[k]) b.hi[k] = o.hi[k]; }
}
inline float tight_extent(const TightBox &b)
{
    float e = 0.0f;
    for (int k = 0; k < 3; k++) if (b.hi[k] - b.lo[k] > e) e = b.hi[k] - b.lo[k];
    return e;
}
// Safety margin of a culling box: 1/GDPT_MARGIN_OWN_DIV of its own size plus 1/GDPT_MARGIN_DIV of the (effective) size of the BLAS it
// belongs to.  The second term is what covers Moller-Trumbore's rounding (absolute error of the order of
// 1e-6 x distance to the ray origin) for rays that start within the reach that goes with it (fast_reach: 7.6 x reserve
// at the reach itself).  Non-finite input boxes become "everything" (never cull).
#ifndef GDPT_MARGIN_DIV
#define GDPT_MARGIN_DIV 4096.0f /* the owner term of the margin is extent / GDPT_MARGIN_DIV; the reach scales with it (fast_reach).
                                   A/B on the B200 (512 | 2048 | 4096): C2 path kernel 0.638 | 0.609 | 0.604 ms, C4 1080p 11.75 | 11.17 | 11.11,
                                   C3 7.86 | 5.75 | 5.37 (triangle tests per frame 114 M -> 57 M); same safety factor at every value */
#endif
#ifndef GDPT_MARGIN_OWN_DIV
#define GDPT_MARGIN_OWN_DIV 2048.0f /* a second, smaller term relative to the box itself (belt and braces: the errors the margin must
                                       absorb scale with the coordinates, which the owner term covers).  A/B (256 | 2048 | none): C2 path
                                       kernel 0.610 | 0.600 | 0.592 ms, C4 1080p 11.16 | 10.94 | 10.90, C3 5.37 | 5.23 | 5.23 */
#endif
inline TightBox tight_inflate(const TightBox &b, float owner_extent)
{
    TightBox r;
    const float m = tight_extent(b) * (1.0f / GDPT_MARGIN_OWN_DIV) + owner_extent * (1.0f / GDPT_MARGIN_DIV) + 1e-30f;
    bool finite = std::isfinite(m);
    for (int k = 0; k < 3; k++) {
        r.lo[k] = b.lo[k] - m; r.hi[k] = b.hi[k] + m;
        finite = finite && std::isfinite(r.lo[k]) && std::isfinite(r.hi[k]) && b.lo[k] <= b.hi[k];
    }
    if (!finite)
        for (int k = 0; k < 3; k++) { r.lo[k] = -FLT_MAX; r.hi[k] = FLT_MAX; }
    return r;
}

// How far out a ray may start for a box's culling margin to be trusted (pt_fast.cuh fast_far_origin).  Two errors grow
// with the size of the coordinates the search works with and must stay inside the margin (>= extent / 512):
//   * Moller-Trumbore accepts hits up to ~1e-6 x (distance to the origin) outside the triangle it tests;
//   * the search's own slab tests (FFMA + hardware reciprocal) move a plane by a few ulp (~2.4e-7) of the larger
//     of |origin| and |plane|.
// With coordinates below (131072 / GDPT_MARGIN_DIV) x extent both stay under a quarter of the margin (256 extents for a
// margin of extent / 512, 32 for extent / 4096).  A box that itself lies farther than half of that from its space's origin
// gets reach 0: every ray that enters it is answered by the exact traversal.  `extent` is the EFFECTIVE extent the margin
// was built on (derive_layout: at least the box's own, more where the scene needs a longer reach).
inline float fast_reach(const TightBox &b, float extent)
{
    float far_corner = 0.0f;
    for (int k = 0; k < 3; k++) { far_corner = std::fmax(far_corner, std::fabs(b.lo[k])); far_corner = std::fmax(far_corner, std::fabs(b.hi[k])); }
    const float reach = (131072.0f / GDPT_MARGIN_DIV) * extent; // 256 extents for the default margin of extent / 512
    return (std::isfinite(reach) && far_corner <= 0.5f * reach) ? reach : 0.0f;
}

// Returns "" on success, otherwise what is wrong with the input arrays.
inline std::string derive_layout(const gdpt_bvh_node *bvh, uint32_t n_nodes, const gdpt_blas_instance *blas, uint32_t n_blas,
                                 const gdpt_tlas_node *tlas, uint32_t n_tlas, const gdpt_triangle_geometry *tris,
                                 uint32_t n_tris, DerivedLayout &out)
{
    char msg[160];
    if (n_tlas == 0 || n_blas == 0) return "empty TLAS / instance buffer";
    if (n_nodes >= LINK_INDEX_MASK) return "too many BVH nodes";

    // ---- BLAS: internal nodes and leaves get their own dense numbering
    std::vector<uint32_t> link(n_nodes);
    uint32_t n_internal = 0, n_leaf = 0;
    for (uint32_t i = 0; i < n_nodes; i++) {
        const gdpt_bvh_node &n = bvh[i];
        if (n.tri_count > 0) {
            if ((uint64_t)n.first_tri_index + n.tri_count > n_tris) {
                std::snprintf(msg, sizeof(msg), "BVH leaf %u references triangles beyond the triangle buffer", i);
                return msg;
            }
            link[i] = LINK_LEAF | n_leaf++;
        } else {
            if (n.left_child >= n_nodes || n.right_child >= n_nodes) {
                std::snprintf(msg, sizeof(msg), "BVH node %u has a child index out of range", i);
                return msg;
            }
            link[i] = n_internal++;
        }
    }

    // ---- true bounding boxes, bottom-up per BLAS (explicit post-order walk from every root)
    std::vector<TightBox> raw(n_nodes, tight_empty());
    std::vector<uint8_t> state(n_nodes, 0); // 0 unseen, 1 children pending, 2 done
    std::vector<float> owner_extent(n_nodes, 0.0f);
    std::vector<uint32_t> walk;
    for (uint32_t b = 0; b < n_blas; b++) {
        const uint32_t root = blas[b].root;
        if (root >= n_nodes) {
            std::snprintf(msg, sizeof(msg), "instance %u root out of range", b);
            return msg;
        }
        if (state[root] == 2) continue;
        walk.assign(1, root);
        while (!walk.empty()) {
            const uint32_t i = walk.back();
            const gdpt_bvh_node &n = bvh[i];
            if (n.tri_count > 0) {
                TightBox t = tight_empty();
                for (uint32_t k = 0; k < n.tri_count; k++)
                    for (int v = 0; v < 3; v++) tight_grow(t, tris[n.first_tri_index + k].v[v]);
                raw[i] = t; state[i] = 2; walk.pop_back();
            } else if (state[i] == 0) {
                state[i] = 1;
                if (state[n.left_child] == 1 || state[n.right_child] == 1) return "BVH child links form a cycle";
                if (state[n.left_child] == 0) walk.push_back(n.left_child);
                if (state[n.right_child] == 0) walk.push_back(n.right_child);
            } else {
                TightBox t = raw[n.left_child];
                tight_merge(t, raw[n.right_child]);
                raw[i] = t; state[i] = 2; walk.pop_back();
            }
        }
    }
    // How far out rays may start, by construction of the margins: four scene radii.  A margin of extent / GDPT_MARGIN_DIV
    // is trusted for origins within (131072 / GDPT_MARGIN_DIV) extents (fast_reach); where that is less than the scene
    // needs -- a small prop in a large room -- the BLAS (or, at the world level, the instance) gets the margin of a larger
    // "effective" extent instead, so no ray inside four scene radii is ever sent to the exact traversal for its origin alone.
    float scene_radius = 0.0f;
    for (uint32_t b = 0; b < n_blas; b++) {
        const TightBox &o = raw[blas[b].root];
        const float *m = blas[b].transform;
        for (int c = 0; c < 8; c++) {
            const double x = (c & 1) ? o.hi[0] : o.lo[0], y = (c & 2) ? o.hi[1] : o.lo[1], z = (c & 4) ? o.hi[2] : o.lo[2];
            for (int k = 0; k < 3; k++) {
                const float w = std::fabs((float)(m[k] * x + m[4 + k] * y + m[8 + k] * z + m[12 + k]));
                if (std::isfinite(w) && w > scene_radius) scene_radius = w;
            }
        }
    }
    const float needed_world = 4.0f * scene_radius;
    std::vector<float> root_extent(n_nodes, 0.0f); // per BLAS root: the effective extent its margins are built on
    for (uint32_t b = 0; b < n_blas; b++) {
        const uint32_t root = blas[b].root;
        const float *a = blas[b].inverse_transform; // local = A w + t
        float norm = 0.0f;
        for (int r = 0; r < 3; r++) norm = std::fmax(norm, std::fabs(a[r]) + std::fabs(a[4 + r]) + std::fabs(a[8 + r]));
        const float shift = std::fmax(std::fabs(a[12]), std::fmax(std::fabs(a[13]), std::fabs(a[14])));
        const float needed_local = norm * needed_world + shift;
        const float eff = std::fmax(tight_extent(raw[root]), std::isfinite(needed_local) ? needed_local * (GDPT_MARGIN_DIV / 131072.0f) : 0.0f);
        if (eff > root_extent[root]) root_extent[root] = eff;
    }
    std::vector<uint8_t> stamped(n_nodes, 0);
    for (uint32_t b = 0; b < n_blas; b++) { // second walk: stamp the owning BLAS's effective extent on every node of its tree
        const uint32_t root = blas[b].root;
        if (stamped[root]) continue;
        walk.assign(1, root);
        while (!walk.empty()) {
            const uint32_t i = walk.back();
            walk.pop_back();
            owner_extent[i] = root_extent[root];
            stamped[i] = 1;
            if (bvh[i].tri_count == 0) { walk.push_back(bvh[i].left_child); walk.push_back(bvh[i].right_child); }
        }
    }
    out.root_extent = root_extent;

    out.wide_nodes.assign(n_internal, WideNode());
    out.leaf_recs.assign(n_leaf, LeafRec());
    for (uint32_t i = 0; i < n_nodes; i++) {
        const gdpt_bvh_node &n = bvh[i];
        if (n.tri_count > 0) {
            LeafRec &l = out.leaf_recs[link[i] & LINK_INDEX_MASK];
            l.first_tri = n.first_tri_index; l.tri_count = n.tri_count; l.orig = i; l.pad = 0;
            continue;
        }
        WideNode &w = out.wide_nodes[link[i]];
        std::memset(&w, 0, sizeof(w));
        const gdpt_bvh_node &L = bvh[n.left_child], &R = bvh[n.right_child];
        // nodes that no instance reaches keep an "everything" culling box
        const TightBox tl = state[n.left_child] == 2 ? tight_inflate(raw[n.left_child], owner_extent[n.left_child]) : tight_inflate(tight_empty(), 0.0f);
        const TightBox tr = state[n.right_child] == 2 ? tight_inflate(raw[n.right_child], owner_extent[n.right_child]) : tight_inflate(tight_empty(), 0.0f);
        for (int k = 0; k < 3; k++) {
            w.lmin[k] = L.aabb_min[k]; w.lmax[k] = L.aabb_max[k];
            w.rmin[k] = R.aabb_min[k]; w.rmax[k] = R.aabb_max[k];
            w.tlmin[k] = tl.lo[k]; w.tlmax[k] = tl.hi[k];
            w.trmin[k] = tr.lo[k]; w.trmax[k] = tr.hi[k];
        }
        w.left = link[n.left_child]; w.right = link[n.right_child]; w.orig = i;
    }

    // ---- TLAS: same split; leaves index the instance table directly
    std::vector<uint32_t> tlink(n_tlas);
    uint32_t nt_internal = 0;
    for (uint32_t i = 0; i < n_tlas; i++) {
        if (tlas[i].left_right == 0) {
            if (tlas[i].blas >= n_blas) {
                std::snprintf(msg, sizeof(msg), "TLAS leaf %u names instance %u of %u", i, tlas[i].blas, n_blas);
                return msg;
            }
            tlink[i] = LINK_TLAS | LINK_LEAF | tlas[i].blas;
        } else {
            const uint32_t l = tlas[i].left_right & 0xFFFFu, r = tlas[i].left_right >> 16;
            if (l >= n_tlas || r >= n_tlas) {
                std::snprintf(msg, sizeof(msg), "TLAS node %u has a child out of range", i);
                return msg;
            }
            tlink[i] = LINK_TLAS | nt_internal++;
        }
    }
    out.wide_tlas.assign(nt_internal, WideNode());
    out.inst_recs.assign(n_blas, InstRec());
    std::vector<TightBox> world(n_blas);
    for (uint32_t b = 0; b < n_blas; b++) {
        InstRec &r = out.inst_recs[b];
        std::memset(&r, 0, sizeof(r));
        std::memcpy(r.inv, blas[b].inverse_transform, sizeof(r.inv));
        r.root_link = link[blas[b].root];
        r.root_orig = blas[b].root;
        r.fast_root = LINK_NONE; // filled by the closest-hit builder (fast_bvh.h)
        const TightBox obj = tight_inflate(raw[blas[b].root], owner_extent[blas[b].root]);
        for (int k = 0; k < 3; k++) { r.tight_min[k] = obj.lo[k]; r.tight_max[k] = obj.hi[k]; }
        r.tight_max[3] = fast_reach(raw[blas[b].root], owner_extent[blas[b].root]);
        // world-space culling box: the eight transformed corners of the object-space one, inflated again
        TightBox w = tight_empty();
        const float *m = blas[b].transform;
        for (int c = 0; c < 8; c++) {
            const double x = (c & 1) ? obj.hi[0] : obj.lo[0], y = (c & 2) ? obj.hi[1] : obj.lo[1], z = (c & 4) ? obj.hi[2] : obj.lo[2];
            const float p[3] = { (float)(m[0] * x + m[4] * y + m[8] * z + m[12]), (float)(m[1] * x + m[5] * y + m[9] * z + m[13]),
                                 (float)(m[2] * x + m[6] * y + m[10] * z + m[14]) };
            tight_grow(w, p);
        }
        const float eff_w = std::fmax(tight_extent(w), needed_world * (GDPT_MARGIN_DIV / 131072.0f)); // see root_extent above
        world[b] = tight_inflate(w, eff_w);
        const float wr = fast_reach(w, eff_w); // world level: the smallest instance decides
        out.world_reach = (b == 0 || wr < out.world_reach) ? wr : out.world_reach;
    }
    // tight boxes of TLAS nodes, bottom-up (TLAS nodes are few: plain recursion-free fixpoint by depth)
    std::vector<TightBox> traw(n_tlas, tight_empty());
    std::vector<uint8_t> tdone(n_tlas, 0);
    for (uint32_t i = 0; i < n_tlas; i++)
        if (tlas[i].left_right == 0) { traw[i] = world[tlas[i].blas]; tdone[i] = 1; }
    for (uint32_t pass = 0; pass < n_tlas; pass++) {
        bool progress = false, all = true;
        for (uint32_t i = 0; i < n_tlas; i++) {
            if (tdone[i]) continue;
            const uint32_t l = tlas[i].left_right & 0xFFFFu, r = tlas[i].left_right >> 16;
            if (tdone[l] && tdone[r]) { traw[i] = traw[l]; tight_merge(traw[i], traw[r]); tdone[i] = 1; progress = true; }
            else all = false;
        }
        if (all) break;
        if (!progress) return "TLAS child links form a cycle";
    }
    for (uint32_t i = 0; i < n_tlas; i++) {
        const gdpt_tlas_node &n = tlas[i];
        if (n.left_right == 0) {
            if (i > 0) out.inst_recs[n.blas].tlas_orig = i; // leaves sit at 1..I in instance order (bvh.cpp:278-287)
            continue;
        }
        const uint32_t l = n.left_right & 0xFFFFu, r = n.left_right >> 16;
        WideNode &w = out.wide_tlas[tlink[i] & LINK_INDEX_MASK];
        std::memset(&w, 0, sizeof(w));
        for (int k = 0; k < 3; k++) {
            w.lmin[k] = tlas[l].aabb_min[k]; w.lmax[k] = tlas[l].aabb_max[k];
            w.rmin[k] = tlas[r].aabb_min[k]; w.rmax[k] = tlas[r].aabb_max[k];
            w.tlmin[k] = traw[l].lo[k]; w.tlmax[k] = traw[l].hi[k];
            w.trmin[k] = traw[r].lo[k]; w.trmax[k] = traw[r].hi[k];
        }
        w.left = tlink[l]; w.right = tlink[r]; w.orig = i;
    }
    // node 0 is a copy of the final root (bvh.cpp:316); with a single instance that root is
    // itself the leaf, and the traversal pops it as node 0
    if (tlas[0].left_right == 0) out.inst_recs[tlas[0].blas].tlas_orig = 0;
    out.tlas_root_link = tlink[0];
    return "";
}

} // namespace gdpt
#endif
