// pt_scene.cuh -- device-side scene view and the derived ("wide") traversal
// layout the backend builds from the uploaded reference arrays when the binding
// table is finalised.
//
// Why a derived layout: the reference reads, per internal-node pop, its own
// 48 B record plus BOTH children's 48 B records to get their boxes
// (main.glsl:277,285-289), and re-reads a child when it is popped.  For
// incoherent rays every such read is its own L1 wavefront, which bounds
// traversal long before issue slots do.  We keep node identity and visiting
// order exactly as in the reference, but store, per INTERNAL node, both child
// boxes and both child links in one 128 B record (one cache line), and per LEAF
// a 16 B descriptor.  Child links are pre-encoded so a popped entry says which
// table it indexes without another load.
//
//   link bits: [31] TLAS level   [30] leaf   [29:0] index into the table
//     BLAS internal -> wide_nodes[]     BLAS leaf -> leaf_recs[]
//     TLAS internal -> wide_tlas[]      TLAS leaf -> inst_recs[]
//
// `orig` fields carry the reference node index for the parity trace.
//
// Tight boxes (culling).  The reference's boxes all contain the planes y=0, z=0
// of their mesh (quirk Q1, bvh.cpp:6-10), so a ray near those planes walks
// thousands of nodes whose triangles it cannot hit.  Next to each reference box
// we keep the TRUE bounding box of the child's triangles, inflated by a safety
// margin.  A child whose tight box the ray misses cannot contain an accepted
// triangle hit, so skipping it changes neither hit.t nor any later decision:
// ordering (d1 < d2) and pruning (d < hit.t) of the surviving children still use
// the reference boxes, and the surviving nodes are visited in reference order.
// Hit records and images are bit-identical; only the visit list gets shorter.
// Trace/parity mode does not cull (the full reference visit order is observable
// there).
#ifndef GDPT_PT_SCENE_CUH
#define GDPT_PT_SCENE_CUH

#include "gdpt_wire.h"
#include <stdint.h>

namespace gdpt {

enum : uint32_t {
    LINK_TLAS = 0x80000000u,
    LINK_LEAF = 0x40000000u,
    LINK_INDEX_MASK = 0x3FFFFFFFu,
    LINK_NONE = 0xFFFFFFFFu
};

// 128 B, 128 B-aligned: eight 128-bit quads.
//   q0 = Lmin.x Lmin.y Lmin.z Lmax.x   q1 = Lmax.y Lmax.z Rmin.x Rmin.y
//   q2 = Rmin.z Rmax.x Rmax.y Rmax.z   q3 = left_link right_link orig_id 0
//   q4..q6 = the same 12 floats for the tight (culling) boxes   q7 = padding
struct __attribute__((aligned(128))) WideNode {
    float lmin[3]; float lmax[3];
    float rmin[3]; float rmax[3];
    uint32_t left, right, orig, pad;
    float tlmin[3]; float tlmax[3];
    float trmin[3]; float trmax[3];
    uint32_t pad2[4];
};
static_assert(sizeof(WideNode) == 128, "WideNode is one cache line");

struct __attribute__((aligned(16))) LeafRec {
    uint32_t first_tri, tri_count, orig, pad;
};
static_assert(sizeof(LeafRec) == 16, "LeafRec is one 128-bit load");

// Per instance: what a TLAS-leaf visit needs (main.glsl:316-323).
struct __attribute__((aligned(16))) InstRec {
    float inv[16];          // BLASInstance.inverse_transform, column-major
    uint32_t root_link;     // pre-encoded link of BLASInstance.root
    uint32_t tlas_orig;     // reference TLAS node index of this leaf
    uint32_t root_orig;     // reference BVH node index of the root
    uint32_t fast_root;     // link of this BLAS's root in the closest-hit tables (fast_bvh.h)
    float tight_min[4];     // tight box of the whole BLAS, object space; w = bits of this BLAS's root link in the four-wide table
    float tight_max[4];     // w = largest |local origin coordinate| for which the search's margins are trusted in this BLAS
};
static_assert(sizeof(InstRec) == 112, "InstRec is seven 128-bit loads");

// ---- closest-hit tables (fast_bvh.h / pt_fast.cuh) -------------------------------------------
// 64 B, four 128-bit quads: both children's true (inflated) boxes and their links.
//   q0 = Lmin.x Lmin.y Lmin.z Lmax.x   q1 = Lmax.y Lmax.z Rmin.x Rmin.y
//   q2 = Rmin.z Rmax.x Rmax.y Rmax.z   q3 = left right - -
// Links: [31] TLAS level, [30] leaf.  BLAS leaf = LINK_LEAF | (count-1) << 27 | first FastTri;
// TLAS leaf = LINK_TLAS | LINK_LEAF | instance; internal = index into fast_nodes (TLAS nodes at fast_tlas_base + index).
struct __attribute__((aligned(64))) FastNode {
    float lmin[3]; float lmax[3];
    float rmin[3]; float rmax[3];
    uint32_t left, right, pad[2];
};
static_assert(sizeof(FastNode) == 64, "FastNode is four 128-bit loads");
// 48 B: the reference's vertices (bit copies) and the triangle's index in the reference arrays.
struct __attribute__((aligned(16))) FastTri {
    float v0[3]; uint32_t orig;
    float v1[3]; uint32_t pad1;
    float v2[3]; uint32_t pad2;
};
static_assert(sizeof(FastTri) == 48, "FastTri is three 128-bit loads");
// 128 B, one cache line, eight 128-bit quads: FOUR children (the two-wide tree collapsed by one level where that
// pays, fast_bvh.h).  q0..q5 = lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] (true boxes, inflated), q6 = links
// (LINK_NONE = no child), q7 unused.  Same link encoding; TLAS-level nodes live in the same table and keep LINK_TLAS.
struct __attribute__((aligned(128))) FastNode4 {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    uint32_t link[4], pad[4];
};
static_assert(sizeof(FastNode4) == 128, "FastNode4 is one cache line");
#define FAST_LEAF_COUNT_SHIFT 27
#define FAST_LEAF_FIRST_MASK 0x07FFFFFFu
#define GDPT_FAST_MAX_DEPTH 120u

struct SceneView {
    // uploaded reference buffers (set 1 bindings 0..5, set 2 binding 0)
    const gdpt_triangle_geometry *tri_geom;
    const gdpt_triangle_data *tri_data;
    const gdpt_material *materials;
    const gdpt_bvh_node *bvh;
    const gdpt_blas_instance *blas;
    const gdpt_tlas_node *tlas;
    const uint8_t *textures;
    int32_t tex_w, tex_h, tex_layers;
    uint32_t n_tris, n_nodes, n_blas, n_tlas, n_materials;
    // derived
    const WideNode *wide_nodes;
    const LeafRec *leaf_recs;
    const WideNode *wide_tlas;
    const InstRec *inst_recs;
    uint32_t tlas_root_link;
    // closest-hit tables; fast_ok == 0 means "use the reference-order traversal only"
    const FastNode *fast_nodes;
    uint32_t fast_tlas_base;      // TLAS internal node i is fast_nodes[fast_tlas_base + i]
    const FastTri *fast_tris;
    const uint32_t *tri_leaf;
    uint32_t fast_ok;
    // four-wide form of the same trees (fast4_ok == 0: not built / too deep for the stack bound)
    const FastNode4 *fast4;
    uint32_t fast4_root;          // link the search starts from (TLAS root)
    uint32_t fast4_ok;
    // largest |origin coordinate| (world space) for which the search's culling margins are trusted; rays from farther
    // out are answered by the exact traversal (derived_layout.h fast_reach)
    float fast_world_reach;
    // "#define GDPT_MATERIAL_EXT" (gdpt_wire.h): roughness / metallic textures, sRGB albedo layers, per-instance material
    // tables of any length.  material_ext == 0: the reference's shading data, bit for bit
    uint32_t material_ext;
    const uint32_t *surface_materials; // set 1 binding 6 or null
    const float *srgb_lut;             // 256 entries: sRGB byte -> linear
};

} // namespace gdpt
#endif
