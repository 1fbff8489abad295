// pt_trace.cuh -- per-ray TLAS/BLAS traversal and ray-triangle intersection.
//
// Semantics are those of ray_trace_tlas / ray_trace_blas / intersectAABB /
// intersectTriangle (main.glsl:224-350) on the reference node arrays: the same
// nodes are visited in the same order, every comparison has the same direction
// and every float op is the same single IEEE operation.  What differs is the
// machinery:
//   * one unified stack for both levels (entries are pre-encoded links, see
//     pt_scene.cuh); the entry that the reference would pop next is kept in a
//     register (`cur`) instead of being pushed and popped again;
//   * the traversal is a resumable state machine (one node per step) so a warp
//     can interleave refilling idle lanes with new rays between steps;
//   * child boxes come from the parent's 64 B WideNode.
// Reference pop order: after testing children (d1 = left, d2 = right) it pushes
// the farther valid child first, so the nearer valid child is popped next; on
// d1 == d2 the RIGHT child is popped first (main.glsl:293-299).
#ifndef GDPT_PT_TRACE_CUH
#define GDPT_PT_TRACE_CUH

#include "pt_math.cuh"
#include "pt_scene.cuh"

namespace gdpt {

#if defined(__CUDACC__)
typedef float4 q4f;
typedef uint4 q4u;
#else
struct alignas(16) q4f { float x, y, z, w; };
struct alignas(16) q4u { uint32_t x, y, z, w; };
#endif

GDPT_HD q4f ldq(const void *base, uint32_t quad_index)
{
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const float4 *>(base) + quad_index);
#else
    return reinterpret_cast<const q4f *>(base)[quad_index];
#endif
}
GDPT_HD q4u ldqu(const void *base, uint32_t quad_index)
{
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4 *>(base) + quad_index);
#else
    return reinterpret_cast<const q4u *>(base)[quad_index];
#endif
}

#define GDPT_NO_INSTANCE 0xFFFFFFFFu
#define GDPT_FRONT_BIT 0x80000000u
// RayState.blas_front = hitInfo.blas | (instance whose local ray the accepted triangle test ran with) << 15 | front << 31.
// The two instances differ only after an equal-t tie across instances: the later triangle test overwrites
// triangle / position / out_dir / front (main.glsl:247-255), but hitInfo.blas changes only when t was lowered
// (main.glsl:322-325), so the reference shades a point of one instance's space with the other instance's transform and
// materials (main.glsl:194-222).  Instance ids fit 15 bits: a TLAS holds at most 65 535 nodes (leftRight, main.glsl:329-330).
#define GDPT_HIT_INST_BITS 15u
#define GDPT_HIT_INST_MASK 0x7FFFu
GDPT_HD uint32_t hit_blas(uint32_t blas_front) { return blas_front & GDPT_HIT_INST_MASK; }
GDPT_HD uint32_t hit_space(uint32_t blas_front) { return (blas_front >> GDPT_HIT_INST_BITS) & GDPT_HIT_INST_MASK; }
#define GDPT_MAX_STACK 128 /* 64 + 64 of the reference (main.glsl:272,307) */

// Optional per-ray parity observables.
struct TraceCounters {
    uint32_t node_pops, box_tests, tri_tests, tlas_leaves, max_stack;
    uint64_t hash;
    uint32_t level_base;     // stack fill when the current instance was entered
    uint32_t *visits; uint32_t visits_cap;
};

GDPT_HD void counters_init(TraceCounters &c, uint32_t *visits, uint32_t cap)
{
    c.node_pops = c.box_tests = c.tri_tests = c.tlas_leaves = 0;
    c.max_stack = 1; // the TLAS root push (main.glsl:309)
    c.hash = GDPT_FNV64_OFFSET;
    c.level_base = 0;
    c.visits = visits; c.visits_cap = cap;
}
GDPT_HD void counters_visit(TraceCounters &c, uint32_t id)
{
    if (c.visits && c.node_pops < c.visits_cap) c.visits[c.node_pops] = id;
    c.node_pops++;
#pragma unroll
    for (int b = 0; b < 4; b++) { c.hash ^= (uint64_t)((id >> (8 * b)) & 0xffu); c.hash *= GDPT_FNV64_PRIME; }
}

struct RayState {
    f3 wo, wd;           // world-space ray (ray.o, ray.d)
    f3 o, d, rd;         // ray in the space being traversed (world or instance-local)
    float t, u, v;       // hitInfo.t, barycentrics
    uint32_t tri;        // hitInfo.triangle
    uint32_t blas_front; // hitInfo.blas | space of the accepted test << 15 | front << 31 (hit_blas / hit_space)
    uint32_t cur;        // link to process next, LINK_NONE when the traversal is over
    uint32_t sp;         // unified stack fill
    uint32_t inst;       // instance being traversed or GDPT_NO_INSTANCE
    uint32_t overflow;   // set if the reference's 64+64 stack entries were exceeded
};

GDPT_HD void ray_begin(RayState &r, const SceneView &sc, f3 o, f3 d)
{
    r.wo = o; r.wd = d;
    r.o = o; r.d = d; r.rd = rcp3(d);
    r.t = 1e9f; r.u = 0.0f; r.v = 0.0f; r.tri = 0u; r.blas_front = 0u;
    r.cur = sc.tlas_root_link; r.sp = 0u; r.inst = GDPT_NO_INSTANCE; r.overflow = 0u;
}

// intersectAABB (main.glsl:259-268)
GDPT_HD float slab_test(const RayState &r, float nx, float ny, float nz, float xx, float xy, float xz)
{
    const float tx1 = (nx - r.o.x) * r.rd.x, tx2 = (xx - r.o.x) * r.rd.x;
    float tmin = min_num(tx1, tx2), tmax = max_num(tx1, tx2);
    const float ty1 = (ny - r.o.y) * r.rd.y, ty2 = (xy - r.o.y) * r.rd.y;
    tmin = max_num(tmin, min_num(ty1, ty2)); tmax = min_num(tmax, max_num(ty1, ty2));
    const float tz1 = (nz - r.o.z) * r.rd.z, tz2 = (xz - r.o.z) * r.rd.z;
    tmin = max_num(tmin, min_num(tz1, tz2)); tmax = min_num(tmax, max_num(tz1, tz2));
    return (tmax >= tmin && tmax > 0.0f) ? tmin : 1e30f;
}

// intersectTriangle (main.glsl:224-257).  `position`/`out_dir` are not stored:
// the shading stage recomputes them from (ray, t) with the same operations.
GDPT_HD void triangle_test_loaded(RayState &r, uint32_t tri_index, const q4f a, const q4f b, const q4f c)
{
    const f3 v0 = mk3(a.x, a.y, a.z);
    const f3 e1 = mk3(b.x, b.y, b.z) - v0, e2 = mk3(c.x, c.y, c.z) - v0;
    const f3 pvec = cross3(r.d, e2);
    const float det = dot3(e1, pvec);
    if (fabsf(det) < 1e-5f) return;
    const float inv_det = 1.0f / det;
    const f3 tvec = r.o - v0;
    const float u = dot3(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return;
    const f3 qvec = cross3(tvec, e1);
    const float v = dot3(r.d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return;
    const float t = dot3(e2, qvec) * inv_det;
    if (t < 0.0f || t > r.t) return;
    // hitInfo.blas is assigned after the BLAS returns iff it lowered t (main.glsl:322-325);
    // since minT == hitInfo.t on entry, that is exactly "a strictly closer triangle was accepted".
    const uint32_t blas = (t < r.t) ? r.inst : hit_blas(r.blas_front);
    const uint32_t front = dot3(cross3(e1, e2), r.d) > 0.0f ? GDPT_FRONT_BIT : 0u;
    r.t = t; r.u = u; r.v = v; r.tri = tri_index; r.blas_front = blas | (r.inst << GDPT_HIT_INST_BITS) | front;
}
GDPT_HD void triangle_test(const SceneView &sc, RayState &r, uint32_t tri_index)
{
    const q4f a = ldq(sc.tri_geom, tri_index * 3u + 0u);
    const q4f b = ldq(sc.tri_geom, tri_index * 3u + 1u);
    const q4f c = ldq(sc.tri_geom, tri_index * 3u + 2u);
    triangle_test_loaded(r, tri_index, a, b, c);
}

template <class Stack> GDPT_HD void stack_push(RayState &r, Stack &st, uint32_t link)
{
    if (r.sp < GDPT_MAX_STACK) st.store(r.sp, link); else r.overflow |= 1u;
    r.sp++;
}
template <class Stack> GDPT_HD uint32_t stack_pop(RayState &r, Stack &st)
{
    if (r.sp == 0u) return LINK_NONE;
    r.sp--;
    return (r.sp < GDPT_MAX_STACK) ? st.load(r.sp) : LINK_NONE;
}

// Shared tail of an internal-node visit: children tests done, pick what comes next.
// keep_l / keep_r: false when the culling box of that child is missed (always true without CULL).
template <bool TRACE, class Stack>
GDPT_HD void choose_next(RayState &r, Stack &st, float d1, float d2, uint32_t left, uint32_t right, bool keep_l, bool keep_r,
                         TraceCounters *tc)
{
    const bool lv = d1 < r.t && keep_l, rv = d2 < r.t && keep_r;
    const bool left_first = d1 < d2;
    const uint32_t first = left_first ? left : right, second = left_first ? right : left;
    const bool fv = left_first ? lv : rv, sv = left_first ? rv : lv;
    if (TRACE) {
        tc->box_tests += 2u;
        const uint32_t fill = (r.sp - tc->level_base) + (fv ? 1u : 0u) + (sv ? 1u : 0u);
        if (fill > tc->max_stack) tc->max_stack = fill;
    }
    if (fv) {
        if (sv) stack_push(r, st, second);
        r.cur = first;
    } else if (sv) {
        r.cur = second;
    } else {
        r.cur = stack_pop(r, st);
    }
}

// Can the subtree inside this (inflated) culling box still change the hit?  Not if the ray misses
// the box, and not if it enters the box beyond the current hit.t: every triangle in there would be
// rejected by `t > hit.t` (main.glsl:247), and hit.t only ever decreases.  Same slab arithmetic;
// NaNs from 0 * inf are dropped by minNum/maxNum, which errs on the side of "touches".
GDPT_HD bool slab_touches(const RayState &r, float nx, float ny, float nz, float xx, float xy, float xz)
{
    const float tx1 = (nx - r.o.x) * r.rd.x, tx2 = (xx - r.o.x) * r.rd.x;
    float tmin = min_num(tx1, tx2), tmax = max_num(tx1, tx2);
    const float ty1 = (ny - r.o.y) * r.rd.y, ty2 = (xy - r.o.y) * r.rd.y;
    tmin = max_num(tmin, min_num(ty1, ty2)); tmax = min_num(tmax, max_num(ty1, ty2));
    const float tz1 = (nz - r.o.z) * r.rd.z, tz2 = (xz - r.o.z) * r.rd.z;
    tmin = max_num(tmin, min_num(tz1, tz2)); tmax = min_num(tmax, max_num(tz1, tz2));
    return !(tmax < tmin) && !(tmax < 0.0f) && !(tmin > r.t);
}

// The two children of a WideNode record: reference distances, and (CULL) whether each child's
// culling box is touched.
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void visit_wide(const void *table, uint32_t idx, uint32_t tag, RayState &r, Stack &st, TraceCounters *tc)
{
    const q4f q0 = ldq(table, idx * 8u + 0u);
    const q4f q1 = ldq(table, idx * 8u + 1u);
    const q4f q2 = ldq(table, idx * 8u + 2u);
    const q4u q3 = ldqu(table, idx * 8u + 3u);
    if (TRACE) counters_visit(*tc, q3.z | tag);
    const float d1 = slab_test(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
    const float d2 = slab_test(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
    bool keep_l = true, keep_r = true;
    if (CULL) {
        const q4f q4 = ldq(table, idx * 8u + 4u);
        const q4f q5 = ldq(table, idx * 8u + 5u);
        const q4f q6 = ldq(table, idx * 8u + 6u);
        keep_l = slab_touches(r, q4.x, q4.y, q4.z, q4.w, q5.x, q5.y);
        keep_r = slab_touches(r, q5.z, q5.w, q6.x, q6.y, q6.z, q6.w);
        if (TRACE) tc->box_tests += 2u; // executed-work accounting: the two extra box tests
    }
    choose_next<TRACE>(r, st, d1, d2, q3.x, q3.y, keep_l, keep_r, tc);
}

// One BLAS internal node (main.glsl:285-300).
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void step_blas_internal(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    visit_wide<TRACE, CULL>(sc.wide_nodes, r.cur & LINK_INDEX_MASK, 0u, r, st, tc);
}

// One BLAS leaf (main.glsl:280-284).
template <bool TRACE, class Stack>
GDPT_HD void step_blas_leaf(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    const q4u leaf = ldqu(sc.leaf_recs, r.cur & LINK_INDEX_MASK);
    if (TRACE) { counters_visit(*tc, leaf.z); tc->tri_tests += leaf.y; }
    for (uint32_t i = 0; i < leaf.y; i++) triangle_test(sc, r, leaf.x + i);
    r.cur = stack_pop(r, st);
}

// Leaf handling split into single-triangle steps, for schedulers that interleave lanes:
// entering the leaf (visit + range fetch + next link) is merged with its first
// triangle test; later calls test one more triangle of the pending range.  The
// next link sits in `cur` but is not processed before the range is exhausted, so
// every test still sees the hit.t the reference would have at that point.
template <bool TRACE, class Stack>
GDPT_HD void step_blas_leaf_one(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc, uint32_t &tri_next,
                                uint32_t &tri_end)
{
    if (tri_next == tri_end) {
        const q4u leaf = ldqu(sc.leaf_recs, r.cur & LINK_INDEX_MASK);
        if (TRACE) { counters_visit(*tc, leaf.z); tc->tri_tests += leaf.y; }
        tri_next = leaf.x; tri_end = leaf.x + leaf.y;
        r.cur = stack_pop(r, st);
    }
    if (tri_next < tri_end) triangle_test(sc, r, tri_next++);
}

// Same, up to N triangles per step: their vertex loads are issued together, the tests then run
// in index order on the running hit.t, so the outcome is that of N single steps.
template <bool TRACE, int N, class Stack>
GDPT_HD void step_blas_leaf_some(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc, uint32_t &tri_next,
                                 uint32_t &tri_end)
{
    if (tri_next == tri_end) {
        const q4u leaf = ldqu(sc.leaf_recs, r.cur & LINK_INDEX_MASK);
        if (TRACE) { counters_visit(*tc, leaf.z); tc->tri_tests += leaf.y; }
        tri_next = leaf.x; tri_end = leaf.x + leaf.y;
        r.cur = stack_pop(r, st);
    }
    q4f va[N], vb[N], vc[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t ti = (tri_next + (uint32_t)i < tri_end) ? tri_next + (uint32_t)i : tri_next; // clamp: a valid address
        va[i] = ldq(sc.tri_geom, ti * 3u + 0u); vb[i] = ldq(sc.tri_geom, ti * 3u + 1u); vc[i] = ldq(sc.tri_geom, ti * 3u + 2u);
    }
#pragma unroll
    for (int i = 0; i < N; i++)
        if (tri_next + (uint32_t)i < tri_end) triangle_test_loaded(r, tri_next + (uint32_t)i, va[i], vb[i], vc[i]);
    tri_next = (tri_end - tri_next > (uint32_t)N) ? tri_next + (uint32_t)N : tri_end;
}

// Back to world space after an instance (main.glsl:316-327: the TLAS loop keeps using `ray`).
template <bool TRACE>
GDPT_HD void leave_instance(RayState &r, TraceCounters *tc)
{
    r.o = r.wo; r.d = r.wd; r.rd = rcp3(r.wd);
    r.inst = GDPT_NO_INSTANCE;
    if (TRACE) tc->level_base = 0u;
}

// r.cur is a TLAS leaf and the ray is in world space: enter the instance (main.glsl:316-323).
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void enter_instance(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    const uint32_t idx = r.cur & LINK_INDEX_MASK;
    const q4f c0 = ldq(sc.inst_recs, idx * 7u + 0u);
    const q4f c1 = ldq(sc.inst_recs, idx * 7u + 1u);
    const q4f c2 = ldq(sc.inst_recs, idx * 7u + 2u);
    const q4f c3 = ldq(sc.inst_recs, idx * 7u + 3u);
    const q4u tail = ldqu(sc.inst_recs, idx * 7u + 4u);
    if (TRACE) {
        counters_visit(*tc, tail.y | GDPT_VISIT_TLAS_TAG);
        tc->tlas_leaves++;
        tc->level_base = r.sp; // BLAS stack starts empty; its root push counts 1 (<= max_stack init)
    }
    // b_ray = inverse_transform * (ray.o, 1), * (ray.d, 0); d is NOT renormalised, so t is shared
    r.o = mk3(((c0.x * r.wo.x + c1.x * r.wo.y) + c2.x * r.wo.z) + c3.x * 1.0f,
              ((c0.y * r.wo.x + c1.y * r.wo.y) + c2.y * r.wo.z) + c3.y * 1.0f,
              ((c0.z * r.wo.x + c1.z * r.wo.y) + c2.z * r.wo.z) + c3.z * 1.0f);
    r.d = mk3(((c0.x * r.wd.x + c1.x * r.wd.y) + c2.x * r.wd.z) + c3.x * 0.0f,
              ((c0.y * r.wd.x + c1.y * r.wd.y) + c2.y * r.wd.z) + c3.y * 0.0f,
              ((c0.z * r.wd.x + c1.z * r.wd.y) + c2.z * r.wd.z) + c3.z * 0.0f);
    r.rd = rcp3(r.d);
    r.inst = idx;
    r.cur = tail.x;
    if (CULL) {
        // the BLAS root is never box-tested upstream (main.glsl:274); with culling, an instance whose
        // true bounds the local ray misses is left again at once (nothing in it can be hit)
        const q4f tmin4 = ldq(sc.inst_recs, idx * 7u + 5u);
        const q4f tmax4 = ldq(sc.inst_recs, idx * 7u + 6u);
        if (TRACE) tc->box_tests += 1u;
        if (!slab_touches(r, tmin4.x, tmin4.y, tmin4.z, tmax4.x, tmax4.y, tmax4.z)) r.cur = stack_pop(r, st);
    }
}

// One TLAS-level entry: leave the instance we were in (if any), then either
// enter an instance (main.glsl:316-323) or test the two TLAS children (:330-346).
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void step_tlas(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    if (r.inst != GDPT_NO_INSTANCE) leave_instance<TRACE>(r, tc);
    if (r.cur & LINK_LEAF) { enter_instance<TRACE, CULL>(sc, r, st, tc); return; }
    visit_wide<TRACE, CULL>(sc.wide_tlas, r.cur & LINK_INDEX_MASK, GDPT_VISIT_TLAS_TAG, r, st, tc);
}

// The compact scheduler's split of the same work: `step_node` takes any internal node -- BLAS, or
// TLAS while the ray is in world space -- through ONE copy of the box code; `step_instance` does the
// space changes (leave and/or enter) and leaves a TLAS-internal link for the next `step_node`.
GDPT_HD bool link_is_node_step(uint32_t l, uint32_t inst)
{
    return (l & LINK_LEAF) == 0u && l != LINK_NONE && ((l & LINK_TLAS) == 0u || inst == GDPT_NO_INSTANCE);
}
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void step_node(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    const bool top = (r.cur & LINK_TLAS) != 0u;
    const void *table = top ? static_cast<const void *>(sc.wide_tlas) : static_cast<const void *>(sc.wide_nodes);
    visit_wide<TRACE, CULL>(table, r.cur & LINK_INDEX_MASK, top ? GDPT_VISIT_TLAS_TAG : 0u, r, st, tc);
}
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void step_instance(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    if (r.inst != GDPT_NO_INSTANCE) leave_instance<TRACE>(r, tc);
    if (r.cur & LINK_LEAF) enter_instance<TRACE, CULL>(sc, r, st, tc);
}

GDPT_HD bool link_is_blas_internal(uint32_t l) { return (l & (LINK_TLAS | LINK_LEAF)) == 0u; }
GDPT_HD bool link_is_blas_leaf(uint32_t l) { return (l & (LINK_TLAS | LINK_LEAF)) == LINK_LEAF; }

// Whole traversal of one ray, one node per iteration (used where no warp-level
// scheduling is wanted, and by the host-side unit check).
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void trace_ray(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    while (r.cur != LINK_NONE) {
        if (link_is_blas_internal(r.cur)) step_blas_internal<TRACE, CULL>(sc, r, st, tc);
        else if (link_is_blas_leaf(r.cur)) step_blas_leaf<TRACE>(sc, r, st, tc);
        else step_tlas<TRACE, CULL>(sc, r, st, tc);
    }
}

// Same traversal with leaves taken one triangle per step (the order a phase-voting warp uses).
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void trace_ray_stepwise(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    uint32_t tri_next = 0, tri_end = 0;
    while (r.cur != LINK_NONE || tri_next < tri_end) {
        if (tri_next < tri_end || link_is_blas_leaf(r.cur)) step_blas_leaf_one<TRACE>(sc, r, st, tc, tri_next, tri_end);
        else if (link_is_blas_internal(r.cur)) step_blas_internal<TRACE, CULL>(sc, r, st, tc);
        else step_tlas<TRACE, CULL>(sc, r, st, tc);
    }
}

// The order the compact path kernel uses: node steps / instance steps / single-triangle steps.
template <bool TRACE, bool CULL, class Stack>
GDPT_HD void trace_ray_compact(const SceneView &sc, RayState &r, Stack &st, TraceCounters *tc)
{
    uint32_t tri_next = 0, tri_end = 0;
    while (r.cur != LINK_NONE || tri_next < tri_end) {
        if (tri_next < tri_end || link_is_blas_leaf(r.cur)) step_blas_leaf_one<TRACE>(sc, r, st, tc, tri_next, tri_end);
        else if (link_is_node_step(r.cur, r.inst)) step_node<TRACE, CULL>(sc, r, st, tc);
        else step_instance<TRACE, CULL>(sc, r, st, tc);
    }
}

} // namespace gdpt
#endif
