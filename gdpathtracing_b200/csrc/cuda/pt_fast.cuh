// pt_fast.cuh -- closest-hit search over the tables of fast_bvh.h, plus the proof that the
// reference traversal (main.glsl:270-350) returns the same record.
//
// What the reference returns for one ray: among the (instance, triangle) pairs it tests, the one
// with the smallest accepted t; on equal t the pair tested last (main.glsl:247 rejects only t > hit.t).
// Every per-pair quantity -- the instance-local ray (main.glsl:316-321), det/u/v/t/front
// (main.glsl:224-257) -- is a function of the ray and the pair alone, so an order-free search that
// evaluates the same expressions gets the same numbers.  The search here finds the minimum t over ALL
// pairs (its boxes are true bounds with a margin) and records whether a second pair reaches exactly that
// t (`RAY_TIE`).  With a unique minimum at pair w, the reference returns w iff it tests w, because
// nothing it tests can have a smaller or equal t:  hit.t stays > t_w until w is tested, so every box
// on the way to w that is entered at d < t_w passes the reference's `d < hit.t` push test
// (main.glsl:293-299, 339-345).  Reference boxes are nested and the slab arithmetic is monotone in the
// box, so it is enough to check the deepest box of each level: the leaf that holds the triangle and
// the TLAS leaf of its instance (`fast_proves_reference_hit`).  Roots are never box-tested upstream.
// A miss of the complete search is a miss of the reference, which tests a subset of the pairs.
// Rays that fail the proof are re-traced in reference order by the caller.
#ifndef GDPT_PT_FAST_CUH
#define GDPT_PT_FAST_CUH

#include "pt_trace.cuh"

namespace gdpt {

#define RAY_OVERFLOW 1u /* RayState.overflow bit 0: traversal stack exceeded */
#define RAY_TIE 2u      /* RayState.overflow bit 1: a second pair reached the current minimum t (or t was NaN) */
#define RAY_FAR 4u      /* RayState.overflow bit 2: the ray starts beyond the reach of the culling margins: the exact traversal answers
                           it WITHOUT tight-box culling, the search is skipped */
#define RAY_AXIAL 8u    /* RayState.overflow bit 3: a direction component is zero: the exact traversal (culling as usual) answers it,
                           the search is skipped */
#define RAY_UNSEARCHED (RAY_FAR | RAY_AXIAL)

// Stack of the search.  The trees' stack need is bounded at upload (fast_bvh.h: max_depth / need4 < GDPT_FAST_MAX_DEPTH
// < GDPT_MAX_STACK), so no overflow test is needed here -- unlike the reference-order traversal, whose 64+64 entries
// (main.glsl:272,307) are a property of the caller's arrays.
template <class Stack> GDPT_HD void fast_push(RayState &r, Stack &st, uint32_t link)
{
    st.store(r.sp, link);
    r.sp++;
}
template <class Stack> GDPT_HD uint32_t fast_pop(RayState &r, Stack &st)
{
    if (r.sp == 0u) return LINK_NONE;
    r.sp--;
    return st.load(r.sp);
}

// Arithmetic of the SEARCH's box tests only.  The boxes are our own, true bounds inflated by a margin (fast_bvh.h,
// derived_layout.h tight_inflate), so their slab tests need not reproduce any reference number: a fused multiply-add and
// the hardware reciprocal move a plane by a few ulp of the coordinates, the margin is 1/512 of the mesh.  (Rays whose
// origin is so far out that a few ulp reach the margin are sent to the exact traversal: fast_far_origin.)  Everything the
// reference computes -- Moller-Trumbore, the instance-local ray, the proof's slab tests -- keeps the exact operations.
GDPT_HD float fast_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
GDPT_HD float fast_rcp(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); // MUFU.RCP; a denormal direction component counts as zero
    return r;
#else
    return 1.0f / x;
#endif
}
GDPT_HD f3 fast_rcp3(f3 a) { return mk3(fast_rcp(a.x), fast_rcp(a.y), fast_rcp(a.z)); }
// true when an origin coordinate is beyond `reach`: the search's culling margins are not trusted there
GDPT_HD bool fast_far_origin(f3 o, float reach) { return !(max_num(max_num(fabsf(o.x), fabsf(o.y)), fabsf(o.z)) <= reach); }

// A direction component that is zero (or so small that its reciprocal times a coordinate overflows; the hardware
// reciprocal flushes denormals): the four-wide step computes plane * rd - origin * rd, which is inf - inf = NaN on that
// axis, and a NaN drops the axis from the test -- nothing is culled along it, and a ray through a million-triangle soup
// visits thousands of nodes (one such camera ray held a whole C3 frame: 8 000 dependent steps).  The proof refuses these
// rays anyway (0 * inf voids its monotonicity argument), so they skip the search and go straight to the exact traversal.
GDPT_HD bool fast_degenerate_dir(f3 d) { return !(fabsf(d.x) >= 1e-30f && fabsf(d.y) >= 1e-30f && fabsf(d.z) >= 1e-30f); }

GDPT_HD void fast_ray_begin(RayState &r, const SceneView &sc, f3 o, f3 d)
{
    ray_begin(r, sc, o, d);
    r.rd = fast_rcp3(d); // the search's own reciprocal; the proof recomputes the exact one
    if (fast_far_origin(o, sc.fast_world_reach)) r.overflow |= RAY_FAR;
    if (fast_degenerate_dir(d)) r.overflow |= RAY_AXIAL;
}
// Link a search starts from: nothing for a ray the search is not trusted with.
GDPT_HD uint32_t fast_start_link(const RayState &r, uint32_t root) { return (r.overflow & RAY_UNSEARCHED) ? LINK_NONE : root; }

// true box of a child: entry distance and whether the subtree can still hold a hit with t <= r.t
GDPT_HD bool fast_slab(const RayState &r, float nx, float ny, float nz, float xx, float xy, float xz, float *entry)
{
    const float tx1 = (nx - r.o.x) * r.rd.x, tx2 = (xx - r.o.x) * r.rd.x;
    float tmin = min_num(tx1, tx2), tmax = max_num(tx1, tx2);
    const float ty1 = (ny - r.o.y) * r.rd.y, ty2 = (xy - r.o.y) * r.rd.y;
    tmin = max_num(tmin, min_num(ty1, ty2)); tmax = min_num(tmax, max_num(ty1, ty2));
    const float tz1 = (nz - r.o.z) * r.rd.z, tz2 = (xz - r.o.z) * r.rd.z;
    tmin = max_num(tmin, min_num(tz1, tz2)); tmax = min_num(tmax, max_num(tz1, tz2));
    *entry = tmin;
    return !(tmax < tmin) && !(tmax < 0.0f) && !(tmin > r.t);
}

// One internal node of either level: nearer child next, farther child pushed.
template <class Stack> GDPT_HD void fast_step_node(const SceneView &sc, RayState &r, Stack &st)
{
    const void *table = sc.fast_nodes; // both levels in one table, TLAS nodes behind the BLAS nodes
    const uint32_t idx = (r.cur & LINK_INDEX_MASK) + ((r.cur & LINK_TLAS) ? sc.fast_tlas_base : 0u);
    const q4f q0 = ldq(table, idx * 4u + 0u);
    const q4f q1 = ldq(table, idx * 4u + 1u);
    const q4f q2 = ldq(table, idx * 4u + 2u);
    const q4u q3 = ldqu(table, idx * 4u + 3u);
    float dl, dr;
    const bool hl = fast_slab(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &dl);
    const bool hr = fast_slab(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &dr);
    const bool left_first = dl <= dr;
    const uint32_t first = left_first ? q3.x : q3.y, second = left_first ? q3.y : q3.x;
    const bool fv = left_first ? hl : hr, sv = left_first ? hr : hl;
    if (fv) {
        if (sv) fast_push(r, st, second);
        r.cur = first;
    } else if (sv) {
        r.cur = second;
    } else {
        r.cur = fast_pop(r, st);
    }
}

GDPT_HD uint32_t fast_bits(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
// One internal node of the four-wide tables: four true-box tests, children visited nearest first.
// The near and the far plane of every axis are picked by the sign of the ray direction (which quad is loaded), so a
// child costs six FFMA, two FMNMX3 + two FMNMX and one compare:  entry = max(near planes, 0), exit = min(far planes, r.t),
// touched iff entry <= exit.  NaNs (0 * inf, inf - inf on axis-parallel rays) are dropped by fmaxf / fminf, which ignores
// that axis: more children visited, never fewer.
// SORT: all touched children in order of entry distance (key = distance bits with the child number in the two low bits,
// five min/max comparators); otherwise the nearest child next and the others pushed in table order.
template <bool SORT = true, class Stack> GDPT_HD void fast_step_node4(const SceneView &sc, RayState &r, Stack &st)
{
    const uint32_t base = (r.cur & LINK_INDEX_MASK) * 8u;
    const bool sx = r.rd.x < 0.0f, sy = r.rd.y < 0.0f, sz = r.rd.z < 0.0f;
    const q4f nx = ldq(sc.fast4, base + (sx ? 3u : 0u)), ny = ldq(sc.fast4, base + (sy ? 4u : 1u)), nz = ldq(sc.fast4, base + (sz ? 5u : 2u));
    const q4f fx = ldq(sc.fast4, base + (sx ? 0u : 3u)), fy = ldq(sc.fast4, base + (sy ? 1u : 4u)), fz = ldq(sc.fast4, base + (sz ? 2u : 5u));
    const q4u lk = ldqu(sc.fast4, base + 6u);
    const float ox = -(r.o.x * r.rd.x), oy = -(r.o.y * r.rd.y), oz = -(r.o.z * r.rd.z);
    const uint32_t kMiss = 0xFFFFFFFFu;
#define GDPT_CHILD(c, n)                                                                                              \
    uint32_t k##n;                                                                                                    \
    {                                                                                                                 \
        const float e = max_num(max_num(fast_fma(nx.c, r.rd.x, ox), fast_fma(ny.c, r.rd.y, oy)),                      \
                                max_num(fast_fma(nz.c, r.rd.z, oz), 0.0f));                                           \
        const float x = min_num(min_num(fast_fma(fx.c, r.rd.x, ox), fast_fma(fy.c, r.rd.y, oy)),                      \
                                min_num(fast_fma(fz.c, r.rd.z, oz), r.t));                                            \
        k##n = (e <= x && lk.c != LINK_NONE) ? ((fast_bits(e) & ~3u) | (uint32_t)n) : kMiss;                          \
    }
    GDPT_CHILD(x, 0) GDPT_CHILD(y, 1) GDPT_CHILD(z, 2) GDPT_CHILD(w, 3)
#undef GDPT_CHILD
#define GDPT_LINK_OF(k) (((k) & 2u) ? (((k) & 1u) ? lk.w : lk.z) : (((k) & 1u) ? lk.y : lk.x))
    if (SORT) {
#define GDPT_CSWAP(a, b) { const uint32_t lo = a < b ? a : b; b = a < b ? b : a; a = lo; }
        GDPT_CSWAP(k0, k1) GDPT_CSWAP(k2, k3) GDPT_CSWAP(k0, k2) GDPT_CSWAP(k1, k3) GDPT_CSWAP(k1, k2)
#undef GDPT_CSWAP
        if (k3 != kMiss) fast_push(r, st, GDPT_LINK_OF(k3)); // farthest first: the nearest pushed child is popped first
        if (k2 != kMiss) fast_push(r, st, GDPT_LINK_OF(k2));
        if (k1 != kMiss) fast_push(r, st, GDPT_LINK_OF(k1));
        r.cur = (k0 != kMiss) ? GDPT_LINK_OF(k0) : fast_pop(r, st);
    } else {
        const uint32_t a = k0 < k1 ? k0 : k1, b = k2 < k3 ? k2 : k3, best = a < b ? a : b;
        if (k3 != kMiss && k3 != best) fast_push(r, st, lk.w);
        if (k2 != kMiss && k2 != best) fast_push(r, st, lk.z);
        if (k1 != kMiss && k1 != best) fast_push(r, st, lk.y);
        if (k0 != kMiss && k0 != best) fast_push(r, st, lk.x);
        r.cur = (best != kMiss) ? GDPT_LINK_OF(best) : fast_pop(r, st);
    }
#undef GDPT_LINK_OF
}

// intersectTriangle (main.glsl:224-257), same operations as triangle_test_loaded; the running minimum
// replaces hit.t, and a pair that reaches the minimum exactly is recorded instead of accepted.
// Where the accepted pair's u / v / triangle / instance go is the caller's choice (`Sink`): into the RayState, or
// straight into a path slot in shared memory so that a lane carries only t through the search (k_path_pool).
struct HitInRay {
    GDPT_HD void accept(RayState &r, float u, float v, uint32_t tri, uint32_t blas_front) const
    {
        r.u = u; r.v = v; r.tri = tri; r.blas_front = blas_front;
    }
};
template <class Sink> GDPT_HD void fast_triangle_test(RayState &r, const q4f a, const q4f b, const q4f c, const Sink &sink)
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
    if (t < r.t) {
        const uint32_t front = dot3(cross3(e1, e2), r.d) > 0.0f ? GDPT_FRONT_BIT : 0u;
        r.t = t;
        sink.accept(r, u, v, fast_bits(a.w), r.inst | (r.inst << GDPT_HIT_INST_BITS) | front);
        r.overflow &= ~RAY_TIE;
    } else {
        r.overflow |= RAY_TIE; // t == r.t (or NaN): the reference's answer would depend on its visiting order
    }
}
GDPT_HD void fast_triangle_test(RayState &r, const q4f a, const q4f b, const q4f c) { fast_triangle_test(r, a, b, c, HitInRay()); }

// One leaf: 1..8 triangles (fast_bvh.h leaf_max(): 2, or 1 for very large meshes).  GDPT_FAST_LEAF_UNROLL 4 issues all loads together (more registers, 4 copies of
// the test in the instruction stream); 1 keeps one copy of the test in a short loop (instruction-cache friendly).
#ifndef GDPT_FAST_LEAF_UNROLL
#define GDPT_FAST_LEAF_UNROLL 1
#endif
template <class Sink> GDPT_HD void fast_leaf_tests(const SceneView &sc, RayState &r, uint32_t leaf_link, const Sink &sink)
{
    const uint32_t first = leaf_link & FAST_LEAF_FIRST_MASK, count = ((leaf_link >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u;
#if GDPT_FAST_LEAF_UNROLL == 4
    q4f va[4], vb[4], vc[4];
#pragma unroll
    for (uint32_t i = 0; i < 4u; i++) {
        const uint32_t ti = first + (i < count ? i : 0u);
        va[i] = ldq(sc.fast_tris, ti * 3u + 0u); vb[i] = ldq(sc.fast_tris, ti * 3u + 1u); vc[i] = ldq(sc.fast_tris, ti * 3u + 2u);
    }
#pragma unroll
    for (uint32_t i = 0; i < 4u; i++)
        if (i < count) fast_triangle_test(r, va[i], vb[i], vc[i], sink);
#elif GDPT_FAST_LEAF_UNROLL == 2
    // two triangles per trip: six loads in flight, one copy of the test in the instruction stream (called twice)
#pragma unroll 1
    for (uint32_t i = 0; i < count; i += 2u) {
        const uint32_t t0 = first + i, t1 = first + (i + 1u < count ? i + 1u : i);
        const q4f a0 = ldq(sc.fast_tris, t0 * 3u + 0u), b0 = ldq(sc.fast_tris, t0 * 3u + 1u), c0 = ldq(sc.fast_tris, t0 * 3u + 2u);
        const q4f a1 = ldq(sc.fast_tris, t1 * 3u + 0u), b1 = ldq(sc.fast_tris, t1 * 3u + 1u), c1 = ldq(sc.fast_tris, t1 * 3u + 2u);
        fast_triangle_test(r, a0, b0, c0, sink);
        if (i + 1u < count) fast_triangle_test(r, a1, b1, c1, sink);
    }
#else
    // software-pipelined by one triangle: the next triangle's vertices are in flight during the current test
    q4f a = ldq(sc.fast_tris, first * 3u + 0u), b = ldq(sc.fast_tris, first * 3u + 1u), c = ldq(sc.fast_tris, first * 3u + 2u);
#pragma unroll 1
    for (uint32_t i = 0; i < count; i++) {
        const uint32_t tn = first + (i + 1u < count ? i + 1u : i);
        const q4f na = ldq(sc.fast_tris, tn * 3u + 0u), nb = ldq(sc.fast_tris, tn * 3u + 1u), nc = ldq(sc.fast_tris, tn * 3u + 2u);
        fast_triangle_test(r, a, b, c, sink);
        a = na; b = nb; c = nc;
    }
#endif
}
GDPT_HD void fast_leaf_tests(const SceneView &sc, RayState &r, uint32_t leaf_link) { fast_leaf_tests(sc, r, leaf_link, HitInRay()); }
template <class Stack> GDPT_HD void fast_step_leaf(const SceneView &sc, RayState &r, Stack &st)
{
    const uint32_t leaf = r.cur;
    r.cur = fast_pop(r, st);
    fast_leaf_tests(sc, r, leaf);
}

// b_ray of main.glsl:316-321 for instance record `c0..c3` (inverse transform columns)
GDPT_HD void fast_local_ray(const q4f c0, const q4f c1, const q4f c2, const q4f c3, f3 wo, f3 wd, f3 *o, f3 *d)
{
    *o = mk3(((c0.x * wo.x + c1.x * wo.y) + c2.x * wo.z) + c3.x * 1.0f,
             ((c0.y * wo.x + c1.y * wo.y) + c2.y * wo.z) + c3.y * 1.0f,
             ((c0.z * wo.x + c1.z * wo.y) + c2.z * wo.z) + c3.z * 1.0f);
    *d = mk3(((c0.x * wd.x + c1.x * wd.y) + c2.x * wd.z) + c3.x * 0.0f,
             ((c0.y * wd.x + c1.y * wd.y) + c2.y * wd.z) + c3.y * 0.0f,
             ((c0.z * wd.x + c1.z * wd.y) + c2.z * wd.z) + c3.z * 0.0f);
}

template <bool WIDE = false, class Stack> GDPT_HD void fast_enter_instance(const SceneView &sc, RayState &r, Stack &st);
// Space changes: back to world space after an instance and/or into the instance `cur` names.
template <class Stack> GDPT_HD void fast_step_instance(const SceneView &sc, RayState &r, Stack &st)
{
    if (r.inst != GDPT_NO_INSTANCE) {
        r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd);
        r.inst = GDPT_NO_INSTANCE;
    }
    if ((r.cur & LINK_LEAF) == 0u) return;
    fast_enter_instance(sc, r, st);
}
// r is in world space and r.cur names an instance: into its space (main.glsl:316-321), or past it if its true box is missed
template <bool WIDE, class Stack> GDPT_HD void fast_enter_instance(const SceneView &sc, RayState &r, Stack &st)
{
    const uint32_t idx = r.cur & LINK_INDEX_MASK;
    const q4f c0 = ldq(sc.inst_recs, idx * 7u + 0u);
    const q4f c1 = ldq(sc.inst_recs, idx * 7u + 1u);
    const q4f c2 = ldq(sc.inst_recs, idx * 7u + 2u);
    const q4f c3 = ldq(sc.inst_recs, idx * 7u + 3u);
    const q4u tail = ldqu(sc.inst_recs, idx * 7u + 4u);
    const q4f tmin4 = ldq(sc.inst_recs, idx * 7u + 5u);
    const q4f tmax4 = ldq(sc.inst_recs, idx * 7u + 6u);
    fast_local_ray(c0, c1, c2, c3, r.wo, r.wd, &r.o, &r.d);
    r.rd = fast_rcp3(r.d);
    r.inst = idx;
    // too far out for this BLAS's margins (derived_layout.h fast_reach), or axis-parallel in this instance's space:
    // the exact traversal answers the ray, the rest of the search is dropped
    const uint32_t unsearched = (fast_far_origin(r.o, tmax4.w) ? RAY_FAR : 0u) | (fast_degenerate_dir(r.d) ? RAY_AXIAL : 0u);
    if (unsearched) { r.overflow |= unsearched; r.cur = LINK_NONE; r.sp = 0u; return; }
    float entry;
    const bool touches = fast_slab(r, tmin4.x, tmin4.y, tmin4.z, tmax4.x, tmax4.y, tmax4.z, &entry);
    const uint32_t root = WIDE ? fast_bits(tmin4.w) : tail.w; // the BLAS root in the table being searched
    r.cur = (touches && root != LINK_NONE) ? root : fast_pop(r, st);
}

GDPT_HD bool fast_link_is_leaf(uint32_t l) { return (l & (LINK_TLAS | LINK_LEAF)) == LINK_LEAF; }
GDPT_HD bool fast_link_is_node(uint32_t l, uint32_t inst)
{
    return (l & LINK_LEAF) == 0u && l != LINK_NONE && ((l & LINK_TLAS) == 0u || inst == GDPT_NO_INSTANCE);
}

// Whole search of one ray (no warp-level scheduling): used by k_primary-style callers and the host check.
template <class Stack> GDPT_HD void fast_trace_ray(const SceneView &sc, RayState &r, Stack &st)
{
    r.cur = fast_start_link(r, r.cur);
    while (r.cur != LINK_NONE) {
        if (fast_link_is_leaf(r.cur)) fast_step_leaf(sc, r, st);
        else if (fast_link_is_node(r.cur, r.inst)) fast_step_node(sc, r, st);
        else fast_step_instance(sc, r, st);
    }
}

// The same search over the four-wide tables (sc.fast4_ok).
template <class Stack> GDPT_HD void fast_trace_ray4(const SceneView &sc, RayState &r, Stack &st)
{
    r.cur = fast_start_link(r, sc.fast4_root);
    while (r.cur != LINK_NONE) {
        if (fast_link_is_leaf(r.cur)) fast_step_leaf(sc, r, st);
        else if (fast_link_is_node(r.cur, r.inst)) fast_step_node4(sc, r, st);
        else {
            if (r.inst != GDPT_NO_INSTANCE) { r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd); r.inst = GDPT_NO_INSTANCE; }
            if (r.cur & LINK_LEAF) fast_enter_instance<true>(sc, r, st);
        }
    }
}

// Is there a pair other than (skip_inst, skip_tri) that the ray hits at t <= bound?  The same complete search, with the
// running minimum started at `bound` and the one pair left out.  Cold path: a few rays per frame get here.
struct FastLocalStack {
    uint32_t slots[GDPT_FAST_MAX_DEPTH + 4u];
    GDPT_HD void store(uint32_t i, uint32_t v) { slots[i] = v; }
    GDPT_HD uint32_t load(uint32_t i) const { return slots[i]; }
};
GDPT_HD bool fast_other_pair_within(const SceneView &sc, f3 wo, f3 wd, float bound, uint32_t skip_inst, uint32_t skip_tri)
{
    FastLocalStack st;
    RayState r;
    fast_ray_begin(r, sc, wo, wd);
    r.t = bound;
    const bool wide = sc.fast4_ok != 0u;
    r.cur = fast_start_link(r, wide ? sc.fast4_root : r.cur);
    while (r.cur != LINK_NONE) {
        if (fast_link_is_leaf(r.cur)) {
            const uint32_t leaf = r.cur;
            r.cur = fast_pop(r, st);
            const uint32_t first = leaf & FAST_LEAF_FIRST_MASK, count = ((leaf >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u;
            for (uint32_t i = 0; i < count; i++) {
                const q4f a = ldq(sc.fast_tris, (first + i) * 3u + 0u), b = ldq(sc.fast_tris, (first + i) * 3u + 1u), c = ldq(sc.fast_tris, (first + i) * 3u + 2u);
                if (fast_bits(a.w) == skip_tri && r.inst == skip_inst) continue;
                fast_triangle_test(r, a, b, c);
            }
        } else if (fast_link_is_node(r.cur, r.inst)) {
            if (wide) fast_step_node4(sc, r, st); else fast_step_node(sc, r, st);
        } else {
            if (r.inst != GDPT_NO_INSTANCE) { r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd); r.inst = GDPT_NO_INSTANCE; }
            if (r.cur & LINK_LEAF) { if (wide) fast_enter_instance<true>(sc, r, st); else fast_enter_instance<false>(sc, r, st); }
        }
    }
    return r.t < bound || (r.overflow & (RAY_TIE | RAY_UNSEARCHED)) != 0u; // a closer pair, one exactly at the bound, or not searched
}

// The search ended with a hit at the unique minimum t_w = r.t (no tie).  Does the reference traversal test this pair?
// It does iff every box on the way to it passes the reference's push test  d < hit.t  at the time its parent is
// visited (main.glsl:290-299, 336-345).  Until the pair is tested, hit.t is the smallest t of the OTHER pairs tested so far,
// hence >= t2 := the smallest t over all other pairs.  So it is enough that, with the reference's own box arithmetic
// (slab_test == intersectAABB, main.glsl:259-268), the TLAS leaf of the instance (world ray) and the BLAS leaf of the
// triangle (instance-local ray) are entered at  d < t_w  (then d < t_w < hit.t whatever was tested before) -- or,
// when rounding puts an entry distance at or a few ulp beyond t_w (a hit on the face of its own leaf box), at  d < t2,
// which one more search with the pair left out decides (fast_other_pair_within).  Reference boxes are nested and the
// slab arithmetic is monotone in the box, so the ancestors of those two boxes are entered no later.  Rays with a zero
// direction component are sent to the exact traversal: 0 * inf in a slab would void the monotonicity argument.
GDPT_HD bool fast_proves_reference_hit(const SceneView &sc, const RayState &r)
{
    const uint32_t inst = hit_blas(r.blas_front); // no tie: also the space of the accepted test
    const q4f c0 = ldq(sc.inst_recs, inst * 7u + 0u);
    const q4f c1 = ldq(sc.inst_recs, inst * 7u + 1u);
    const q4f c2 = ldq(sc.inst_recs, inst * 7u + 2u);
    const q4f c3 = ldq(sc.inst_recs, inst * 7u + 3u);
    const q4u tail = ldqu(sc.inst_recs, inst * 7u + 4u);
    RayState w;
    w.o = r.wo; w.rd = rcp3(r.wd);
    if (!(r.wd.x != 0.0f && r.wd.y != 0.0f && r.wd.z != 0.0f)) return false;
    float entered = 0.0f; // the later of the two entry distances
    if ((sc.tlas_root_link & LINK_LEAF) == 0u) {
        const q4f n0 = ldq(sc.tlas, tail.y * 2u + 0u);
        const q4f n1 = ldq(sc.tlas, tail.y * 2u + 1u);
        entered = slab_test(w, n0.x, n0.y, n0.z, n1.x, n1.y, n1.z);
    }
    f3 ld;
    fast_local_ray(c0, c1, c2, c3, r.wo, r.wd, &w.o, &ld);
    w.rd = rcp3(ld);
    if (!(ld.x != 0.0f && ld.y != 0.0f && ld.z != 0.0f)) return false;
#if defined(__CUDA_ARCH__)
    const uint32_t leaf = __ldg(sc.tri_leaf + r.tri);
#else
    const uint32_t leaf = sc.tri_leaf[r.tri];
#endif
    if (leaf != tail.z) { // a root that is itself the leaf is never box-tested (main.glsl:274)
        const q4f b0 = ldq(sc.bvh, leaf * 3u + 0u);
        const q4f b1 = ldq(sc.bvh, leaf * 3u + 1u);
        const float d = slab_test(w, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z);
        entered = d > entered ? d : entered;
    }
    if (entered < r.t) return true;
    if (!(entered < 1e29f)) return false; // the reference's arithmetic misses one of the boxes outright
    return !fast_other_pair_within(sc, r.wo, r.wd, entered, inst, r.tri);
}

// Verdict on a finished search: true = the record in `r` is the reference's record.
GDPT_HD bool fast_result_is_reference(const SceneView &sc, const RayState &r)
{
    if (r.overflow & (RAY_TIE | RAY_UNSEARCHED)) return false;
    if (!(r.t < 1e9f)) return true;
    return fast_proves_reference_hit(sc, r);
}

} // namespace gdpt
#endif
