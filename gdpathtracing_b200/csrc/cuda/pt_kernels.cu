// pt_kernels.cu -- the sm_100a kernels behind "main.glsl" (K1) and
// "progressive_rendering.glsl" (K2).  Compile with -fmad=false (see pt_math.cuh).
//
// K1 is a wavefront pipeline over a structure-of-arrays path queue:
//
//   primary      ray generation + traversal of every pixel's camera ray.  Misses
//                are finished in place (sky, colour + depth store); hits are
//                appended to the queue (warp-aggregated atomics).
//   shade(i)     one thread per surviving hit: material fetch, emission, BRDF
//                sample/eval, throughput update; the continuation ray is appended
//                to the other queue, terminated paths store their pixel.
//   trace(i)     persistent warps over the continuation rays with dynamic
//                refill: lanes whose ray finished pull the next ray while the
//                rest keep traversing; hit fields are written in place, misses
//                are finished, hits go to a compact index list for shade(i).
//
// Traversal-stage scheduling: each warp claims chunks of work with one atomic,
// keeps a 16-entry per-lane stack in shared memory (entry-major, so lane i always
// hits bank i: conflict-free) with a local-memory spill above it, walks BLAS
// internal nodes in a tight loop (box tests only) and handles leaves / TLAS
// entries at a common reconvergence point, and re-checks the number of live
// lanes with a ballot every step to decide when to refill.
#include "pt_kernels.cuh"
#include "pt_fast.cuh"
#include "pt_post.cuh"
#include "pt_shade.cuh"
#include "pt_trace.cuh"

#include <cstdio>

#ifndef GDPT_OOL_COLD
#define GDPT_OOL_COLD 1
#endif

namespace gdpt {

namespace {

constexpr int kTraceThreads = 128;
constexpr int kSmemStack = 16;
constexpr int kShadeThreads = 128;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kChunkPrimary = 32; // one 8x4-pixel tile: small chunks keep the expensive tiles spread over many warps
constexpr uint32_t kChunkBounce = 32;
constexpr uint32_t kChunkHeavy = 4;  // work items taken at a time while the long-path classes last
constexpr int kLeafTris = 4;         // triangle tests per leaf step of the path kernel

struct SmemStack {
    uint32_t *col;   // this thread's column of the CTA's shared stack array
    uint32_t *spill; // local-memory continuation
    __device__ __forceinline__ void store(uint32_t i, uint32_t v)
    {
        if (i < (uint32_t)kSmemStack) col[i * kTraceThreads] = v; else spill[i - kSmemStack] = v;
    }
    __device__ __forceinline__ uint32_t load(uint32_t i) const
    {
        return (i < (uint32_t)kSmemStack) ? col[i * kTraceThreads] : spill[i - kSmemStack];
    }
};

__device__ __forceinline__ bool work_to_pixel(const FrameArgs &a, uint32_t w, int *px, int *py)
{
    const uint32_t tile = w >> 5, lane = w & 31u;
    const uint32_t tiles_x = ((uint32_t)a.width + 7u) >> 3;
    const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
    const int lx = (int)(tx * 8u + (lane & 7u)), ly = (int)(ty * 4u + (lane >> 3));
    if (lx >= a.width || ly >= a.local_rows) return false;
    int y = ly;
    if (a.shard_parts > 1) {
        const int lb = ly / a.shard_band;
        y = (lb * a.shard_parts + a.shard_part) * a.shard_band + (ly - lb * a.shard_band);
    }
    if (y >= a.height) return false;
    *px = lx; *py = y;
    return true;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Optional out-of-line form of the cold code (shading, ray generation).  Measured on the demo frame:
// the call ABI's register spills cost more (k_path 1.67 -> 2.04 ms) than the smaller hot-loop
// footprint saves, so it is off; kept for the compact-loop rework.
constexpr bool kOutOfLineCold = GDPT_OOL_COLD != 0;
struct PrimaryRay { f3 o, d; u2 seed; };
__device__ __noinline__ void shade_and_bounce_ool(const SceneView *sc, f3 wo, f3 wd, float t, float u, float v, uint32_t tri,
                                                  uint32_t blas_front, f3 radiance, f3 throughput, u2 *seed, BounceResult *out)
{
    *out = shade_and_bounce(*sc, wo, wd, t, u, v, tri, blas_front, radiance, throughput, *seed);
}
__device__ __noinline__ void generate_primary_ray_ool(const gdpt_camera *cam, int width, int height, int px, int py, PrimaryRay *out)
{
    out->seed = generate_primary_ray(*cam, width, height, px, py, &out->o, &out->d);
}

__device__ __forceinline__ uint32_t class_base(const FrameArgs &a, int c)
{
    return c == 0 ? 0u : a.queue_cap + (uint32_t)(c - 1) * a.heavy_cap;
}
__device__ __forceinline__ int cost_class(uint32_t cost)
{
    if (cost < 16u) return 0;
    const int c = 28 - __clz(cost); // 16..31 -> 1, 32..63 -> 2, ... 1024.. -> 7
    return c > kCostClasses - 1 ? kCostClasses - 1 : c;
}
// Work item w of the survivor lists, heaviest class first.  end[c] = items in classes >= c... see k_path.
struct SurvivorLists {
    uint32_t n[kCostClasses];
    uint32_t total, heavy_total;
    __device__ __forceinline__ void load(const FrameArgs &a)
    {
        total = 0u;
#pragma unroll
        for (int c = 0; c < kCostClasses; c++) {
            n[c] = min(a.counters->qcount[c], c == 0 ? a.queue_cap : a.heavy_cap);
            total += n[c];
        }
        heavy_total = 0u; // classes fetched a few items at a time: paths of 128 steps and more
#pragma unroll
        for (int c = kFirstHeavyClass; c < kCostClasses; c++) heavy_total += n[c];
    }
    __device__ __forceinline__ uint32_t pixel(const FrameArgs &a, uint32_t w) const
    {
#pragma unroll
        for (int c = kCostClasses - 1; c > 0; c--) {
            if (w < n[c]) return a.hit_list[class_base(a, c) + w];
            w -= n[c];
        }
        return a.hit_list[w];
    }
};

__device__ __forceinline__ float4 *plane(const FrameArgs &a, int q, int p) { return a.queue[q] + (size_t)p * a.queue_cap; }

__device__ __forceinline__ void write_trace_record(const FrameArgs &a, int segment, uint32_t pixel, const RayState &r,
                                                   const TraceCounters &tc)
{
    if (!a.trace || segment >= a.trace_segments) return;
    gdpt_trace_record rec;
    const bool hit = r.t < 1e9f;
    rec.hit = hit ? 1u : 0u;
    rec.triangle = hit ? r.tri : 0u;
    rec.blas = hit ? (r.blas_front & ~GDPT_FRONT_BIT) : 0u;
    rec.front = hit ? (r.blas_front >> 31) : 0u;
    rec.t = r.t; rec.u = hit ? r.u : 0.0f; rec.v = hit ? r.v : 0.0f;
    rec.node_pops = tc.node_pops; rec.box_tests = tc.box_tests; rec.tri_tests = tc.tri_tests;
    rec.tlas_leaves = tc.tlas_leaves; rec.max_stack = tc.max_stack;
    rec.visit_hash_lo = (uint32_t)tc.hash; rec.visit_hash_hi = (uint32_t)(tc.hash >> 32);
    a.trace[(size_t)segment * a.width * a.height + pixel] = rec;
}

// MODE 0: primary rays generated from pixel work items.  MODE 1: rays read from queue `src`.
template <bool TRACE, int MODE, bool CULL>
__global__ void __launch_bounds__(kTraceThreads) k_trace(const FrameArgs a, const int segment, const int src)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    uint32_t spill[GDPT_MAX_STACK - kSmemStack];
    SmemStack st;
    st.col = s_stack + threadIdx.x;
    st.spill = spill;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    uint32_t *cursor = &cnt->cursor[MODE == 0 ? 0 : 2 * segment];
    const uint32_t total = (MODE == 0) ? a.n_work : min(cnt->qcount[segment], a.queue_cap);
    const uint32_t chunk = (MODE == 0) ? kChunkPrimary : kChunkBounce;
    __shared__ gdpt_camera s_cam;
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    __syncthreads();
    const gdpt_camera &cam = s_cam;
    const int refill_below = max(a.refill_below, 1);

    RayState r;
    r.cur = LINK_NONE; r.sp = 0; r.overflow = 0; r.t = 1e9f;
    bool has = false;
    uint32_t item = 0;  // MODE 0: pixel index; MODE 1: queue slot
    u2 seed; seed.x = seed.y = 0u;
    TraceCounters tc;
    if (TRACE) counters_init(tc, nullptr, 0);

    uint32_t tri_next = 0, tri_end = 0;     // pending triangle range of the leaf being tested (schedule 1)
    uint32_t chunk_next = 0, chunk_end = 0; // warp-uniform
    bool exhausted = (total == 0u);
    unsigned long long my_pops = 0, my_boxes = 0, my_tris = 0, my_leaves = 0, my_phits = 0;
    uint32_t my_max_stack = 0, my_overflow = 0;

    for (;;) {
        // ---------------- refill idle lanes ----------------
        unsigned idle = __ballot_sync(kFull, !has);
        const int live = 32 - __popc(idle);
        if (!exhausted && live < refill_below) {
            for (int round = 0; round < 2 && idle != 0u && !exhausted; round++) {
                if (chunk_next == chunk_end) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(cursor, chunk);
                    base = __shfl_sync(kFull, base, 0);
                    if (base >= total) { exhausted = true; break; }
                    chunk_next = base;
                    chunk_end = min(base + chunk, total);
                }
                const uint32_t avail = chunk_end - chunk_next;
                const uint32_t rank = __popc(idle & lanemask_lt);
                const bool take = !has && rank < avail;
                if (take) {
                    const uint32_t w = chunk_next + rank;
                    f3 o, d;
                    bool valid = true;
                    if (MODE == 0) {
                        int px, py;
                        valid = work_to_pixel(a, w, &px, &py);
                        if (valid) {
                            seed = generate_primary_ray(cam, a.width, a.height, px, py, &o, &d);
                            item = (uint32_t)py * (uint32_t)a.width + (uint32_t)px;
                        }
                    } else {
                        const float4 p0 = plane(a, src, 0)[w], p1 = plane(a, src, 1)[w];
                        o = mk3(p0.x, p0.y, p0.z); d = mk3(p1.x, p1.y, p1.z);
                        item = w;
                    }
                    if (valid) {
                        ray_begin(r, a.sc, o, d);
                        has = true;
                        if (TRACE) {
                            uint32_t *vis = nullptr; // the visit list is kept for primary rays only
                            if (MODE == 0 && a.visits) vis = a.visits + (size_t)item * a.visits_per_ray;
                            counters_init(tc, vis, a.visits_per_ray);
                        }
                    }
                }
                chunk_next += min((uint32_t)__popc(idle), avail);
                idle = __ballot_sync(kFull, !has); // lanes that drew padding (or nothing) try again
            }
        }
        if (__ballot_sync(kFull, has) == 0u) {
            if (exhausted) break;
            continue; // nothing taken this round (all items were padding); try again
        }

        // ---------------- traverse ----------------
        const int keep_going = exhausted ? 1 : refill_below;
        if (a.schedule == 0) {
            // while-while: descend internal nodes in a tight loop, then one leaf / TLAS step
            for (int it = 0; it < a.burst; ++it) {
                while (link_is_blas_internal(r.cur)) step_blas_internal<TRACE, CULL>(a.sc, r, st, &tc);
                if (r.cur != LINK_NONE) {
                    if (link_is_blas_leaf(r.cur)) step_blas_leaf<TRACE>(a.sc, r, st, &tc);
                    else step_tlas<TRACE, CULL>(a.sc, r, st, &tc);
                }
                const unsigned walking = __ballot_sync(kFull, r.cur != LINK_NONE);
                if (__popc(walking) < keep_going) break;
            }
        } else {
            // phase voting: every lane is in one of three phases (L: one triangle test, I: one
            // internal node = two box tests, T: one TLAS-level entry).  Each iteration the warp
            // executes only the phase most lanes are in, so an instruction stream is not issued
            // for a handful of lanes while the rest could have joined it a step later.
            for (int it = 0; it < a.burst; ++it) {
                const bool in_l = tri_next < tri_end || link_is_blas_leaf(r.cur);
                const bool in_i = !in_l && link_is_blas_internal(r.cur);
                const bool in_t = !in_l && !in_i && r.cur != LINK_NONE;
                const int n_l = __popc(__ballot_sync(kFull, in_l)), n_i = __popc(__ballot_sync(kFull, in_i)),
                          n_t = __popc(__ballot_sync(kFull, in_t));
                if (n_l + n_i + n_t < keep_going) break;
                if (n_l >= n_i && n_l >= n_t) {
                    if (in_l) step_blas_leaf_one<TRACE>(a.sc, r, st, &tc, tri_next, tri_end);
                } else if (n_i >= n_t) {
                    if (in_i) step_blas_internal<TRACE, CULL>(a.sc, r, st, &tc);
                } else {
                    if (in_t) step_tlas<TRACE, CULL>(a.sc, r, st, &tc);
                }
            }
        }

        // ---------------- retire finished rays ----------------
        const bool fin = has && r.cur == LINK_NONE && tri_next == tri_end;
        const bool is_hit = fin && r.t < 1e9f;
        if (MODE == 0) {
            // hits join the path queue
            const unsigned hm = __ballot_sync(kFull, is_hit && !a.debug_steps);
            uint32_t base = 0;
            if (hm != 0u) {
                const int leader = __ffs(hm) - 1;
                if ((int)lane == leader) base = atomicAdd(&cnt->qcount[0], (uint32_t)__popc(hm));
                base = __shfl_sync(kFull, base, leader);
            }
            if (fin) {
                if (TRACE) write_trace_record(a, 0, item, r, tc);
                if (a.debug_steps) { // main.glsl:358-361,423-427
                    float e = TRACE ? (float)tc.tri_tests / 256.0f : 0.0f;
                    e = e < 0.0f ? 0.0f : (e > 1.0f ? 1.0f : e);
                    a.out_rgba8[item] = pack_rgba8(mk3(e, e, e));
                    a.out_depth[item] = encode_depth(cam, cam.z_far);
                } else if (is_hit) {
                    const uint32_t slot = base + __popc(hm & lanemask_lt);
                    if (slot < a.queue_cap) {
                        plane(a, 0, 0)[slot] = make_float4(r.wo.x, r.wo.y, r.wo.z, __uint_as_float(item));
                        plane(a, 0, 1)[slot] = make_float4(r.wd.x, r.wd.y, r.wd.z, r.t);
                        plane(a, 0, 2)[slot] = make_float4(1.0f, 1.0f, 1.0f, r.u);
                        plane(a, 0, 3)[slot] = make_float4(0.0f, 0.0f, 0.0f, r.v);
                        plane(a, 0, 4)[slot] = make_float4(__uint_as_float(seed.x), __uint_as_float(seed.y),
                                                           __uint_as_float(r.tri), __uint_as_float(r.blas_front));
                    }
                    my_phits++;
                } else {
                    const f3 radiance = mk3(0.0f, 0.0f, 0.0f) + mk3(1.0f, 1.0f, 1.0f) * sample_sky(r.wd);
                    a.out_rgba8[item] = pack_rgba8(radiance);
                    a.out_depth[item] = encode_depth(cam, cam.z_far);
                }
            }
        } else {
            const unsigned hm = __ballot_sync(kFull, is_hit);
            uint32_t base = 0;
            if (hm != 0u) {
                const int leader = __ffs(hm) - 1;
                if ((int)lane == leader) base = atomicAdd(&cnt->lcount[segment], (uint32_t)__popc(hm));
                base = __shfl_sync(kFull, base, leader);
            }
            if (fin) {
                const float4 p0 = plane(a, src, 0)[item];
                const uint32_t pixel = __float_as_uint(p0.w);
                if (TRACE) write_trace_record(a, segment, pixel, r, tc);
                if (is_hit) {
                    plane(a, src, 1)[item].w = r.t;
                    plane(a, src, 2)[item].w = r.u;
                    plane(a, src, 3)[item].w = r.v;
                    float4 *p4 = plane(a, src, 4) + item;
                    p4->z = __uint_as_float(r.tri);
                    p4->w = __uint_as_float(r.blas_front);
                    a.hit_list[base + __popc(hm & lanemask_lt)] = item;
                } else {
                    const float4 p2 = plane(a, src, 2)[item], p3 = plane(a, src, 3)[item];
                    const f3 radiance = mk3(p3.x, p3.y, p3.z) + mk3(p2.x, p2.y, p2.z) * sample_sky(r.wd);
                    a.out_rgba8[pixel] = pack_rgba8(radiance);
                }
            }
        }
        if (fin) {
            if (TRACE) {
                my_pops += tc.node_pops; my_boxes += tc.box_tests; my_tris += tc.tri_tests; my_leaves += tc.tlas_leaves;
                if (tc.max_stack > my_max_stack) my_max_stack = tc.max_stack;
            }
            my_overflow |= r.overflow;
            has = false;
        }
    }

    // ---------------- per-warp statistics ----------------
    if (MODE == 0) {
        for (int off = 16; off > 0; off >>= 1) my_phits += __shfl_down_sync(kFull, my_phits, off);
        if (lane == 0 && my_phits) atomicAdd(&cnt->primary_hits, my_phits);
    }
    if (TRACE) {
        for (int off = 16; off > 0; off >>= 1) {
            my_pops += __shfl_down_sync(kFull, my_pops, off);
            my_boxes += __shfl_down_sync(kFull, my_boxes, off);
            my_tris += __shfl_down_sync(kFull, my_tris, off);
            my_leaves += __shfl_down_sync(kFull, my_leaves, off);
            my_max_stack = max(my_max_stack, __shfl_down_sync(kFull, my_max_stack, off));
        }
        if (lane == 0) {
            atomicAdd(&cnt->node_pops, my_pops); atomicAdd(&cnt->box_tests, my_boxes);
            atomicAdd(&cnt->tri_tests, my_tris); atomicAdd(&cnt->tlas_leaves, my_leaves);
            atomicMax(&cnt->max_stack, my_max_stack);
        }
    }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

// Single-kernel schedule: every lane carries one whole path (main.glsl:372-401) from its camera ray
// to termination, so there is no barrier between bounces and no queue traffic.  A warp is a small
// scheduler over five phases -- L (one triangle test), I (one internal node), T (one TLAS-level
// entry), S (shade the finished segment and start the next, or finish the path) and R (refill idle
// lanes with new pixels).  Each iteration it executes the one phase that pays most: S once
// `shade_at` lanes hold a finished ray (or nothing is walking), R once enough lanes are idle,
// otherwise the traversal phase most lanes are in.
//
// SRC 0: work items are 8x4-pixel tiles of the whole (sharded) image.  SRC 1: work items are the
// pixels k_primary_cull left in `hit_list` (camera rays that touch some instance's tight box).
// COMPACT: the hot loop is written for the instruction cache (32 KB L1.5, ~6 KB L0): one copy of the
// box code serves BLAS and TLAS internal nodes, leaves run one triangle-test body in a short loop,
// instance entry/exit is its own small phase, the phase census is one REDUX, and shading / ray
// generation are called out of line with by-value arguments.
template <bool TRACE, bool CULL, int SRC, int MINB = 4, bool COMPACT = false>
__global__ void __launch_bounds__(kTraceThreads, MINB) k_path(const FrameArgs a)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    __shared__ gdpt_camera s_cam;
    uint32_t spill[GDPT_MAX_STACK - kSmemStack];
    SmemStack st;
    st.col = s_stack + threadIdx.x;
    st.spill = spill;
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    __syncthreads();
    const gdpt_camera &cam = s_cam;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    SurvivorLists lists;
    if (SRC == 1) lists.load(a);
    const uint32_t total = (SRC == 0) ? a.n_work : lists.total;
    const int refill_below = max(a.refill_below, 1);
    const int shade_at = min(max(a.shade_at, 1), 32);
    const uint32_t lead_min = a.lead_min > 0 ? (uint32_t)a.lead_min : 0xFFFFFFFFu;
    uint32_t steps = 0; // scheduler iterations this lane's current path took part in
    uint32_t pred = 0;  // what its pixel's path cost in the previous frame (0 = unknown / light)
    bool heavy_done = false; // warp-uniform: the heavy classes are handed out
    const int last_segment = a.debug_steps ? 0 : a.max_depth - 1;

    RayState r;
    r.cur = LINK_NONE; r.sp = 0; r.overflow = 0; r.t = 1e9f;
    f3 throughput = mk3(1.0f, 1.0f, 1.0f), radiance = mk3(0.0f, 0.0f, 0.0f);
    u2 seed; seed.x = seed.y = 0u;
    uint32_t pixel = 0;
    int segment = 0;
    bool has = false;
    uint32_t tri_next = 0, tri_end = 0;
    TraceCounters tc;
    if (TRACE) counters_init(tc, nullptr, 0);
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = (total == 0u);
    unsigned long long my_rays = 0, my_phits = 0, my_pops = 0, my_boxes = 0, my_tris = 0, my_leaves = 0;
    uint32_t my_max_stack = 0, my_overflow = 0;
    const bool prof = a.warp_prof != nullptr;
    const unsigned long long t_start = prof ? global_ns() : 0ull;
    uint32_t it_i = 0, it_l = 0, it_t = 0, it_f = 0, it_e = 0, n_started = 0;

    for (;;) {
        const bool in_l = has && (tri_next < tri_end || link_is_blas_leaf(r.cur));
        const bool in_i = has && !in_l && (COMPACT ? link_is_node_step(r.cur, r.inst) : link_is_blas_internal(r.cur));
        const bool in_t = has && !in_l && !in_i && r.cur != LINK_NONE;
        const bool fin = has && !in_l && !in_i && !in_t;
        int n_l, n_i, n_t, n_fin, n_idle;
        unsigned idle;
        if (COMPACT) {
            // one REDUX: every lane adds 1 into the 6-bit field of its phase
            const uint32_t census = __reduce_add_sync(kFull, 1u << (in_l ? 0 : (in_i ? 6 : (in_t ? 12 : (fin ? 18 : 24)))));
            n_l = (int)(census & 63u); n_i = (int)((census >> 6) & 63u); n_t = (int)((census >> 12) & 63u);
            n_fin = (int)((census >> 18) & 63u); n_idle = (int)(census >> 24);
            idle = 0u; // taken by ballot only where it is needed (refill)
        } else {
            n_l = __popc(__ballot_sync(kFull, in_l)); n_i = __popc(__ballot_sync(kFull, in_i));
            n_t = __popc(__ballot_sync(kFull, in_t)); n_fin = __popc(__ballot_sync(kFull, fin));
            idle = __ballot_sync(kFull, !has);
            n_idle = __popc(idle);
        }
        const int n_walk = n_l + n_i + n_t;
        // critical-path-first: once some path is long, the longest one picks the phase, so the path
        // that decides when the kernel ends moves every iteration
        int lead_phase = -1; // 0 L, 1 I, 2 T, 3 finished
        if (COMPACT) {
            // the lane whose path is expected to run longest (previous frame's cost) picks the phase, so the
            // path that decides when this warp ends advances every iteration instead of every other one
            const uint32_t key = (has && pred >= lead_min) ? ((pred << 5) | lane) : 0u; // pred < 2^27: unique per lane
            const uint32_t most = __reduce_max_sync(kFull, key);
            if (most != 0u) lead_phase = __shfl_sync(kFull, in_l ? 0 : (in_i ? 1 : (in_t ? 2 : 3)), most & 31u);
        } else {
            const uint32_t key = has ? steps : 0u;
            const uint32_t most = __reduce_max_sync(kFull, key);
            if (most >= lead_min) {
                const unsigned who = __ballot_sync(kFull, has && steps == most);
                lead_phase = __shfl_sync(kFull, in_l ? 0 : (in_i ? 1 : (in_t ? 2 : 3)), __ffs(who) - 1);
            }
        }

        if (n_fin > 0 && (n_fin >= shade_at || n_walk == 0 || lead_phase == 3)) {
            // ---------------- S: finish a segment ----------------
            it_f++;
            if (fin) {
                const bool hit = r.t < 1e9f;
                my_rays++;
                if (segment == 0 && hit) my_phits++;
                if (TRACE) {
                    write_trace_record(a, segment, pixel, r, tc);
                    my_pops += tc.node_pops; my_boxes += tc.box_tests; my_tris += tc.tri_tests; my_leaves += tc.tlas_leaves;
                    if (tc.max_stack > my_max_stack) my_max_stack = tc.max_stack;
                }
                my_overflow |= r.overflow;
                bool alive = false;
                if (a.debug_steps) { // main.glsl:358-361,423-427
                    float e = TRACE ? (float)tc.tri_tests / 256.0f : 0.0f;
                    e = e < 0.0f ? 0.0f : (e > 1.0f ? 1.0f : e);
                    radiance = mk3(e, e, e);
                    a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                } else if (!hit) {
                    radiance = radiance + throughput * sample_sky(r.wd);
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                } else {
                    BounceResult br;
                    if (COMPACT && kOutOfLineCold) {
                        u2 sd = seed; // by value: nothing the hot loop keeps in registers has its address taken
                        shade_and_bounce_ool(&a.sc, r.wo, r.wd, r.t, r.u, r.v, r.tri, r.blas_front, radiance, throughput, &sd, &br);
                        seed = sd;
                    } else {
                        br = shade_and_bounce(a.sc, r.wo, r.wd, r.t, r.u, r.v, r.tri, r.blas_front, radiance, throughput, seed);
                    }
                    radiance = br.radiance;
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
                    alive = br.alive && segment < last_segment;
                    if (alive) {
                        throughput = br.throughput;
                        ray_begin(r, a.sc, br.next_o, br.next_d);
                        segment++;
                        if (TRACE) counters_init(tc, nullptr, 0);
                    }
                }
                if (!alive) {
                    a.out_rgba8[pixel] = pack_rgba8(radiance);
                    if (SRC == 1) a.cost[pixel] = steps;
                    has = false;
                }
            }
            continue;
        }
        if (!exhausted && n_idle > 0 && (32 - n_idle < refill_below || n_walk + n_fin == 0)) {
            // ---------------- R: new camera rays for idle lanes ----------------
            it_e++;
            if (chunk_next == chunk_end) {
                uint32_t base = 0, len = kChunkPrimary;
                if (lane == 0) {
                    if (SRC == 1 && !heavy_done) { // the long paths (front of the survivor order) are dealt a few per warp
                        base = atomicAdd(&cnt->cursor[0], kChunkHeavy);
                        len = kChunkHeavy;
                        if (base >= lists.heavy_total) base = 0xFFFFFFFFu;
                        else if (base + len > lists.heavy_total) len = lists.heavy_total - base;
                    }
                    if (SRC == 0 || heavy_done || base == 0xFFFFFFFFu) {
                        const uint32_t first = (SRC == 1) ? lists.heavy_total : 0u;
                        base = first + atomicAdd(&cnt->cursor[1], kChunkPrimary);
                        len = kChunkPrimary | 0x80000000u; // flag: came from the light cursor
                    }
                }
                base = __shfl_sync(kFull, base, 0);
                len = __shfl_sync(kFull, len, 0);
                if (len & 0x80000000u) { heavy_done = true; len &= 0x7FFFFFFFu; }
                if (base >= total) { exhausted = true; continue; }
                chunk_next = base;
                chunk_end = min(base + len, total);
            }
            if (COMPACT) idle = __ballot_sync(kFull, !has);
            const uint32_t avail = chunk_end - chunk_next;
            const uint32_t rank = __popc(idle & lanemask_lt);
            if (!has && rank < avail) {
                int px = 0, py = 0;
                bool valid;
                if (SRC == 0) valid = work_to_pixel(a, chunk_next + rank, &px, &py);
                else {
                    const uint32_t p = lists.pixel(a, chunk_next + rank);
                    py = (int)(p / (uint32_t)a.width); px = (int)(p - (uint32_t)py * (uint32_t)a.width);
                    pred = min(a.cost[p], (1u << 27) - 1u);
                    valid = true;
                }
                if (valid) {
                    f3 o, d;
                    if (COMPACT && kOutOfLineCold) {
                        PrimaryRay pr;
                        generate_primary_ray_ool(&cam, a.width, a.height, px, py, &pr);
                        o = pr.o; d = pr.d; seed = pr.seed;
                    } else {
                        seed = generate_primary_ray(cam, a.width, a.height, px, py, &o, &d);
                    }
                    pixel = (uint32_t)py * (uint32_t)a.width + (uint32_t)px;
                    throughput = mk3(1.0f, 1.0f, 1.0f); radiance = mk3(0.0f, 0.0f, 0.0f);
                    segment = 0;
                    steps = 0;
                    ray_begin(r, a.sc, o, d);
                    has = true;
                    if (TRACE) counters_init(tc, a.visits ? a.visits + (size_t)pixel * a.visits_per_ray : nullptr, a.visits_per_ray);
                }
            }
            n_started += min((uint32_t)n_idle, avail);
            chunk_next += min((uint32_t)n_idle, avail);
            continue;
        }
        if (n_walk == 0) {
            if (exhausted && n_fin == 0) break;
            continue;
        }
        // ---------------- L / I / T: one traversal step of the leading path's phase, else the most popular ----------------
        int run = (n_l >= n_i && n_l >= n_t) ? 0 : (n_i >= n_t ? 1 : 2);
        if (lead_phase >= 0 && lead_phase < 3) run = lead_phase;
        if (COMPACT && run == 1 && a.burst > 1) {
            // node burst: keep descending while at least half of the lanes that started stay on internal nodes
            // (no census in between): the long paths spend most of their steps here
            it_i++;
            bool go = in_i;
            int need = (n_i + 1) >> 1;
#pragma unroll 1
            for (int b = 0; b < a.burst; b++) {
                if (go) { step_node<TRACE, CULL>(a.sc, r, st, &tc); steps++; go = link_is_node_step(r.cur, r.inst); }
                if (__popc(__ballot_sync(kFull, go)) < need) break;
            }
            continue;
        }
        if (run == 0) {
            it_l++;
            if (in_l) {
                if (COMPACT) {
                    step_blas_leaf_one<TRACE>(a.sc, r, st, &tc, tri_next, tri_end); // enters the leaf if needed + first test
#pragma unroll 1
                    for (int i = 1; i < kLeafTris && tri_next < tri_end; i++) triangle_test(a.sc, r, tri_next++);
                } else {
                    step_blas_leaf_some<TRACE, kLeafTris>(a.sc, r, st, &tc, tri_next, tri_end);
                }
                steps++;
            }
        } else if (run == 1) {
            it_i++;
            if (in_i) {
                if (COMPACT) step_node<TRACE, CULL>(a.sc, r, st, &tc);
                else step_blas_internal<TRACE, CULL>(a.sc, r, st, &tc);
                steps++;
            }
        } else {
            it_t++;
            if (in_t) {
                if (COMPACT) step_instance<TRACE, CULL>(a.sc, r, st, &tc);
                else step_tlas<TRACE, CULL>(a.sc, r, st, &tc);
                steps++;
            }
        }
    }
    if (prof && lane == 0) {
        unsigned long long *w = a.warp_prof + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        w[0] = t_start; w[1] = global_ns(); w[2] = it_i; w[3] = it_l; w[4] = it_t; w[5] = it_f; w[6] = it_e; w[7] = n_started;
    }

    for (int off = 16; off > 0; off >>= 1) {
        my_rays += __shfl_down_sync(kFull, my_rays, off);
        my_phits += __shfl_down_sync(kFull, my_phits, off);
    }
    if (lane == 0) { atomicAdd(&cnt->rays, my_rays); atomicAdd(&cnt->primary_hits, my_phits); }
    if (TRACE) {
        for (int off = 16; off > 0; off >>= 1) {
            my_pops += __shfl_down_sync(kFull, my_pops, off);
            my_boxes += __shfl_down_sync(kFull, my_boxes, off);
            my_tris += __shfl_down_sync(kFull, my_tris, off);
            my_leaves += __shfl_down_sync(kFull, my_leaves, off);
            my_max_stack = max(my_max_stack, __shfl_down_sync(kFull, my_max_stack, off));
        }
        if (lane == 0) {
            atomicAdd(&cnt->node_pops, my_pops); atomicAdd(&cnt->box_tests, my_boxes);
            atomicAdd(&cnt->tri_tests, my_tris); atomicAdd(&cnt->tlas_leaves, my_leaves);
            atomicMax(&cnt->max_stack, my_max_stack);
        }
    }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

// ------------------------------------------------------------------------------------------------
// Path kernel of schedule 5: k_path's scheduler (whole paths per lane, phase census, longest path
// first) over the closest-hit tables.  A ray is answered by the order-free search of pt_fast.cuh --
// I: one internal node (two true-box tests, four LDG.128), L: one leaf (a few triangles), T: instance
// entry/exit -- and, when the search ends, by the proof that the reference traversal returns the same
// record.  The few rays without a proof (ties, hits that sit on a reference box face) are re-traced in
// reference order by an out-of-line call, so every record equals the reference's bit for bit.
struct LocalStack {
    uint32_t *slots;
    __device__ __forceinline__ void store(uint32_t i, uint32_t v) { slots[i] = v; }
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return slots[i]; }
};
struct ExactHit { float t, u, v; uint32_t tri, blas_front, overflow; };
__device__ __noinline__ void exact_retrace(const SceneView *sc, f3 wo, f3 wd, ExactHit *out)
{
    uint32_t slots[GDPT_MAX_STACK];
    LocalStack st;
    st.slots = slots;
    RayState r;
    ray_begin(r, *sc, wo, wd);
    trace_ray_compact<false, true>(*sc, r, st, nullptr);
    out->t = r.t; out->u = r.u; out->v = r.v; out->tri = r.tri; out->blas_front = r.blas_front; out->overflow = r.overflow;
}

__device__ __forceinline__ void write_hit_record(const FrameArgs &a, int segment, uint32_t pixel, float t, float u, float v, uint32_t tri,
                                                 uint32_t blas_front)
{
    if (segment >= a.trace_segments) return;
    gdpt_trace_record rec;
    const bool hit = t < 1e9f;
    rec.hit = hit ? 1u : 0u;
    rec.triangle = hit ? tri : 0u;
    rec.blas = hit ? (blas_front & ~GDPT_FRONT_BIT) : 0u;
    rec.front = hit ? (blas_front >> 31) : 0u;
    rec.t = t; rec.u = hit ? u : 0.0f; rec.v = hit ? v : 0.0f;
    rec.node_pops = rec.box_tests = rec.tri_tests = rec.tlas_leaves = rec.max_stack = 0u;
    rec.visit_hash_lo = rec.visit_hash_hi = 0u;
    a.trace[(size_t)segment * a.width * a.height + pixel] = rec;
}

// Verdict on a finished search, out of line: executed once per ray, not per step.
__device__ __noinline__ bool fast_verdict_ool(const SceneView *sc, f3 wo, f3 wd, float t, uint32_t tri, uint32_t blas_front, uint32_t flags)
{
    RayState r;
    r.wo = wo; r.wd = wd; r.t = t; r.tri = tri; r.blas_front = blas_front; r.overflow = flags;
    return fast_result_is_reference(*sc, r);
}

// Scheduler phases of k_path_fast.  A lane may be able to join more than one: a leaf it reaches is parked in
// `pend` (speculative traversal: the search is order-free, so the leaf can wait) and the lane keeps descending;
// the warp runs the leaf phase when enough lanes hold one.
//   I  one internal node of either level (back to world space first if the link is a TLAS one)
//   L  one leaf (the parked one, else the current link): <= 4 triangle tests
//   T  enter the instance the current link names
__device__ __forceinline__ bool lane_can_node(uint32_t cur, uint32_t pend)
{
    return cur != LINK_NONE && (cur & LINK_LEAF) == 0u && ((cur & LINK_TLAS) == 0u || pend == LINK_NONE);
}
__device__ __forceinline__ bool lane_can_enter(uint32_t cur, uint32_t pend)
{
    return cur != LINK_NONE && (cur & (LINK_TLAS | LINK_LEAF)) == (LINK_TLAS | LINK_LEAF) && pend == LINK_NONE;
}
__device__ __forceinline__ bool lane_can_leaf(uint32_t cur, uint32_t pend) { return pend != LINK_NONE || fast_link_is_leaf(cur); }

template <bool REC, int MINB>
__global__ void __launch_bounds__(kTraceThreads, MINB) k_path_fast(const FrameArgs a)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    __shared__ gdpt_camera s_cam;
    uint32_t spill[GDPT_MAX_STACK - kSmemStack];
    SmemStack st;
    st.col = s_stack + threadIdx.x;
    st.spill = spill;
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    __syncthreads();
    const gdpt_camera &cam = s_cam;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    SurvivorLists lists;
    lists.load(a);
    const uint32_t total = lists.total;
    const int refill_below = max(a.refill_below, 1);
    const int shade_at = min(max(a.shade_at, 1), 32);
    const uint32_t lead_min = a.lead_min > 0 ? (uint32_t)a.lead_min : 0xFFFFFFFFu;
    uint32_t steps = 0, pred = 0;
    bool heavy_done = false;
    const int last_segment = a.max_depth - 1;

    RayState r;
    r.cur = LINK_NONE; r.sp = 0; r.overflow = 0; r.t = 1e9f; r.inst = GDPT_NO_INSTANCE;
    f3 wrd = mk3(0.0f, 0.0f, 0.0f);   // 1 / world direction (ray.rD, main.glsl:421)
    uint32_t pend = LINK_NONE;        // parked leaf (instance-local: flushed before the space changes)
    f3 throughput = mk3(1.0f, 1.0f, 1.0f), radiance = mk3(0.0f, 0.0f, 0.0f);
    u2 seed; seed.x = seed.y = 0u;
    uint32_t pixel = 0;
    int segment = 0;
    bool has = false;
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = (total == 0u);
    unsigned long long my_rays = 0, my_phits = 0, my_retraced = 0;
    uint32_t my_overflow = 0;
    const bool prof = a.warp_prof != nullptr;
    const unsigned long long t_start = prof ? global_ns() : 0ull;
    uint32_t it_i = 0, it_l = 0, it_t = 0, it_f = 0, it_e = 0, n_started = 0;

    for (;;) {
        const bool can_i = has && lane_can_node(r.cur, pend);
        const bool can_l = has && lane_can_leaf(r.cur, pend);
        const bool can_t = has && lane_can_enter(r.cur, pend);
        const bool fin = has && r.cur == LINK_NONE && pend == LINK_NONE;
        const uint32_t census = __reduce_add_sync(kFull, (can_i ? 1u : 0u) | (can_l ? 1u << 6 : 0u) | (can_t ? 1u << 12 : 0u) |
                                                             (fin ? 1u << 18 : 0u) | (has ? 0u : 1u << 24));
        const int n_i = (int)(census & 63u), n_l = (int)((census >> 6) & 63u), n_t = (int)((census >> 12) & 63u),
                  n_fin = (int)((census >> 18) & 63u), n_idle = (int)(census >> 24);
        const int n_walk = 32 - n_idle - n_fin;
        int lead_phase = -1; // 0 L, 1 I, 2 T, 3 finished: what the path expected to run longest needs next
        {
            const uint32_t key = (has && pred >= lead_min) ? ((pred << 5) | lane) : 0u;
            const uint32_t most = __reduce_max_sync(kFull, key);
            if (most != 0u) lead_phase = __shfl_sync(kFull, fin ? 3 : (can_i ? 1 : (can_l ? 0 : 2)), most & 31u);
        }

        if (n_fin > 0 && (n_fin >= shade_at || n_walk == 0 || lead_phase == 3)) {
            // ---------------- S: prove, then finish the segment ----------------
            it_f++;
            if (fin) {
                if (!fast_verdict_ool(&a.sc, r.wo, r.wd, r.t, r.tri, r.blas_front, r.overflow)) {
                    ExactHit eh; // rare: exact reference-order traversal of this ray
                    exact_retrace(&a.sc, r.wo, r.wd, &eh);
                    r.t = eh.t; r.u = eh.u; r.v = eh.v; r.tri = eh.tri; r.blas_front = eh.blas_front; r.overflow = eh.overflow & RAY_OVERFLOW;
                    my_retraced++;
                }
                const bool hit = r.t < 1e9f;
                my_rays++;
                if (segment == 0 && hit) my_phits++;
                if (REC) write_hit_record(a, segment, pixel, r.t, r.u, r.v, r.tri, r.blas_front);
                my_overflow |= r.overflow & RAY_OVERFLOW;
                bool alive = false;
                if (!hit) {
                    radiance = radiance + throughput * sample_sky(r.wd);
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                } else {
                    BounceResult br;
                    u2 sd = seed;
                    shade_and_bounce_ool(&a.sc, r.wo, r.wd, r.t, r.u, r.v, r.tri, r.blas_front, radiance, throughput, &sd, &br);
                    seed = sd;
                    radiance = br.radiance;
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
                    alive = br.alive && segment < last_segment;
                    if (alive) {
                        throughput = br.throughput;
                        fast_ray_begin(r, a.sc, br.next_o, br.next_d);
                        wrd = r.rd;
                        segment++;
                    }
                }
                if (!alive) {
                    a.out_rgba8[pixel] = pack_rgba8(radiance);
                    a.cost[pixel] = steps;
                    has = false;
                }
            }
            continue;
        }
        if (!exhausted && n_idle > 0 && (32 - n_idle < refill_below || n_walk + n_fin == 0)) {
            // ---------------- R: new camera rays for idle lanes ----------------
            it_e++;
            if (chunk_next == chunk_end) {
                uint32_t base = 0, len = kChunkPrimary;
                if (lane == 0) {
                    if (!heavy_done) {
                        base = atomicAdd(&cnt->cursor[0], kChunkHeavy);
                        len = kChunkHeavy;
                        if (base >= lists.heavy_total) base = 0xFFFFFFFFu;
                        else if (base + len > lists.heavy_total) len = lists.heavy_total - base;
                    }
                    if (heavy_done || base == 0xFFFFFFFFu) {
                        base = lists.heavy_total + atomicAdd(&cnt->cursor[1], kChunkPrimary);
                        len = kChunkPrimary | 0x80000000u;
                    }
                }
                base = __shfl_sync(kFull, base, 0);
                len = __shfl_sync(kFull, len, 0);
                if (len & 0x80000000u) { heavy_done = true; len &= 0x7FFFFFFFu; }
                if (base >= total) { exhausted = true; continue; }
                chunk_next = base;
                chunk_end = min(base + len, total);
            }
            const unsigned idle = __ballot_sync(kFull, !has);
            const uint32_t avail = chunk_end - chunk_next;
            const uint32_t rank = __popc(idle & lanemask_lt);
            if (!has && rank < avail) {
                const uint32_t p = lists.pixel(a, chunk_next + rank);
                const int py = (int)(p / (uint32_t)a.width), px = (int)(p - (uint32_t)py * (uint32_t)a.width);
                pred = min(a.cost[p], (1u << 27) - 1u);
                PrimaryRay pr;
                generate_primary_ray_ool(&cam, a.width, a.height, px, py, &pr);
                seed = pr.seed;
                pixel = p;
                throughput = mk3(1.0f, 1.0f, 1.0f); radiance = mk3(0.0f, 0.0f, 0.0f);
                segment = 0;
                steps = 0;
                fast_ray_begin(r, a.sc, pr.o, pr.d);
                wrd = r.rd;
                has = true;
            }
            n_started += min((uint32_t)n_idle, avail);
            chunk_next += min((uint32_t)n_idle, avail);
            continue;
        }
        if (n_walk == 0) {
            if (exhausted && n_fin == 0) break;
            continue;
        }
        // ---------------- I / L / T: the phase that advances most lanes per instruction, or the leading path's ----------------
        int run = (n_i * 3 >= n_l && n_i * 3 >= n_t * 2) ? 1 : (n_l >= n_t * 2 ? 0 : 2);
        if (lead_phase >= 0 && lead_phase < 3) run = lead_phase;
        if (run == 1) {
            it_i++;
            bool go = can_i;
            const int need = (n_i + 1) >> 1;
#pragma unroll 1
            for (int b = 0; b < a.burst; b++) { // node burst: no census while at least half of the starters stay on internal nodes
                if (go) {
                    if ((r.cur & LINK_TLAS) != 0u && r.inst != GDPT_NO_INSTANCE) { // back to world space (main.glsl:316-327)
                        r.o = r.wo; r.d = r.wd; r.rd = wrd; r.inst = GDPT_NO_INSTANCE;
                    }
                    fast_step_node(a.sc, r, st);
                    steps++;
                    if (pend == LINK_NONE && fast_link_is_leaf(r.cur)) { pend = r.cur; r.cur = stack_pop(r, st); } // park the leaf, keep descending
                    go = lane_can_node(r.cur, pend);
                }
                if (__popc(__ballot_sync(kFull, go)) < need) break;
            }
        } else if (run == 0) {
            it_l++;
            if (can_l) {
                uint32_t leaf = pend;
                if (pend != LINK_NONE) pend = LINK_NONE;
                else { leaf = r.cur; r.cur = stack_pop(r, st); }
                fast_leaf_tests(a.sc, r, leaf);
                steps++;
            }
        } else {
            it_t++;
            if (can_t) {
                if (r.inst != GDPT_NO_INSTANCE) { r.o = r.wo; r.d = r.wd; r.rd = wrd; r.inst = GDPT_NO_INSTANCE; }
                fast_enter_instance(a.sc, r, st);
                steps++;
            }
        }
    }
    if (prof && lane == 0) {
        unsigned long long *w = a.warp_prof + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        w[0] = t_start; w[1] = global_ns(); w[2] = it_i; w[3] = it_l; w[4] = it_t; w[5] = it_f; w[6] = it_e; w[7] = n_started;
    }
    for (int off = 16; off > 0; off >>= 1) {
        my_rays += __shfl_down_sync(kFull, my_rays, off);
        my_phits += __shfl_down_sync(kFull, my_phits, off);
        my_retraced += __shfl_down_sync(kFull, my_retraced, off);
    }
    if (lane == 0) {
        atomicAdd(&cnt->rays, my_rays); atomicAdd(&cnt->primary_hits, my_phits);
        if (my_retraced) atomicAdd(&cnt->retraced, my_retraced);
    }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

// ------------------------------------------------------------------------------------------------
// Path kernel of schedule 6: k_path_fast's search (phases I / L / T over the closest-hit tables, proof,
// exact re-trace) with the PATHS taken out of the lanes.  A warp owns a pool of kPoolSlots path slots in
// shared memory (world ray, hit, throughput, radiance, seed, pixel, segment); a lane only ever holds the
// traversal state of one RAY and the number of its slot.  A lane whose search ends writes the hit into
// the slot, queues the slot for shading and takes the next ready ray, so lanes do not wait for a shading
// quorum; shading (verdict, material, BRDF sample, continuation ray) runs when 32 finished rays wait --
// a full warp per instruction instead of the 10-16 lanes of the quorum scheme -- and camera-ray generation
// refills free slots up to 32 at a time.  Every number a path produces is the one k_path_fast produces:
// the per-ray and per-path arithmetic is shared, only the lane that executes it differs.
constexpr int kPoolSlotsDefault = 64; // path slots per warp (template argument kPoolSlots)
constexpr int kPoolParkDefault = 1; // leaves a lane may park while it keeps descending (template argument kPoolPark)
enum PoolField {
    PF_WOX, PF_WOY, PF_WOZ, PF_WDX, PF_WDY, PF_WDZ,   // ray.o, ray.d (world)
    PF_T, PF_U, PF_V, PF_TRI, PF_BF, PF_FLAGS,        // finished search
    PF_THR, PF_THG, PF_THB, PF_RAR, PF_RAG, PF_RAB,   // throughput, radiance
    PF_SEEDX, PF_SEEDY, PF_PIXEL, PF_SEGMENT, PF_STEPS,
    PF_COUNT
};

// Phases of k_path_pool.  The node step never changes space: a TLAS link that comes up while the lane is inside an
// instance is a crossing (phase T), like an instance entry, so the node-step code carries no space restore.
__device__ __forceinline__ bool pool_can_node(uint32_t cur, uint32_t inst)
{
    return cur != LINK_NONE && (cur & LINK_LEAF) == 0u && ((cur & LINK_TLAS) == 0u || inst == GDPT_NO_INSTANCE);
}
__device__ __forceinline__ bool pool_can_cross(uint32_t cur, uint32_t pend, uint32_t inst)
{
    return cur != LINK_NONE && (cur & LINK_TLAS) != 0u && pend == LINK_NONE && ((cur & LINK_LEAF) != 0u || inst != GDPT_NO_INSTANCE);
}

template <bool REC, int MINB, int kPoolPark, int kPoolSlots, bool WIDE>
__global__ void __launch_bounds__(kTraceThreads, MINB) k_path_pool(const FrameArgs a)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    __shared__ uint32_t s_pool[kTraceThreads / 32][PF_COUNT * kPoolSlots];
    __shared__ uint8_t s_lists[kTraceThreads / 32][3 * kPoolSlots];
    __shared__ gdpt_camera s_cam;
    uint32_t spill[GDPT_MAX_STACK - kSmemStack];
    SmemStack st;
    st.col = s_stack + threadIdx.x;
    st.spill = spill;
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    static_assert(kPoolSlots >= 32 && kPoolSlots <= 255 && kPoolSlots % 8 == 0, "pool: 32..248 slots");
    const unsigned lane = threadIdx.x & 31u;
    uint32_t *const pool = s_pool[threadIdx.x >> 5];
    uint8_t *const ready_list = s_lists[threadIdx.x >> 5];
    uint8_t *const done_list = ready_list + kPoolSlots;
    uint8_t *const free_list = done_list + kPoolSlots;
    for (unsigned i = lane; i < (unsigned)kPoolSlots; i += 32u) free_list[i] = (uint8_t)(kPoolSlots - 1 - i);
    __syncthreads();
    const gdpt_camera &cam = s_cam;
#define PF(f, sl) pool[(f) * kPoolSlots + (sl)]
#define PFF(f, sl) __uint_as_float(pool[(f) * kPoolSlots + (sl)])

    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    SurvivorLists lists;
    lists.load(a);
    const uint32_t total = lists.total;
    const int swap_at = min(max(a.refill_below, 1), 32);   // lanes without a walking ray before the pool is serviced ...
    const uint32_t swap_wait = a.pool_wait > 0 ? (uint32_t)a.pool_wait : 0xFFFFFFFFu; // ... or this many lane-iterations spent waiting
    uint32_t waited = 0;
    const uint32_t min_free = (a.pool_alive >= 32 && a.pool_alive < kPoolSlots) ? (uint32_t)(kPoolSlots - a.pool_alive) : 0u;
    const int shade_low = min(max(a.shade_at, 1), 32);     // finished rays that justify a partial shading batch
    bool heavy_done = false;
    const int last_segment = a.max_depth - 1;

    RayState r;
    r.cur = LINK_NONE; r.sp = 0; r.overflow = 0; r.t = 1e9f; r.inst = GDPT_NO_INSTANCE;
    f3 wrd = mk3(0.0f, 0.0f, 0.0f);
    uint32_t park[kPoolPark];  // parked leaves, oldest first (instance-local: flushed before the space changes)
    uint32_t n_park = 0;
#pragma unroll
    for (int k = 0; k < kPoolPark; k++) park[k] = LINK_NONE;
    uint32_t slot = 0, steps = 0;
    bool has = false;
    uint32_t ready_count = 0, done_count = 0, free_count = (uint32_t)kPoolSlots; // warp-uniform
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = (total == 0u);
    unsigned long long my_rays = 0, my_phits = 0, my_retraced = 0;
    uint32_t my_overflow = 0;
    const bool prof = a.warp_prof != nullptr;
    const unsigned long long t_start = prof ? global_ns() : 0ull;
    uint32_t it_i = 0, it_l = 0, it_t = 0, it_f = 0, it_e = 0, n_started = 0;

    for (;;) {
        const uint32_t pend = n_park ? park[0] : LINK_NONE;
        const bool can_i = has && pool_can_node(r.cur, r.inst);
        const bool can_l = has && lane_can_leaf(r.cur, pend);
        const bool can_t = has && pool_can_cross(r.cur, pend, r.inst);
        const bool fin = has && r.cur == LINK_NONE && pend == LINK_NONE;
        const uint32_t census = __reduce_add_sync(kFull, (can_i ? 1u : 0u) | (can_l ? 1u << 6 : 0u) | (can_t ? 1u << 12 : 0u) |
                                                             (fin ? 1u << 18 : 0u) | (has ? 0u : 1u << 24));
        const int n_i = (int)(census & 63u), n_l = (int)((census >> 6) & 63u), n_t = (int)((census >> 12) & 63u),
                  n_fin = (int)((census >> 18) & 63u), n_idle = (int)(census >> 24);
        const int n_walk = 32 - n_idle - n_fin;
        const bool can_refill = !exhausted && free_count > min_free; // alive paths (slots in use) stay below the cap

        waited += (uint32_t)n_fin;
        if (done_count >= 32u || n_walk == 0 || n_fin >= swap_at || (n_fin > 0 && waited >= swap_wait) ||
            (n_idle >= swap_at && (ready_count > 0u || can_refill || done_count >= (uint32_t)shade_low))) {
            waited = 0;
            // ---------------- pool service ----------------
            // 1. retire: finished searches go to their slots, the slots to the shading queue
            if (n_fin > 0) {
                const unsigned m = __ballot_sync(kFull, fin);
                if (fin) {
                    PF(PF_T, slot) = __float_as_uint(r.t); PF(PF_U, slot) = __float_as_uint(r.u); PF(PF_V, slot) = __float_as_uint(r.v);
                    PF(PF_TRI, slot) = r.tri; PF(PF_BF, slot) = r.blas_front; PF(PF_FLAGS, slot) = r.overflow;
                    PF(PF_STEPS, slot) += steps;
                    done_list[done_count + (uint32_t)__popc(m & lanemask_lt)] = (uint8_t)slot;
                    has = false;
                }
                done_count += (uint32_t)__popc(m);
                __syncwarp();
            }
            const uint32_t n_out = (uint32_t)(n_fin + n_idle); // lanes that hold no walking ray now
            const bool shade = done_count >= 32u ||
                               (done_count > 0u && n_out > ready_count && !can_refill && (done_count >= (uint32_t)shade_low || n_walk == 0));
            if (shade) {
                // 2. S: prove and shade up to 32 finished rays, oldest first
                it_f++;
                const uint32_t n = min(done_count, 32u);
                const bool mine = lane < n;
                const uint32_t sl = mine ? (uint32_t)done_list[lane] : 0u;
                const uint32_t rest = done_count - n; // < 32
                const uint32_t moved = lane < rest ? (uint32_t)done_list[n + lane] : 0u;
                __syncwarp();
                if (lane < rest) done_list[lane] = (uint8_t)moved;
                done_count = rest;
                bool alive = false, dead = false;
                if (mine) {
                    const f3 wo = mk3(PFF(PF_WOX, sl), PFF(PF_WOY, sl), PFF(PF_WOZ, sl));
                    const f3 wd = mk3(PFF(PF_WDX, sl), PFF(PF_WDY, sl), PFF(PF_WDZ, sl));
                    float ht = PFF(PF_T, sl), hu = PFF(PF_U, sl), hv = PFF(PF_V, sl);
                    uint32_t htri = PF(PF_TRI, sl), hbf = PF(PF_BF, sl), hflags = PF(PF_FLAGS, sl);
                    const uint32_t pixel = PF(PF_PIXEL, sl);
                    const int segment = (int)PF(PF_SEGMENT, sl);
                    if (!fast_verdict_ool(&a.sc, wo, wd, ht, htri, hbf, hflags)) {
                        ExactHit eh; // rare: exact reference-order traversal of this ray
                        exact_retrace(&a.sc, wo, wd, &eh);
                        ht = eh.t; hu = eh.u; hv = eh.v; htri = eh.tri; hbf = eh.blas_front; hflags = eh.overflow & RAY_OVERFLOW;
                        my_retraced++;
                    }
                    const bool hit = ht < 1e9f;
                    my_rays++;
                    if (segment == 0 && hit) my_phits++;
                    if (REC) write_hit_record(a, segment, pixel, ht, hu, hv, htri, hbf);
                    my_overflow |= hflags & RAY_OVERFLOW;
                    f3 radiance = mk3(PFF(PF_RAR, sl), PFF(PF_RAG, sl), PFF(PF_RAB, sl));
                    const f3 throughput = mk3(PFF(PF_THR, sl), PFF(PF_THG, sl), PFF(PF_THB, sl));
                    if (!hit) {
                        radiance = radiance + throughput * sample_sky(wd);
                        if (segment == 0) a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                    } else {
                        BounceResult br;
                        u2 sd; sd.x = PF(PF_SEEDX, sl); sd.y = PF(PF_SEEDY, sl);
                        shade_and_bounce_ool(&a.sc, wo, wd, ht, hu, hv, htri, hbf, radiance, throughput, &sd, &br);
                        radiance = br.radiance;
                        if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
                        alive = br.alive && segment < last_segment;
                        if (alive) {
                            PF(PF_SEEDX, sl) = sd.x; PF(PF_SEEDY, sl) = sd.y;
                            PF(PF_THR, sl) = __float_as_uint(br.throughput.x); PF(PF_THG, sl) = __float_as_uint(br.throughput.y);
                            PF(PF_THB, sl) = __float_as_uint(br.throughput.z);
                            PF(PF_RAR, sl) = __float_as_uint(radiance.x); PF(PF_RAG, sl) = __float_as_uint(radiance.y);
                            PF(PF_RAB, sl) = __float_as_uint(radiance.z);
                            PF(PF_WOX, sl) = __float_as_uint(br.next_o.x); PF(PF_WOY, sl) = __float_as_uint(br.next_o.y);
                            PF(PF_WOZ, sl) = __float_as_uint(br.next_o.z);
                            PF(PF_WDX, sl) = __float_as_uint(br.next_d.x); PF(PF_WDY, sl) = __float_as_uint(br.next_d.y);
                            PF(PF_WDZ, sl) = __float_as_uint(br.next_d.z);
                            PF(PF_SEGMENT, sl) = (uint32_t)(segment + 1);
                        }
                    }
                    if (!alive) {
                        a.out_rgba8[pixel] = pack_rgba8(radiance);
                        // scheduling hint for the next frame: running mean of the path's cost at this pixel (a single
                        // frame's cost is one random walk; the mean says what the pixel usually sees)
                        if (a.cost_ema) {
                            const uint32_t old = a.cost[pixel];
                            a.cost[pixel] = old ? (old * 3u + PF(PF_STEPS, sl) + 2u) >> 2 : PF(PF_STEPS, sl);
                        } else {
                            a.cost[pixel] = PF(PF_STEPS, sl);
                        }
                        dead = true;
                    }
                }
                const unsigned m_alive = __ballot_sync(kFull, alive), m_dead = __ballot_sync(kFull, dead);
                if (alive) ready_list[ready_count + (uint32_t)__popc(m_alive & lanemask_lt)] = (uint8_t)sl;
                if (dead) free_list[free_count + (uint32_t)__popc(m_dead & lanemask_lt)] = (uint8_t)sl;
                ready_count += (uint32_t)__popc(m_alive);
                free_count += (uint32_t)__popc(m_dead);
                __syncwarp();
            } else if (n_out > ready_count && can_refill) {
                // 3. R: camera rays into free slots
                it_e++;
                if (chunk_next == chunk_end) {
                    uint32_t base = 0, len = kChunkPrimary;
                    if (lane == 0) {
                        if (!heavy_done) {
                            base = atomicAdd(&cnt->cursor[0], kChunkHeavy);
                            len = kChunkHeavy;
                            if (base >= lists.heavy_total) base = 0xFFFFFFFFu;
                            else if (base + len > lists.heavy_total) len = lists.heavy_total - base;
                        }
                        if (heavy_done || base == 0xFFFFFFFFu) {
                            base = lists.heavy_total + atomicAdd(&cnt->cursor[1], kChunkPrimary);
                            len = kChunkPrimary | 0x80000000u;
                        }
                    }
                    base = __shfl_sync(kFull, base, 0);
                    len = __shfl_sync(kFull, len, 0);
                    if (len & 0x80000000u) { heavy_done = true; len &= 0x7FFFFFFFu; }
                    if (base >= total) { exhausted = true; continue; }
                    chunk_next = base;
                    chunk_end = min(base + len, total);
                }
                const uint32_t g = min(min(chunk_end - chunk_next, free_count - min_free), 32u);
                if (lane < g) {
                    const uint32_t sl = (uint32_t)free_list[free_count - 1u - lane];
                    const uint32_t p = lists.pixel(a, chunk_next + lane);
                    const int py = (int)(p / (uint32_t)a.width), px = (int)(p - (uint32_t)py * (uint32_t)a.width);
                    PrimaryRay pr;
                    generate_primary_ray_ool(&cam, a.width, a.height, px, py, &pr);
                    PF(PF_WOX, sl) = __float_as_uint(pr.o.x); PF(PF_WOY, sl) = __float_as_uint(pr.o.y); PF(PF_WOZ, sl) = __float_as_uint(pr.o.z);
                    PF(PF_WDX, sl) = __float_as_uint(pr.d.x); PF(PF_WDY, sl) = __float_as_uint(pr.d.y); PF(PF_WDZ, sl) = __float_as_uint(pr.d.z);
                    PF(PF_THR, sl) = __float_as_uint(1.0f); PF(PF_THG, sl) = __float_as_uint(1.0f); PF(PF_THB, sl) = __float_as_uint(1.0f);
                    PF(PF_RAR, sl) = 0u; PF(PF_RAG, sl) = 0u; PF(PF_RAB, sl) = 0u;
                    PF(PF_SEEDX, sl) = pr.seed.x; PF(PF_SEEDY, sl) = pr.seed.y;
                    PF(PF_PIXEL, sl) = p; PF(PF_SEGMENT, sl) = 0u; PF(PF_STEPS, sl) = 0u;
                    ready_list[ready_count + lane] = (uint8_t)sl;
                }
                free_count -= g; ready_count += g; chunk_next += g; n_started += g;
                __syncwarp();
            }
            // 4. feed: lanes without a ray take ready slots
            if (ready_count > 0u) {
                const unsigned idle = __ballot_sync(kFull, !has);
                const uint32_t rank = (uint32_t)__popc(idle & lanemask_lt);
                if (!has && rank < ready_count) {
                    slot = (uint32_t)ready_list[ready_count - 1u - rank];
                    fast_ray_begin(r, a.sc, mk3(PFF(PF_WOX, slot), PFF(PF_WOY, slot), PFF(PF_WOZ, slot)),
                                   mk3(PFF(PF_WDX, slot), PFF(PF_WDY, slot), PFF(PF_WDZ, slot)));
                    if (WIDE) r.cur = a.sc.fast4_root; // the four-wide tables have their own root link
                    wrd = r.rd;
                    n_park = 0;
                    steps = 0;
                    has = true;
                }
                ready_count -= min((uint32_t)__popc(idle), ready_count);
                __syncwarp();
            }
            if (exhausted && ready_count == 0u && done_count == 0u && __ballot_sync(kFull, has) == 0u) break;
            continue;
        }
        // ---------------- I / L / T: the phase that advances most lanes per instruction ----------------
        const int run = (n_i * 3 >= n_l && n_i * 3 >= n_t * 2) ? 1 : (n_l >= n_t * 2 ? 0 : 2);
        if (run == 1) {
            it_i++;
            bool go = can_i;
            const int need = (n_i + 1) >> 1;
#pragma unroll 1
            for (int b = 0; b < a.burst; b++) { // node burst: no census while at least half of the starters stay on internal nodes
                if (go) {
                    if (WIDE) fast_step_node4(a.sc, r, st); else fast_step_node(a.sc, r, st);
                    steps++;
                    if (n_park < (uint32_t)kPoolPark && fast_link_is_leaf(r.cur)) { // park the leaf, keep descending
#pragma unroll
                        for (int k = 0; k < kPoolPark; k++)
                            if (n_park == (uint32_t)k) park[k] = r.cur;
                        n_park++;
                        r.cur = fast_pop(r, st);
                    }
                    go = pool_can_node(r.cur, r.inst);
                }
                if (__popc(__ballot_sync(kFull, go)) < need) break;
            }
        } else if (run == 0) {
            it_l++;
            if (can_l) {
                uint32_t leaf = park[0];
                if (n_park) {
#pragma unroll
                    for (int k = 0; k + 1 < kPoolPark; k++) park[k] = park[k + 1];
                    n_park--;
                } else { leaf = r.cur; r.cur = fast_pop(r, st); }
                fast_leaf_tests(a.sc, r, leaf);
                steps++;
            }
        } else {
            it_t++;
            if (can_t) { // back to world space (main.glsl:316-327) and/or into the instance the link names
                if (r.inst != GDPT_NO_INSTANCE) { r.o = r.wo; r.d = r.wd; r.rd = wrd; r.inst = GDPT_NO_INSTANCE; }
                if (r.cur & LINK_LEAF) fast_enter_instance<WIDE>(a.sc, r, st);
                steps++;
            }
        }
    }
#undef PF
#undef PFF
    if (prof && lane == 0) {
        unsigned long long *w = a.warp_prof + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        w[0] = t_start; w[1] = global_ns(); w[2] = it_i; w[3] = it_l; w[4] = it_t; w[5] = it_f; w[6] = it_e; w[7] = n_started;
    }
    for (int off = 16; off > 0; off >>= 1) {
        my_rays += __shfl_down_sync(kFull, my_rays, off);
        my_phits += __shfl_down_sync(kFull, my_phits, off);
        my_retraced += __shfl_down_sync(kFull, my_retraced, off);
    }
    if (lane == 0) {
        atomicAdd(&cnt->rays, my_rays); atomicAdd(&cnt->primary_hits, my_phits);
        if (my_retraced) atomicAdd(&cnt->retraced, my_retraced);
    }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

// Camera-ray classification (first kernel of the two-kernel schedule): one thread per pixel in
// 8x4-tile order, so a warp is one tile and runs in lockstep.  The thread generates its camera ray
// (main.glsl:405-421) and walks the TLAS level with tight-box culling.  A ray that reaches no
// instance whose tight box it touches cannot hit anything: it is finished here (sky colour +
// far depth, main.glsl:366-369,430-435) exactly as the full traversal would finish it.  The other
// pixels are appended, warp-aggregated, to `hit_list` for k_path<SRC=1>, which restarts them from
// the camera (the seed and the ray are functions of (pixel, frame_index) only).
__global__ void __launch_bounds__(kTraceThreads) k_primary_cull(const FrameArgs a)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    __shared__ gdpt_camera s_cam;
    uint32_t spill[GDPT_MAX_STACK - kSmemStack];
    SmemStack st;
    st.col = s_stack + threadIdx.x;
    st.spill = spill;
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    __syncthreads();
    const gdpt_camera &cam = s_cam;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    unsigned long long my_done = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t w0 = blockIdx.x * blockDim.x; w0 < a.n_work; w0 += stride) { // warp-uniform trip count
        const uint32_t w = w0 + threadIdx.x;
        int px, py;
        const bool valid = w < a.n_work && work_to_pixel(a, w, &px, &py);
        bool survivor = false;
        uint32_t pixel = 0;
        if (valid) {
            f3 o, d;
            generate_primary_ray(cam, a.width, a.height, px, py, &o, &d);
            pixel = (uint32_t)py * (uint32_t)a.width + (uint32_t)px;
            RayState r;
            ray_begin(r, a.sc, o, d);
            while (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) step_tlas<false, true>(a.sc, r, st, nullptr);
            survivor = r.cur != LINK_NONE;
            if (!survivor) {
                a.out_rgba8[pixel] = pack_rgba8(mk3(0.0f, 0.0f, 0.0f) + mk3(1.0f, 1.0f, 1.0f) * sample_sky(d));
                a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                if (a.trace) write_hit_record(a, 0, pixel, 1e9f, 0.0f, 0.0f, 0u, 0u); // GDPT_RECORD_HITS: a miss
                my_done++;
            }
        }
        // append to the list of the pixel's cost class (previous frame), one atomic per class present in the warp
        int cls = survivor ? cost_class(a.cost[pixel]) : -1;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) { // pass 1: lanes whose heavy list was full fall back to the light list
            const unsigned peers = __match_any_sync(kFull, cls);
            if (cls >= 0) {
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(&cnt->qcount[cls], (uint32_t)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                const uint32_t slot = base + __popc(peers & lanemask_lt);
                const uint32_t cap = cls == 0 ? a.queue_cap : a.heavy_cap;
                if (slot < cap) { a.hit_list[class_base(a, cls) + slot] = pixel; cls = -1; }
                else cls = 0;
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) my_done += __shfl_down_sync(kFull, my_done, off);
    if (lane == 0 && my_done) atomicAdd(&cnt->rays, my_done);
}

// ------------------------------------------------------------------------------------------------
// Lane-multiplexed path kernel (second kernel of schedule 4).
//
// k_path keeps one path per lane in registers; a warp then issues every traversal instruction for
// the handful of lanes that happen to be in the phase being executed (ncu: ~7 of 32 lanes in the
// box/triangle code) and its 114 registers leave 4 warps per scheduler to hide a dependent ALU
// chain.  Here every lane owns K path contexts ("slots") that live in shared memory -- word f of
// slot k of thread t sits at s_state[(f*K + k)*kMuxThreads + t], so a lane only ever touches bank
// (t mod 32) whatever slot it picks: conflict-free without any coordination.  Each iteration the
// warp counts, with one REDUX, how many lanes own a slot in each phase
//     I  one internal node (BLAS, or TLAS while outside an instance): two reference + two tight box tests
//     L  one triangle test of the pending leaf range
//     T  instance bookkeeping: leave the current instance and/or enter the popped one
//     F  segment finished: shade + sample the continuation ray, or finish the pixel
//     E  empty: take the next surviving pixel from k_primary_cull's list
// executes the phase most lanes can join, and every lane works on its first slot in that phase.
// A lane idles only when none of its K slots is in the chosen phase.  Registers hold no path state
// between iterations.  The steps are the same device functions (same node order, same float ops)
// the other schedules and the CPU unit check use, so results stay bit-identical.
//
// Cold per-path state (world ray, throughput, radiance, seed, pixel, segment) is an 80 B record in
// global memory (L2-resident), touched at T and F steps only.
constexpr int kMuxThreads = 128;
constexpr int kMuxStack = 8; // smem-resident stack entries per slot; deeper entries go to local memory
enum MuxField { MF_OX, MF_OY, MF_OZ, MF_DX, MF_DY, MF_DZ, MF_RX, MF_RY, MF_RZ, MF_T, MF_U, MF_V, MF_TRI, MF_BF,
                MF_CUR, MF_SP, MF_INST, MF_TN, MF_TE, MF_STEPS, MF_STACK, MF_COUNT = MF_STACK + kMuxStack };
enum MuxPhase : uint32_t { PH_E = 0, PH_I = 1, PH_L = 2, PH_T = 3, PH_F = 4 };

template <int K> struct MuxSlot {
    uint32_t *base; // &s_state[k * kMuxThreads + tid]
    uint32_t *spill;
    __device__ __forceinline__ uint32_t ldu(int f) const { return base[f * K * kMuxThreads]; }
    __device__ __forceinline__ float ldf(int f) const { return __uint_as_float(base[f * K * kMuxThreads]); }
    __device__ __forceinline__ void stu(int f, uint32_t v) const { base[f * K * kMuxThreads] = v; }
    __device__ __forceinline__ void stf(int f, float v) const { base[f * K * kMuxThreads] = __float_as_uint(v); }
    // stack interface of pt_trace.cuh
    __device__ __forceinline__ void store(uint32_t i, uint32_t v)
    {
        if (i < (uint32_t)kMuxStack) base[(MF_STACK + i) * K * kMuxThreads] = v; else spill[i - kMuxStack] = v;
    }
    __device__ __forceinline__ uint32_t load(uint32_t i) const
    {
        return (i < (uint32_t)kMuxStack) ? base[(MF_STACK + i) * K * kMuxThreads] : spill[i - kMuxStack];
    }
};

__device__ __forceinline__ uint32_t mux_phase_of(uint32_t cur, bool pending, uint32_t inst)
{
    if (pending || link_is_blas_leaf(cur)) return PH_L;
    if (cur == LINK_NONE) return PH_F;
    if (cur & LINK_TLAS) return (inst != GDPT_NO_INSTANCE || (cur & LINK_LEAF)) ? PH_T : PH_I;
    return PH_I;
}

// global path record: five 128-bit quads
//   q0 = wo.xyz, pixel   q1 = wd.xyz, segment   q2 = wrd.xyz, seed.x   q3 = throughput.rgb, seed.y   q4 = radiance.rgb, -
template <int K>
__device__ __forceinline__ void mux_begin_segment(const SceneView &sc, const MuxSlot<K> &m, float4 *rec, f3 o, f3 d, f3 thr, f3 rad,
                                                  u2 seed, uint32_t pixel, uint32_t segment)
{
    const f3 rd = rcp3(d);
    rec[0] = make_float4(o.x, o.y, o.z, __uint_as_float(pixel));
    rec[1] = make_float4(d.x, d.y, d.z, __uint_as_float(segment));
    rec[2] = make_float4(rd.x, rd.y, rd.z, __uint_as_float(seed.x));
    rec[3] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(seed.y));
    rec[4] = make_float4(rad.x, rad.y, rad.z, 0.0f);
    m.stf(MF_OX, o.x); m.stf(MF_OY, o.y); m.stf(MF_OZ, o.z);
    m.stf(MF_DX, d.x); m.stf(MF_DY, d.y); m.stf(MF_DZ, d.z);
    m.stf(MF_RX, rd.x); m.stf(MF_RY, rd.y); m.stf(MF_RZ, rd.z);
    m.stf(MF_T, 1e9f); m.stf(MF_U, 0.0f); m.stf(MF_V, 0.0f); m.stu(MF_TRI, 0u); m.stu(MF_BF, 0u);
    m.stu(MF_CUR, sc.tlas_root_link); m.stu(MF_SP, 0u); m.stu(MF_INST, GDPT_NO_INSTANCE);
    m.stu(MF_TN, 0u); m.stu(MF_TE, 0u);
    if (segment == 0u) m.stu(MF_STEPS, 0u);
}

template <int K, int MINB>
__global__ void __launch_bounds__(kMuxThreads, MINB) k_path_mux(const FrameArgs a)
{
    extern __shared__ uint32_t s_state[]; // MF_COUNT * K * kMuxThreads words
    __shared__ gdpt_camera s_cam;
    uint32_t spill[K * (GDPT_MAX_STACK - kMuxStack)];
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    __syncthreads();
    const gdpt_camera &cam = s_cam;
    const SceneView &sc = a.sc;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    SurvivorLists lists;
    lists.load(a);
    const uint32_t total = lists.total;
    float4 *my_recs = a.path_recs + ((size_t)blockIdx.x * kMuxThreads + threadIdx.x) * (size_t)(K * 5);
    const int last_segment = a.max_depth - 1;
    const int refill_at = min(max(a.refill_below, 1), 32); // R once this many lanes own an empty slot
    const int shade_at = min(max(a.shade_at, 1), 32);      // F once this many lanes own a finished segment

    uint32_t ph = 0u; // 4 bits per slot, all PH_E
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = (total == 0u);
    unsigned long long my_rays = 0, my_phits = 0;
    uint32_t my_overflow = 0;
    const bool prof = a.warp_prof != nullptr;
    const unsigned long long t_start = prof ? global_ns() : 0ull;
    uint32_t it_count[5] = { 0, 0, 0, 0, 0 }, n_started = 0; // indexed by MuxPhase

    for (;;) {
        // ---- which phases do my slots offer?  one REDUX gives the per-phase lane counts
        uint32_t offer = 0u; // bit p: some slot of this lane is in phase p
#pragma unroll
        for (int k = 0; k < K; k++) offer |= 1u << ((ph >> (4 * k)) & 15u);
        uint32_t packed = 0u;
#pragma unroll
        for (int p = 0; p < 5; p++) packed |= ((offer >> p) & 1u) << (6 * p);
        packed = __reduce_add_sync(kFull, packed);
        const int n_e = (int)(packed & 63u), n_i = (int)((packed >> 6) & 63u), n_l = (int)((packed >> 12) & 63u),
                  n_t = (int)((packed >> 18) & 63u), n_f = (int)((packed >> 24) & 63u);
        const bool can_fill = !exhausted && n_e > 0;
        const int n_walk_best = max(n_i, max(n_l, n_t));

        uint32_t phase;
        if (can_fill && (n_e >= refill_at || (n_walk_best == 0 && n_f == 0))) phase = PH_E;
        else if (n_f > 0 && (n_f >= shade_at || n_walk_best == 0)) phase = PH_F;
        else if (n_walk_best == 0) break; // nothing walking, nothing to shade, no work left
        else if (n_l >= n_i && n_l >= n_t) phase = PH_L;
        else if (n_i >= n_t) phase = PH_I;
        else phase = PH_T;
        if (prof) {
#pragma unroll
            for (int p = 0; p < 5; p++) it_count[p] += (phase == (uint32_t)p) ? 1u : 0u;
        }

        // ---- my first slot in that phase
        int k = -1;
#pragma unroll
        for (int j = K - 1; j >= 0; j--)
            if (((ph >> (4 * j)) & 15u) == phase) k = j;
        MuxSlot<K> m;
        m.base = s_state + (k < 0 ? 0 : k) * kMuxThreads + threadIdx.x;
        m.spill = spill + (k < 0 ? 0 : k) * (GDPT_MAX_STACK - kMuxStack);
        uint32_t next_phase = phase;

        if (phase == PH_I) {
            if (k >= 0) {
                RayState r;
                r.o = mk3(m.ldf(MF_OX), m.ldf(MF_OY), m.ldf(MF_OZ));
                r.rd = mk3(m.ldf(MF_RX), m.ldf(MF_RY), m.ldf(MF_RZ));
                r.t = m.ldf(MF_T); r.cur = m.ldu(MF_CUR); r.sp = m.ldu(MF_SP); r.overflow = 0u;
                const void *table = (r.cur & LINK_TLAS) ? (const void *)sc.wide_tlas : (const void *)sc.wide_nodes;
                visit_wide<false, true>(table, r.cur & LINK_INDEX_MASK, 0u, r, m, nullptr);
                m.stu(MF_CUR, r.cur); m.stu(MF_SP, r.sp);
                my_overflow |= r.overflow;
                uint32_t inst = GDPT_NO_INSTANCE;
                if (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) inst = m.ldu(MF_INST);
                next_phase = mux_phase_of(r.cur, false, inst);
            }
        } else if (phase == PH_L) {
            if (k >= 0) {
                RayState r;
                uint32_t tn = m.ldu(MF_TN), te = m.ldu(MF_TE);
                r.cur = m.ldu(MF_CUR);
                r.inst = m.ldu(MF_INST);
                if (tn == te) { // entering the leaf: fetch its range, pop what comes after it
                    r.sp = m.ldu(MF_SP);
                    const q4u leaf = ldqu(sc.leaf_recs, r.cur & LINK_INDEX_MASK);
                    tn = leaf.x; te = leaf.x + leaf.y;
                    r.cur = stack_pop(r, m);
                    m.stu(MF_CUR, r.cur); m.stu(MF_SP, r.sp); m.stu(MF_TE, te);
                }
                if (tn < te) {
                    r.o = mk3(m.ldf(MF_OX), m.ldf(MF_OY), m.ldf(MF_OZ));
                    r.d = mk3(m.ldf(MF_DX), m.ldf(MF_DY), m.ldf(MF_DZ));
                    r.t = m.ldf(MF_T); r.blas_front = m.ldu(MF_BF); r.tri = 0xFFFFFFFFu;
                    triangle_test(sc, r, tn);
                    if (r.tri != 0xFFFFFFFFu) { // accepted
                        m.stf(MF_T, r.t); m.stf(MF_U, r.u); m.stf(MF_V, r.v); m.stu(MF_TRI, r.tri); m.stu(MF_BF, r.blas_front);
                    }
                    tn++;
                }
                m.stu(MF_TN, tn);
                next_phase = mux_phase_of(r.cur, tn < te, r.inst);
            }
        } else if (phase == PH_T) {
            if (k >= 0) {
                RayState r;
                r.cur = m.ldu(MF_CUR); r.sp = m.ldu(MF_SP); r.inst = m.ldu(MF_INST); r.overflow = 0u;
                float4 *rec = my_recs + k * 5;
                if (r.inst != GDPT_NO_INSTANCE) { // back to world space (main.glsl:316-327: the TLAS loop uses `ray`)
                    const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2];
                    r.o = mk3(q0.x, q0.y, q0.z); r.d = mk3(q1.x, q1.y, q1.z); r.rd = mk3(q2.x, q2.y, q2.z);
                    r.inst = GDPT_NO_INSTANCE;
                    if (!(r.cur & LINK_LEAF)) {
                        m.stf(MF_OX, r.o.x); m.stf(MF_OY, r.o.y); m.stf(MF_OZ, r.o.z);
                        m.stf(MF_DX, r.d.x); m.stf(MF_DY, r.d.y); m.stf(MF_DZ, r.d.z);
                        m.stf(MF_RX, r.rd.x); m.stf(MF_RY, r.rd.y); m.stf(MF_RZ, r.rd.z);
                    }
                } else {
                    r.o = mk3(m.ldf(MF_OX), m.ldf(MF_OY), m.ldf(MF_OZ));
                    r.d = mk3(m.ldf(MF_DX), m.ldf(MF_DY), m.ldf(MF_DZ));
                }
                if (r.cur & LINK_LEAF) {
                    r.wo = r.o; r.wd = r.d;
                    r.t = m.ldf(MF_T); // distance culling of the instance's tight box
                    step_tlas<false, true>(sc, r, m, nullptr); // inst == NONE here: enters the instance (and may leave it at once)
                    m.stf(MF_OX, r.o.x); m.stf(MF_OY, r.o.y); m.stf(MF_OZ, r.o.z);
                    m.stf(MF_DX, r.d.x); m.stf(MF_DY, r.d.y); m.stf(MF_DZ, r.d.z);
                    m.stf(MF_RX, r.rd.x); m.stf(MF_RY, r.rd.y); m.stf(MF_RZ, r.rd.z);
                }
                m.stu(MF_CUR, r.cur); m.stu(MF_SP, r.sp); m.stu(MF_INST, r.inst);
                my_overflow |= r.overflow;
                next_phase = mux_phase_of(r.cur, false, r.inst);
            }
        } else if (phase == PH_F) {
            if (k >= 0) {
                float4 *rec = my_recs + k * 5;
                const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3], q4 = rec[4];
                const f3 wo = mk3(q0.x, q0.y, q0.z), wd = mk3(q1.x, q1.y, q1.z);
                const uint32_t pixel = __float_as_uint(q0.w);
                const int segment = (int)__float_as_uint(q1.w);
                u2 seed; seed.x = __float_as_uint(q2.w); seed.y = __float_as_uint(q3.w);
                f3 throughput = mk3(q3.x, q3.y, q3.z), radiance = mk3(q4.x, q4.y, q4.z);
                const float t = m.ldf(MF_T);
                const bool hit = t < 1e9f;
                my_rays++;
                if (segment == 0 && hit) my_phits++;
                bool alive = false;
                if (!hit) {
                    radiance = radiance + throughput * sample_sky(wd);
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                } else {
                    BounceResult br;
                    if (kOutOfLineCold) shade_and_bounce_ool(&sc, wo, wd, t, m.ldf(MF_U), m.ldf(MF_V), m.ldu(MF_TRI), m.ldu(MF_BF), radiance, throughput, &seed, &br);
                    else br = shade_and_bounce(sc, wo, wd, t, m.ldf(MF_U), m.ldf(MF_V), m.ldu(MF_TRI), m.ldu(MF_BF), radiance, throughput, seed);
                    radiance = br.radiance;
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
                    alive = br.alive && segment < last_segment;
                    if (alive) mux_begin_segment<K>(sc, m, rec, br.next_o, br.next_d, br.throughput, radiance, seed, pixel, (uint32_t)segment + 1u);
                }
                if (!alive) {
                    a.out_rgba8[pixel] = pack_rgba8(radiance);
                    a.cost[pixel] = m.ldu(MF_STEPS);
                    next_phase = PH_E;
                } else {
                    next_phase = mux_phase_of(sc.tlas_root_link, false, GDPT_NO_INSTANCE);
                }
            }
        } else { // PH_E: one new pixel for every lane that owns an empty slot
            if (chunk_next == chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt->cursor[1], kChunkPrimary);
                base = __shfl_sync(kFull, base, 0);
                if (base >= total) { exhausted = true; continue; }
                chunk_next = base;
                chunk_end = min(base + kChunkPrimary, total);
            }
            const unsigned want = __ballot_sync(kFull, k >= 0);
            const uint32_t avail = chunk_end - chunk_next;
            const uint32_t rank = __popc(want & lanemask_lt);
            if (k >= 0 && rank < avail) {
                const uint32_t pixel = lists.pixel(a, chunk_next + rank);
                const int py = (int)(pixel / (uint32_t)a.width), px = (int)(pixel - (uint32_t)py * (uint32_t)a.width);
                f3 o, d;
                u2 seed;
                if (kOutOfLineCold) {
                    PrimaryRay pr;
                    generate_primary_ray_ool(&cam, a.width, a.height, px, py, &pr);
                    o = pr.o; d = pr.d; seed = pr.seed;
                } else {
                    seed = generate_primary_ray(cam, a.width, a.height, px, py, &o, &d);
                }
                mux_begin_segment<K>(sc, m, my_recs + k * 5, o, d, mk3(1.0f, 1.0f, 1.0f), mk3(0.0f, 0.0f, 0.0f), seed, pixel, 0u);
                next_phase = mux_phase_of(sc.tlas_root_link, false, GDPT_NO_INSTANCE);
            }
            n_started += min((uint32_t)__popc(want), avail);
            chunk_next += min((uint32_t)__popc(want), avail);
        }
        if (k >= 0) {
            ph = (ph & ~(15u << (4 * k))) | (next_phase << (4 * k));
            if (phase >= PH_I && phase <= PH_T) m.stu(MF_STEPS, m.ldu(MF_STEPS) + 1u);
        }
    }
    if (prof && lane == 0) {
        unsigned long long *w = a.warp_prof + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        w[0] = t_start; w[1] = global_ns(); w[2] = it_count[PH_I]; w[3] = it_count[PH_L]; w[4] = it_count[PH_T];
        w[5] = it_count[PH_F]; w[6] = it_count[PH_E]; w[7] = n_started;
    }

    for (int off = 16; off > 0; off >>= 1) {
        my_rays += __shfl_down_sync(kFull, my_rays, off);
        my_phits += __shfl_down_sync(kFull, my_phits, off);
    }
    if (lane == 0) { atomicAdd(&cnt->rays, my_rays); atomicAdd(&cnt->primary_hits, my_phits); }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

// shade(segment): entries come from queue `src` (through the hit list for segment > 0),
// continuation rays go to queue `src ^ 1`.
__global__ void __launch_bounds__(kShadeThreads) k_shade(const FrameArgs a, const int segment, const int src)
{
    FrameCounters *cnt = a.counters;
    const uint32_t n = (segment == 0) ? min(cnt->qcount[0], a.queue_cap) : min(cnt->lcount[segment], a.queue_cap);
    const bool last = (segment == a.max_depth - 1);
    const int dst = src ^ 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lanemask_lt = (1u << lane) - 1u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const gdpt_camera cam = *a.camera;

    for (uint32_t base = warp_id * 32u; base < n; base += warps_total * 32u) {
        const uint32_t k = base + lane;
        const bool valid = k < n;
        bool alive = false;
        BounceResult br;
        uint32_t pixel = 0;
        u2 seed; seed.x = seed.y = 0u;
        if (valid) {
            const uint32_t e = (segment == 0) ? k : a.hit_list[k];
            const float4 p0 = plane(a, src, 0)[e], p1 = plane(a, src, 1)[e], p2 = plane(a, src, 2)[e],
                         p3 = plane(a, src, 3)[e], p4 = plane(a, src, 4)[e];
            pixel = __float_as_uint(p0.w);
            seed.x = __float_as_uint(p4.x); seed.y = __float_as_uint(p4.y);
            br = shade_and_bounce(a.sc, mk3(p0.x, p0.y, p0.z), mk3(p1.x, p1.y, p1.z), p1.w, p2.w, p3.w,
                                  __float_as_uint(p4.z), __float_as_uint(p4.w), mk3(p3.x, p3.y, p3.z),
                                  mk3(p2.x, p2.y, p2.z), seed);
            if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
            alive = br.alive && !last;
            if (!alive) a.out_rgba8[pixel] = pack_rgba8(br.radiance);
        }
        const unsigned am = __ballot_sync(kFull, alive);
        if (am != 0u) {
            uint32_t slot0 = 0;
            const int leader = __ffs(am) - 1;
            if ((int)lane == leader) slot0 = atomicAdd(&cnt->qcount[segment + 1], (uint32_t)__popc(am));
            slot0 = __shfl_sync(kFull, slot0, leader);
            if (alive) {
                const uint32_t slot = slot0 + __popc(am & lanemask_lt);
                if (slot < a.queue_cap) {
                    plane(a, dst, 0)[slot] = make_float4(br.next_o.x, br.next_o.y, br.next_o.z, __uint_as_float(pixel));
                    plane(a, dst, 1)[slot] = make_float4(br.next_d.x, br.next_d.y, br.next_d.z, 0.0f);
                    plane(a, dst, 2)[slot] = make_float4(br.throughput.x, br.throughput.y, br.throughput.z, 0.0f);
                    plane(a, dst, 3)[slot] = make_float4(br.radiance.x, br.radiance.y, br.radiance.z, 0.0f);
                    plane(a, dst, 4)[slot] = make_float4(__uint_as_float(seed.x), __uint_as_float(seed.y), 0.0f, 0.0f);
                }
            }
        }
    }
}

// K2 (progressive_rendering.glsl:28-46): acc = (frame_count > 1 ? acc : 0) + screen;
// screen = rgba8(ACES(acc / frame_count)).  Four pixels per thread: one 128-bit
// load of the RGBA8 quad, four 128-bit loads/stores of the RGBA32F accumulator.
// `peers`: RGBA8 images of the other GPUs of a row-band frame (peer memory over NVLink, CUDA IPC).  Every
// tone-mapped quad this GPU owns is also stored there, so the presented frame assembles itself in every
// GPU's image while the kernel runs -- the exchange step of the row-band partition fused into its producer
// instead of an all-gather after it.
// `raw` is the image K1 wrote (the bound screen image itself in the reference's call sequence; a frame-private image
// when two pipelined frames overlap), `screen` receives the tone-mapped result.
__global__ void __launch_bounds__(256) k_progressive(const uint32_t *raw, uint32_t *screen, float4 *__restrict__ accum,
                                                     const gdpt_progressive_params *__restrict__ params, int width,
                                                     int height, int shard_part, int shard_parts, int shard_band,
                                                     const PeerScreens peers)
{
    // imageLoad of an rgba8 texel = byte / 255.0f (progressive_rendering.glsl:33): every block divides each of the 256
    // byte values once and looks the quotients up afterwards -- the same IEEE quotients, three divisions per pixel fewer
    __shared__ float s_unorm[256];
    s_unorm[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const uint32_t frame_count = params->frame_count;
    const float fc = (float)frame_count;
    const size_t n_quads = ((size_t)width * height) >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += stride) {
        const size_t p = q << 2;
        if (shard_parts > 1) { // width is a multiple of 4 here, so a quad never straddles rows
            const int y = (int)(p / (size_t)width);
            if ((y / shard_band) % shard_parts != shard_part) continue;
        }
        const uint4 s4 = reinterpret_cast<const uint4 *>(raw)[q];
        const uint32_t in[4] = { s4.x, s4.y, s4.z, s4.w };
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            f3 rad = mk3(s_unorm[in[k] & 0xffu], s_unorm[(in[k] >> 8) & 0xffu], s_unorm[(in[k] >> 16) & 0xffu]);
            if (frame_count > 1u) {
                const float4 acc = accum[p + k];
                rad = rad + mk3(acc.x, acc.y, acc.z);
            }
            accum[p + k] = make_float4(rad.x, rad.y, rad.z, 1.0f);
            const f3 avg = (rad / fc) * 1.0f;
            out[k] = pack_rgba8(mk3(aces_channel(avg.x), aces_channel(avg.y), aces_channel(avg.z)));
        }
        const uint4 o4 = make_uint4(out[0], out[1], out[2], out[3]);
        reinterpret_cast<uint4 *>(screen)[q] = o4;
        for (int k = 0; k < peers.n; k++) reinterpret_cast<uint4 *>(peers.p[k])[q] = o4;
    }
    // tail pixels when W*H is not a multiple of 4
    if (blockIdx.x == 0 && threadIdx.x < (((size_t)width * height) & 3)) {
        const size_t p = (n_quads << 2) + threadIdx.x;
        const int y = (int)(p / (size_t)width);
        if (shard_parts <= 1 || (y / shard_band) % shard_parts == shard_part) {
            const uint32_t in = raw[p];
            f3 rad = mk3(s_unorm[in & 0xffu], s_unorm[(in >> 8) & 0xffu], s_unorm[(in >> 16) & 0xffu]);
            if (frame_count > 1u) { const float4 acc = accum[p]; rad = rad + mk3(acc.x, acc.y, acc.z); }
            accum[p] = make_float4(rad.x, rad.y, rad.z, 1.0f);
            const f3 avg = (rad / fc) * 1.0f;
            const uint32_t o = pack_rgba8(mk3(aces_channel(avg.x), aces_channel(avg.y), aces_channel(avg.z)));
            screen[p] = o;
            for (int k = 0; k < peers.n; k++) peers.p[k][p] = o;
        }
    }
}

// K3 (temporal_reprojection.glsl:31-71): one thread per pixel, rows of 32 x 8 pixel tiles.  Per pixel: 4 B + 4 B of
// the current frame, a 4 B depth and a 16 B history gather at the reprojected position (the same position while the
// camera rests), 16 B + 4 B of stores.  `history` / `next` are chosen by the host from frameCount's parity (:46).
__global__ void __launch_bounds__(256) k_temporal(uint32_t *__restrict__ screen, const float *__restrict__ depth,
                                                  const float *__restrict__ history, float *__restrict__ next,
                                                  const gdpt_temporal_params *__restrict__ params)
{
    __shared__ gdpt_temporal_params p;
    if (threadIdx.x < sizeof(gdpt_temporal_params) / 4u)
        reinterpret_cast<uint32_t *>(&p)[threadIdx.x] = reinterpret_cast<const uint32_t *>(params)[threadIdx.x];
    __syncthreads();
    const int tiles_x = (p.width + 31) >> 5, tiles_y = (p.height + 7) >> 3;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int x = (tile % tiles_x) * 32 + (int)(threadIdx.x & 31u), y = (tile / tiles_x) * 8 + (int)(threadIdx.x >> 5);
        if (x < p.width && y < p.height) temporal_pixel(p, x, y, screen, depth, history, next);
    }
}

struct Shapes {
    bool ready = false;
    int sms = 148;
    int trace_blocks[2][2][2] = {}; // [TRACE][MODE][CULL]
    int path_blocks[2][2] = {};     // [TRACE][CULL], SRC 0
    int path_list_blocks[2] = {};   // [TRACE], CULL, SRC 1
    int path_list_blocks_minb[9] = {}; // untraced, by MINB (register cap variants)
    int mux_blocks[5] = {};         // k_path_mux<K>, index K
    int fast_blocks[2][9] = {};     // k_path_fast<REC, MINB>
    int pool_blocks[2][9] = {};     // k_path_pool<REC, MINB>
    int cull_blocks_per_sm = 1;
    int shade_blocks = 0;
    int prog_blocks = 0;
};
Shapes g_shapes[16];

template <bool TRACE, int MODE, bool CULL> int trace_grid(int sms)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<TRACE, MODE, CULL>, kTraceThreads, 0);
    return sms * (per_sm > 0 ? per_sm : 1);
}

constexpr size_t mux_smem_bytes(int k) { return (size_t)MF_COUNT * k * kMuxThreads * sizeof(uint32_t); }

template <int K, int MINB> int mux_grid(int sms)
{
    cudaFuncSetAttribute(k_path_mux<K, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mux_smem_bytes(K));
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path_mux<K, MINB>, kMuxThreads, mux_smem_bytes(K));
    return sms * (per_sm > 0 ? per_sm : 1);
}

template <int MODE> void launch_trace_kernel(const Shapes &sh, const FrameArgs &a, bool trace, int segment, int src, cudaStream_t s)
{
    const bool cull = a.cull != 0;
    const int blocks = sh.trace_blocks[trace ? 1 : 0][MODE][cull ? 1 : 0];
    if (trace) {
        if (cull) k_trace<true, MODE, true><<<blocks, kTraceThreads, 0, s>>>(a, segment, src);
        else k_trace<true, MODE, false><<<blocks, kTraceThreads, 0, s>>>(a, segment, src);
    } else {
        if (cull) k_trace<false, MODE, true><<<blocks, kTraceThreads, 0, s>>>(a, segment, src);
        else k_trace<false, MODE, false><<<blocks, kTraceThreads, 0, s>>>(a, segment, src);
    }
}

} // namespace

void init_launch_shapes(int device)
{
    Shapes &s = g_shapes[device & 15];
    if (s.ready) return;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    s.sms = prop.multiProcessorCount;
    s.trace_blocks[0][0][0] = trace_grid<false, 0, false>(s.sms); s.trace_blocks[0][0][1] = trace_grid<false, 0, true>(s.sms);
    s.trace_blocks[0][1][0] = trace_grid<false, 1, false>(s.sms); s.trace_blocks[0][1][1] = trace_grid<false, 1, true>(s.sms);
    s.trace_blocks[1][0][0] = trace_grid<true, 0, false>(s.sms); s.trace_blocks[1][0][1] = trace_grid<true, 0, true>(s.sms);
    s.trace_blocks[1][1][0] = trace_grid<true, 1, false>(s.sms); s.trace_blocks[1][1][1] = trace_grid<true, 1, true>(s.sms);
    int per_sm = 0;
    auto grid_of = [&](auto kernel, int threads) {
        per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
        return s.sms * (per_sm > 0 ? per_sm : 1);
    };
    s.path_blocks[0][0] = grid_of(k_path<false, false, 0>, kTraceThreads);
    s.path_blocks[0][1] = grid_of(k_path<false, true, 0>, kTraceThreads);
    s.path_blocks[1][0] = grid_of(k_path<true, false, 0>, kTraceThreads);
    s.path_blocks[1][1] = grid_of(k_path<true, true, 0>, kTraceThreads);
    s.path_list_blocks[0] = grid_of(k_path<false, true, 1>, kTraceThreads);
    s.path_list_blocks[1] = grid_of(k_path<true, true, 1>, kTraceThreads);
    s.path_list_blocks_minb[1] = grid_of(k_path<false, true, 1, 4, true>, kTraceThreads);
    s.path_list_blocks_minb[2] = grid_of(k_path<false, true, 1, 6, true>, kTraceThreads);
    s.path_list_blocks_minb[5] = grid_of(k_path<false, true, 1, 5>, kTraceThreads);
    s.path_list_blocks_minb[6] = grid_of(k_path<false, true, 1, 6>, kTraceThreads);
    s.path_list_blocks_minb[8] = grid_of(k_path<false, true, 1, 8>, kTraceThreads);
    s.fast_blocks[0][4] = grid_of(k_path_fast<false, 4>, kTraceThreads);
    s.fast_blocks[0][5] = grid_of(k_path_fast<false, 5>, kTraceThreads);
    s.fast_blocks[0][6] = grid_of(k_path_fast<false, 6>, kTraceThreads);
    s.fast_blocks[0][8] = grid_of(k_path_fast<false, 8>, kTraceThreads);
    s.fast_blocks[1][4] = grid_of(k_path_fast<true, 4>, kTraceThreads);
    s.pool_blocks[0][0] = grid_of(k_path_pool<false, 4, kPoolParkDefault, kPoolSlotsDefault, false>, kTraceThreads);
    s.pool_blocks[0][1] = grid_of(k_path_pool<false, 4, kPoolParkDefault, kPoolSlotsDefault, true>, kTraceThreads);
    s.pool_blocks[1][1] = grid_of(k_path_pool<true, 4, kPoolParkDefault, kPoolSlotsDefault, true>, kTraceThreads);
    s.pool_blocks[0][2] = grid_of(k_path_pool<false, 3, kPoolParkDefault, kPoolSlotsDefault, true>, kTraceThreads);
    s.pool_blocks[1][0] = grid_of(k_path_pool<true, 4, kPoolParkDefault, kPoolSlotsDefault, false>, kTraceThreads);
    s.mux_blocks[1] = mux_grid<1, 8>(s.sms);
    s.mux_blocks[2] = mux_grid<2, 8>(s.sms);
    s.mux_blocks[3] = mux_grid<3, 5>(s.sms);
    s.mux_blocks[4] = mux_grid<4, 4>(s.sms);
    grid_of(k_primary_cull, kTraceThreads);
    s.cull_blocks_per_sm = per_sm > 0 ? per_sm : 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_shade, kShadeThreads, 0); s.shade_blocks = s.sms * (per_sm > 0 ? per_sm : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_progressive, 256, 0); s.prog_blocks = s.sms * (per_sm > 0 ? per_sm : 1);
    s.ready = true;
}

static Shapes &shapes_for_current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    init_launch_shapes(dev);
    return g_shapes[dev & 15];
}

void launch_path(const FrameArgs &a, bool trace, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    const bool cull = a.cull != 0;
    const int blocks = sh.path_blocks[trace ? 1 : 0][cull ? 1 : 0];
    if (trace) {
        if (cull) k_path<true, true, 0><<<blocks, kTraceThreads, 0, s>>>(a);
        else k_path<true, false, 0><<<blocks, kTraceThreads, 0, s>>>(a);
    } else {
        if (cull) k_path<false, true, 0><<<blocks, kTraceThreads, 0, s>>>(a);
        else k_path<false, false, 0><<<blocks, kTraceThreads, 0, s>>>(a);
    }
}

// grid of a persistent kernel: all resident blocks, or a.blocks_per_sm per SM when that is smaller
static int persistent_grid(const Shapes &sh, const FrameArgs &a, int resident)
{
    if (a.blocks_per_sm > 0 && a.blocks_per_sm * sh.sms < resident) return a.blocks_per_sm * sh.sms;
    return resident;
}

void launch_primary_cull(const FrameArgs &a, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    const int needed = (int)((a.n_work + kTraceThreads - 1) / kTraceThreads);
    const int resident = sh.sms * sh.cull_blocks_per_sm;
    k_primary_cull<<<needed < resident ? (needed > 0 ? needed : 1) : resident, kTraceThreads, 0, s>>>(a);
}

void launch_path_list(const FrameArgs &a, bool trace, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    if (trace) k_path<true, true, 1><<<persistent_grid(sh, a, sh.path_list_blocks[1]), kTraceThreads, 0, s>>>(a);
    else if (a.path_minb == 1) k_path<false, true, 1, 4, true><<<persistent_grid(sh, a, sh.path_list_blocks_minb[1]), kTraceThreads, 0, s>>>(a);
    else if (a.path_minb == 2) k_path<false, true, 1, 6, true><<<persistent_grid(sh, a, sh.path_list_blocks_minb[2]), kTraceThreads, 0, s>>>(a);
    else if (a.path_minb == 5) k_path<false, true, 1, 5><<<persistent_grid(sh, a, sh.path_list_blocks_minb[5]), kTraceThreads, 0, s>>>(a);
    else if (a.path_minb == 6) k_path<false, true, 1, 6><<<persistent_grid(sh, a, sh.path_list_blocks_minb[6]), kTraceThreads, 0, s>>>(a);
    else if (a.path_minb == 8) k_path<false, true, 1, 8><<<persistent_grid(sh, a, sh.path_list_blocks_minb[8]), kTraceThreads, 0, s>>>(a);
    else k_path<false, true, 1><<<persistent_grid(sh, a, sh.path_list_blocks[0]), kTraceThreads, 0, s>>>(a);
}

void launch_primary(const FrameArgs &a, bool trace, cudaStream_t s)
{
    launch_trace_kernel<0>(shapes_for_current_device(), a, trace, 0, 0, s);
}

void launch_shade(const FrameArgs &a, int segment, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    k_shade<<<sh.shade_blocks, kShadeThreads, 0, s>>>(a, segment, segment & 1);
}

void launch_trace(const FrameArgs &a, int segment, bool trace, cudaStream_t s)
{
    // rays of segment i were written by shade(i-1) into queue (i-1)&1 ^ 1 == i&1
    launch_trace_kernel<1>(shapes_for_current_device(), a, trace, segment, segment & 1, s);
}

void launch_progressive(const uint32_t *raw_rgba8, uint32_t *screen_rgba8, float4 *accum, const gdpt_progressive_params *params_dev,
                        int width, int height, int shard_part, int shard_parts, int shard_band, const PeerScreens &peers, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    k_progressive<<<sh.prog_blocks, 256, 0, s>>>(raw_rgba8, screen_rgba8, accum, params_dev, width, height, shard_part, shard_parts,
                                                 shard_band, peers);
}

__global__ void __launch_bounds__(128) k_small_copies(const SmallCopies sc)
{
#pragma unroll
    for (int k = 0; k < 3; k++)
        for (uint32_t i = threadIdx.x; i < sc.c[k].words; i += blockDim.x) sc.c[k].dst[i] = sc.c[k].src[i];
}
void launch_small_copies(const SmallCopies &sc, cudaStream_t s) { k_small_copies<<<1, 128, 0, s>>>(sc); }

void launch_temporal(uint32_t *screen_rgba8, const float *depth, const float *history, float *next,
                     const gdpt_temporal_params *params_dev, int width, int height, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    const int tiles = ((width + 31) / 32) * ((height + 7) / 8);
    const int blocks = tiles < sh.sms * 8 ? tiles : sh.sms * 8; // 8 resident blocks of 256 threads per SM
    k_temporal<<<blocks, 256, 0, s>>>(screen_rgba8, depth, history, next, params_dev);
}

static int fast_minb(const FrameArgs &a) { return (a.path_minb == 5 || a.path_minb == 6 || a.path_minb == 8) ? a.path_minb : 4; }

void launch_path_fast(const FrameArgs &a, bool record, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    if (record) { k_path_fast<true, 4><<<persistent_grid(sh, a, sh.fast_blocks[1][4]), kTraceThreads, 0, s>>>(a); return; }
    const int m = fast_minb(a);
    const int grid = persistent_grid(sh, a, sh.fast_blocks[0][m]);
    switch (m) {
    case 5: k_path_fast<false, 5><<<grid, kTraceThreads, 0, s>>>(a); break;
    case 6: k_path_fast<false, 6><<<grid, kTraceThreads, 0, s>>>(a); break;
    case 8: k_path_fast<false, 8><<<grid, kTraceThreads, 0, s>>>(a); break;
    default: k_path_fast<false, 4><<<grid, kTraceThreads, 0, s>>>(a); break;
    }
}

// Compile-time shape of k_path_pool, chosen by A/B on the B200 (C2 demo frame / C4 instanced at 1080p, ms):
//   parked leaves 1 | 2 | 3 | 4          0.859 | 0.892 | 0.908 | 0.918      20.3 | 20.9 | 21.4 | 21.8
//   slots per warp 40 | 64 | 96          0.961 | 0.878 | 0.896              21.5 | 20.5 | 20.9
//   blocks per SM 4 | 5 | 6 | 8          0.854 | 0.960 | 1.124 | 1.320      19.0 | 18.6 | 21.8 | 24.8   (registers 128 | 96 | 80 | 64)
void launch_path_pool(const FrameArgs &a, bool record, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    const bool wide = a.wide_bvh != 0 && a.sc.fast4_ok != 0; // four-wide tables (fast_bvh.h Collapse)
    if (record) {
        if (wide) k_path_pool<true, 4, kPoolParkDefault, kPoolSlotsDefault, true><<<persistent_grid(sh, a, sh.pool_blocks[1][1]), kTraceThreads, 0, s>>>(a);
        else k_path_pool<true, 4, kPoolParkDefault, kPoolSlotsDefault, false><<<persistent_grid(sh, a, sh.pool_blocks[1][0]), kTraceThreads, 0, s>>>(a);
        return;
    }
    if (wide && a.path_minb == 3) // A/B: 3 blocks per SM, no register cap (GDPT_PATH_MINB=3)
        k_path_pool<false, 3, kPoolParkDefault, kPoolSlotsDefault, true><<<persistent_grid(sh, a, sh.pool_blocks[0][2]), kTraceThreads, 0, s>>>(a);
    else if (wide) k_path_pool<false, 4, kPoolParkDefault, kPoolSlotsDefault, true><<<persistent_grid(sh, a, sh.pool_blocks[0][1]), kTraceThreads, 0, s>>>(a);
    else k_path_pool<false, 4, kPoolParkDefault, kPoolSlotsDefault, false><<<persistent_grid(sh, a, sh.pool_blocks[0][0]), kTraceThreads, 0, s>>>(a);
}

void launch_path_mux(const FrameArgs &a, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    const int k = (a.mux_k >= 1 && a.mux_k <= 4) ? a.mux_k : 2;
    const int grid = persistent_grid(sh, a, sh.mux_blocks[k]);
    switch (k) {
    case 1: k_path_mux<1, 8><<<grid, kMuxThreads, mux_smem_bytes(1), s>>>(a); break;
    case 3: k_path_mux<3, 5><<<grid, kMuxThreads, mux_smem_bytes(3), s>>>(a); break;
    case 4: k_path_mux<4, 4><<<grid, kMuxThreads, mux_smem_bytes(4), s>>>(a); break;
    default: k_path_mux<2, 8><<<grid, kMuxThreads, mux_smem_bytes(2), s>>>(a); break;
    }
}

size_t mux_path_record_quads()
{
    Shapes &sh = shapes_for_current_device();
    size_t most = 0;
    for (int k = 1; k <= 4; k++) {
        const size_t n = (size_t)sh.mux_blocks[k] * kMuxThreads * k * 5;
        if (n > most) most = n;
    }
    return most;
}

size_t path_kernel_warps(const FrameArgs &a)
{
    Shapes &sh = shapes_for_current_device();
    if (a.schedule == 3 && (a.path_minb == 1 || a.path_minb == 2 || a.path_minb == 5 || a.path_minb == 6 || a.path_minb == 8)) return (size_t)sh.path_list_blocks_minb[a.path_minb] * (kTraceThreads / 32);
    if (a.schedule == 5) return (size_t)sh.fast_blocks[0][fast_minb(a)] * (kTraceThreads / 32);
    if (a.schedule == 6) return (size_t)sh.pool_blocks[0][(a.wide_bvh != 0 && a.sc.fast4_ok != 0) ? 1 : 0] * (kTraceThreads / 32);
    if (a.schedule == 4) return (size_t)sh.mux_blocks[(a.mux_k >= 1 && a.mux_k <= 4) ? a.mux_k : 2] * (kMuxThreads / 32);
    int most = 0;
    for (int t = 0; t < 2; t++) {
        for (int c = 0; c < 2; c++) most = most > sh.path_blocks[t][c] ? most : sh.path_blocks[t][c];
        most = most > sh.path_list_blocks[t] ? most : sh.path_list_blocks[t];
    }
    return (size_t)most * (kTraceThreads / 32);
}

int k1_launch_count(int schedule, int max_depth, bool debug_steps)
{
    if (schedule == 2) return 1;
    if (schedule >= 3) return 2;
    if (debug_steps) return 1;
    return 1 + max_depth + (max_depth - 1); // primary + shade(0..D-1) + trace(1..D-1)
}

} // namespace gdpt
