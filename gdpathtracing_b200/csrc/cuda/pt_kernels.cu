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
#include "pt_shade.cuh"
#include "pt_trace.cuh"

#include <cstdio>

namespace gdpt {

namespace {

constexpr int kTraceThreads = 128;
constexpr int kSmemStack = 16;
constexpr int kShadeThreads = 128;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kChunkPrimary = 32; // one 8x4-pixel tile: small chunks keep the expensive tiles spread over many warps
constexpr uint32_t kChunkBounce = 32;

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
template <bool TRACE, bool CULL>
__global__ void __launch_bounds__(kTraceThreads) k_path(const FrameArgs a)
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
    const uint32_t total = a.n_work;
    const int refill_below = max(a.refill_below, 1);
    const int shade_at = min(max(a.shade_at, 1), 32);
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

    for (;;) {
        const bool in_l = has && (tri_next < tri_end || link_is_blas_leaf(r.cur));
        const bool in_i = has && !in_l && link_is_blas_internal(r.cur);
        const bool in_t = has && !in_l && !in_i && r.cur != LINK_NONE;
        const bool fin = has && !in_l && !in_i && !in_t;
        const int n_l = __popc(__ballot_sync(kFull, in_l)), n_i = __popc(__ballot_sync(kFull, in_i)),
                  n_t = __popc(__ballot_sync(kFull, in_t)), n_fin = __popc(__ballot_sync(kFull, fin));
        const unsigned idle = __ballot_sync(kFull, !has);
        const int n_walk = n_l + n_i + n_t, n_idle = __popc(idle);

        if (n_fin > 0 && (n_fin >= shade_at || n_walk == 0)) {
            // ---------------- S: finish a segment ----------------
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
                    const BounceResult br = shade_and_bounce(a.sc, r.wo, r.wd, r.t, r.u, r.v, r.tri, r.blas_front, radiance,
                                                             throughput, seed);
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
                    has = false;
                }
            }
            continue;
        }
        if (!exhausted && n_idle > 0 && (32 - n_idle < refill_below || n_walk + n_fin == 0)) {
            // ---------------- R: new camera rays for idle lanes ----------------
            if (chunk_next == chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt->cursor[0], kChunkPrimary);
                base = __shfl_sync(kFull, base, 0);
                if (base >= total) { exhausted = true; continue; }
                chunk_next = base;
                chunk_end = min(base + kChunkPrimary, total);
            }
            const uint32_t avail = chunk_end - chunk_next;
            const uint32_t rank = __popc(idle & lanemask_lt);
            if (!has && rank < avail) {
                int px, py;
                if (work_to_pixel(a, chunk_next + rank, &px, &py)) {
                    f3 o, d;
                    seed = generate_primary_ray(cam, a.width, a.height, px, py, &o, &d);
                    pixel = (uint32_t)py * (uint32_t)a.width + (uint32_t)px;
                    throughput = mk3(1.0f, 1.0f, 1.0f); radiance = mk3(0.0f, 0.0f, 0.0f);
                    segment = 0;
                    ray_begin(r, a.sc, o, d);
                    has = true;
                    if (TRACE) counters_init(tc, a.visits ? a.visits + (size_t)pixel * a.visits_per_ray : nullptr, a.visits_per_ray);
                }
            }
            chunk_next += min((uint32_t)n_idle, avail);
            continue;
        }
        if (n_walk == 0) {
            if (exhausted && n_fin == 0) break;
            continue;
        }
        // ---------------- L / I / T: one traversal step of the most popular phase ----------------
        if (n_l >= n_i && n_l >= n_t) {
            if (in_l) step_blas_leaf_one<TRACE>(a.sc, r, st, &tc, tri_next, tri_end);
        } else if (n_i >= n_t) {
            if (in_i) step_blas_internal<TRACE, CULL>(a.sc, r, st, &tc);
        } else {
            if (in_t) step_tlas<TRACE, CULL>(a.sc, r, st, &tc);
        }
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
__global__ void __launch_bounds__(256) k_progressive(uint32_t *__restrict__ screen, float4 *__restrict__ accum,
                                                     const gdpt_progressive_params *__restrict__ params, int width,
                                                     int height, int shard_part, int shard_parts, int shard_band)
{
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
        const uint4 s4 = reinterpret_cast<const uint4 *>(screen)[q];
        const uint32_t in[4] = { s4.x, s4.y, s4.z, s4.w };
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            f3 rad = mk3((float)(in[k] & 0xffu) / 255.0f, (float)((in[k] >> 8) & 0xffu) / 255.0f,
                         (float)((in[k] >> 16) & 0xffu) / 255.0f);
            if (frame_count > 1u) {
                const float4 acc = accum[p + k];
                rad = rad + mk3(acc.x, acc.y, acc.z);
            }
            accum[p + k] = make_float4(rad.x, rad.y, rad.z, 1.0f);
            const f3 avg = (rad / fc) * 1.0f;
            out[k] = pack_rgba8(mk3(aces_channel(avg.x), aces_channel(avg.y), aces_channel(avg.z)));
        }
        reinterpret_cast<uint4 *>(screen)[q] = make_uint4(out[0], out[1], out[2], out[3]);
    }
    // tail pixels when W*H is not a multiple of 4
    if (blockIdx.x == 0 && threadIdx.x < (((size_t)width * height) & 3)) {
        const size_t p = (n_quads << 2) + threadIdx.x;
        const int y = (int)(p / (size_t)width);
        if (shard_parts <= 1 || (y / shard_band) % shard_parts == shard_part) {
            const uint32_t in = screen[p];
            f3 rad = mk3((float)(in & 0xffu) / 255.0f, (float)((in >> 8) & 0xffu) / 255.0f, (float)((in >> 16) & 0xffu) / 255.0f);
            if (frame_count > 1u) { const float4 acc = accum[p]; rad = rad + mk3(acc.x, acc.y, acc.z); }
            accum[p] = make_float4(rad.x, rad.y, rad.z, 1.0f);
            const f3 avg = (rad / fc) * 1.0f;
            screen[p] = pack_rgba8(mk3(aces_channel(avg.x), aces_channel(avg.y), aces_channel(avg.z)));
        }
    }
}

struct Shapes {
    bool ready = false;
    int sms = 148;
    int trace_blocks[2][2][2] = {}; // [TRACE][MODE][CULL]
    int path_blocks[2][2] = {};     // [TRACE][CULL]
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
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path<false, false>, kTraceThreads, 0); s.path_blocks[0][0] = s.sms * (per_sm > 0 ? per_sm : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path<false, true>, kTraceThreads, 0); s.path_blocks[0][1] = s.sms * (per_sm > 0 ? per_sm : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path<true, false>, kTraceThreads, 0); s.path_blocks[1][0] = s.sms * (per_sm > 0 ? per_sm : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path<true, true>, kTraceThreads, 0); s.path_blocks[1][1] = s.sms * (per_sm > 0 ? per_sm : 1);
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
        if (cull) k_path<true, true><<<blocks, kTraceThreads, 0, s>>>(a);
        else k_path<true, false><<<blocks, kTraceThreads, 0, s>>>(a);
    } else {
        if (cull) k_path<false, true><<<blocks, kTraceThreads, 0, s>>>(a);
        else k_path<false, false><<<blocks, kTraceThreads, 0, s>>>(a);
    }
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

void launch_progressive(uint32_t *screen_rgba8, float4 *accum, const gdpt_progressive_params *params_dev, int width,
                        int height, int shard_part, int shard_parts, int shard_band, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    k_progressive<<<sh.prog_blocks, 256, 0, s>>>(screen_rgba8, accum, params_dev, width, height, shard_part, shard_parts,
                                                 shard_band);
}

int k1_launch_count(int max_depth, bool debug_steps)
{
    if (debug_steps) return 1;
    return 1 + max_depth + (max_depth - 1); // primary + shade(0..D-1) + trace(1..D-1)
}

} // namespace gdpt
