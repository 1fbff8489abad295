// k_path_ref.cuh -- path kernel in the REFERENCE'S visiting order (part of pt_kernels.cu's translation unit).
// Schedule 2 (trace / DEBUG_STEPS / no culling: the observable is the reference's own work) and schedule 3
// (camera-ray classification + culled reference order; also the fallback when the closest-hit tables cannot be built).
#ifndef GDPT_K_PATH_REF_CUH
#define GDPT_K_PATH_REF_CUH
// (included inside namespace gdpt { namespace { ... } } of pt_kernels.cu)

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

#endif
