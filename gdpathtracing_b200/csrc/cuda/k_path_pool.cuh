// k_path_pool.cuh -- path kernel of schedule 6 (part of pt_kernels.cu's translation unit): closest-hit search +
// proof (pt_fast.cuh), paths pooled per warp in shared memory.
#ifndef GDPT_K_PATH_POOL_CUH
#define GDPT_K_PATH_POOL_CUH
// (included inside namespace gdpt { namespace { ... } } of pt_kernels.cu)

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
#ifndef GDPT_POOL_MINB
#define GDPT_POOL_MINB 4
#endif
constexpr int kPoolMinBlocks = GDPT_POOL_MINB; // resident blocks per SM the register allocation is capped for (launch bounds)
constexpr int kPoolDenseMinBlocks = 5; // the second build of the timed instantiation, for throughput-bound scenes (FrameArgs::pool_dense)
#ifndef GDPT_POOL_SLOTS
#define GDPT_POOL_SLOTS 80
#endif
constexpr int kPoolSlotsDefault = GDPT_POOL_SLOTS; // path slots per warp (template argument kPoolSlots)
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
__device__ __forceinline__ bool pool_can_cross(uint32_t cur, uint32_t inst)
{
    return cur != LINK_NONE && (cur & LINK_TLAS) != 0u && ((cur & LINK_LEAF) != 0u || inst != GDPT_NO_INSTANCE);
}

// Where an accepted triangle's u / v / triangle / instance go: into the ray's path slot (field-major pool), so a lane holds
// only t while it searches.
template <int kPoolSlots> struct HitInPool {
    uint32_t *slot0; // &pool[slot]
    __device__ __forceinline__ void accept(RayState &, float u, float v, uint32_t tri, uint32_t blas_front) const
    {
        slot0[PF_U * kPoolSlots] = __float_as_uint(u); slot0[PF_V * kPoolSlots] = __float_as_uint(v);
        slot0[PF_TRI * kPoolSlots] = tri; slot0[PF_BF * kPoolSlots] = blas_front;
    }
};

// Launch-time form of the scheduling knobs (launch_path_pool): the kernel reads them from the constant bank as they are.
//   refill_below = lanes without a walking ray before the pool is serviced (1..32)
//   pool_wait    = ... or this many lane-iterations spent waiting (0xFFFFFFFF = off)
//   pool_alive   = free slots a warp keeps unused (cap on alive paths), shade_at = finished rays that justify a partial batch
template <bool REC, int MINB, int kPoolSlots, bool WIDE, bool COUNT = false, bool PROF = false>
__global__ void __launch_bounds__(kTraceThreads, MINB) k_path_pool(const FrameArgs a)
{
    __shared__ uint32_t s_stack[kSmemStack * kTraceThreads];
    __shared__ uint32_t s_pool[kTraceThreads / 32][PF_COUNT * kPoolSlots];
    __shared__ uint8_t s_lists[kTraceThreads / 32][3 * kPoolSlots];
    __shared__ gdpt_camera s_cam;
    __shared__ SurvivorLists s_surv;
    uint32_t spill[GDPT_MAX_STACK - kSmemStack];
    SmemStack st;
    st.col = s_stack + threadIdx.x;
    st.spill = spill;
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(&s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    if (threadIdx.x == kTraceThreads - 1) s_surv.load(a);
    static_assert(kPoolSlots >= 32 && kPoolSlots <= 255 && kPoolSlots % 8 == 0, "pool: 32..248 slots");
    const unsigned lane = threadIdx.x & 31u;
    uint32_t *const pool = s_pool[threadIdx.x >> 5];
    uint8_t *const ready_list = s_lists[threadIdx.x >> 5];
    uint8_t *const done_list = ready_list + kPoolSlots;
    uint8_t *const free_list = done_list + kPoolSlots;
    for (unsigned i = lane; i < (unsigned)kPoolSlots; i += 32u) free_list[i] = (uint8_t)(kPoolSlots - 1 - i);
    __syncthreads();
    const gdpt_camera &cam = s_cam;
    const float far_depth = encode_depth(cam, cam.z_far); // what a primary ray that escapes stores (main.glsl:430-431)
#define PF(f, sl) pool[(f) * kPoolSlots + (sl)]
#define PFF(f, sl) __uint_as_float(pool[(f) * kPoolSlots + (sl)])

    const unsigned lanemask_lt = (1u << lane) - 1u;
    FrameCounters *cnt = a.counters;
    const SurvivorLists &lists = s_surv;
#define swap_at a.refill_below
#define swap_wait ((uint32_t)a.pool_wait)
#define min_free ((uint32_t)a.pool_alive)
#define shade_low a.shade_at
    uint32_t waited = 0;
    bool heavy_done = false;

    RayState r;
    r.cur = LINK_NONE; r.sp = 0; r.overflow = 0; r.t = 1e9f; r.inst = GDPT_NO_INSTANCE;
    uint32_t slot = 0, steps = 0;
    bool has = false;
    uint32_t ready_count = 0, done_count = 0, free_count = (uint32_t)kPoolSlots; // warp-uniform
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = (lists.total == 0u);
    uint32_t my_rays = 0, my_phits = 0, my_retraced = 0; // per lane and launch: far below 2^32
    unsigned long long own_nodes = 0, own_boxes = 0, own_tris = 0, own_insts = 0, own_proofs = 0; // COUNT only
    uint32_t my_overflow = 0;
    const bool prof = PROF && a.warp_prof != nullptr; // the per-warp schedule profile has its own instantiation
    const unsigned long long t_start = prof ? global_ns() : 0ull;
    uint32_t it_i = 0, it_l = 0, it_t = 0, it_f = 0, it_e = 0, n_started = 0;

    for (;;) {
        const bool can_i = has && pool_can_node(r.cur, r.inst);
        const bool fin = has && r.cur == LINK_NONE;
        const uint32_t census = __reduce_add_sync(kFull, (can_i ? 1u : 0u) | (fin ? 1u << 8 : 0u) | (has ? 0u : 1u << 16));
        const int n_i = (int)(census & 63u), n_fin = (int)((census >> 8) & 63u), n_idle = (int)(census >> 16);
        const int n_walk = 32 - n_idle - n_fin;
        const bool can_refill = !exhausted && free_count > min_free; // alive paths (slots in use) stay below the cap

        waited += (uint32_t)n_fin;
        if (done_count >= 32u || n_walk == 0 || n_fin >= swap_at || (n_fin > 0 && waited >= swap_wait) ||
            (n_idle >= swap_at && (ready_count > 0u || can_refill || done_count >= (uint32_t)shade_low))) {
            waited = 0;
            // ---------------- pool service ----------------
            // 1. retire: finished searches go to their slots, the slots to the shading queue
            if (n_fin > 0) {
                // a.miss_now: a ray that provably escaped (a miss of the complete search, no flag) is finished by its own lane
                // right here -- sky colour, pixel, cost hint, slot freed -- so the shading batches hold hits only (the one
                // ordering of the shading work by the branch it takes that this path offers: hit / sky)
                const bool escaped = fin && a.miss_now != 0 && !(r.t < 1e9f) && r.overflow == 0u;
                const unsigned m = __ballot_sync(kFull, fin && !escaped), m_esc = __ballot_sync(kFull, escaped);
                if (escaped) {
                    const f3 wd = mk3(PFF(PF_WDX, slot), PFF(PF_WDY, slot), PFF(PF_WDZ, slot));
                    const f3 radiance = mk3(PFF(PF_RAR, slot), PFF(PF_RAG, slot), PFF(PF_RAB, slot)) +
                                        mk3(PFF(PF_THR, slot), PFF(PF_THG, slot), PFF(PF_THB, slot)) * sample_sky(wd);
                    const uint32_t pixel = PF(PF_PIXEL, slot);
                    const int segment = (int)PF(PF_SEGMENT, slot);
                    if (segment == 0) a.out_depth[pixel] = far_depth;
                    if (REC) write_hit_record(a, segment, pixel, 1e9f, 0.0f, 0.0f, 0u, 0u);
                    a.out_rgba8[pixel] = pack_rgba8(radiance);
                    const uint32_t cost_now = PF(PF_STEPS, slot) + steps;
                    if (a.cost_ema) {
                        const uint32_t old = a.cost[pixel];
                        a.cost[pixel] = old ? (old * 3u + cost_now + 2u) >> 2 : cost_now;
                    } else {
                        a.cost[pixel] = cost_now;
                    }
                    my_rays++;
                    if (COUNT) own_proofs++;
                    free_list[free_count + (uint32_t)__popc(m_esc & lanemask_lt)] = (uint8_t)slot;
                    has = false;
                } else if (fin) { // u / v / triangle / instance are in the slot already (HitInPool)
                    PF(PF_T, slot) = __float_as_uint(r.t); PF(PF_FLAGS, slot) = r.overflow;
                    PF(PF_STEPS, slot) += steps;
                    done_list[done_count + (uint32_t)__popc(m & lanemask_lt)] = (uint8_t)slot;
                    has = false;
                }
                done_count += (uint32_t)__popc(m);
                free_count += (uint32_t)__popc(m_esc);
                __syncwarp();
            }
            const uint32_t n_out = (uint32_t)(n_fin + n_idle); // lanes that hold no walking ray now
            const bool shade = done_count >= 32u ||
                               (done_count > 0u && n_out > ready_count && !can_refill && (done_count >= (uint32_t)shade_low || n_walk == 0));
            if (shade) {
                // 2. S: prove and shade up to 32 finished rays, oldest first
                if (PROF) it_f++;
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
                    if (COUNT) own_proofs++;
                    if (!fast_verdict_ool(&a.sc, wo, wd, ht, htri, hbf, hflags)) {
                        ExactHit eh; // rare: exact reference-order traversal of this ray
                        exact_retrace(&a.sc, wo, wd, (hflags & RAY_FAR) == 0u, &eh);
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
                        if (segment == 0) a.out_depth[pixel] = far_depth;
                    } else {
                        BounceResult br;
                        u2 sd; sd.x = PF(PF_SEEDX, sl); sd.y = PF(PF_SEEDY, sl);
                        shade_and_bounce_ool(&a.sc, wo, wd, ht, hu, hv, htri, hbf, radiance, throughput, &sd, &br);
                        radiance = br.radiance;
                        if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
                        alive = br.alive && segment < a.max_depth - 1;
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
                if (PROF) it_e++;
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
                    if (base >= lists.total) { exhausted = true; continue; }
                    chunk_next = base;
                    chunk_end = min(base + len, lists.total);
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
                free_count -= g; ready_count += g; chunk_next += g; if (PROF) n_started += g;
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
                    r.cur = fast_start_link(r, WIDE ? a.sc.fast4_root : r.cur); // the four-wide tables have their own root link
                    steps = 0;
                    has = true;
                }
                ready_count -= min((uint32_t)__popc(idle), ready_count);
                __syncwarp();
            }
            if (exhausted && ready_count == 0u && done_count == 0u && __ballot_sync(kFull, has) == 0u) break;
            continue;
        }
        // ---------------- I / L / T ----------------
        // Every phase that has a lane runs, in the order a ray meets them (nodes, then the leaf it reached, then the
        // crossing), so a ray advances in every iteration -- a frame is one wave of paths whose length is the latency of its
        // longest path, and a lane that waits for its phase to win a vote lengthens exactly that.  (The majority-phase loop
        // this replaced, and thresholds of 2 / 4 / 8 lanes for the other phases, were slower on every config: DESIGN.md 4.)
        if (n_i > 0) {
            if (PROF) it_i++;
            bool go = can_i;
            const int need = (n_i + 1) >> 1;
#pragma unroll 1
            for (int b = 0; b < a.burst; b++) { // node burst: no census while at least half of the starters stay on internal nodes
                if (go) {
                    if (WIDE) fast_step_node4(a.sc, r, st); else fast_step_node(a.sc, r, st);
                    steps++;
                    if (COUNT) { own_nodes++; own_boxes += WIDE ? 4u : 2u; }
                    go = pool_can_node(r.cur, r.inst);
                }
                if (__popc(__ballot_sync(kFull, go)) < need) break;
            }
        }
        const bool now_l = has && fast_link_is_leaf(r.cur);
        if (__ballot_sync(kFull, now_l)) {
            if (PROF) it_l++;
            if (now_l) {
                const uint32_t leaf = r.cur;
                r.cur = fast_pop(r, st);
                HitInPool<kPoolSlots> sink;
                sink.slot0 = pool + slot;
                fast_leaf_tests(a.sc, r, leaf, sink);
                steps++;
                if (COUNT) own_tris += ((leaf >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u;
            }
        }
        const bool now_t = has && pool_can_cross(r.cur, r.inst);
        if (__ballot_sync(kFull, now_t)) {
            if (PROF) it_t++;
            if (now_t) { // back to world space (main.glsl:316-327) and/or into the instance the link names
                r.wo = mk3(PFF(PF_WOX, slot), PFF(PF_WOY, slot), PFF(PF_WOZ, slot)); // the world ray lives in the slot, not in the lane
                r.wd = mk3(PFF(PF_WDX, slot), PFF(PF_WDY, slot), PFF(PF_WDZ, slot));
                if (r.inst != GDPT_NO_INSTANCE) { r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd); r.inst = GDPT_NO_INSTANCE; }
                if (r.cur & LINK_LEAF) { fast_enter_instance<WIDE>(a.sc, r, st); if (COUNT) own_insts++; }
                steps++;
            }
        }
    }
#undef PF
#undef PFF
#undef swap_at
#undef swap_wait
#undef min_free
#undef shade_low
    if (prof && lane == 0) {
        unsigned long long *w = a.warp_prof + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        w[0] = t_start; w[1] = global_ns(); w[2] = it_i; w[3] = it_l; w[4] = it_t; w[5] = it_f; w[6] = it_e; w[7] = n_started;
    }
    my_rays = __reduce_add_sync(kFull, my_rays); my_phits = __reduce_add_sync(kFull, my_phits); // per warp: still far below 2^32
    my_retraced = __reduce_add_sync(kFull, my_retraced);
    if (lane == 0) {
        atomicAdd(&cnt->rays, (unsigned long long)my_rays); atomicAdd(&cnt->primary_hits, (unsigned long long)my_phits);
        if (my_retraced) atomicAdd(&cnt->retraced, (unsigned long long)my_retraced);
    }
    if (COUNT) { // own work of the search (roofline numerator of bench.py)
        for (int off = 16; off > 0; off >>= 1) {
            own_nodes += __shfl_down_sync(kFull, own_nodes, off); own_boxes += __shfl_down_sync(kFull, own_boxes, off);
            own_tris += __shfl_down_sync(kFull, own_tris, off); own_insts += __shfl_down_sync(kFull, own_insts, off);
            own_proofs += __shfl_down_sync(kFull, own_proofs, off);
        }
        if (lane == 0) {
            atomicAdd(&cnt->own_node_steps, own_nodes); atomicAdd(&cnt->own_box_tests, own_boxes); atomicAdd(&cnt->own_tri_tests, own_tris);
            atomicAdd(&cnt->own_inst_entries, own_insts); atomicAdd(&cnt->own_proofs, own_proofs);
        }
    }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

#endif
