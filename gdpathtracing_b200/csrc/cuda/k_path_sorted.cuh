// k_path_sorted.cuh -- path kernel of schedule 7 (part of pt_kernels.cu's translation unit): the closest-hit search +
// proof of pt_fast.cuh with the RAYS taken out of the lanes.
#ifndef GDPT_K_PATH_SORTED_CUH
#define GDPT_K_PATH_SORTED_CUH
// (included inside namespace gdpt { namespace { ... } } of pt_kernels.cu)

// k_path_pool keeps a ray's traversal state in the registers of one lane, so a warp issues every node-step
// instruction for the lanes that happen to be on an internal node (ncu: 8 of 32) and every triangle-test
// instruction for the lanes that happen to be on a leaf.  Here a BLOCK owns 32 x kRows rays whose whole state -- world
// ray, local ray, hit, link, stack, path -- lives in shared memory, field-major (word f of slot s at f*kSlots + s).
// Slot s belongs to COLUMN s mod 32 = its shared-memory bank, and column c is served by lane c of whichever warp
// comes by, so no access to ray state ever has a bank conflict and no ray is ever copied.  Every slot carries a
// 4-bit phase tag (the kRows tags of a column share one word):
//     I  next link is an internal node of the space the ray is in      L  next link is a leaf (<= 3 triangles)
//     T  next link crosses spaces (leave an instance / enter one)      F  search over: prove, shade, bounce
//     FREE / BUSY (claimed by a warp)
// Each iteration a warp counts the block's tags (one word per lane, two REDUX), picks the phase with the most rays,
// every lane claims one ray of that phase in its column (compare-and-swap on the column's tag word) and the warp runs
// ONE phase body: the node step on ~30 lanes instead of 8.  With four warps sharing 256 rays a phase almost always has
// a ray in every column -- a warp-private pool of the same size (64 rays) left 9 lanes per node step, because rays
// waiting for a shading batch or a refill held most of its slots.  A lane keeps nothing between iterations, so the
// register budget is that of the largest phase body, not of their union.  Shading runs on 32 finished rays at a time;
// free slots are refilled with camera rays 32 at a time, longest paths first (cost classes of k_primary_cull).
// Every number a ray or a path produces comes from the device functions k_path_pool uses (fast_step_node4,
// fast_leaf_tests, fast_enter_instance, the verdict, the exact re-trace, shade_and_bounce): which lane runs them,
// and when, is all that differs.
enum SortField {
    SF_WOX, SF_WOY, SF_WOZ, SF_WDX, SF_WDY, SF_WDZ,              // world ray
    SF_OX, SF_OY, SF_OZ, SF_DX, SF_DY, SF_DZ, SF_RX, SF_RY, SF_RZ, // ray in the space being searched, reciprocal direction
    SF_T, SF_U, SF_V, SF_TRI, SF_BF, SF_FLAGS,                   // best hit so far
    SF_CUR, SF_SP, SF_INST,                                      // next link, stack fill, instance
    SF_THR, SF_THG, SF_THB, SF_RAR, SF_RAG, SF_RAB,              // path: throughput, radiance
    SF_SEEDX, SF_SEEDY, SF_PIXEL, SF_SEGMENT, SF_STEPS,
    SF_COUNT
};
enum SortTag : uint32_t { TAG_FREE = 0, TAG_I = 1, TAG_L = 2, TAG_T = 3, TAG_F = 4, TAG_BUSY = 15 };
constexpr int kSortStack = 14;   // stack entries per slot in shared memory; deeper ones go to the global spill area
constexpr int kSortRowsDefault = 8; // rays per column: 32 x 8 = 256 rays per block of four warps

template <int kRows> constexpr size_t sorted_smem_bytes()
{
    return (size_t)(SF_COUNT + kSortStack) * (32u * kRows) * 4u + 32u * 4u + 8u * 4u + sizeof(gdpt_camera) + 16u;
}

// nibbles of `w` that equal `tag`: bit 4*row set for every match
__device__ __forceinline__ uint32_t sorted_match(uint32_t w, uint32_t tag, uint32_t rows_mask)
{
    const uint32_t x = w ^ (tag * 0x11111111u);
    return ~(x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u & rows_mask;
}
// Claim one slot of `tag` in this lane's column: its nibble becomes TAG_BUSY.  Returns the row, or -1.
__device__ __forceinline__ int sorted_claim(uint32_t *col_word, uint32_t tag, uint32_t first_row, uint32_t rows_mask)
{
    uint32_t w = *reinterpret_cast<volatile uint32_t *>(col_word);
    for (;;) {
        const uint32_t z = sorted_match(w, tag, rows_mask);
        if (z == 0u) return -1;
        const uint32_t from = z & (0xFFFFFFFFu << (4u * first_row)); // rotate which row goes first
        const int bit = __ffs((int)(from ? from : z)) - 1;
        const uint32_t seen = atomicCAS(col_word, w, w | (0xFu << bit));
        if (seen == w) return bit >> 2;
        w = seen;
    }
}
// Hand a claimed slot back with its new tag (other warps change other nibbles of the word at the same time).
__device__ __forceinline__ void sorted_release(uint32_t *col_word, int row, uint32_t tag)
{
    atomicAnd(col_word, ~(0xFu << (4 * row)) | (tag << (4 * row)));
}

template <int kSlots> struct SlotStack {
    uint32_t *col;   // entry 0 of this slot's shared-memory column (stride kSlots words)
    uint32_t *spill; // this slot's part of the global spill area
    __device__ __forceinline__ void store(uint32_t i, uint32_t v)
    {
        if (i < (uint32_t)kSortStack) col[i * kSlots] = v; else spill[i - kSortStack] = v;
    }
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return (i < (uint32_t)kSortStack) ? col[i * kSlots] : spill[i - kSortStack]; }
};

__device__ __forceinline__ uint32_t sorted_tag_of(uint32_t cur, uint32_t inst)
{
    if (cur == LINK_NONE) return TAG_F;
    if ((cur & (LINK_TLAS | LINK_LEAF)) == LINK_LEAF) return TAG_L;
    if ((cur & LINK_LEAF) == 0u && ((cur & LINK_TLAS) == 0u || inst == GDPT_NO_INSTANCE)) return TAG_I;
    return TAG_T;
}

// Own work of the search (roofline numerator of bench.py): what THIS kernel executed, not what the reference would.
struct OwnWork { unsigned long long node_steps, box_tests, tri_tests, inst_entries, proofs; };

template <bool REC, bool COUNT, int kRows, bool SORT4>
__global__ void __launch_bounds__(kTraceThreads, 4) k_path_sorted(const FrameArgs a)
{
    static_assert(kRows >= 2 && kRows <= 8, "2..8 rays per column (their tags share one word)");
    constexpr int kSlots = 32 * kRows;
    constexpr uint32_t kRowsMask = kRows == 8 ? 0xFFFFFFFFu : ((1u << (4 * kRows)) - 1u);
    extern __shared__ __align__(16) uint8_t s_raw[];
    const unsigned lane = threadIdx.x & 31u;
    uint32_t *const pool = reinterpret_cast<uint32_t *>(s_raw);
    uint32_t *const stk = pool + SF_COUNT * kSlots;
    uint32_t *const col_tags = stk + kSortStack * kSlots;      // one word per column: kRows 4-bit tags
    uint32_t *const counts = col_tags + 32;                    // rays per tag (index = tag), kept by claim / release: the census
    gdpt_camera *const s_cam = reinterpret_cast<gdpt_camera *>(counts + 8);
    if (threadIdx.x < sizeof(gdpt_camera) / 4u)
        reinterpret_cast<uint32_t *>(s_cam)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.camera)[threadIdx.x];
    if (threadIdx.x < 32u) col_tags[threadIdx.x] = 0u; // every slot TAG_FREE
    if (threadIdx.x < 8u) counts[threadIdx.x] = threadIdx.x == TAG_FREE ? (uint32_t)kSlots : 0u;
    __syncthreads();
    const gdpt_camera &cam = *s_cam;
    uint32_t *const my_col = col_tags + lane;
#define SF(f, sl) pool[(f) * kSlots + (sl)]
#define SFF(f, sl) __uint_as_float(pool[(f) * kSlots + (sl)])
#define SFW(f, sl, x) pool[(f) * kSlots + (sl)] = __float_as_uint(x)

    FrameCounters *cnt = a.counters;
    SurvivorLists lists;
    lists.load(a);
    const uint32_t total = lists.total;
    uint32_t *const spill_base = a.sorted_spill + (size_t)blockIdx.x * kSlots * a.sorted_spill_depth;
    const int burst = max(a.burst, 1);
    const int shade_low = min(max(a.shade_at, 1), 32);
    const int refill_at = min(max(a.refill_below, 1), 32);
    const int affinity = a.lead_min; // rays of a warp's own phase that keep it there (0 = no affinity)
    const uint32_t warp_id = threadIdx.x >> 5;
    const int last_segment = a.max_depth - 1;
    bool heavy_done = false, exhausted = (total == 0u);
    uint32_t chunk_next = 0, chunk_end = 0, iter = warp_id;
    unsigned long long my_rays = 0, my_phits = 0, my_retraced = 0;
    uint32_t my_overflow = 0;
    OwnWork own = { 0, 0, 0, 0, 0 };
    const bool prof = a.warp_prof != nullptr;
    const unsigned long long t_start = prof ? global_ns() : 0ull;
    uint32_t it_i = 0, it_l = 0, it_t = 0, it_f = 0, it_e = 0, n_started = 0;
    uint32_t ln_i = 0, ln_l = 0, ln_t = 0, ln_f = 0, ln_e = 0, it_idle = 0; // lanes that held a ray, per phase; iterations without work

    for (;; iter++) {
        // ---------------- census: rays of the block per tag (claimed slots count for nobody) ----------------
        // The counts trail the tag words by a few instructions: a hint for choosing the phase, the claim below decides.
        const volatile uint32_t *const vc = counts;
        const int n_free = (int)vc[TAG_FREE], n_i = (int)vc[TAG_I], n_l = (int)vc[TAG_L], n_t = (int)vc[TAG_T], n_f = (int)vc[TAG_F];
        const int most = max(n_i, max(n_l, n_t));
        const uint32_t walk = (n_i >= n_l && n_i >= n_t) ? TAG_I : (n_l >= n_t ? TAG_L : TAG_T);
        // Which phase this warp runs.  Warps have an affinity -- 0 and 1 the node step, 2 leaves and space crossings, 3
        // shading and refill -- so the warps that share an SM sub-partition (warp w of every resident block) mostly run
        // the same few hundred instructions and its instruction cache keeps them; any warp takes any phase when its own
        // has fewer than `affinity` rays waiting or another phase piles up.
        uint32_t phase = TAG_BUSY; // TAG_I / TAG_L / TAG_T / TAG_F, TAG_FREE = refill, TAG_BUSY = undecided
        if (affinity > 0) {
            if (warp_id == 3u) {
                if (n_f >= shade_low) phase = TAG_F;
                else if (!exhausted && n_free >= refill_at) phase = TAG_FREE;
            } else if (warp_id == 2u) {
                if (n_l >= affinity && n_l >= n_t) phase = TAG_L;
                else if (n_t >= affinity) phase = TAG_T;
                else if (n_l >= affinity) phase = TAG_L;
            } else if (n_i >= affinity) phase = TAG_I;
        }
        if (phase == TAG_BUSY) {
            if (n_f >= (affinity > 0 ? 64 : 32)) phase = TAG_F;
            else if (!exhausted && n_free >= (affinity > 0 ? 64 : 32)) phase = TAG_FREE;
            else if (most >= 24) phase = walk;
            else if (n_f >= shade_low && affinity == 0) phase = TAG_F;
            else if (!exhausted && n_free >= refill_at && affinity == 0) phase = TAG_FREE;
            else if (most > 0) phase = walk;
            else if (n_f > 0) phase = TAG_F;
            else if (!exhausted && n_free > 0) phase = TAG_FREE;
            else if (exhausted && n_free == kSlots) break; // nothing left anywhere in the block
            else { it_idle++; __nanosleep(200); continue; } // the other warps hold what is left
        }

        // ---------------- every lane claims one slot of that phase in its own column ----------------
        int row = sorted_claim(my_col, phase, iter % (uint32_t)kRows, kRowsMask);
        __threadfence_block(); // the slot's state was written by whoever released it
        unsigned claimed = __ballot_sync(kFull, row >= 0);
        if (claimed == 0u) { it_idle++; continue; }
        if (prof) {
            const uint32_t n = (uint32_t)__popc(claimed);
            if (phase == TAG_I) ln_i += n; else if (phase == TAG_L) ln_l += n; else if (phase == TAG_T) ln_t += n;
            else if (phase == TAG_F) ln_f += n; else ln_e += n;
        }
        if (lane == 0u) atomicSub(&counts[phase], (uint32_t)__popc(claimed));
        const unsigned lanemask_lt = (1u << lane) - 1u;

        if (phase == TAG_FREE) {
            // ---------------- R: camera rays into free slots ----------------
            it_e++;
            if (chunk_next == chunk_end) {
                uint32_t cb = 0, len = kChunkPrimary;
                if (lane == 0) {
                    if (!heavy_done) {
                        cb = atomicAdd(&cnt->cursor[0], kChunkHeavy);
                        len = kChunkHeavy;
                        if (cb >= lists.heavy_total) cb = 0xFFFFFFFFu;
                        else if (cb + len > lists.heavy_total) len = lists.heavy_total - cb;
                    }
                    if (heavy_done || cb == 0xFFFFFFFFu) {
                        cb = lists.heavy_total + atomicAdd(&cnt->cursor[1], kChunkPrimary);
                        len = kChunkPrimary | 0x80000000u;
                    }
                }
                cb = __shfl_sync(kFull, cb, 0);
                len = __shfl_sync(kFull, len, 0);
                if (len & 0x80000000u) { heavy_done = true; len &= 0x7FFFFFFFu; }
                if (cb >= total) { // the survivor lists are handed out: give the claimed slots back
                    exhausted = true;
                    if (row >= 0) sorted_release(my_col, row, TAG_FREE);
                    if (lane == 0u) atomicAdd(&counts[TAG_FREE], (uint32_t)__popc(claimed));
                    continue;
                }
                chunk_next = cb;
                chunk_end = min(cb + len, total);
            }
            const uint32_t rank = (uint32_t)__popc(claimed & lanemask_lt);
            const uint32_t n_act = min((uint32_t)__popc(claimed), chunk_end - chunk_next);
            if (row >= 0 && rank >= n_act) { sorted_release(my_col, row, TAG_FREE); row = -1; } // more free slots than pixels in this chunk
            if (lane == 0u && (uint32_t)__popc(claimed) > n_act) atomicAdd(&counts[TAG_FREE], (uint32_t)__popc(claimed) - n_act);
            if (row >= 0) {
                const uint32_t sl = (uint32_t)row * 32u + lane;
                const uint32_t p = lists.pixel(a, chunk_next + rank);
                const int py = (int)(p / (uint32_t)a.width), px = (int)(p - (uint32_t)py * (uint32_t)a.width);
                PrimaryRay pr;
                generate_primary_ray_ool(&cam, a.width, a.height, px, py, &pr);
                SFW(SF_WOX, sl, pr.o.x); SFW(SF_WOY, sl, pr.o.y); SFW(SF_WOZ, sl, pr.o.z);
                SFW(SF_WDX, sl, pr.d.x); SFW(SF_WDY, sl, pr.d.y); SFW(SF_WDZ, sl, pr.d.z);
                SFW(SF_OX, sl, pr.o.x); SFW(SF_OY, sl, pr.o.y); SFW(SF_OZ, sl, pr.o.z);
                SFW(SF_DX, sl, pr.d.x); SFW(SF_DY, sl, pr.d.y); SFW(SF_DZ, sl, pr.d.z);
                const f3 rd = fast_rcp3(pr.d);
                SFW(SF_RX, sl, rd.x); SFW(SF_RY, sl, rd.y); SFW(SF_RZ, sl, rd.z);
                SFW(SF_T, sl, 1e9f); SF(SF_U, sl) = 0u; SF(SF_V, sl) = 0u; SF(SF_TRI, sl) = 0u; SF(SF_BF, sl) = 0u;
                SF(SF_FLAGS, sl) = fast_far_origin(pr.o, a.sc.fast_world_reach) ? RAY_FAR : 0u;
                SF(SF_CUR, sl) = a.sc.fast4_root; SF(SF_SP, sl) = 0u; SF(SF_INST, sl) = GDPT_NO_INSTANCE;
                SFW(SF_THR, sl, 1.0f); SFW(SF_THG, sl, 1.0f); SFW(SF_THB, sl, 1.0f);
                SF(SF_RAR, sl) = 0u; SF(SF_RAG, sl) = 0u; SF(SF_RAB, sl) = 0u;
                SF(SF_SEEDX, sl) = pr.seed.x; SF(SF_SEEDY, sl) = pr.seed.y;
                SF(SF_PIXEL, sl) = p; SF(SF_SEGMENT, sl) = 0u; SF(SF_STEPS, sl) = 0u;
                __threadfence_block();
                sorted_release(my_col, row, sorted_tag_of(a.sc.fast4_root, GDPT_NO_INSTANCE));
            }
            if (lane == 0u) atomicAdd(&counts[sorted_tag_of(a.sc.fast4_root, GDPT_NO_INSTANCE)], n_act);
            chunk_next += n_act; n_started += n_act;
            continue;
        }
        const bool act = row >= 0;
        const uint32_t sl = act ? (uint32_t)row * 32u + lane : 0u;
        uint32_t new_tag = TAG_FREE;

        if (phase == TAG_I) {
            // ---------------- I: node steps ----------------
            it_i++;
            if (act) {
                SlotStack<kSlots> st;
                st.col = stk + sl;
                st.spill = spill_base + (size_t)sl * a.sorted_spill_depth;
                RayState r;
                r.o = mk3(SFF(SF_OX, sl), SFF(SF_OY, sl), SFF(SF_OZ, sl));
                r.rd = mk3(SFF(SF_RX, sl), SFF(SF_RY, sl), SFF(SF_RZ, sl));
                r.t = SFF(SF_T, sl);
                r.cur = SF(SF_CUR, sl); r.sp = SF(SF_SP, sl);
                uint32_t inst = SF(SF_INST, sl);
                uint32_t tag = TAG_I, steps = 0;
#pragma unroll 1
                for (int b = 0; b < burst && tag == TAG_I; b++) {
                    fast_step_node4<SORT4>(a.sc, r, st);
                    steps++;
                    if (inst != GDPT_NO_INSTANCE && r.cur != LINK_NONE && (r.cur & LINK_TLAS)) {
                        // the instance is done: back to world space here (main.glsl:322-327), not in an iteration of its own
                        const f3 wo = mk3(SFF(SF_WOX, sl), SFF(SF_WOY, sl), SFF(SF_WOZ, sl));
                        const f3 wd = mk3(SFF(SF_WDX, sl), SFF(SF_WDY, sl), SFF(SF_WDZ, sl));
                        r.o = wo; r.rd = fast_rcp3(wd);
                        SFW(SF_OX, sl, wo.x); SFW(SF_OY, sl, wo.y); SFW(SF_OZ, sl, wo.z);
                        SFW(SF_DX, sl, wd.x); SFW(SF_DY, sl, wd.y); SFW(SF_DZ, sl, wd.z);
                        SFW(SF_RX, sl, r.rd.x); SFW(SF_RY, sl, r.rd.y); SFW(SF_RZ, sl, r.rd.z);
                        inst = GDPT_NO_INSTANCE;
                        SF(SF_INST, sl) = inst;
                    }
                    tag = sorted_tag_of(r.cur, inst);
                }
                SF(SF_CUR, sl) = r.cur; SF(SF_SP, sl) = r.sp;
                SF(SF_STEPS, sl) += steps;
                new_tag = tag;
                if (COUNT) { own.node_steps += steps; own.box_tests += 4u * steps; }
            }
        } else if (phase == TAG_L) {
            // ---------------- L: one leaf ----------------
            it_l++;
            if (act) {
                SlotStack<kSlots> st;
                st.col = stk + sl;
                st.spill = spill_base + (size_t)sl * a.sorted_spill_depth;
                RayState r;
                r.o = mk3(SFF(SF_OX, sl), SFF(SF_OY, sl), SFF(SF_OZ, sl));
                r.d = mk3(SFF(SF_DX, sl), SFF(SF_DY, sl), SFF(SF_DZ, sl));
                r.t = SFF(SF_T, sl);
                r.inst = SF(SF_INST, sl);
                r.overflow = SF(SF_FLAGS, sl);
                r.sp = SF(SF_SP, sl);
                const uint32_t leaf = SF(SF_CUR, sl);
                const float t_in = r.t;
                const uint32_t flags_in = r.overflow;
                r.u = 0.0f; r.v = 0.0f; r.tri = 0u; r.blas_front = 0u;
                r.cur = fast_pop(r, st);
                fast_leaf_tests(a.sc, r, leaf);
                if (r.t < t_in) {
                    SFW(SF_T, sl, r.t); SFW(SF_U, sl, r.u); SFW(SF_V, sl, r.v); SF(SF_TRI, sl) = r.tri; SF(SF_BF, sl) = r.blas_front;
                }
                if (r.overflow != flags_in) SF(SF_FLAGS, sl) = r.overflow;
                SF(SF_CUR, sl) = r.cur; SF(SF_SP, sl) = r.sp;
                SF(SF_STEPS, sl) += 1u;
                if (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) { // the instance is done: back to world space (leaves exist inside instances only)
                    const f3 wo = mk3(SFF(SF_WOX, sl), SFF(SF_WOY, sl), SFF(SF_WOZ, sl));
                    const f3 wd = mk3(SFF(SF_WDX, sl), SFF(SF_WDY, sl), SFF(SF_WDZ, sl));
                    const f3 wrd = fast_rcp3(wd);
                    SFW(SF_OX, sl, wo.x); SFW(SF_OY, sl, wo.y); SFW(SF_OZ, sl, wo.z);
                    SFW(SF_DX, sl, wd.x); SFW(SF_DY, sl, wd.y); SFW(SF_DZ, sl, wd.z);
                    SFW(SF_RX, sl, wrd.x); SFW(SF_RY, sl, wrd.y); SFW(SF_RZ, sl, wrd.z);
                    r.inst = GDPT_NO_INSTANCE;
                    SF(SF_INST, sl) = r.inst;
                }
                new_tag = sorted_tag_of(r.cur, r.inst);
                if (COUNT) own.tri_tests += ((leaf >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u;
            }
        } else if (phase == TAG_T) {
            // ---------------- T: into the instance the link names (main.glsl:316-321) ----------------
            it_t++;
            if (act) {
                SlotStack<kSlots> st;
                st.col = stk + sl;
                st.spill = spill_base + (size_t)sl * a.sorted_spill_depth;
                RayState r;
                r.wo = mk3(SFF(SF_WOX, sl), SFF(SF_WOY, sl), SFF(SF_WOZ, sl));
                r.wd = mk3(SFF(SF_WDX, sl), SFF(SF_WDY, sl), SFF(SF_WDZ, sl));
                r.t = SFF(SF_T, sl);
                r.cur = SF(SF_CUR, sl); r.sp = SF(SF_SP, sl); r.inst = SF(SF_INST, sl);
                r.overflow = SF(SF_FLAGS, sl);
                const uint32_t flags_in = r.overflow;
                r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd);
                r.inst = GDPT_NO_INSTANCE;
                if (r.cur & LINK_LEAF) {
                    fast_enter_instance<true>(a.sc, r, st);
                    if (COUNT) own.inst_entries++;
                    if (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) { // its box is missed and the next link is a world-level one
                        r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd);
                        r.inst = GDPT_NO_INSTANCE;
                    }
                }
                SFW(SF_OX, sl, r.o.x); SFW(SF_OY, sl, r.o.y); SFW(SF_OZ, sl, r.o.z);
                SFW(SF_DX, sl, r.d.x); SFW(SF_DY, sl, r.d.y); SFW(SF_DZ, sl, r.d.z);
                SFW(SF_RX, sl, r.rd.x); SFW(SF_RY, sl, r.rd.y); SFW(SF_RZ, sl, r.rd.z);
                SF(SF_CUR, sl) = r.cur; SF(SF_SP, sl) = r.sp; SF(SF_INST, sl) = r.inst;
                if (r.overflow != flags_in) SF(SF_FLAGS, sl) = r.overflow;
                SF(SF_STEPS, sl) += 1u;
                new_tag = sorted_tag_of(r.cur, r.inst);
            }
        } else {
            // ---------------- F: prove, shade, bounce ----------------
            it_f++;
            if (act) {
                const f3 wo = mk3(SFF(SF_WOX, sl), SFF(SF_WOY, sl), SFF(SF_WOZ, sl));
                const f3 wd = mk3(SFF(SF_WDX, sl), SFF(SF_WDY, sl), SFF(SF_WDZ, sl));
                float ht = SFF(SF_T, sl), hu = SFF(SF_U, sl), hv = SFF(SF_V, sl);
                uint32_t htri = SF(SF_TRI, sl), hbf = SF(SF_BF, sl), hflags = SF(SF_FLAGS, sl);
                const uint32_t pixel = SF(SF_PIXEL, sl);
                const int segment = (int)SF(SF_SEGMENT, sl);
                if (COUNT) own.proofs++;
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
                f3 radiance = mk3(SFF(SF_RAR, sl), SFF(SF_RAG, sl), SFF(SF_RAB, sl));
                const f3 throughput = mk3(SFF(SF_THR, sl), SFF(SF_THG, sl), SFF(SF_THB, sl));
                bool alive = false;
                if (!hit) {
                    radiance = radiance + throughput * sample_sky(wd);
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, cam.z_far);
                } else {
                    BounceResult br;
                    u2 sd; sd.x = SF(SF_SEEDX, sl); sd.y = SF(SF_SEEDY, sl);
                    shade_and_bounce_ool(&a.sc, wo, wd, ht, hu, hv, htri, hbf, radiance, throughput, &sd, &br);
                    radiance = br.radiance;
                    if (segment == 0) a.out_depth[pixel] = encode_depth(cam, br.first_hit_distance);
                    alive = br.alive && segment < last_segment;
                    if (alive) {
                        SF(SF_SEEDX, sl) = sd.x; SF(SF_SEEDY, sl) = sd.y;
                        SFW(SF_THR, sl, br.throughput.x); SFW(SF_THG, sl, br.throughput.y); SFW(SF_THB, sl, br.throughput.z);
                        SFW(SF_RAR, sl, radiance.x); SFW(SF_RAG, sl, radiance.y); SFW(SF_RAB, sl, radiance.z);
                        SFW(SF_WOX, sl, br.next_o.x); SFW(SF_WOY, sl, br.next_o.y); SFW(SF_WOZ, sl, br.next_o.z);
                        SFW(SF_WDX, sl, br.next_d.x); SFW(SF_WDY, sl, br.next_d.y); SFW(SF_WDZ, sl, br.next_d.z);
                        SFW(SF_OX, sl, br.next_o.x); SFW(SF_OY, sl, br.next_o.y); SFW(SF_OZ, sl, br.next_o.z);
                        SFW(SF_DX, sl, br.next_d.x); SFW(SF_DY, sl, br.next_d.y); SFW(SF_DZ, sl, br.next_d.z);
                        const f3 rd = fast_rcp3(br.next_d);
                        SFW(SF_RX, sl, rd.x); SFW(SF_RY, sl, rd.y); SFW(SF_RZ, sl, rd.z);
                        SFW(SF_T, sl, 1e9f); SF(SF_U, sl) = 0u; SF(SF_V, sl) = 0u; SF(SF_TRI, sl) = 0u; SF(SF_BF, sl) = 0u;
                        SF(SF_FLAGS, sl) = fast_far_origin(br.next_o, a.sc.fast_world_reach) ? RAY_FAR : 0u;
                        SF(SF_CUR, sl) = a.sc.fast4_root; SF(SF_SP, sl) = 0u; SF(SF_INST, sl) = GDPT_NO_INSTANCE;
                        SF(SF_SEGMENT, sl) = (uint32_t)(segment + 1);
                        new_tag = sorted_tag_of(a.sc.fast4_root, GDPT_NO_INSTANCE);
                    }
                }
                if (!alive) {
                    a.out_rgba8[pixel] = pack_rgba8(radiance);
                    // scheduling hint for the next frame: running mean of the path's cost at this pixel
                    const uint32_t cost = SF(SF_STEPS, sl);
                    if (a.cost_ema) {
                        const uint32_t old = a.cost[pixel];
                        a.cost[pixel] = old ? (old * 3u + cost + 2u) >> 2 : cost;
                    } else {
                        a.cost[pixel] = cost;
                    }
                    new_tag = TAG_FREE;
                }
            }
        }
        if (act) {
            __threadfence_block(); // the slot's state before its tag
            sorted_release(my_col, row, new_tag);
            const unsigned same = __match_any_sync(claimed, new_tag); // one counter update per tag value
            if (lane == (unsigned)(__ffs((int)same) - 1)) atomicAdd(&counts[new_tag], (uint32_t)__popc(same));
        }
    }
#undef SF
#undef SFF
#undef SFW
    if (prof && lane == 0) {
        unsigned long long *w = a.warp_prof + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u;
        // low word: iterations of the phase; high word: lanes that held a ray in them ([7]: paths started | idle iterations)
        w[0] = t_start; w[1] = global_ns();
        w[2] = it_i | ((unsigned long long)ln_i << 32); w[3] = it_l | ((unsigned long long)ln_l << 32);
        w[4] = it_t | ((unsigned long long)ln_t << 32); w[5] = it_f | ((unsigned long long)ln_f << 32);
        w[6] = it_e | ((unsigned long long)ln_e << 32); w[7] = n_started | ((unsigned long long)it_idle << 32);
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
    if (COUNT) {
        for (int off = 16; off > 0; off >>= 1) {
            own.node_steps += __shfl_down_sync(kFull, own.node_steps, off);
            own.box_tests += __shfl_down_sync(kFull, own.box_tests, off);
            own.tri_tests += __shfl_down_sync(kFull, own.tri_tests, off);
            own.inst_entries += __shfl_down_sync(kFull, own.inst_entries, off);
            own.proofs += __shfl_down_sync(kFull, own.proofs, off);
        }
        if (lane == 0) {
            atomicAdd(&cnt->own_node_steps, own.node_steps); atomicAdd(&cnt->own_box_tests, own.box_tests);
            atomicAdd(&cnt->own_tri_tests, own.tri_tests); atomicAdd(&cnt->own_inst_entries, own.inst_entries);
            atomicAdd(&cnt->own_proofs, own.proofs);
        }
    }
    if (my_overflow) atomicOr(&cnt->overflow, 1u);
}

#endif
