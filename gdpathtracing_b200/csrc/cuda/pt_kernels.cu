// pt_kernels.cu -- the sm_100a kernels behind "main.glsl" (K1), "progressive_rendering.glsl" (K2) and
// "temporal_reprojection.glsl" (K3).  Compile with -fmad=false (see pt_math.cuh).
//
// K1 = a camera-ray classification kernel (k_primary_cull) + one persistent path kernel:
//   schedule 6 (default)  k_path_pool (k_path_pool.cuh): closest-hit search over our own four-wide SAH tables +
//                         proof that the reference traversal returns the same record (pt_fast.cuh), paths pooled
//                         per warp in shared memory
//   schedule 3            k_path<SRC 1> (k_path_ref.cuh): the reference's own visiting order with tight-box culling
//   schedule 2            k_path<SRC 0>: one kernel over all pixels, no classification; what trace mode,
//                         DEBUG_STEPS and GDPT_CULL 0 run (full reference visit order, work counters)
// Every schedule produces the same bytes.
#include "pt_kernels.cuh"
#include "pt_fast.cuh"
#include "pt_post.cuh"
#include "pt_shade.cuh"
#include "pt_trace.cuh"

#include <algorithm>
#include <cstdio>

#ifndef GDPT_OOL_COLD
#define GDPT_OOL_COLD 1
#endif

namespace gdpt {

namespace {

constexpr int kTraceThreads = 128;
constexpr int kSmemStack = 16;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kChunkPrimary = 32; // one 8x4-pixel tile: small chunks keep the expensive tiles spread over many warps
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

// tiles_x_magic = ceil(2^32 / tiles_x): tile / tiles_x = umulhi(tile, magic) while tile * tiles_x < 2^32 (the error of the
// magic number is below tiles_x / 2^32 per unit of tile); 0 = use the division
__device__ __forceinline__ uint32_t tiles_x_magic(const FrameArgs &a)
{
    const uint32_t tiles_x = ((uint32_t)a.width + 7u) >> 3;
    const unsigned long long tiles = (unsigned long long)(a.n_work >> 5);
    return (tiles_x > 1u && tiles * tiles_x < 0xFFFFFFFFull) ? 0xFFFFFFFFu / tiles_x + 1u : 0u;
}
__device__ __forceinline__ bool work_to_pixel(const FrameArgs &a, uint32_t w, int *px, int *py, uint32_t magic = 0u)
{
    const uint32_t tile = w >> 5, lane = w & 31u;
    const uint32_t tiles_x = ((uint32_t)a.width + 7u) >> 3;
    const uint32_t ty = magic ? __umulhi(tile, magic) : tile / tiles_x, tx = tile - ty * tiles_x;
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


__device__ __forceinline__ void write_trace_record(const FrameArgs &a, int segment, uint32_t pixel, const RayState &r,
                                                   const TraceCounters &tc)
{
    if (!a.trace || segment >= a.trace_segments) return;
    gdpt_trace_record rec;
    const bool hit = r.t < 1e9f;
    rec.hit = hit ? 1u : 0u;
    rec.triangle = hit ? r.tri : 0u;
    rec.blas = hit ? hit_blas(r.blas_front) : 0u;
    rec.front = hit ? (r.blas_front >> 31) : 0u;
    rec.t = r.t; rec.u = hit ? r.u : 0.0f; rec.v = hit ? r.v : 0.0f;
    rec.node_pops = tc.node_pops; rec.box_tests = tc.box_tests; rec.tri_tests = tc.tri_tests;
    rec.tlas_leaves = tc.tlas_leaves; rec.max_stack = tc.max_stack;
    rec.visit_hash_lo = (uint32_t)tc.hash; rec.visit_hash_hi = (uint32_t)(tc.hash >> 32);
    a.trace[(size_t)segment * a.width * a.height + pixel] = rec;
}

#include "k_path_ref.cuh"

// ------------------------------------------------------------------------------------------------
// Shared by the closest-hit path kernels: the exact re-trace of a ray in reference order (ties, failed proofs),
// hit records of the rendering kernels, the out-of-line verdict.
struct LocalStack {
    uint32_t *slots;
    __device__ __forceinline__ void store(uint32_t i, uint32_t v) { slots[i] = v; }
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return slots[i]; }
};
struct ExactHit { float t, u, v; uint32_t tri, blas_front, overflow; };
// `trust_margins` false (the ray starts beyond the reach the culling margins are sized for, RAY_FAR): the full reference
// visiting order, no tight-box culling -- exact whatever the distances.
__device__ __noinline__ void exact_retrace(const SceneView *sc, f3 wo, f3 wd, bool trust_margins, ExactHit *out)
{
    uint32_t slots[GDPT_MAX_STACK];
    LocalStack st;
    st.slots = slots;
    RayState r;
    ray_begin(r, *sc, wo, wd);
    if (trust_margins) trace_ray_compact<false, true>(*sc, r, st, nullptr);
    else trace_ray_compact<false, false>(*sc, r, st, nullptr);
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
    rec.blas = hit ? hit_blas(blas_front) : 0u;
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

#include "k_path_pool.cuh"

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
    const uint32_t magic = tiles_x_magic(a);
    const float far_depth = encode_depth(cam, cam.z_far); // what every finished pixel stores (main.glsl:430-431)
    for (uint32_t w0 = blockIdx.x * blockDim.x; w0 < a.n_work; w0 += stride) { // warp-uniform trip count
        const uint32_t w = w0 + threadIdx.x;
        int px, py;
        const bool valid = w < a.n_work && work_to_pixel(a, w, &px, &py, magic);
        bool survivor = false;
        uint32_t pixel = 0;
        if (valid) {
            f3 o, d;
            generate_primary_ray(cam, a.width, a.height, px, py, &o, &d);
            pixel = (uint32_t)py * (uint32_t)a.width + (uint32_t)px;
            RayState r;
            if (a.sc.fast4_ok && a.schedule == 6) {
                // What the path kernel's search would do at the TLAS level, as an any-hit query: four-wide steps over the
                // instances' true world boxes, then the touched instance's own box in its space; the first instance whose
                // box the ray touches makes the pixel a survivor.  Rays the search is not trusted with survive unseen.
                fast_ray_begin(r, a.sc, o, d);
                const uint32_t unsearched = r.overflow & RAY_UNSEARCHED;
                if (unsearched) {
                    // the path kernel will answer this ray with the exact traversal: classify it the same way -- the
                    // reference's own TLAS walk, with the tight boxes only where their margins are trusted (not RAY_FAR)
                    ray_begin(r, a.sc, o, d);
                    if (unsearched & RAY_FAR) { while (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) step_tlas<false, false>(a.sc, r, st, nullptr); }
                    else { while (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) step_tlas<false, true>(a.sc, r, st, nullptr); }
                    survivor = r.cur != LINK_NONE;
                    r.cur = LINK_NONE;
                } else {
                    r.cur = a.sc.fast4_root;
                }
                while (!survivor && r.cur != LINK_NONE) {
                    if (r.cur & LINK_LEAF) {
                        fast_enter_instance<true>(a.sc, r, st);
                        survivor = (r.overflow & RAY_UNSEARCHED) != 0u || (r.cur != LINK_NONE && (r.cur & LINK_TLAS) == 0u);
                        r.o = r.wo; r.d = r.wd; r.rd = fast_rcp3(r.wd); r.inst = GDPT_NO_INSTANCE; // missed: back to world space
                    } else {
                        fast_step_node4(a.sc, r, st);
                    }
                }
            } else {
                ray_begin(r, a.sc, o, d);
                while (r.cur != LINK_NONE && (r.cur & LINK_TLAS)) step_tlas<false, true>(a.sc, r, st, nullptr);
                survivor = r.cur != LINK_NONE;
            }
            if (!survivor) {
                a.out_rgba8[pixel] = pack_rgba8(mk3(0.0f, 0.0f, 0.0f) + mk3(1.0f, 1.0f, 1.0f) * sample_sky(d));
                a.out_depth[pixel] = far_depth;
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

// K2 (progressive_rendering.glsl:28-46): acc = (frame_count > 1 ? acc : 0) + screen;
// screen = rgba8(ACES(acc / frame_count)).  HBM-bound: 36 B per pixel (4 R raw + 16 R + 16 W accumulation + 4 W screen;
// the first frame reads no accumulation).  One pixel per lane: a warp's accumulator load / store is one 512 B
// instruction over four full lines, its RGBA8 load / store one 128 B line -- no half-used sector.
// `peers`: RGBA8 images of the other GPUs of a row-band frame (peer memory over NVLink, CUDA IPC).  Every
// tone-mapped pixel this GPU owns is also stored there, so the presented frame assembles itself in every
// GPU's image while the kernel runs -- the exchange step of the row-band partition fused into its producer
// instead of an all-gather after it.
// `raw` is the image K1 wrote (the bound screen image itself in the reference's call sequence; a frame-private image
// when two pipelined frames overlap), `screen` receives the tone-mapped result.  frame_count comes from the device
// Params block (stream-ordered uploads) or, when `params` is null, from the argument.
//
// The six quotients of a pixel (acc / frame_count and the ACES fraction, three channels each) are IEEE divisions in the
// reference's arithmetic.  K2 runs the reciprocal + FMA sequence of the compiler's own division fast path itself
// (pt_shade.cuh, `quotient_in_window`) behind ONE test per pixel: the three sums are +0 or inside [2^-40, 2^40] and
// frame_count >= 1.  Then x = sum / frame_count is +0 or inside [2^-72, 2^40], the ACES denominator inside [0.14, 2^82],
// its numerator +0 or at least 2^-78, and every product, residual and reciprocal of the six sequences is +0 or a normal
// number, so each returns the correctly rounded quotient (+0 for a +0 numerator: 0 * r = +0, residual +0).  The
// reciprocal of frame_count is shared by the three channels.  Any other pixel (a negative, tiny, huge or non-finite
// value somebody stored in the accumulation buffer) takes the plain `/` path, out of line.
// (tests/test_gpu_k2.py: crafted sums on both sides of the window's edges against the oracle's divisions.)
#ifndef GDPT_K2_UNROLL
#define GDPT_K2_UNROLL 1
#endif
#ifndef GDPT_K2_MINB
#define GDPT_K2_MINB 8
#endif
constexpr int kProgressiveUnroll = GDPT_K2_UNROLL;
__device__ __noinline__ uint32_t progressive_tone_map_generic(f3 rad, float fc)
{
    const f3 avg = (rad / fc) * 1.0f;
    return pack_rgba8(mk3(aces_channel(avg.x), aces_channel(avg.y), aces_channel(avg.z)));
}
__global__ void __launch_bounds__(256, GDPT_K2_MINB) k_progressive(const uint32_t *__restrict__ raw, uint32_t *screen, float4 *accum,
                                                     const gdpt_progressive_params *__restrict__ params, uint32_t frame_count_arg,
                                                     int width, int height, int shard_part, int shard_parts, int shard_band,
                                                     const PeerScreens peers)
{
    // imageLoad of an rgba8 texel = byte / 255.0f (progressive_rendering.glsl:33): every block divides each of the 256
    // byte values once and looks the quotients up afterwards -- the same IEEE quotients, three divisions per pixel fewer
    __shared__ float s_unorm[256];
    s_unorm[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const uint32_t frame_count = params ? params->frame_count : frame_count_arg;
    const float fc = (float)frame_count;
    const float r_fc = refined_rcp(fc); // (unused garbage when frame_count is 0: those pixels take the generic path)
    const size_t n = (size_t)width * height;
    // every block takes one contiguous, equally long range of pixels (a multiple of 128, so warps stay line-aligned):
    // the grid is one wave of resident blocks and they all finish together
    const size_t per_block = ((n + gridDim.x - 1) / gridDim.x + 127u) & ~(size_t)127u;
    const size_t begin = (size_t)blockIdx.x * per_block, end = begin + per_block < n ? begin + per_block : n;
    const size_t tile = (size_t)blockDim.x * kProgressiveUnroll;
    for (size_t base = begin; base < end; base += tile) {
        uint32_t in[kProgressiveUnroll];
        float4 acc[kProgressiveUnroll];
        bool mine[kProgressiveUnroll];
#pragma unroll
        for (int k = 0; k < kProgressiveUnroll; k++) {
            const size_t p = base + (size_t)k * blockDim.x + threadIdx.x;
            mine[k] = p < end;
            if (mine[k] && shard_parts > 1) mine[k] = ((int)(p / (size_t)width) / shard_band) % shard_parts == shard_part;
            in[k] = 0u;
            acc[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (mine[k]) {
                in[k] = __ldg(raw + p);
                if (frame_count > 1u) acc[k] = accum[p];
            }
        }
#pragma unroll
        for (int k = 0; k < kProgressiveUnroll; k++) {
            if (!mine[k]) continue;
            const size_t p = base + (size_t)k * blockDim.x + threadIdx.x;
            f3 rad = mk3(s_unorm[in[k] & 0xffu], s_unorm[(in[k] >> 8) & 0xffu], s_unorm[(in[k] >> 16) & 0xffu]);
            if (frame_count > 1u) rad = rad + mk3(acc[k].x, acc[k].y, acc[k].z);
            accum[p] = make_float4(rad.x, rad.y, rad.z, 1.0f);
            const bool window = frame_count != 0u && in_quotient_window(rad);
            uint32_t o;
            if (window) {
                o = pack_rgba8(mk3(aces_channel_in_window(quotient_in_window(rad.x, fc, r_fc)),
                                   aces_channel_in_window(quotient_in_window(rad.y, fc, r_fc)),
                                   aces_channel_in_window(quotient_in_window(rad.z, fc, r_fc))));
            } else {
                o = progressive_tone_map_generic(rad, fc);
            }
            screen[p] = o;
            if (peers.n > 0) {
#pragma unroll
                for (int j = 0; j < kMaxPeerScreens; j++) // (constant indices: the pointer table stays in the parameter bank)
                    if (j < peers.n) peers.p[j][p] = o;
            }
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
    __shared__ float s_unorm[256]; // imageLoad of an rgba8 texel = byte / 255.0f (temporal_reprojection.glsl:36), as in k_progressive
    if (threadIdx.x < sizeof(gdpt_temporal_params) / 4u)
        reinterpret_cast<uint32_t *>(&p)[threadIdx.x] = reinterpret_cast<const uint32_t *>(params)[threadIdx.x];
    s_unorm[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const int tiles_x = (p.width + 31) >> 5, tiles_y = (p.height + 7) >> 3;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int x = (tile % tiles_x) * 32 + (int)(threadIdx.x & 31u), y = (tile / tiles_x) * 8 + (int)(threadIdx.x >> 5);
        if (x < p.width && y < p.height) temporal_pixel(p, x, y, screen, depth, history, next, s_unorm);
    }
}

struct Shapes {
    bool ready = false;
    int sms = 148;
    int path_blocks[2][2] = {};  // k_path<TRACE, CULL, SRC 0>
    int path_list_blocks = 0;    // k_path<false, true, SRC 1, 4, COMPACT>
    int path_list_record_blocks = 0; // k_path<true, true, SRC 1> (hit records of schedule 3)
    int pool_blocks[2][2] = {};  // k_path_pool<REC, .., WIDE>
    int pool_count_blocks = 0;   // k_path_pool<.., WIDE, COUNT>
    int pool_dense_blocks = 0;   // k_path_pool<false, kPoolDenseMinBlocks, .., WIDE>: 96 registers, five blocks per SM
    int cull_blocks_per_sm = 1;
    int prog_blocks = 0;
};
Shapes g_shapes[16];

} // namespace

void init_launch_shapes(int device)
{
    Shapes &s = g_shapes[device & 15];
    if (s.ready) return;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    s.sms = prop.multiProcessorCount;
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
    s.path_list_blocks = grid_of(k_path<false, true, 1, 4, true>, kTraceThreads);
    s.path_list_record_blocks = grid_of(k_path<true, true, 1>, kTraceThreads);
    s.pool_blocks[0][0] = grid_of(k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, false>, kTraceThreads);
    s.pool_blocks[0][1] = grid_of(k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, true>, kTraceThreads);
    s.pool_blocks[1][0] = grid_of(k_path_pool<true, kPoolMinBlocks, kPoolSlotsDefault, false>, kTraceThreads);
    s.pool_blocks[1][1] = grid_of(k_path_pool<true, kPoolMinBlocks, kPoolSlotsDefault, true>, kTraceThreads);
    s.pool_count_blocks = grid_of(k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, true, true>, kTraceThreads);
    s.pool_dense_blocks = grid_of(k_path_pool<false, kPoolDenseMinBlocks, kPoolSlotsDefault, true>, kTraceThreads);
    grid_of(k_primary_cull, kTraceThreads);
    s.cull_blocks_per_sm = per_sm > 0 ? per_sm : 1;
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

void launch_path_list(const FrameArgs &a, bool record, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    if (record) k_path<true, true, 1><<<persistent_grid(sh, a, sh.path_list_record_blocks), kTraceThreads, 0, s>>>(a);
    else k_path<false, true, 1, 4, true><<<persistent_grid(sh, a, sh.path_list_blocks), kTraceThreads, 0, s>>>(a);
}

void launch_progressive(const uint32_t *raw_rgba8, uint32_t *screen_rgba8, float4 *accum, const gdpt_progressive_params *params_dev,
                        uint32_t frame_count, int width, int height, int shard_part, int shard_parts, int shard_band,
                        const PeerScreens &peers, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    const size_t warps = ((size_t)width * height + 127u) / 128u; // no more blocks than 128-pixel pieces
    const int blocks = (int)(warps < (size_t)sh.prog_blocks ? (warps > 0 ? warps : 1) : (size_t)sh.prog_blocks);
    k_progressive<<<blocks, 256, 0, s>>>(raw_rgba8, screen_rgba8, accum, params_dev, frame_count, width, height, shard_part,
                                         shard_parts, shard_band, peers);
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

// Compile-time shape of k_path_pool, chosen by A/B on the B200 (C2 demo frame / C4 instanced at 1080p, ms):
//   parked leaves 1 | 2 | 3 | 4          0.859 | 0.892 | 0.908 | 0.918      20.3 | 20.9 | 21.4 | 21.8
//   (a lane that reached a leaf kept descending with the leaf parked; under the all-phases loop the code this takes costs more
//   than the descent it saves -- 1 | 0 parked leaves: C2 0.559 | 0.514, C4 9.58 | 9.13 -- and it is gone)
//   slots per warp 40 | 64 | 96          0.961 | 0.878 | 0.896              21.5 | 20.5 | 20.9
//   (with every phase per iteration, 64 | 80 | 96: C2 0.595 | 0.591 | 0.590, end to end 4 580 | 4 790 | 4 757 Mrays/s; C4 9.94 | 9.75 | 9.88 -> 80)
//   blocks per SM 4 | 5 | 6 | 8          0.854 | 0.960 | 1.124 | 1.320      19.0 | 18.6 | 21.8 | 24.8   (registers 128 | 96 | 80 | 64)
void launch_path_pool(const FrameArgs &a_in, bool record, cudaStream_t s)
{
    Shapes &sh = shapes_for_current_device();
    FrameArgs a = a_in; // scheduling knobs in the form the kernel reads them (k_path_pool.cuh)
    a.refill_below = std::min(std::max(a_in.refill_below, 1), 32);
    a.pool_wait = a_in.pool_wait > 0 ? a_in.pool_wait : -1; // read as uint32: off
    a.pool_alive = (a_in.pool_alive >= 32 && a_in.pool_alive < kPoolSlotsDefault) ? kPoolSlotsDefault - a_in.pool_alive : 0; // slots kept free
    a.shade_at = std::min(std::max(a_in.shade_at, 1), 32);
    const bool wide = a.wide_bvh != 0 && a.sc.fast4_ok != 0; // four-wide tables (fast_bvh.h Collapse)
    const int grid = persistent_grid(sh, a, sh.pool_blocks[record ? 1 : 0][wide ? 1 : 0]);
    if (a.count_work && wide && !record) { // the instantiation that also counts its own work (bench.py's roofline numerator)
        k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, true, true><<<persistent_grid(sh, a, sh.pool_count_blocks), kTraceThreads, 0, s>>>(a);
        return;
    }
    if (a.warp_prof && wide && !record) { // per-warp schedule profile (tools/warp_profile.py): same grid as the timed instantiation
        k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, true, false, true><<<grid, kTraceThreads, 0, s>>>(a);
        return;
    }
    if (a.pool_dense && wide && !record) { // throughput-bound scenes: five blocks per SM at 96 registers (a few spilled words)
        k_path_pool<false, kPoolDenseMinBlocks, kPoolSlotsDefault, true>
            <<<persistent_grid(sh, a, sh.pool_dense_blocks), kTraceThreads, 0, s>>>(a);
        return;
    }
    if (record) {
        if (wide) k_path_pool<true, kPoolMinBlocks, kPoolSlotsDefault, true><<<grid, kTraceThreads, 0, s>>>(a);
        else k_path_pool<true, kPoolMinBlocks, kPoolSlotsDefault, false><<<grid, kTraceThreads, 0, s>>>(a);
    } else {
        if (wide) k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, true><<<grid, kTraceThreads, 0, s>>>(a);
        else k_path_pool<false, kPoolMinBlocks, kPoolSlotsDefault, false><<<grid, kTraceThreads, 0, s>>>(a);
    }
}

size_t path_kernel_warps(const FrameArgs &a)
{
    Shapes &sh = shapes_for_current_device();
    if (a.schedule == 6) return (size_t)sh.pool_blocks[0][(a.wide_bvh != 0 && a.sc.fast4_ok != 0) ? 1 : 0] * (kTraceThreads / 32);
    if (a.schedule == 3) return (size_t)sh.path_list_blocks * (kTraceThreads / 32);
    int most = 0;
    for (int t = 0; t < 2; t++)
        for (int c = 0; c < 2; c++) most = most > sh.path_blocks[t][c] ? most : sh.path_blocks[t][c];
    return (size_t)most * (kTraceThreads / 32);
}

int k1_launch_count(int schedule) { return schedule == 2 ? 1 : 2; }

} // namespace gdpt
