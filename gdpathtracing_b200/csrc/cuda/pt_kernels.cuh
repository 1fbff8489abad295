// pt_kernels.cuh -- host-visible launch interface of the path-tracing kernels.
#ifndef GDPT_PT_KERNELS_CUH
#define GDPT_PT_KERNELS_CUH

#include "pt_scene.cuh"
#include <cuda_runtime.h>

namespace gdpt {

enum { kMaxDepth = 32 };
// Surviving pixels are binned by the cost (lane steps) their path had in the previous frame and started
// longest first: class 0 = < 16 steps or unknown, 1: >= 16, 2: >= 32, ... 7: >= 1024.  Classes from
// kFirstHeavyClass (>= 128 steps) up are handed out a few pixels at a time so they spread over many warps; the
// light classes in tile-sized chunks, shortest last, which bounds the drain after the list is exhausted
// by the latency of a short path.
enum { kCostClasses = 8, kFirstHeavyClass = 4 };

// Device-resident counters of one frame; zeroed with one memset per dispatch.
struct FrameCounters {
    uint32_t cursor[2 * kMaxDepth + 2]; // dynamic work-fetch cursor per stage
    uint32_t qcount[kMaxDepth + 1];     // entries appended to the path queue by primary / shade(i)
    uint32_t lcount[kMaxDepth + 1];     // hit-list length produced by trace stage i
    uint32_t max_stack;
    uint32_t overflow;                  // a ray exceeded the reference's stack capacity
    unsigned long long primary_hits;
    unsigned long long rays;            // ray_trace() calls, counted by the single-kernel schedule
    unsigned long long node_pops, box_tests, tri_tests, tlas_leaves; // TRACE builds only
    unsigned long long retraced;        // closest-hit schedules: rays re-traced in reference order (proof failed, tie, far origin)
    // own work of the closest-hit search (COUNT instantiations only): what the timed kernel itself executes
    unsigned long long own_node_steps, own_box_tests, own_tri_tests, own_inst_entries, own_proofs;
};

struct FrameArgs {
    SceneView sc;
    const gdpt_camera *camera; // bound Camera storage buffer (device)
    int width, height, max_depth;
    int shard_part, shard_parts, shard_band; // row bands rendered by this device
    int local_rows;                          // rows this shard owns
    uint32_t n_work;                         // primary work items (8x4-pixel tiles x 32)
    uint32_t *out_rgba8;                     // bound RGBA8 image
    float *out_depth;                        // bound R32F image
    uint32_t queue_cap;                      // pixels of the frame = capacity of the light survivor list
    uint32_t *hit_list;                      // surviving pixels of k_primary_cull, kCostClasses lists (see class_base)
    uint32_t heavy_cap;                      // capacity of each of the heavy-class lists behind the light list
    uint32_t *cost;                          // per pixel: scheduler iterations its path took in the previous frame
    int blocks_per_sm;                       // cap on resident path-kernel blocks per SM (0 = occupancy limit)
    int lead_min;                            // schedule 3: the path with the most iterations picks the phase once it has this many
    int wide_bvh;                            // schedule 6: search the four-wide tables when they exist (1) or the two-wide ones (0)
    int cost_ema;                            // schedule 6: per-pixel cost hint is a running mean over frames (1) or the last frame's (0)
    int pool_alive;                          // schedule 6: cap on the paths a warp keeps alive (32..slots; 0 = all slots)
    int pool_wait;                           // schedule 6: lane-iterations finished rays may wait before a pool service (0 = off)
    int sort4;                               // schedule 6: children of a four-wide node in full distance order (1) or nearest first (0)
    int count_work;                          // schedule 6: run the instantiation that counts its own work (untimed frames of bench.py)
    int pool_dense;                          // schedule 6: the 96-register build, five blocks per SM (many instances / large meshes)
    int miss_now;                            // schedule 6: escaped rays are finished at retire time, shading batches hold hits only
    FrameCounters *counters;
    // scheduling knobs (results do not depend on them)
    int refill_below;  // refill idle lanes when fewer than this many lanes are traversing
    int burst;         // node steps between refill checks
    int schedule;      // 2: single path kernel over all pixels in reference order (trace mode, DEBUG_STEPS, no culling),
                       // 3: camera-ray classification kernel + reference-order path kernel over the surviving pixels,
                       // 6: classification kernel + closest-hit path kernel with pooled paths (pt_fast.cuh, k_path_pool.cuh)
    int shade_at;      // schedule 2: shade once this many lanes wait with a finished ray
    int cull;          // 1: skip children whose tight box the ray misses (pt_scene.cuh); results identical
    // optional per-warp schedule profile (gdpt_shader_set_warp_profile): 8 x u64 per warp of the path kernel
    //   [0] globaltimer at start (ns)  [1] globaltimer at end  [2..6] iterations spent in phase I, L, T, F(shade), E(refill)
    //   [7] paths started
    unsigned long long *warp_prof;
    // parity outputs (TRACE builds only)
    gdpt_trace_record *trace; int trace_segments;
    uint32_t *visits; uint32_t visits_per_ray;
    int debug_steps;
};

struct LaunchShape { int blocks; int threads; };

// RGBA8 images of the other GPUs that K2 mirrors its rows into (gdpt_shader_set_peer_screens)
enum { kMaxPeerScreens = 15 };
struct PeerScreens { uint32_t *p[kMaxPeerScreens]; int n; };

// K1.  `trace` selects the instrumented instantiation.
// Schedule 2: whole paths per lane over all pixels, reference visiting order, no stage barriers.
void launch_path(const FrameArgs &a, bool trace, cudaStream_t s);
// Schedules 3 and 6 start with the camera-ray classification against the tight boxes ...
void launch_primary_cull(const FrameArgs &a, cudaStream_t s);
// ... then schedule 3 runs whole paths per lane, in reference order with culling, for the pixels that can hit something
// (`record` also writes the hit records of the first a.trace_segments segments) ...
void launch_path_list(const FrameArgs &a, bool record, cudaStream_t s);
// ... and schedule 6 (default when rendering) answers every ray by the closest-hit search of pt_fast.cuh (+ proof, exact
// re-trace where it fails), paths kept in per-warp shared-memory pools (k_path_pool): lanes swap rays instead of
// waiting for a shading quorum, shading and camera-ray generation run on full warps.
void launch_path_pool(const FrameArgs &a, bool record, cudaStream_t s);
size_t path_kernel_warps(const FrameArgs &a); // warps of the path kernel enqueue_k1 would launch for `a`
// K2: progressive accumulation + ACES (progressive_rendering.glsl:28-46).
// frame_count is read from params_dev (stream-ordered Params block) or, when that is null, taken from the argument.
void launch_progressive(const uint32_t *raw_rgba8, uint32_t *screen_rgba8, float4 *accum, const gdpt_progressive_params *params_dev,
                        uint32_t frame_count, int width, int height, int shard_part, int shard_parts, int shard_band,
                        const PeerScreens &peers, cudaStream_t s);
// K3: temporal reprojection (temporal_reprojection.glsl:31-71); `history` is the frame buffer the previous dispatch wrote.
void launch_temporal(uint32_t *screen_rgba8, const float *depth, const float *history, float *next,
                     const gdpt_temporal_params *params_dev, int width, int height, cudaStream_t s);
// Up to three word copies in one launch: the per-frame blocks (camera, post-process params) read straight from
// page-locked host memory, counters written straight into it -- keeps copy-engine operations (and the engine
// switches around them) out of the frame's stream.
struct SmallCopies { struct { uint32_t *dst; const uint32_t *src; uint32_t words; } c[3]; };
void launch_small_copies(const SmallCopies &sc, cudaStream_t s);
// Number of kernels one K1 dispatch launches.
int k1_launch_count(int schedule);
// One-time per device: query SM count / occupancy for the persistent grids.
void init_launch_shapes(int device);

} // namespace gdpt
#endif
