// pt_kernels.cuh -- host-visible launch interface of the path-tracing kernels.
#ifndef GDPT_PT_KERNELS_CUH
#define GDPT_PT_KERNELS_CUH

#include "pt_scene.cuh"
#include <cuda_runtime.h>

namespace gdpt {

enum { kMaxDepth = 32 };

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
};

// Path queue: structure-of-arrays of five 16 B planes per entry.
//   plane 0: ray.o.xyz, pixel          plane 1: ray.d.xyz, hit t
//   plane 2: throughput.rgb, hit u     plane 3: radiance.rgb, hit v
//   plane 4: seed.x, seed.y, hit triangle, hit blas|front<<31
struct FrameArgs {
    SceneView sc;
    const gdpt_camera *camera; // bound Camera storage buffer (device)
    int width, height, max_depth;
    int shard_part, shard_parts, shard_band; // row bands rendered by this device
    int local_rows;                          // rows this shard owns
    uint32_t n_work;                         // primary work items (8x4-pixel tiles x 32)
    uint32_t *out_rgba8;                     // bound RGBA8 image
    float *out_depth;                        // bound R32F image
    float4 *queue[2];
    uint32_t queue_cap;
    uint32_t *hit_list;
    FrameCounters *counters;
    // scheduling knobs (results do not depend on them)
    int refill_below;  // refill idle lanes when fewer than this many lanes are traversing
    int burst;         // node steps between refill checks
    int schedule;      // 0: wavefront + while-while descent, 1: wavefront + phase voting, 2: single path kernel
    int shade_at;      // schedule 2: shade once this many lanes wait with a finished ray
    int cull;          // 1: skip children whose tight box the ray misses (pt_scene.cuh); results identical
    // parity outputs (TRACE builds only)
    gdpt_trace_record *trace; int trace_segments;
    uint32_t *visits; uint32_t visits_per_ray;
    int debug_steps;
};

struct LaunchShape { int blocks; int threads; };

// K1 stages.  `trace` selects the instrumented instantiation.
// Single-kernel schedule (a.schedule == 2): whole paths per lane, no stage barriers.
void launch_path(const FrameArgs &a, bool trace, cudaStream_t s);
void launch_primary(const FrameArgs &a, bool trace, cudaStream_t s);
void launch_shade(const FrameArgs &a, int segment, cudaStream_t s);
void launch_trace(const FrameArgs &a, int segment, bool trace, cudaStream_t s);
// K2: progressive accumulation + ACES (progressive_rendering.glsl:28-46).
void launch_progressive(uint32_t *screen_rgba8, float4 *accum, const gdpt_progressive_params *params_dev,
                        int width, int height, int shard_part, int shard_parts, int shard_band, cudaStream_t s);
// Number of kernels one K1 dispatch launches for a given depth.
int k1_launch_count(int max_depth, bool debug_steps);
// One-time per device: query SM count / occupancy for the persistent grids.
void init_launch_shapes(int device);

} // namespace gdpt
#endif
