/* gdpt.h -- C-ABI of libgdpt_cuda.so, the B200 backend that replaces the gdcs
 * RenderingDevice compute-shader dispatch underneath PathTracingCamera.
 *
 * Every entry point is what a maintainer of the reference would bind in place
 * of one gdcs `ComputeShader` method (src/gdcs/include/gdcs.h:21-72,
 * src/gdcs/src/gdcs.cpp); the replaced reference interface is cited on each
 * declaration.  Plain pointers and sizes only: no C++ types, no torch types,
 * no exceptions cross this boundary.
 *
 * Conventions
 *  - Ownership follows the reference (PackedByteArray passed by value): the
 *    caller keeps its host memory, the callee copies at call time.  All device
 *    memory belongs to the gdpt_device and is released by gdpt_device_destroy.
 *  - gdpt_rid is the analogue of godot::RID: an opaque, device-scoped handle.
 *    0 is the invalid RID (RID::is_valid() == false).
 *  - Functions returning int return GDPT_OK (0) or a negative gdpt_status and
 *    record a message retrievable with gdpt_last_error().  The reference has no
 *    return codes (printerr + early return, gdcs.cpp:22-67); the host adapter
 *    maps non-zero to printerr + no-op to preserve that behaviour.
 *  - Not re-entrant per device; the reference is single-threaded
 *    (everything runs inside Node::_notification on the main thread).
 *  - There is no CPU fallback anywhere behind this API: without a CUDA device
 *    gdpt_device_create fails with GDPT_ERR_NO_DEVICE.
 */
#ifndef GDPT_H
#define GDPT_H

#include "gdpt_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GDPT_API __declspec(dllexport)
#else
#define GDPT_API __attribute__((visibility("default")))
#endif

typedef enum gdpt_status {
    GDPT_OK                = 0,
    GDPT_ERR_NO_DEVICE     = -1,  /* no CUDA device / ordinal out of range */
    GDPT_ERR_CUDA          = -2,  /* a CUDA runtime call failed (message has the detail) */
    GDPT_ERR_INVALID_ARG   = -3,
    GDPT_ERR_NOT_READY     = -4,  /* check_ready() would be false (gdcs.cpp:258-273) */
    GDPT_ERR_UNKNOWN_SHADER= -5,  /* shader_path does not name a kernel we implement */
    GDPT_ERR_BAD_BINDING   = -6,  /* binding table incomplete / wrong size for the shader */
    GDPT_ERR_UNSUPPORTED   = -7
} gdpt_status;

typedef struct gdpt_device gdpt_device;   /* stands in for the local RenderingDevice* */
typedef struct gdpt_shader gdpt_shader;   /* stands in for gdcs ComputeShader */
typedef uint64_t gdpt_rid;                /* stands in for godot::RID */

/* RenderingDevice::DataFormat values the path uses (path_tracing_camera.cpp:148,
 * 160,182; progressive_rendering.cpp:36). Numeric values are ours. */
typedef enum gdpt_data_format {
    GDPT_FORMAT_R8G8B8A8_UNORM      = 1,
    GDPT_FORMAT_R32_SFLOAT          = 2,
    GDPT_FORMAT_R32G32B32A32_SFLOAT = 3
} gdpt_data_format;

/* RenderingDevice::UniformType values used with add_existing_buffer
 * (progressive_rendering.cpp:30). */
typedef enum gdpt_uniform_type {
    GDPT_UNIFORM_TYPE_IMAGE          = 3,
    GDPT_UNIFORM_TYPE_STORAGE_BUFFER = 8
} gdpt_uniform_type;

/* PathTracingCamera::Denoising (path_tracing_camera.h:30-34). */
typedef enum gdpt_denoising {
    GDPT_DENOISE_PROGRESSIVE_RENDERING = 0,
    GDPT_DENOISE_TEMPORAL_REPROJECTION = 1,
    GDPT_DENOISE_NONE                  = 2
} gdpt_denoising;

/* ------------------------------------------------------------------ device */

/* Replaces RenderingServer::create_local_rendering_device()
 * (path_tracing_camera.cpp:114).  cuda_ordinal: GPU index; one device object
 * per process-visible GPU. */
GDPT_API int  gdpt_device_create(int cuda_ordinal, gdpt_device **out_device);
/* Replaces `delete _rd` + free_rid loop in ~ComputeShader (gdcs.cpp:76-87),
 * without the shared-device double-delete hazard: shaders do not own the device. */
GDPT_API void gdpt_device_destroy(gdpt_device *device);
/* Last error text of this device (or of the failed create when device==NULL). */
GDPT_API const char *gdpt_last_error(const gdpt_device *device);
/* Library/ABI version, compile-time constant. */
GDPT_API uint32_t gdpt_abi_version(void);

/* ------------------------------------------------------------------ shader */

/* Replaces ComputeShader::ComputeShader(shader_path, rd, args) (gdcs.cpp:15-74).
 * The basename of shader_path selects the hand-written kernel set:
 *   "main.glsl"                   -> path-trace kernels (K1)
 *   "progressive_rendering.glsl"  -> accumulate + ACES kernel (K2)
 *   "temporal_reprojection.glsl"  -> depth-tested reprojection + blend + ACES kernel (K3);
 *                                    bindings set0 b0 Params 88 B (gdpt_temporal_params), b1 the main
 *                                    shader's rgba8 image, b2 its r32f depth image, b3/b4 two rgba32f
 *                                    frame buffers (post_processing/temporal_reprojection.cpp:27-48)
 * `args` are the "#define ..." strings gdcs injects into the GLSL source
 * (gdcs.cpp:277-322).  Recognised defines (all optional):
 *   "#define DEBUG_STEPS"      heat-map mode of main.glsl:4,358-361,423-427
 *   "#define MAX_DEPTH n"      path segments; default 5 = main.glsl:377
 *   "#define GDPT_TRACE"       also emit per-ray parity records (gdpt_shader_read_trace)
 *   "#define GDPT_REFERENCE_ORDER"  visit every node the reference visits (no tight-box culling);
 *                              results are identical either way, only the work differs.  Implied
 *                              by GDPT_TRACE unless "#define GDPT_CULL 1" is also given
 *   "#define GDPT_RECORD_HITS n"  the rendering kernels themselves also write the hit records (no work
 *                              counters) of the first n segments; read with gdpt_shader_read_trace
 *   "#define GDPT_VARIANT 3"   keep the reference visiting order (with culling) when rendering; the default
 *                              (variant 6) answers rays with an order-free closest-hit search plus a proof that
 *                              the reference reaches the same triangle, and re-traces the rest;
 *                              variant 2 is one kernel over all pixels in reference order.  Identical results
 *   "#define GDPT_COUNT_WORK"  the closest-hit path kernel also counts the work it executes itself
 *                              (gdpt_frame_stats own_*); slower, for the roofline numerator of bench.py
 *   "#define GDPT_TUNE_<NAME> n"  scheduling knob of the path kernels (BURST, SHADE_AT, REFILL_BELOW, POOL_WAIT,
 *                              ...; A/B measurements).  Results do not depend on them.  The process environment
 *                              is never consulted: what runs is decided by this list alone
 * Unknown defines are ignored, as a GLSL compiler would ignore an unused macro. */
GDPT_API int  gdpt_shader_create(gdpt_device *device, const char *shader_path,
                                 const char *const *args, int n_args,
                                 gdpt_shader **out_shader);
/* Replaces ComputeShader::~ComputeShader (gdcs.cpp:76-87); frees the RIDs this
 * shader created, never the device. */
GDPT_API void gdpt_shader_destroy(gdpt_shader *shader);

/* Replaces ComputeShader::create_storage_buffer_uniform (gdcs.cpp:91-106). */
GDPT_API gdpt_rid gdpt_shader_create_storage_buffer_uniform(gdpt_shader *shader,
                                 const void *data, uint64_t size, int binding, int set);
/* Replaces ComputeShader::update_storage_buffer_uniform (gdcs.cpp:108-111). */
GDPT_API int  gdpt_shader_update_storage_buffer_uniform(gdpt_shader *shader, gdpt_rid rid,
                                 const void *data, uint64_t size);
/* Replaces ComputeShader::get_storage_buffer_uniform (gdcs.cpp:113-116). */
GDPT_API int  gdpt_shader_get_storage_buffer_uniform(gdpt_shader *shader, gdpt_rid rid,
                                 void *out, uint64_t capacity);

/* Replaces create_texture_format + create_image_uniform (gdcs.cpp:118-167): a
 * W x H storage image, initial contents `pixels` (may be NULL = zeros). */
GDPT_API gdpt_rid gdpt_shader_create_image_uniform(gdpt_shader *shader, const void *pixels,
                                 int width, int height, gdpt_data_format format,
                                 int binding, int set);
/* Replaces ComputeShader::create_layered_image_uniform (gdcs.cpp:174-207): a
 * sampler2DArray of n_layers RGBA8 layers, default RDSamplerState (nearest,
 * clamp-to-edge). */
GDPT_API gdpt_rid gdpt_shader_create_layered_image_uniform(gdpt_shader *shader,
                                 const void *const *layers, int n_layers,
                                 int width, int height, gdpt_data_format format,
                                 int binding, int set);
/* Replaces ComputeShader::get_image_uniform_buffer (gdcs.cpp:169-172):
 * blocking device->host read of one layer. */
GDPT_API int  gdpt_shader_get_image_uniform_buffer(gdpt_shader *shader, gdpt_rid rid, int layer,
                                 void *out, uint64_t capacity);

/* Replaces ComputeShader::add_existing_buffer (gdcs.cpp:209-223): bind a RID
 * created by another shader of the same device (the shared screen image). */
GDPT_API int  gdpt_shader_add_existing_buffer(gdpt_shader *shader, gdpt_rid rid,
                                 gdpt_uniform_type uniform_type, int binding, int set);
/* Replaces ComputeShader::finish_create_uniforms (gdcs.cpp:225-235): validates
 * the binding table against the kernel set and builds the device-side derived
 * scene layout. */
GDPT_API int  gdpt_shader_finish_create_uniforms(gdpt_shader *shader);
/* Replaces ComputeShader::check_ready (gdcs.cpp:258-273): 1 ready, 0 not. */
GDPT_API int  gdpt_shader_check_ready(const gdpt_shader *shader);
/* Replaces ComputeShader::compute(groups) (gdcs.cpp:237-251): launch + block
 * until done (submit(); sync();).  The group counts are accepted for signature
 * compatibility and validated against ceil(W/32) x ceil(H/32) x 1
 * (path_tracing_camera.cpp:204); the CUDA grid is chosen by the backend. */
GDPT_API int  gdpt_shader_compute(gdpt_shader *shader, int groups_x, int groups_y, int groups_z);

/* ---------------------------------------------------------- extensions (ours)
 * Not in gdcs; they exist so the same kernels can be driven without the
 * per-dispatch submit+sync and full-frame round trips of the reference. */

/* One frame of PathTracingCamera::render() (path_tracing_camera.cpp:193-232) in
 * one call: camera upload (:200), K1 (:204), the selected post process
 * (:206-226; `progressive` is the post-process shader of `mode`: the progressive
 * one, the temporal-reprojection one, or NULL for GDPT_DENOISE_NONE) and the read-back
 * (:228-229) into caller memory.  frame_count is the value the host policy of
 * progressive_rendering.cpp:53-60 computed; in temporal mode it is ignored and the
 * kernel runs with the Params block last uploaded or staged (gdpt_shader_stage_params).  out_rgba8 (W*H*4 B) may be pinned
 * or pageable host memory; out_depth (W*H floats) may be NULL. Blocking. */
GDPT_API int  gdpt_render_frame(gdpt_shader *main_shader, gdpt_shader *progressive,
                                 const gdpt_camera *camera, gdpt_denoising mode,
                                 uint32_t frame_count, void *out_rgba8, float *out_depth);

/* Same work, but nothing leaves the GPU: enqueue on the device stream and
 * return.  Used with device-resident presentation (multi-GPU tile gather) and
 * by the benchmark's device-timed leg. */
GDPT_API int  gdpt_render_frame_async(gdpt_shader *main_shader, gdpt_shader *progressive,
                                 const gdpt_camera *camera, gdpt_denoising mode,
                                 uint32_t frame_count);
GDPT_API int  gdpt_device_synchronize(gdpt_device *device);

/* Pipelined form of gdpt_render_frame: `begin` enqueues the camera upload, K1, the post process AND the
 * read-back of the finished frame into out_rgba8 / out_depth (page-locked memory from gdpt_host_alloc
 * keeps the copy asynchronous), then returns without waiting.  The read-back runs on a second stream from
 * a device-side copy of the images, so the next frame's kernels overlap it; consecutive progressive / un-denoised
 * frames also overlap each other (two compute streams, frame-private K1 images, lists and counters; only the
 * accumulate/tone-map steps are ordered), which fills the SMs a frame's last long paths leave idle.  At most
 * GDPT_MAX_FRAMES_IN_FLIGHT frames may be in flight; `wait` blocks until the OLDEST of them is complete in its
 * caller buffers and reports its stats (out_stats may be NULL).  Results are byte-identical to gdpt_render_frame
 * called frame by frame. */
#define GDPT_MAX_FRAMES_IN_FLIGHT 4
struct gdpt_frame_stats;
GDPT_API int  gdpt_render_frame_begin(gdpt_shader *main_shader, gdpt_shader *progressive,
                                 const gdpt_camera *camera, gdpt_denoising mode, uint32_t frame_count,
                                 void *out_rgba8, float *out_depth);
GDPT_API int  gdpt_render_frame_wait(gdpt_shader *main_shader, struct gdpt_frame_stats *out_stats);

/* Non-blocking form of update_storage_buffer_uniform (gdcs.cpp:108-111) for the Params block
 * (set0 b0) of a post-process shader: the bytes are kept and uploaded, in stream order, by the
 * next gdpt_render_frame / _async / _begin call that runs this shader (TemporalReprojection::render
 * uploads its 88 B block every frame, temporal_reprojection.cpp:61-62).  size <= 256. */
GDPT_API int  gdpt_shader_stage_params(gdpt_shader *post_shader, const void *data, uint64_t size);

/* Restrict K1/K2 to image rows [row_begin,row_end) interleaved in bands:
 * a pixel row y is rendered iff ((y / band_rows) % n_parts) == part.  Used to
 * tile-shard a frame over several GPUs; (0,1) restores the full frame. */
GDPT_API int  gdpt_shader_set_shard(gdpt_shader *main_shader, int part, int n_parts, int band_rows);

/* Raw device address + byte size of a RID, for zero-copy hand-off to NCCL /
 * torch (the caller must not free it). */
GDPT_API int  gdpt_rid_device_pointer(gdpt_device *device, gdpt_rid rid,
                                 uint64_t *out_ptr, uint64_t *out_size);
/* K2 (progressive_rendering.glsl:28-46) on images the caller holds on this device (ours; upstream dispatches it only on
 * its own bound images): raw_rgba8 is one frame as K1 stored it, accum_rgba32f the accumulation, screen_rgba8 receives
 * the tone-mapped frame (may alias raw_rgba8), frame_count is Params.frame_count.  Enqueued on the device's stream.
 * What the sample-index partition needs to stay bit-identical with the sequential accumulation: the per-frame
 * images of all GPUs are accumulated in frame order, each GPU taking a block of rows (multigpu.SampleIndexAccumulator). */
GDPT_API int  gdpt_progressive_accumulate(gdpt_device *device, uint64_t raw_rgba8, uint64_t screen_rgba8,
                                          uint64_t accum_rgba32f, int width, int height, uint32_t frame_count);
/* Row-band frames without a gather (ours; upstream is single-GPU).  A frame sharded with gdpt_shader_set_shard is
 * rendered by several GPUs, one process each; the presented image is the union of their bands.  Instead of an
 * all-gather after the frame, the accumulate/tone-map kernel of `progressive_shader` stores every pixel it owns into
 * the RGBA8 image of each peer as well (peer memory over NVLink), so every GPU's image holds the whole frame once all
 * GPUs have finished the dispatch (the caller separates frames with a barrier).  Images cross the process boundary
 * through CUDA IPC: export the handle of the screen RID (set 0 binding 0 of main.glsl, gdcs.cpp:122-131), open the
 * peers' handles, hand the addresses to the progressive shader.  n = 0 clears.  Pointers of images on the same
 * device are accepted too (single-process tests). */
#define GDPT_IPC_HANDLE_BYTES 64
/* A plain device buffer owned by the device object (zero-filled, no host copy): frames a process keeps for the other
 * per-GPU processes to read over NVLink (gdpt_rid_ipc_export + gdpt_rid_device_pointer; multigpu.PeerFrameStore).
 * Returns 0 on failure.  Freed by gdpt_device_free_buffer or with the device. */
GDPT_API gdpt_rid gdpt_device_create_buffer(gdpt_device *device, uint64_t size);
GDPT_API int  gdpt_device_free_buffer(gdpt_device *device, gdpt_rid rid);
GDPT_API int  gdpt_rid_ipc_export(gdpt_device *device, gdpt_rid rid, void *out_handle);
GDPT_API int  gdpt_device_ipc_open(gdpt_device *device, const void *handle, uint64_t *out_ptr);
GDPT_API int  gdpt_device_ipc_close(gdpt_device *device, uint64_t ptr);
GDPT_API int  gdpt_shader_set_peer_screens(gdpt_shader *progressive_shader, const uint64_t *ptrs, int n);
/* Page-locked host memory for the read-back target (stands in for the
 * PackedByteArray that texture_get_data returns, gdcs.cpp:169-172): D2H copies
 * into it run at full PCIe rate.  Free with gdpt_host_free. */
GDPT_API void *gdpt_host_alloc(uint64_t size);
GDPT_API void  gdpt_host_free(void *ptr);
/* cudaStream_t of the device as an integer, so a caller can order its own work. */
GDPT_API uint64_t gdpt_device_stream(gdpt_device *device);

typedef struct gdpt_frame_stats {
    uint64_t rays;          /* ray_trace() calls of the last K1 dispatch (main.glsl:352) */
    uint64_t primary_hits;
    uint64_t node_pops;     /* filled only under GDPT_TRACE */
    uint64_t box_tests;
    uint64_t tri_tests;
    uint64_t tlas_leaves;
    uint32_t kernel_launches; /* kernels launched by the last dispatch */
    uint32_t max_stack;
    float    k1_ms;         /* CUDA-event time of the last K1 dispatch */
    float    k2_ms;
    uint64_t retraced;      /* rays the closest-hit search handed to the exact reference-order traversal */
    /* own work of the closest-hit search, filled only under "#define GDPT_COUNT_WORK": what the rendering kernel itself
     * executed (four-wide node steps, slab tests = 4 per step, triangle tests, instance entries, proofs) */
    uint64_t own_node_steps, own_box_tests, own_tri_tests, own_inst_entries, own_proofs;
} gdpt_frame_stats;
GDPT_API int  gdpt_shader_get_stats(gdpt_shader *main_shader, gdpt_frame_stats *out);
/* Which kernel schedule finish_create_uniforms chose for this main shader (2, 3 or 6: see GDPT_VARIANT above); -1 before. */
GDPT_API int  gdpt_shader_get_schedule(const gdpt_shader *main_shader);

/* Profiling aid: when on, an event is recorded between the stage kernels of every K1
 * dispatch; get_stage_times returns the number of stages and their device times in launch
 * order: primary, shade(0), trace(1), shade(1), ..., shade(D-1).  Blocks until the frame is done. */
GDPT_API int  gdpt_shader_set_stage_timing(gdpt_shader *main_shader, int on);
GDPT_API int  gdpt_shader_get_stage_times(gdpt_shader *main_shader, float *out_ms, int capacity);

/* Profiling aid: when on, the path kernel of every K1 dispatch then
 * leaves one 8 x uint64 record per warp: globaltimer ns at start and end, the number of scheduler
 * iterations it spent on internal nodes, triangle tests, instance entries/exits, shading and
 * refilling, and the number of paths it started.  read_warp_profile returns the number of warps
 * (or a negative gdpt_status) and blocks until the frame is done. */
GDPT_API int  gdpt_shader_set_warp_profile(gdpt_shader *main_shader, int on);
GDPT_API int64_t gdpt_shader_read_warp_profile(gdpt_shader *main_shader, uint64_t *out, uint64_t capacity_words);

/* Under "#define GDPT_TRACE": per-pixel parity record of path segment `segment`
 * (0 = primary ray) of the last K1 dispatch; capacity in records (W*H needed).
 * Pixels whose path ended before `segment` have hit = 0xFFFFFFFF. */
GDPT_API int  gdpt_shader_read_trace(gdpt_shader *main_shader, int segment,
                                 gdpt_trace_record *out, uint64_t capacity);
/* Under GDPT_TRACE: the first `max_per_ray` popped node ids of every primary ray
 * (TLAS pops carry GDPT_VISIT_TLAS_TAG), row-major [pixel][max_per_ray]. */
GDPT_API int  gdpt_shader_read_visits(gdpt_shader *main_shader, uint32_t *out,
                                 uint32_t max_per_ray, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* GDPT_H */
