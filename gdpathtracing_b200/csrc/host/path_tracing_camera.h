// path_tracing_camera.h -- Godot-free twins of the reference's PathTracingCamera
// node (src/path_tracing/path_tracing_camera.{h,cpp}), its Camera uniform block
// (src/path_tracing/render_parameters.h:14-38) and the ProgressiveRendering post
// process driver (src/path_tracing/post_processing/progressive_rendering.{h,cpp}).
// Same properties, same init()/render() call sequence against ComputeShader;
// the engine services a Node gets from Godot (window size, global transform,
// the TextureRect that shows the frame) are explicit setters/getters here.
#ifndef GDPT_PATH_TRACING_CAMERA_H
#define GDPT_PATH_TRACING_CAMERA_H

#include "compute_shader.h"
#include "geometry_group3d.h"
#include "xform_math.h"

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace gdpt {

// Camera::set_camera_transform (render_parameters.h:23-38)
struct CameraBlock : gdpt_camera {
    CameraBlock();
    void set_camera_transform(const Xform3 &model, const Mat4 &projection);
};

class ProgressiveRendering {
public:
    ProgressiveRendering() {}
    ~ProgressiveRendering();
    void init(gdpt_device *rd, gdpt_rid original_screen_texture_rid, int width, int height);
    // frame_count policy of progressive_rendering.cpp:53-60; returns the value dispatched with
    uint32_t advance(const Xform3 &camera_transform);
    void render(const Xform3 &camera_transform);
    ComputeShader *shader() const { return cs_; }
    uint32_t frame_count() const { return params_.frame_count; }
    gdpt_rid frame_buffer_rid() const { return frame_buffer_rid_; }

private:
    ComputeShader *cs_ = nullptr;
    gdpt_progressive_params params_ = { 0, 0, 1 };
    Xform3 previous_transform_; // identity, like a default-constructed Transform3D
    gdpt_rid params_rid_ = 0, screen_rid_ = 0, frame_buffer_rid_ = 0;
};

// post_processing/temporal_reprojection.{h,cpp}
class TemporalReprojection {
public:
    TemporalReprojection();
    ~TemporalReprojection();
    void init(gdpt_device *rd, gdpt_rid original_screen_texture_rid, gdpt_rid original_depth_texture_rid, int width, int height);
    // parameter update of temporal_reprojection.cpp:57-61 (delta matrix, previous_vp, ++frame_count); returns the block dispatched with
    const gdpt_temporal_params &advance(const Xform3 &view_matrix, const Mat4 &projection_matrix);
    void render(const Xform3 &view_matrix, const Mat4 &projection_matrix);
    ComputeShader *shader() const { return cs_; }
    const gdpt_temporal_params &params() const { return params_; }
    gdpt_rid frame_buffer_rid(int which) const { return which == 0 ? frame_buffer_rid_1_ : frame_buffer_rid_2_; }

private:
    ComputeShader *cs_ = nullptr;
    gdpt_temporal_params params_;
    Mat4 previous_vp_; // identity, like a default-constructed Projection
    gdpt_rid params_rid_ = 0, screen_rid_ = 0, frame_buffer_rid_1_ = 0, frame_buffer_rid_2_ = 0;
};

class PathTracingCamera {
public:
    enum Denoising { PROGRESSIVE_RENDERING = 0, TEMPORAL_REPROJECTION = 1, NONE = 2 };

    PathTracingCamera() {}
    ~PathTracingCamera();

    // --- reference properties (path_tracing_camera.cpp:3-31) ---
    float get_fov() const { return fov_; }
    void set_fov(float v) { fov_ = v; }
    GeometryGroup3D *get_geometry_group() const { return geometry_group_; }
    void set_geometry_group(GeometryGroup3D *g) { geometry_group_ = g; }
    Denoising get_denoising_mode() const { return denoising_mode_; }
    void set_denoising_mode(Denoising m) { denoising_mode_ = m; }
    // output_texture: the frame lands in get_output_image() instead of a TextureRect
    const uint8_t *get_output_image() const { return output_image_; }

    // --- what Godot supplies to a node ---
    void set_window_size(int w, int h) { window_w_ = w; window_h_ = h; }   // DisplayServer::window_get_size()
    void set_global_transform(const Xform3 &t) { global_transform_ = t; }   // Node3D::get_global_transform()

    // --- backend parameters (no upstream counterpart) ---
    void set_max_depth(int d) { max_depth_ = d; }          // the disabled num_bounces property (:9-11); 5 = main.glsl:377
    void set_cuda_device(int ordinal) { cuda_ordinal_ = ordinal; }
    void set_frame_index(uint32_t f) { camera_.frame_index = f; } // upstream leaves it uninitialised (render_parameters.h:19)
    void set_shard(int part, int parts, int band_rows) { shard_part_ = part; shard_parts_ = parts; shard_band_ = band_rows; }
    void set_trace(int segments, uint32_t visits_per_ray) { trace_segments_ = segments; visits_per_ray_ = visits_per_ray; }
    void set_debug_steps(bool on) { debug_steps_ = on; }
    // -1 backend default (cull when rendering, reference order when tracing), 0 reference order, 1 cull
    void set_cull(int mode) { cull_ = mode; }
    // hit records (no work counters) of the first n segments, written by the RENDERING kernels (parity check of the fast path)
    void set_record_hits(int segments) { record_hits_ = segments; }
    // kernel schedule ("#define GDPT_VARIANT n", include/gdpt.h); -1 = backend default.  Results do not depend on it.
    void set_variant(int variant) { variant_ = variant; }
    // scheduling knob of the path kernels ("#define GDPT_TUNE_<NAME> n", include/gdpt.h), for A/B measurements only
    void set_tuning(const std::string &name, int value) { tuning_[name] = value; }
    // "#define GDPT_COUNT_WORK": the path kernel also counts its own work (gdpt_frame_stats own_*); measurement aid
    void set_count_work(bool on) { count_work_ = on; }
    // true: one gdpt_render_frame call per frame; false: the reference's dispatch-by-dispatch sequence
    void set_fused_frame(bool on) { fused_frame_ = on; }

    // NOTIFICATION_READY / NOTIFICATION_INTERNAL_PROCESS bodies (path_tracing_camera.cpp:111-232)
    bool init();
    void render();
    // render without the device->host read-back (frame stays in the output RID)
    void render_device_only();
    // pipelined render(): begin() enqueues a frame including its read-back and returns; wait() blocks until the
    // oldest frame in flight (at most two) is in host memory and returns it (nullptr if none / on error).
    // Frame N's read-back overlaps frame N+1's kernels.
    bool render_begin();
    const uint8_t *render_wait(gdpt_frame_stats *stats = nullptr);

    // upstream creates the post-process object inside the first render() (path_tracing_camera.cpp:207-223); callers that
    // need its shader before the first frame (peer screens of a row-band frame) create it here, state untouched
    void prepare_post()
    {
        if (cs_ == nullptr) return;
        if (denoising_mode_ == PROGRESSIVE_RENDERING) ensure_progressive();
        else if (denoising_mode_ == TEMPORAL_REPROJECTION) ensure_temporal();
    }

    ComputeShader *compute_shader() const { return cs_; }
    ProgressiveRendering *progressive() const { return progressive_renderer_; }
    TemporalReprojection *temporal() const { return temporal_reprojection_; }
    gdpt_rid output_texture_rid() const { return output_texture_rid_; }
    gdpt_rid depth_texture_rid() const { return depth_texture_rid_; }
    int width() const { return render_parameters_.width; }
    int height() const { return render_parameters_.height; }
    const gdpt_camera &camera_block() const { return camera_; }
    uint32_t last_frame_count() const { return last_frame_count_; }

private:
    void ensure_progressive();
    void ensure_temporal();
    // post-process bookkeeping of one frame (progressive_rendering.cpp:53-60 / temporal_reprojection.cpp:57-61);
    // returns the shader to run after K1 (nullptr for NONE) and sets *frame_count
    gdpt_shader *advance_post(uint32_t *frame_count);
    float fov_ = 90.0f;
    ComputeShader *cs_ = nullptr;
    ProgressiveRendering *progressive_renderer_ = nullptr;
    TemporalReprojection *temporal_reprojection_ = nullptr;
    GeometryGroup3D *geometry_group_ = nullptr;
    uint8_t *output_image_ = nullptr; // pinned, W*H*4
    uint8_t *pipeline_image_[GDPT_MAX_FRAMES_IN_FLIGHT] = {}; // pinned targets of the frames in flight
    unsigned pipe_head_ = 0, pipe_tail_ = 0;
    gdpt_render_params render_parameters_ = {};
    CameraBlock camera_;
    Mat4 projection_matrix_;
    Xform3 global_transform_;
    int window_w_ = 0, window_h_ = 0;
    gdpt_rid output_texture_rid_ = 0, depth_texture_rid_ = 0, render_parameters_rid_ = 0, camera_rid_ = 0;
    gdpt_rid triangles_geometry_rid_ = 0, triangles_data_rid_ = 0, materials_rid_ = 0, bvh_tree_rid_ = 0, blas_rid_ = 0, tlas_rid_ = 0,
             texture_array_rid_ = 0;
    gdpt_device *rd_ = nullptr;
    Denoising denoising_mode_ = PROGRESSIVE_RENDERING;
    int max_depth_ = 5, cuda_ordinal_ = 0;
    int shard_part_ = 0, shard_parts_ = 1, shard_band_ = 4;
    int trace_segments_ = 0; uint32_t visits_per_ray_ = 0;
    bool debug_steps_ = false, fused_frame_ = true;
    int cull_ = -1;
    int record_hits_ = 0;
    int variant_ = -1;
    std::map<std::string, int> tuning_;
    bool count_work_ = false;
    uint32_t last_frame_count_ = 0;
};

} // namespace gdpt
#endif
