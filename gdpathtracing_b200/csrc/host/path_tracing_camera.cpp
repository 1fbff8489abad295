// path_tracing_camera.cpp -- see path_tracing_camera.h.
#include "path_tracing_camera.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace gdpt {

CameraBlock::CameraBlock()
{
    std::memset(static_cast<gdpt_camera *>(this), 0, sizeof(gdpt_camera));
    z_near = 0.01f; // render_parameters.h:20-21
    z_far = 1000.0f;
    frame_index = 0; // indeterminate upstream (render_parameters.h:19); the first render() makes it 1
}

void CameraBlock::set_camera_transform(const Xform3 &model, const Mat4 &projection)
{
    position[0] = model.origin.x; position[1] = model.origin.y; position[2] = model.origin.z; position[3] = 1.0f;
    const Mat4 t = projection * Mat4(model.affine_inverse());
    t.to_float16(vp);
    t.inverse().to_float16(ivp);
}

// ------------------------------------------------------------ ProgressiveRendering

ProgressiveRendering::~ProgressiveRendering() { delete cs_; }

void ProgressiveRendering::init(gdpt_device *rd, gdpt_rid original_screen_texture_rid, int width, int height)
{
    screen_rid_ = original_screen_texture_rid;
    params_.width = width; params_.height = height; params_.frame_count = 1;
    cs_ = new ComputeShader("res://addons/jar_path_tracing/src/shaders/progressive_rendering.glsl", rd);
    params_rid_ = cs_->create_storage_buffer_uniform(&params_, sizeof(params_), 0, 0);
    cs_->add_existing_buffer(screen_rid_, GDPT_UNIFORM_TYPE_IMAGE, 1, 0);
    frame_buffer_rid_ = cs_->create_image_uniform(nullptr, width, height, GDPT_FORMAT_R32G32B32A32_SFLOAT, 2, 0);
    cs_->finish_create_uniforms();
}

uint32_t ProgressiveRendering::advance(const Xform3 &camera_transform)
{
    const bool camera_moved = !previous_transform_.is_equal_approx(camera_transform);
    previous_transform_ = camera_transform;
    if (camera_moved) params_.frame_count = 1;
    else params_.frame_count++;
    return params_.frame_count;
}

void ProgressiveRendering::render(const Xform3 &camera_transform)
{
    if (cs_ == nullptr || !cs_->check_ready()) return;
    advance(camera_transform);
    cs_->update_storage_buffer_uniform(params_rid_, &params_, sizeof(params_));
    cs_->compute((int)std::ceil(params_.width / 32.0f), (int)std::ceil(params_.height / 32.0f), 1);
}

// ------------------------------------------------------------ TemporalReprojection

TemporalReprojection::TemporalReprojection()
{
    // deltaMatrix is uninitialised upstream until the first render() fills it (temporal_reprojection.h:17); zeros here
    std::memset(&params_, 0, sizeof(params_));
    params_.blend_factor = 0.75f; // temporal_reprojection.h:21-23; none of the three is read by the shader
    params_.near_plane = 0.01f;
    params_.far_plane = 1000.0f;
}

TemporalReprojection::~TemporalReprojection() { delete cs_; }

void TemporalReprojection::init(gdpt_device *rd, gdpt_rid original_screen_texture_rid, gdpt_rid original_depth_texture_rid, int width,
                                int height)
{
    screen_rid_ = original_screen_texture_rid;
    params_.width = width; params_.height = height; params_.frame_count = 1;
    cs_ = new ComputeShader("res://addons/jar_path_tracing/src/shaders/temporal_reprojection.glsl", rd);
    params_rid_ = cs_->create_storage_buffer_uniform(&params_, sizeof(params_), 0, 0);
    cs_->add_existing_buffer(screen_rid_, GDPT_UNIFORM_TYPE_IMAGE, 1, 0);
    cs_->add_existing_buffer(original_depth_texture_rid, GDPT_UNIFORM_TYPE_IMAGE, 2, 0);
    // Image::create(..., FORMAT_RGBAF) is zero-filled (temporal_reprojection.cpp:41-46)
    frame_buffer_rid_1_ = cs_->create_image_uniform(nullptr, width, height, GDPT_FORMAT_R32G32B32A32_SFLOAT, 3, 0);
    frame_buffer_rid_2_ = cs_->create_image_uniform(nullptr, width, height, GDPT_FORMAT_R32G32B32A32_SFLOAT, 4, 0);
    cs_->finish_create_uniforms();
}

const gdpt_temporal_params &TemporalReprojection::advance(const Xform3 &view_matrix, const Mat4 &projection_matrix)
{
    const Mat4 vp = projection_matrix * Mat4(view_matrix);
    // `Transform3D deltaMatrix = previous_vp * vp.inverse();` keeps the affine part only (temporal_reprojection.cpp:58)
    const Mat4 delta = (previous_vp_ * vp.inverse()).affine_part();
    previous_vp_ = vp;
    params_.frame_count++;
    delta.to_float16(params_.delta_matrix);
    return params_;
}

void TemporalReprojection::render(const Xform3 &view_matrix, const Mat4 &projection_matrix)
{
    if (cs_ == nullptr || !cs_->check_ready()) return;
    advance(view_matrix, projection_matrix);
    cs_->update_storage_buffer_uniform(params_rid_, &params_, sizeof(params_));
    cs_->compute((int)std::ceil(params_.width / 32.0f), (int)std::ceil(params_.height / 32.0f), 1);
}

// ------------------------------------------------------------ PathTracingCamera

PathTracingCamera::~PathTracingCamera()
{
    delete progressive_renderer_;
    delete temporal_reprojection_;
    delete cs_;
    if (output_image_) gdpt_host_free(output_image_);
    for (int i = 1; i < GDPT_MAX_FRAMES_IN_FLIGHT; i++) // [0] is output_image_
        if (pipeline_image_[i]) gdpt_host_free(pipeline_image_[i]);
    if (rd_) gdpt_device_destroy(rd_);
}

bool PathTracingCamera::init()
{
    // one device for every shader of this camera (path_tracing_camera.cpp:113-114)
    if (gdpt_device_create(cuda_ordinal_, &rd_) != GDPT_OK) {
        std::fprintf(stderr, "Failed to create rendering device: %s\n", gdpt_last_error(nullptr));
        rd_ = nullptr;
        return false;
    }
    if (geometry_group_ == nullptr) {
        std::fprintf(stderr, "No geometry group set.\n");
        return false;
    }
    if (window_w_ <= 0 || window_h_ <= 0) {
        std::fprintf(stderr, "No window size set.\n");
        return false;
    }
    geometry_group_->build();

    render_parameters_.width = window_w_;
    render_parameters_.height = window_h_;
    render_parameters_.fov = fov_;
    render_parameters_.triangle_count = (uint32_t)geometry_group_->get_triangle_count();
    render_parameters_.blas_count = (uint32_t)geometry_group_->get_blas_count();
    projection_matrix_ = Mat4::perspective(fov_, static_cast<float>(window_w_) / window_h_, 0.01f, 1000.0f);
    camera_.set_camera_transform(global_transform_, projection_matrix_);

    std::vector<std::string> defines = { "#define TESTe" }; // the dummy define upstream passes (:139)
    defines.push_back("#define MAX_DEPTH " + std::to_string(max_depth_));
    if (debug_steps_) defines.push_back("#define DEBUG_STEPS");
    if (trace_segments_ > 0) defines.push_back("#define GDPT_TRACE " + std::to_string(trace_segments_));
    if (visits_per_ray_ > 0) defines.push_back("#define GDPT_TRACE_VISITS " + std::to_string(visits_per_ray_));
    if (cull_ >= 0) defines.push_back("#define GDPT_CULL " + std::to_string(cull_));
    if (variant_ >= 0) defines.push_back("#define GDPT_VARIANT " + std::to_string(variant_));
    if (record_hits_ > 0) defines.push_back("#define GDPT_RECORD_HITS " + std::to_string(record_hits_));
    if (count_work_) defines.push_back("#define GDPT_COUNT_WORK");
    if (geometry_group_ && geometry_group_->get_material_ext()) defines.push_back("#define GDPT_MATERIAL_EXT");
    for (const auto &kv : tuning_) defines.push_back("#define GDPT_TUNE_" + kv.first + " " + std::to_string(kv.second));
    cs_ = new ComputeShader("res://addons/jar_path_tracing/src/shaders/main.glsl", rd_, defines);

    render_parameters_rid_ = cs_->create_storage_buffer_uniform(&render_parameters_, sizeof(render_parameters_), 2, 0);
    camera_rid_ = cs_->create_storage_buffer_uniform(static_cast<gdpt_camera *>(&camera_), sizeof(gdpt_camera), 3, 0);

    const size_t out_bytes = (size_t)window_w_ * window_h_ * 4;
    output_image_ = static_cast<uint8_t *>(gdpt_host_alloc(out_bytes));
    if (!output_image_) {
        std::fprintf(stderr, "No output texture set.\n");
        return false;
    }
    std::memset(output_image_, 0, out_bytes);
    output_texture_rid_ = cs_->create_image_uniform(output_image_, window_w_, window_h_, GDPT_FORMAT_R8G8B8A8_UNORM, 0, 0);
    depth_texture_rid_ = cs_->create_image_uniform(nullptr, window_w_, window_h_, GDPT_FORMAT_R32_SFLOAT, 1, 0);

    const auto &tg = geometry_group_->get_triangles_geometry_buffer();
    const auto &td = geometry_group_->get_triangles_data_buffer();
    const auto &mt = geometry_group_->get_materials_buffer();
    const auto &bv = geometry_group_->get_bvh_buffer();
    const auto &bl = geometry_group_->get_blas_buffer();
    const auto &tl = geometry_group_->get_tlas_buffer();
    triangles_geometry_rid_ = cs_->create_storage_buffer_uniform(tg.data(), tg.size() * sizeof(tg[0]), 0, 1);
    triangles_data_rid_ = cs_->create_storage_buffer_uniform(td.data(), td.size() * sizeof(td[0]), 1, 1);
    materials_rid_ = cs_->create_storage_buffer_uniform(mt.data(), mt.size() * sizeof(mt[0]), 2, 1);
    bvh_tree_rid_ = cs_->create_storage_buffer_uniform(bv.data(), bv.size() * sizeof(bv[0]), 3, 1);
    blas_rid_ = cs_->create_storage_buffer_uniform(bl.data(), bl.size() * sizeof(bl[0]), 4, 1);
    tlas_rid_ = cs_->create_storage_buffer_uniform(tl.data(), tl.size() * sizeof(tl[0]), 5, 1);
    if (geometry_group_->get_material_ext()) { // material table for any number of surfaces per instance (gdpt_wire.h)
        const auto &sm = geometry_group_->get_surface_materials_buffer();
        cs_->create_storage_buffer_uniform(sm.data(), sm.size() * sizeof(sm[0]), 6, 1);
    }

    std::vector<const void *> layers;
    for (const auto &layer : geometry_group_->get_textures_buffer()) layers.push_back(layer.data());
    const int res = geometry_group_->get_texture_array_resolution();
    texture_array_rid_ = cs_->create_layered_image_uniform(layers, res, res, GDPT_FORMAT_R8G8B8A8_UNORM, 0, 2);

    cs_->finish_create_uniforms();
    if (cs_->check_ready() && shard_parts_ > 1 &&
        gdpt_shader_set_shard(cs_->handle(), shard_part_, shard_parts_, shard_band_) != GDPT_OK)
        std::fprintf(stderr, "set_shard: %s\n", gdpt_last_error(rd_));
    return cs_->check_ready();
}

void PathTracingCamera::ensure_progressive()
{
    if (progressive_renderer_ == nullptr) {
        progressive_renderer_ = new ProgressiveRendering();
        progressive_renderer_->init(rd_, output_texture_rid_, render_parameters_.width, render_parameters_.height);
    }
}

void PathTracingCamera::ensure_temporal()
{
    if (temporal_reprojection_ == nullptr) {
        temporal_reprojection_ = new TemporalReprojection();
        temporal_reprojection_->init(rd_, output_texture_rid_, depth_texture_rid_, render_parameters_.width, render_parameters_.height);
    }
}

gdpt_shader *PathTracingCamera::advance_post(uint32_t *frame_count)
{
    *frame_count = 0;
    if (denoising_mode_ == PROGRESSIVE_RENDERING) {
        ensure_progressive();
        *frame_count = progressive_renderer_->advance(global_transform_);
        return progressive_renderer_->shader()->handle();
    }
    if (denoising_mode_ == TEMPORAL_REPROJECTION) {
        ensure_temporal();
        const gdpt_temporal_params &tp = temporal_reprojection_->advance(global_transform_.affine_inverse(), projection_matrix_);
        *frame_count = tp.frame_count;
        gdpt_shader *h = temporal_reprojection_->shader()->handle();
        if (gdpt_shader_stage_params(h, &tp, sizeof(tp)) != GDPT_OK) std::fprintf(stderr, "stage_params: %s\n", gdpt_last_error(rd_));
        return h;
    }
    return nullptr;
}

void PathTracingCamera::render_device_only()
{
    if (cs_ == nullptr || !cs_->check_ready()) return;
    camera_.set_camera_transform(global_transform_, projection_matrix_);
    camera_.frame_index++;
    uint32_t frame_count = 0;
    gdpt_shader *prog = advance_post(&frame_count);
    last_frame_count_ = frame_count;
    if (gdpt_render_frame_async(cs_->handle(), prog, &camera_, (gdpt_denoising)denoising_mode_, frame_count) != GDPT_OK)
        std::fprintf(stderr, "render_frame_async: %s\n", gdpt_last_error(rd_));
}

bool PathTracingCamera::render_begin()
{
    if (cs_ == nullptr || !cs_->check_ready()) return false;
    if (pipe_head_ - pipe_tail_ >= GDPT_MAX_FRAMES_IN_FLIGHT) {
        std::fprintf(stderr, "render_begin: %d frames are already in flight\n", GDPT_MAX_FRAMES_IN_FLIGHT);
        return false;
    }
    if (!pipeline_image_[1]) {
        pipeline_image_[0] = output_image_;
        for (int i = 1; i < GDPT_MAX_FRAMES_IN_FLIGHT; i++) {
            pipeline_image_[i] = static_cast<uint8_t *>(gdpt_host_alloc((size_t)window_w_ * window_h_ * 4));
            if (!pipeline_image_[i]) return false;
        }
    }
    camera_.set_camera_transform(global_transform_, projection_matrix_);
    camera_.frame_index++;
    uint32_t frame_count = 0;
    gdpt_shader *prog = advance_post(&frame_count);
    last_frame_count_ = frame_count;
    if (gdpt_render_frame_begin(cs_->handle(), prog, &camera_, (gdpt_denoising)denoising_mode_, frame_count,
                                pipeline_image_[pipe_head_ % GDPT_MAX_FRAMES_IN_FLIGHT], nullptr) != GDPT_OK) {
        std::fprintf(stderr, "render_frame_begin: %s\n", gdpt_last_error(rd_));
        return false;
    }
    pipe_head_++;
    return true;
}

const uint8_t *PathTracingCamera::render_wait(gdpt_frame_stats *stats)
{
    if (cs_ == nullptr || pipe_head_ == pipe_tail_) return nullptr;
    if (gdpt_render_frame_wait(cs_->handle(), stats) != GDPT_OK) {
        std::fprintf(stderr, "render_frame_wait: %s\n", gdpt_last_error(rd_));
        pipe_tail_++;
        return nullptr;
    }
    return pipeline_image_[pipe_tail_++ % GDPT_MAX_FRAMES_IN_FLIGHT];
}

void PathTracingCamera::render()
{
    if (cs_ == nullptr || !cs_->check_ready()) return;
    const int W = render_parameters_.width, H = render_parameters_.height;
    if (fused_frame_) {
        camera_.set_camera_transform(global_transform_, projection_matrix_);
        camera_.frame_index++;
        uint32_t frame_count = 0;
        gdpt_shader *prog = advance_post(&frame_count);
        last_frame_count_ = frame_count;
        if (gdpt_render_frame(cs_->handle(), prog, &camera_, (gdpt_denoising)denoising_mode_, frame_count, output_image_, nullptr) != GDPT_OK)
            std::fprintf(stderr, "render_frame: %s\n", gdpt_last_error(rd_));
        return;
    }
    // dispatch-by-dispatch, as path_tracing_camera.cpp:197-230
    camera_.set_camera_transform(global_transform_, projection_matrix_);
    camera_.frame_index++;
    cs_->update_storage_buffer_uniform(camera_rid_, static_cast<gdpt_camera *>(&camera_), sizeof(gdpt_camera));
    cs_->compute((int)std::ceil(W / 32.0f), (int)std::ceil(H / 32.0f), 1);
    switch (denoising_mode_) {
    case PROGRESSIVE_RENDERING:
        ensure_progressive();
        progressive_renderer_->render(global_transform_);
        last_frame_count_ = progressive_renderer_->frame_count();
        break;
    case TEMPORAL_REPROJECTION: // path_tracing_camera.cpp:215-221
        ensure_temporal();
        temporal_reprojection_->render(global_transform_.affine_inverse(), projection_matrix_);
        last_frame_count_ = temporal_reprojection_->params().frame_count;
        break;
    case NONE:
        break;
    }
    cs_->get_image_uniform_buffer_into(output_texture_rid_, output_image_, (uint64_t)W * H * 4);
}

} // namespace gdpt
