// host_capi.cpp -- flat C exports of the host layer (include/gdpt_host.h).
#include "gdpt_host.h"

#include "geometry_group3d.h"
#include "path_tracing_camera.h"

#include <cstring>
#include <vector>

struct gdpt_geometry_group { gdpt::GeometryGroup3D impl; };
struct gdpt_camera_node { gdpt::PathTracingCamera impl; };

extern "C" {

gdpt_geometry_group *gdpt_group_create(void) { return new gdpt_geometry_group(); }
void gdpt_group_destroy(gdpt_geometry_group *g) { delete g; }

int gdpt_group_add_texture(gdpt_geometry_group *g, const uint8_t *rgba8, int width, int height)
{
    return g->impl.add_texture(rgba8, width, height);
}

int gdpt_group_add_material(gdpt_geometry_group *g, const gdpt_standard_material *m)
{
    gdpt::StandardMaterial s;
    std::memcpy(s.albedo, m->albedo, sizeof(s.albedo));
    s.metallic = m->metallic; s.roughness = m->roughness;
    std::memcpy(s.emission, m->emission, sizeof(s.emission));
    s.emission_energy_multiplier = m->emission_energy_multiplier;
    s.albedo_texture = m->albedo_texture;
    s.is_standard = m->is_standard != 0;
    return g->impl.add_material(s);
}

int gdpt_group_add_material_ext(gdpt_geometry_group *g, const gdpt_standard_material *m, int roughness_texture, int metallic_texture,
                                int albedo_srgb)
{
    gdpt::StandardMaterial s;
    std::memcpy(s.albedo, m->albedo, sizeof(s.albedo));
    s.metallic = m->metallic; s.roughness = m->roughness;
    std::memcpy(s.emission, m->emission, sizeof(s.emission));
    s.emission_energy_multiplier = m->emission_energy_multiplier;
    s.albedo_texture = m->albedo_texture;
    s.is_standard = m->is_standard != 0;
    s.roughness_texture = roughness_texture; s.metallic_texture = metallic_texture; s.albedo_srgb = albedo_srgb != 0;
    return g->impl.add_material(s);
}
void gdpt_group_set_material_ext(gdpt_geometry_group *g, int on) { g->impl.set_material_ext(on != 0); }
int gdpt_group_get_material_ext(const gdpt_geometry_group *g) { return g->impl.get_material_ext() ? 1 : 0; }

int gdpt_group_add_mesh(gdpt_geometry_group *g, int n_surfaces, const int32_t *vertex_counts, const int32_t *index_counts,
                        const float *positions, const float *normals, const float *uvs, const int32_t *indices)
{
    std::vector<gdpt::SurfaceArrays> surf((size_t)n_surfaces);
    int64_t voff = 0, ioff = 0;
    for (int s = 0; s < n_surfaces; s++) {
        surf[s].positions = positions + 3 * voff; surf[s].normals = normals + 3 * voff; surf[s].uvs = uvs + 2 * voff;
        surf[s].vertex_count = vertex_counts[s];
        surf[s].indices = indices + ioff; surf[s].index_count = index_counts[s];
        voff += vertex_counts[s]; ioff += index_counts[s];
    }
    return g->impl.add_mesh(surf.data(), n_surfaces);
}

void gdpt_group_add_mesh_instance(gdpt_geometry_group *g, int mesh, const float *transform12, int material_override,
                                  const int32_t *surface_overrides, int n_surface_overrides)
{
    std::vector<int> ov(surface_overrides, surface_overrides + (surface_overrides ? n_surface_overrides : 0));
    g->impl.add_mesh_instance(mesh, gdpt::Xform3::from_rows12(transform12), material_override, ov.data(), (int)ov.size());
}

void gdpt_group_set_default_material(gdpt_geometry_group *g, int material) { g->impl.set_default_material(material); }
void gdpt_group_set_texture_array_resolution(gdpt_geometry_group *g, int r) { g->impl.set_texture_array_resolution(r); }
int gdpt_group_get_texture_array_resolution(const gdpt_geometry_group *g) { return g->impl.get_texture_array_resolution(); }
void gdpt_group_set_build_threads(gdpt_geometry_group *g, int n) { g->impl.set_build_threads(n); }
int gdpt_group_get_build_threads(const gdpt_geometry_group *g) { return g->impl.get_build_threads(); }
void gdpt_group_build(gdpt_geometry_group *g) { g->impl.build(); }
double gdpt_group_last_build_seconds(const gdpt_geometry_group *g) { return g->impl.last_build_seconds(); }

uint64_t gdpt_group_buffer_size(const gdpt_geometry_group *g, int which)
{
    const gdpt::GeometryGroup3D &i = g->impl;
    switch (which) {
    case 0: return i.get_triangles_geometry_buffer().size() * sizeof(gdpt_triangle_geometry);
    case 1: return i.get_triangles_data_buffer().size() * sizeof(gdpt_triangle_data);
    case 2: return i.get_materials_buffer().size() * sizeof(gdpt_material);
    case 3: return i.get_bvh_buffer().size() * sizeof(gdpt_bvh_node);
    case 4: return i.get_blas_buffer().size() * sizeof(gdpt_blas_instance);
    case 5: return i.get_tlas_buffer().size() * sizeof(gdpt_tlas_node);
    case 6: return i.get_surface_materials_buffer().size() * sizeof(uint32_t);
    }
    return 0;
}

const void *gdpt_group_buffer_data(const gdpt_geometry_group *g, int which)
{
    const gdpt::GeometryGroup3D &i = g->impl;
    switch (which) {
    case 0: return i.get_triangles_geometry_buffer().data();
    case 1: return i.get_triangles_data_buffer().data();
    case 2: return i.get_materials_buffer().data();
    case 3: return i.get_bvh_buffer().data();
    case 4: return i.get_blas_buffer().data();
    case 5: return i.get_tlas_buffer().data();
    case 6: return i.get_surface_materials_buffer().data();
    }
    return nullptr;
}

int gdpt_group_texture_layer_count(const gdpt_geometry_group *g) { return (int)g->impl.get_textures_buffer().size(); }
const uint8_t *gdpt_group_texture_layer(const gdpt_geometry_group *g, int layer)
{
    const auto &t = g->impl.get_textures_buffer();
    return (layer >= 0 && layer < (int)t.size()) ? t[layer].data() : nullptr;
}

gdpt_camera_node *gdpt_camera_create(void) { return new gdpt_camera_node(); }
void gdpt_camera_destroy(gdpt_camera_node *c) { delete c; }
void gdpt_camera_set_fov(gdpt_camera_node *c, float fov) { c->impl.set_fov(fov); }
float gdpt_camera_get_fov(const gdpt_camera_node *c) { return c->impl.get_fov(); }
void gdpt_camera_set_geometry_group(gdpt_camera_node *c, gdpt_geometry_group *g) { c->impl.set_geometry_group(g ? &g->impl : nullptr); }
void gdpt_camera_set_denoising_mode(gdpt_camera_node *c, int mode) { c->impl.set_denoising_mode((gdpt::PathTracingCamera::Denoising)mode); }
int gdpt_camera_get_denoising_mode(const gdpt_camera_node *c) { return (int)c->impl.get_denoising_mode(); }
void gdpt_camera_set_window_size(gdpt_camera_node *c, int w, int h) { c->impl.set_window_size(w, h); }
void gdpt_camera_set_global_transform(gdpt_camera_node *c, const float *t) { c->impl.set_global_transform(gdpt::Xform3::from_rows12(t)); }
void gdpt_camera_set_max_depth(gdpt_camera_node *c, int d) { c->impl.set_max_depth(d); }
void gdpt_camera_set_cuda_device(gdpt_camera_node *c, int o) { c->impl.set_cuda_device(o); }
void gdpt_camera_set_frame_index(gdpt_camera_node *c, uint32_t f) { c->impl.set_frame_index(f); }
void gdpt_camera_set_shard(gdpt_camera_node *c, int part, int parts, int band) { c->impl.set_shard(part, parts, band); }
void gdpt_camera_set_trace(gdpt_camera_node *c, int segments, uint32_t visits) { c->impl.set_trace(segments, visits); }
void gdpt_camera_set_debug_steps(gdpt_camera_node *c, int on) { c->impl.set_debug_steps(on != 0); }
void gdpt_camera_set_cull(gdpt_camera_node *c, int mode) { c->impl.set_cull(mode); }
void gdpt_camera_set_variant(gdpt_camera_node *c, int variant) { c->impl.set_variant(variant); }
void gdpt_camera_set_count_work(gdpt_camera_node *c, int on) { c->impl.set_count_work(on != 0); }
void gdpt_camera_set_tuning(gdpt_camera_node *c, const char *name, int value) { if (name) c->impl.set_tuning(name, value); }
void gdpt_camera_set_record_hits(gdpt_camera_node *c, int segments) { c->impl.set_record_hits(segments); }
void gdpt_camera_set_fused_frame(gdpt_camera_node *c, int on) { c->impl.set_fused_frame(on != 0); }
int gdpt_camera_init(gdpt_camera_node *c) { return c->impl.init() ? 1 : 0; }
void gdpt_camera_render(gdpt_camera_node *c) { c->impl.render(); }
void gdpt_camera_render_device_only(gdpt_camera_node *c) { c->impl.render_device_only(); }
int gdpt_camera_render_begin(gdpt_camera_node *c) { return c->impl.render_begin() ? 1 : 0; }
const uint8_t *gdpt_camera_render_wait(gdpt_camera_node *c, gdpt_frame_stats *stats) { return c->impl.render_wait(stats); }
const uint8_t *gdpt_camera_output_image(const gdpt_camera_node *c) { return c->impl.get_output_image(); }
gdpt_shader *gdpt_camera_main_shader(const gdpt_camera_node *c) { return c->impl.compute_shader() ? c->impl.compute_shader()->handle() : nullptr; }
gdpt_shader *gdpt_camera_progressive_shader(const gdpt_camera_node *c)
{
    return (c->impl.progressive() && c->impl.progressive()->shader()) ? c->impl.progressive()->shader()->handle() : nullptr;
}
void gdpt_camera_prepare_post(gdpt_camera_node *c) { c->impl.prepare_post(); }
gdpt_shader *gdpt_camera_temporal_shader(const gdpt_camera_node *c)
{
    return (c->impl.temporal() && c->impl.temporal()->shader()) ? c->impl.temporal()->shader()->handle() : nullptr;
}
gdpt_rid gdpt_camera_temporal_rid(const gdpt_camera_node *c, int which) { return c->impl.temporal() ? c->impl.temporal()->frame_buffer_rid(which) : 0; }
int gdpt_camera_get_temporal_params(const gdpt_camera_node *c, gdpt_temporal_params *out)
{
    if (!c->impl.temporal()) return 0;
    *out = c->impl.temporal()->params();
    return 1;
}
gdpt_device *gdpt_camera_device(const gdpt_camera_node *c) { return c->impl.compute_shader() ? c->impl.compute_shader()->get_rendering_device() : nullptr; }
gdpt_rid gdpt_camera_output_rid(const gdpt_camera_node *c) { return c->impl.output_texture_rid(); }
gdpt_rid gdpt_camera_depth_rid(const gdpt_camera_node *c) { return c->impl.depth_texture_rid(); }
gdpt_rid gdpt_camera_accum_rid(const gdpt_camera_node *c) { return c->impl.progressive() ? c->impl.progressive()->frame_buffer_rid() : 0; }
void gdpt_camera_get_camera_block(const gdpt_camera_node *c, gdpt_camera *out) { *out = c->impl.camera_block(); }
uint32_t gdpt_camera_last_frame_count(const gdpt_camera_node *c) { return c->impl.last_frame_count(); }

void gdpt_make_camera_block(const float *transform12, float fov_degrees, int width, int height, uint32_t frame_index, gdpt_camera *out)
{
    gdpt::CameraBlock cb;
    const gdpt::Mat4 proj = gdpt::Mat4::perspective(fov_degrees, static_cast<float>(width) / height, 0.01f, 1000.0f);
    cb.set_camera_transform(gdpt::Xform3::from_rows12(transform12), proj);
    cb.frame_index = frame_index;
    *out = cb;
}

void gdpt_make_temporal_delta(const float *previous_vp16, const float *transform12, float fov_degrees, int width, int height,
                              float *out_vp16, float *out_delta16)
{
    const gdpt::Mat4 proj = gdpt::Mat4::perspective(fov_degrees, static_cast<float>(width) / height, 0.01f, 1000.0f);
    const gdpt::Xform3 view = gdpt::Xform3::from_rows12(transform12).affine_inverse(); // path_tracing_camera.cpp:220
    gdpt::Mat4 prev;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) prev.m[c][r] = previous_vp16[c * 4 + r];
    const gdpt::Mat4 vp = proj * gdpt::Mat4(view);
    (prev * vp.inverse()).affine_part().to_float16(out_delta16);
    vp.to_float16(out_vp16);
}

} // extern "C"
