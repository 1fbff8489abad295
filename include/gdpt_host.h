/* gdpt_host.h -- flat C view of the host layer that sits above libgdpt_cuda.so
 * (libgdpt_host.so): the Godot-free twins of GeometryGroup3D and
 * PathTracingCamera.  Inside Godot these classes are driven by the scene tree;
 * this header is how the standalone harness (Python via ctypes, or any C
 * program) drives the very same code.  Each function cites the reference
 * member it stands for.  Transforms are 12 floats: Basis rows (row-major, as
 * godot::Basis stores them) followed by the origin.
 */
#ifndef GDPT_HOST_H
#define GDPT_HOST_H

#include "gdpt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gdpt_geometry_group gdpt_geometry_group; /* GeometryGroup3D */
typedef struct gdpt_camera_node gdpt_camera_node;       /* PathTracingCamera */

/* The StandardMaterial3D fields GeometryGroup3D::build reads
 * (src/path_tracing/geometry_group3d.cpp:271-292). */
typedef struct gdpt_standard_material {
    float   albedo[3];
    float   metallic;
    float   roughness;
    float   emission[3];
    float   emission_energy_multiplier;
    int32_t albedo_texture;   /* texture handle from gdpt_group_add_texture, -1 = none */
    int32_t is_standard;      /* 0: not a StandardMaterial3D -> default material (:128-130) */
} gdpt_standard_material;

/* ---- GeometryGroup3D ---------------------------------------------------- */
GDPT_API gdpt_geometry_group *gdpt_group_create(void);
GDPT_API void gdpt_group_destroy(gdpt_geometry_group *g);
/* resources (Ref<Texture2D> / Ref<Material> / Ref<ArrayMesh>); return handles */
GDPT_API int  gdpt_group_add_texture(gdpt_geometry_group *g, const uint8_t *rgba8, int width, int height);
GDPT_API int  gdpt_group_add_material(gdpt_geometry_group *g, const gdpt_standard_material *m);
/* Material breadth beyond geometry_group3d.cpp:271-292, converted only while the group's material_ext
 * property is on: roughness / metallic texture handles (-1 = none; the red channel scales the scalar) and whether the
 * albedo texture holds sRGB-encoded colour.  With material_ext on the group also emits buffer 6, the material table for
 * instances with any number of surfaces (gdpt_wire.h), and PathTracingCamera creates its shader with
 * "#define GDPT_MATERIAL_EXT".  Off (the default): the reference's bytes and the reference's shading. */
GDPT_API int  gdpt_group_add_material_ext(gdpt_geometry_group *g, const gdpt_standard_material *m, int roughness_texture,
                                          int metallic_texture, int albedo_srgb);
GDPT_API void gdpt_group_set_material_ext(gdpt_geometry_group *g, int on);
GDPT_API int  gdpt_group_get_material_ext(const gdpt_geometry_group *g);
/* one ArrayMesh: per-surface vertex/index counts, arrays concatenated surface after surface
 * (Mesh::surface_get_arrays, bvh.cpp:193-199) */
GDPT_API int  gdpt_group_add_mesh(gdpt_geometry_group *g, int n_surfaces, const int32_t *vertex_counts,
                                  const int32_t *index_counts, const float *positions, const float *normals,
                                  const float *uvs, const int32_t *indices);
/* one MeshInstance3D child (geometry_group3d.cpp:160-206) */
GDPT_API void gdpt_group_add_mesh_instance(gdpt_geometry_group *g, int mesh, const float *transform12,
                                  int material_override, const int32_t *surface_overrides, int n_surface_overrides);
/* properties (geometry_group3d.cpp:75-105) */
GDPT_API void gdpt_group_set_default_material(gdpt_geometry_group *g, int material);
GDPT_API void gdpt_group_set_texture_array_resolution(gdpt_geometry_group *g, int resolution);
GDPT_API int  gdpt_group_get_texture_array_resolution(const gdpt_geometry_group *g);
/* ours: threads used by the BLAS build (bvh.cpp:108-185 is single-threaded); 0 = all hardware threads, 1 = one.
 * The arrays that come out are byte-identical for every value. */
GDPT_API void gdpt_group_set_build_threads(gdpt_geometry_group *g, int threads);
GDPT_API int  gdpt_group_get_build_threads(const gdpt_geometry_group *g);
/* GeometryGroup3D::build (geometry_group3d.cpp:228-366) */
GDPT_API void gdpt_group_build(gdpt_geometry_group *g);
GDPT_API double gdpt_group_last_build_seconds(const gdpt_geometry_group *g);
/* get_*_count / get_*_buffer (geometry_group3d.cpp:7-73): which = 0 triangle geometry, 1 triangle data,
 * 2 materials, 3 bvh, 4 blas instances, 5 tlas; texture layers through gdpt_group_texture_layer */
GDPT_API uint64_t gdpt_group_buffer_size(const gdpt_geometry_group *g, int which);
GDPT_API const void *gdpt_group_buffer_data(const gdpt_geometry_group *g, int which);
GDPT_API int  gdpt_group_texture_layer_count(const gdpt_geometry_group *g);
GDPT_API const uint8_t *gdpt_group_texture_layer(const gdpt_geometry_group *g, int layer);

/* ---- PathTracingCamera -------------------------------------------------- */
GDPT_API gdpt_camera_node *gdpt_camera_create(void);
GDPT_API void gdpt_camera_destroy(gdpt_camera_node *c);
/* properties (path_tracing_camera.cpp:3-31) */
GDPT_API void  gdpt_camera_set_fov(gdpt_camera_node *c, float fov);
GDPT_API float gdpt_camera_get_fov(const gdpt_camera_node *c);
GDPT_API void  gdpt_camera_set_geometry_group(gdpt_camera_node *c, gdpt_geometry_group *g);
GDPT_API void  gdpt_camera_set_denoising_mode(gdpt_camera_node *c, int mode);
GDPT_API int   gdpt_camera_get_denoising_mode(const gdpt_camera_node *c);
/* engine services */
GDPT_API void  gdpt_camera_set_window_size(gdpt_camera_node *c, int width, int height);
GDPT_API void  gdpt_camera_set_global_transform(gdpt_camera_node *c, const float *transform12);
/* backend parameters */
GDPT_API void  gdpt_camera_set_max_depth(gdpt_camera_node *c, int max_depth);
GDPT_API void  gdpt_camera_set_cuda_device(gdpt_camera_node *c, int ordinal);
GDPT_API void  gdpt_camera_set_frame_index(gdpt_camera_node *c, uint32_t frame_index);
GDPT_API void  gdpt_camera_set_shard(gdpt_camera_node *c, int part, int n_parts, int band_rows);
GDPT_API void  gdpt_camera_set_trace(gdpt_camera_node *c, int segments, uint32_t visits_per_ray);
GDPT_API void  gdpt_camera_set_debug_steps(gdpt_camera_node *c, int on);
/* -1 backend default, 0 reference visit order, 1 tight-box culling (results identical) */
GDPT_API void  gdpt_camera_set_cull(gdpt_camera_node *c, int mode);
/* hit records (gdpt_trace_record without work counters) of the first `segments` path segments, written by the
 * rendering kernels themselves; read back with gdpt_shader_read_trace */
GDPT_API void  gdpt_camera_set_record_hits(gdpt_camera_node *c, int segments);
/* kernel schedule ("#define GDPT_VARIANT n", gdpt.h); -1 = backend default; results are identical for every value */
GDPT_API void  gdpt_camera_set_variant(gdpt_camera_node *c, int variant);
/* scheduling knob of the path kernels ("#define GDPT_TUNE_<NAME> n", gdpt.h): A/B measurements; results are identical */
GDPT_API void  gdpt_camera_set_count_work(gdpt_camera_node *c, int on);
GDPT_API void  gdpt_camera_set_tuning(gdpt_camera_node *c, const char *name, int value);
GDPT_API void  gdpt_camera_set_fused_frame(gdpt_camera_node *c, int on);
/* init() / render() (path_tracing_camera.cpp:111-232); init returns 1 when check_ready() */
GDPT_API int   gdpt_camera_init(gdpt_camera_node *c);
GDPT_API void  gdpt_camera_render(gdpt_camera_node *c);
GDPT_API void  gdpt_camera_render_device_only(gdpt_camera_node *c);
/* pipelined render(): begin returns 1 once a frame (incl. its read-back) is enqueued; wait blocks for the oldest
 * frame in flight (at most GDPT_MAX_FRAMES_IN_FLIGHT) and returns its RGBA8 pixels in page-locked host memory (NULL: none / error). */
GDPT_API int   gdpt_camera_render_begin(gdpt_camera_node *c);
GDPT_API const uint8_t *gdpt_camera_render_wait(gdpt_camera_node *c, gdpt_frame_stats *stats);
GDPT_API const uint8_t *gdpt_camera_output_image(const gdpt_camera_node *c);
/* handles for tests / benchmarks */
GDPT_API gdpt_shader *gdpt_camera_main_shader(const gdpt_camera_node *c);
GDPT_API gdpt_shader *gdpt_camera_progressive_shader(const gdpt_camera_node *c);
/* creates the post-process object of the current denoising_mode now instead of inside the first render()
 * (path_tracing_camera.cpp:207-223 creates it lazily); frame counters are untouched */
GDPT_API void gdpt_camera_prepare_post(gdpt_camera_node *c);
/* the TemporalReprojection post process (created by the first frame rendered in that mode): its shader, its two
 * ping-pong frame buffers (which = 0 / 1 -> frameBuffer1 / frameBuffer2) and the Params block last dispatched with */
GDPT_API gdpt_shader *gdpt_camera_temporal_shader(const gdpt_camera_node *c);
GDPT_API gdpt_rid gdpt_camera_temporal_rid(const gdpt_camera_node *c, int which);
GDPT_API int   gdpt_camera_get_temporal_params(const gdpt_camera_node *c, gdpt_temporal_params *out);
GDPT_API gdpt_device *gdpt_camera_device(const gdpt_camera_node *c);
GDPT_API gdpt_rid gdpt_camera_output_rid(const gdpt_camera_node *c);
GDPT_API gdpt_rid gdpt_camera_depth_rid(const gdpt_camera_node *c);
GDPT_API gdpt_rid gdpt_camera_accum_rid(const gdpt_camera_node *c);
GDPT_API void  gdpt_camera_get_camera_block(const gdpt_camera_node *c, gdpt_camera *out);
GDPT_API uint32_t gdpt_camera_last_frame_count(const gdpt_camera_node *c);

/* Camera::set_camera_transform + Projection::create_perspective, standalone
 * (render_parameters.h:23-38, path_tracing_camera.cpp:134): fills vp/ivp/position. */
GDPT_API void gdpt_make_camera_block(const float *transform12, float fov_degrees, int width, int height,
                                  uint32_t frame_index, gdpt_camera *out);
/* The parameter update of TemporalReprojection::render, standalone (temporal_reprojection.cpp:57-61 with the
 * view matrix of path_tracing_camera.cpp:220): vp = projection * Projection(camera_transform.affine_inverse()),
 * delta = the affine part of previous_vp * vp.inverse().  Matrices are 16 floats, column-major. */
GDPT_API void gdpt_make_temporal_delta(const float *previous_vp16, const float *transform12, float fov_degrees,
                                  int width, int height, float *out_vp16, float *out_delta16);

#ifdef __cplusplus
}
#endif
#endif /* GDPT_HOST_H */
