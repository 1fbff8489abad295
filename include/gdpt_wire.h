/* gdpt_wire.h -- wire formats that cross the drop-in boundary.
 *
 * These are the POD records the reference memcpy's into PackedByteArrays and
 * binds as std430 storage buffers.  The CUDA backend consumes exactly these
 * bytes; nothing here is invented by us.  Each struct cites the reference
 * definition it mirrors (paths relative to the reference checkout).
 *
 * Matrix memory order is column-major, m[col*4+row] (src/utils.h:15-49).
 */
#ifndef GDPT_WIRE_H
#define GDPT_WIRE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
#define GDPT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define GDPT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* set 0 binding 2 of main.glsl (main.glsl:101-108), host twin
 * PathTracingCamera::RenderParameters (src/path_tracing/path_tracing_camera.h:36-44).
 * The kernels read only width/height, like the shader. */
typedef struct gdpt_render_params {
    float    background[4];
    int32_t  width;
    int32_t  height;
    float    fov;
    uint32_t triangle_count;
    uint32_t blas_count;
} gdpt_render_params;
GDPT_STATIC_ASSERT(sizeof(gdpt_render_params) == 36, "Params is 36 B");

/* set 0 binding 3 (main.glsl:110-117), host twin Camera
 * (src/path_tracing/render_parameters.h:14-22). */
typedef struct gdpt_camera {
    float    vp[16];
    float    ivp[16];
    float    position[4];
    uint32_t frame_index;
    float    z_near;
    float    z_far;
    uint32_t _tail_pad;      /* struct is alignas(16) upstream: 160 B */
} gdpt_camera;
GDPT_STATIC_ASSERT(sizeof(gdpt_camera) == 160, "Camera is 160 B");
GDPT_STATIC_ASSERT(offsetof(gdpt_camera, ivp) == 64, "ivp @64");
GDPT_STATIC_ASSERT(offsetof(gdpt_camera, position) == 128, "position @128");
GDPT_STATIC_ASSERT(offsetof(gdpt_camera, frame_index) == 144, "frame_index @144");

/* set 1 binding 3: BVHNode (src/bvh/bvh.h:46-54, main.glsl:45-52). */
typedef struct gdpt_bvh_node {
    float    aabb_min[4];
    float    aabb_max[4];
    uint32_t left_child;
    uint32_t right_child;
    uint32_t first_tri_index;
    uint32_t tri_count;      /* >0 <=> leaf */
} gdpt_bvh_node;
GDPT_STATIC_ASSERT(sizeof(gdpt_bvh_node) == 48, "BVHNode is 48 B");

/* set 1 binding 5: TLASNode (src/bvh/bvh.h:56-62, main.glsl:54-60). */
typedef struct gdpt_tlas_node {
    float    aabb_min[3];
    uint32_t left_right;     /* lo16 = left, hi16 = right, 0 <=> leaf */
    float    aabb_max[3];
    uint32_t blas;
} gdpt_tlas_node;
GDPT_STATIC_ASSERT(sizeof(gdpt_tlas_node) == 32, "TLASNode is 32 B");

/* set 1 binding 4: BLASInstance (src/bvh/bvh.h:64-71, main.glsl:84-93). */
typedef struct gdpt_blas_instance {
    float    transform[16];
    float    inverse_transform[16];
    float    aabb_min[4];
    float    aabb_max[4];
    uint32_t root;           /* global index into the BVH node array */
    uint32_t materials[3];   /* surface index -> material index */
} gdpt_blas_instance;
GDPT_STATIC_ASSERT(sizeof(gdpt_blas_instance) == 176, "BLASInstance is 176 B");
GDPT_STATIC_ASSERT(offsetof(gdpt_blas_instance, root) == 160, "root @160");

/* set 1 binding 0: GpuTriangleGeometry (render_parameters.h:59-62, main.glsl:14-16). */
typedef struct gdpt_triangle_geometry {
    float v[3][4];           /* w = 1 padding */
} gdpt_triangle_geometry;
GDPT_STATIC_ASSERT(sizeof(gdpt_triangle_geometry) == 48, "TriangleGeometry is 48 B");

/* set 1 binding 1: GpuTriangleData (render_parameters.h:64-71, main.glsl:18-24). */
typedef struct gdpt_triangle_data {
    float    n0[3];
    uint32_t material_index; /* = surface index inside the mesh */
    float    n1[4];
    float    n2[4];
    float    uvs[3][2];
    float    _tail_pad[2];
} gdpt_triangle_data;
GDPT_STATIC_ASSERT(sizeof(gdpt_triangle_data) == 80, "TriangleData is 80 B");
GDPT_STATIC_ASSERT(offsetof(gdpt_triangle_data, uvs) == 48, "uvs @48");

/* set 1 binding 2: GpuMaterial (render_parameters.h:49-57, main.glsl:32-43). */
typedef struct gdpt_material {
    float   albedo[4];
    float   emission[4];     /* rgb colour, w = energy multiplier */
    float   metallic;
    float   roughness;
    int32_t albedo_texture_index; /* -1: none */
    /* The reference pads the record with five unused floats (main.glsl:39).  A shader created with
     * "#define GDPT_MATERIAL_EXT" (material breadth beyond geometry_group3d.cpp:271-292) reads three of
     * them; without the define they are ignored, as upstream.  Texture fields hold layer + 1, 0 = none. */
    uint32_t ext_roughness_texture; /* roughness *= red channel of that layer  */
    uint32_t ext_metallic_texture;  /* metallic  *= red channel of that layer  */
    uint32_t ext_flags;             /* GDPT_MATERIAL_ALBEDO_SRGB: the albedo layer holds sRGB-encoded colour (upstream samples
                                     * it as UNORM, path_tracing_camera.cpp:182) */
    float   _pad[2];
} gdpt_material;
#define GDPT_MATERIAL_ALBEDO_SRGB 1u
/* Optional set 1 binding 6 of a GDPT_MATERIAL_EXT shader: material ids for ANY number of surfaces per instance
 * (BLASInstance.materials holds three, bvh.h:71).  uint32 words: offset[n_instances + 1] (in words, from the start of
 * the buffer), then the ids; the material of surface s of instance i is word offset[i] + s. */
GDPT_STATIC_ASSERT(sizeof(gdpt_material) == 64, "Material is 64 B");

/* progressive_rendering.glsl:12-16, host twin
 * ProgressiveRendering::RenderParameters (post_processing/progressive_rendering.h:14-19). */
typedef struct gdpt_progressive_params {
    int32_t  width;
    int32_t  height;
    uint32_t frame_count;
} gdpt_progressive_params;
GDPT_STATIC_ASSERT(sizeof(gdpt_progressive_params) == 12, "progressive Params is 12 B");

/* temporal_reprojection.glsl:4-12, host twin TemporalReprojection::RenderParameters
 * (post_processing/temporal_reprojection.h:15-23).  delta_matrix is column-major
 * (Utils::projection_to_float, src/utils.h:39-49); blend_factor, near_plane and
 * far_plane are uploaded but never read by the shader (it blends with a literal 0.75). */
typedef struct gdpt_temporal_params {
    float    delta_matrix[16];
    int32_t  width;
    int32_t  height;
    uint32_t frame_count;
    float    blend_factor;
    float    near_plane;
    float    far_plane;
} gdpt_temporal_params;
GDPT_STATIC_ASSERT(sizeof(gdpt_temporal_params) == 88, "temporal Params is 88 B");
GDPT_STATIC_ASSERT(offsetof(gdpt_temporal_params, frame_count) == 72, "frameCount @72");

/* Host-side triangle record handed to the BLAS builder (src/bvh/bvh.h:22-29). */
typedef struct gdpt_build_triangle {
    float    vertices[3][4];
    float    centroid[4];
    float    normals[3][4];
    float    uvs[3][2];
    uint32_t material_index;
    uint32_t _tail_pad;
} gdpt_build_triangle;
GDPT_STATIC_ASSERT(sizeof(gdpt_build_triangle) == 144, "BVH::Triangle is 144 B");
GDPT_STATIC_ASSERT(offsetof(gdpt_build_triangle, material_index) == 136, "materialIndex @136");

/* ---- parity observables; emitted only in trace mode ---- */

/* One record per traced ray segment. */
typedef struct gdpt_trace_record {
    uint32_t hit;            /* hitInfo.t < 1e9 (main.glsl:349) */
    uint32_t triangle;       /* global triangle index (main.glsl:251) */
    uint32_t blas;           /* instance index (main.glsl:325) */
    uint32_t front;          /* main.glsl:255 */
    float    t;              /* hitInfo.t */
    float    u, v;           /* barycentrics */
    uint32_t node_pops;      /* TLAS + BLAS stack pops, roots included */
    uint32_t box_tests;      /* intersectAABB calls */
    uint32_t tri_tests;      /* hitInfo.steps (main.glsl:225) */
    uint32_t tlas_leaves;    /* instance ray transforms */
    uint32_t max_stack;      /* deepest stack fill seen on either stack */
    uint32_t visit_hash_lo;  /* FNV-1a 64 over the pop sequence (TLAS pops tagged bit 31) */
    uint32_t visit_hash_hi;
} gdpt_trace_record;
GDPT_STATIC_ASSERT(sizeof(gdpt_trace_record) == 56, "trace record is 56 B");

#define GDPT_VISIT_TLAS_TAG 0x80000000u
#define GDPT_FNV64_OFFSET 0xcbf29ce484222325ull
#define GDPT_FNV64_PRIME  0x100000001b3ull

#endif /* GDPT_WIRE_H */
