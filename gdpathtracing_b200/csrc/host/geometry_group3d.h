// geometry_group3d.h -- Godot-free twin of the reference's GeometryGroup3D node
// (src/path_tracing/geometry_group3d.{h,cpp}): same properties
// (default_material, texture_array_resolution), same build() pipeline, same
// get_*_buffer() accessors.  The Godot scene tree is replaced by explicit
// resource/instance registration: a "resource handle" plays the role of a
// Ref<> pointer identity (the reference de-duplicates meshes, materials and
// textures by pointer, geometry_group3d.cpp:119-148,161-175).
#ifndef GDPT_GEOMETRY_GROUP3D_H
#define GDPT_GEOMETRY_GROUP3D_H

#include "accel_build.h"

#include <cstdint>
#include <vector>

namespace gdpt {

// The StandardMaterial3D fields build() reads (geometry_group3d.cpp:271-292).
// Defaults are Godot's StandardMaterial3D defaults.
struct StandardMaterial {
    float albedo[3] = { 1.0f, 1.0f, 1.0f };
    float metallic = 0.0f;
    float roughness = 1.0f;
    float emission[3] = { 0.0f, 0.0f, 0.0f };
    float emission_energy_multiplier = 1.0f;
    int albedo_texture = -1; // texture resource handle, -1 = none
    // material breadth beyond geometry_group3d.cpp:271-292; converted only when the group's material_ext is on
    int roughness_texture = -1, metallic_texture = -1; // StandardMaterial3D.roughness_texture / metallic_texture (red channel)
    bool albedo_srgb = false;                          // the albedo texture holds sRGB-encoded colour
    bool is_standard = true; // false: some other Material subclass -> resolves to index 0 (:128-130)
};

class GeometryGroup3D {
public:
    GeometryGroup3D();

    // --- resources (stand-ins for Ref<Texture2D>, Ref<Material>, Ref<Mesh>) ---
    int add_texture(const uint8_t *rgba8, int width, int height);
    int add_material(const StandardMaterial &m);
    int add_mesh(const SurfaceArrays *surfaces, int n_surfaces); // arrays are copied

    // --- scene: one call per MeshInstance3D found by the reference's BFS (:150-214) ---
    // material_override < 0: none.  surface_overrides[i] < 0: no override on surface i.
    void add_mesh_instance(int mesh, const Xform3 &global_transform, int material_override,
                           const int *surface_overrides, int n_surface_overrides);

    // --- properties ---
    int get_default_material() const { return default_material_; }
    void set_default_material(int material_handle) { default_material_ = material_handle; }
    int get_texture_array_resolution() const { return texture_array_resolution_; }
    void set_texture_array_resolution(int v) { texture_array_resolution_ = v; }
    // ours (no upstream twin): threads of the BLAS build, 0 = all hardware threads, 1 = upstream's single thread; same bytes
    int get_build_threads() const { return build_threads_; }
    void set_build_threads(int v) { build_threads_ = v < 0 ? 0 : v; }

    // ours: convert roughness / metallic textures and the sRGB flag into the extension words of the material
    // record, and emit the material table for instances with any number of surfaces (gdpt_wire.h).  Off = upstream's bytes.
    bool get_material_ext() const { return material_ext_; }
    void set_material_ext(bool on) { material_ext_ = on; }
    const std::vector<uint32_t> &get_surface_materials_buffer() const { return surface_materials_; }

    void build();

    int get_blas_count() const { return (int)blas_instances_.size(); }
    int get_material_count() const { return (int)materials_.size(); }
    int get_triangle_count() const { return (int)triangles_.size(); }
    int get_bvh_node_count() const { return (int)bvh_nodes_.size(); }
    int get_tlas_node_count() const { return (int)tlas_nodes_.size(); }

    const std::vector<gdpt_triangle_geometry> &get_triangles_geometry_buffer() const { return triangles_geometry_; }
    const std::vector<gdpt_triangle_data> &get_triangles_data_buffer() const { return triangles_data_; }
    const std::vector<gdpt_material> &get_materials_buffer() const { return materials_; }
    const std::vector<gdpt_bvh_node> &get_bvh_buffer() const { return bvh_nodes_; }
    const std::vector<gdpt_blas_instance> &get_blas_buffer() const { return blas_instances_; }
    const std::vector<gdpt_tlas_node> &get_tlas_buffer() const { return tlas_nodes_; }
    // layers of texture_array_resolution^2 RGBA8 texels
    const std::vector<std::vector<uint8_t>> &get_textures_buffer() const { return textures_; }

    double last_build_seconds() const { return build_seconds_; }

private:
    struct TextureRes { std::vector<uint8_t> rgba; int w, h; };
    struct MeshRes {
        std::vector<std::vector<float>> positions, normals, uvs;
        std::vector<std::vector<int32_t>> indices;
    };
    struct NodeRef { int mesh_slot; std::vector<int> material_ids; Xform3 transform; };
    struct PendingInstance { int mesh; Xform3 transform; int material_override; std::vector<int> surface_overrides; };

    unsigned get_material_index(int material_handle);
    int get_texture_index(int texture_handle);
    static std::vector<uint8_t> resize_bilinear(const TextureRes &src, int res, int threads = 1);
    static void resize_rows(const TextureRes &src, int res, int y_begin, int y_end, uint8_t *out);

    std::vector<TextureRes> texture_pool_;
    std::vector<StandardMaterial> material_pool_;
    std::vector<MeshRes> mesh_pool_;
    std::vector<PendingInstance> scene_;

    int default_material_ = -1;
    int texture_array_resolution_;
    int build_threads_ = 0;

    std::vector<int> mesh_refs_;         // initial_geometry_references
    std::vector<int> material_refs_;     // initial_material_references (handle, -2 = built-in default)
    std::vector<int> texture_refs_;      // texture_references
    std::vector<NodeRef> node_refs_;

    std::vector<gdpt_bvh_node> bvh_nodes_;
    std::vector<gdpt_tlas_node> tlas_nodes_;
    std::vector<gdpt_build_triangle> triangles_;
    std::vector<gdpt_triangle_geometry> triangles_geometry_;
    std::vector<gdpt_triangle_data> triangles_data_;
    std::vector<gdpt_blas_instance> blas_instances_;
    std::vector<gdpt_material> materials_;
    bool material_ext_ = false;
    std::vector<uint32_t> surface_materials_; // offset[n_instances + 1], ids (set 1 binding 6 of a GDPT_MATERIAL_EXT shader)
    std::vector<std::vector<uint8_t>> textures_;
    double build_seconds_ = 0.0;
};

} // namespace gdpt
#endif
