// geometry_group3d.cpp -- see geometry_group3d.h.  Pipeline order follows
// GeometryGroup3D::build (src/path_tracing/geometry_group3d.cpp:228-366).
#include "geometry_group3d.h"

#include <chrono>
#include <thread>
#include <cmath>
#include <cstring>

namespace gdpt {

GeometryGroup3D::GeometryGroup3D() : texture_array_resolution_(512) {} // ctor default upstream (geometry_group3d.cpp:3)

int GeometryGroup3D::add_texture(const uint8_t *rgba8, int width, int height)
{
    TextureRes t;
    t.w = width; t.h = height;
    t.rgba.assign(rgba8, rgba8 + (size_t)width * height * 4);
    texture_pool_.push_back(std::move(t));
    return (int)texture_pool_.size() - 1;
}

int GeometryGroup3D::add_material(const StandardMaterial &m)
{
    material_pool_.push_back(m);
    return (int)material_pool_.size() - 1;
}

int GeometryGroup3D::add_mesh(const SurfaceArrays *surfaces, int n_surfaces)
{
    MeshRes m;
    for (int s = 0; s < n_surfaces; s++) {
        const SurfaceArrays &a = surfaces[s];
        m.positions.emplace_back(a.positions, a.positions + a.vertex_count * 3);
        m.normals.emplace_back(a.normals, a.normals + a.vertex_count * 3);
        m.uvs.emplace_back(a.uvs, a.uvs + a.vertex_count * 2);
        m.indices.emplace_back(a.indices, a.indices + a.index_count);
    }
    mesh_pool_.push_back(std::move(m));
    return (int)mesh_pool_.size() - 1;
}

void GeometryGroup3D::add_mesh_instance(int mesh, const Xform3 &global_transform, int material_override,
                                        const int *surface_overrides, int n_surface_overrides)
{
    PendingInstance p;
    p.mesh = mesh; p.transform = global_transform; p.material_override = material_override;
    if (surface_overrides) p.surface_overrides.assign(surface_overrides, surface_overrides + n_surface_overrides);
    scene_.push_back(std::move(p));
}

// get_material_index (geometry_group3d.cpp:119-135): identity lookup first, then
// "is it a StandardMaterial3D", else the default at index 0.
unsigned GeometryGroup3D::get_material_index(int handle)
{
    if (handle >= 0)
        for (size_t i = 0; i < material_refs_.size(); i++)
            if (material_refs_[i] == handle) return (unsigned)i;
    if (handle < 0 || handle >= (int)material_pool_.size() || !material_pool_[handle].is_standard) return 0;
    material_refs_.push_back(handle);
    return (unsigned)material_refs_.size() - 1;
}

// get_texture_index (geometry_group3d.cpp:137-148)
int GeometryGroup3D::get_texture_index(int handle)
{
    if (handle < 0 || handle >= (int)texture_pool_.size()) return -1;
    for (size_t i = 0; i < texture_refs_.size(); i++)
        if (texture_refs_[i] == handle) return (int)i;
    texture_refs_.push_back(handle);
    return (int)texture_refs_.size() - 1;
}

// Stand-in for Image::resize(res, res) with Godot's default bilinear filter
// (geometry_group3d.cpp:295-299).  Godot's own resampler lives in the engine,
// which is not part of the reference tree, so this is NOT pinned to it: pixel
// centres are mapped with the usual (i+0.5)*scale-0.5 rule, edges clamp.
// Rows are independent: `threads` workers take equal bands of rows (0 = all hardware threads); the bytes do not depend on it.
std::vector<uint8_t> GeometryGroup3D::resize_bilinear(const TextureRes &src, int res, int threads)
{
    std::vector<uint8_t> out((size_t)res * res * 4);
    if (src.w == res && src.h == res) { out = src.rgba; return out; }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min(threads, std::max(1, res / 64));
    if (threads > 1) {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] { resize_rows(src, res, (int)((long long)res * t / threads), (int)((long long)res * (t + 1) / threads), out.data()); });
        for (std::thread &th : pool) th.join();
    } else {
        resize_rows(src, res, 0, res, out.data());
    }
    return out;
}

void GeometryGroup3D::resize_rows(const TextureRes &src, int res, int y_begin, int y_end, uint8_t *out)
{
    const float sx = (float)src.w / (float)res, sy = (float)src.h / (float)res;
    for (int y = y_begin; y < y_end; y++) {
        float fy = ((float)y + 0.5f) * sy - 0.5f;
        if (fy < 0) fy = 0;
        int y0 = (int)fy, y1 = y0 + 1 < src.h ? y0 + 1 : src.h - 1;
        const float wy = fy - (float)y0;
        for (int x = 0; x < res; x++) {
            float fx = ((float)x + 0.5f) * sx - 0.5f;
            if (fx < 0) fx = 0;
            int x0 = (int)fx, x1 = x0 + 1 < src.w ? x0 + 1 : src.w - 1;
            const float wx = fx - (float)x0;
            for (int c = 0; c < 4; c++) {
                const float p00 = src.rgba[((size_t)y0 * src.w + x0) * 4 + c], p01 = src.rgba[((size_t)y0 * src.w + x1) * 4 + c];
                const float p10 = src.rgba[((size_t)y1 * src.w + x0) * 4 + c], p11 = src.rgba[((size_t)y1 * src.w + x1) * 4 + c];
                const float top = p00 + (p01 - p00) * wx, bot = p10 + (p11 - p10) * wx;
                const float v = top + (bot - top) * wy;
                out[((size_t)y * res + x) * 4 + c] = (uint8_t)std::lrintf(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
}

void GeometryGroup3D::build()
{
    const auto t0 = std::chrono::steady_clock::now();
    // Upstream forgets to clear triangles/textures/texture_references between builds
    // (geometry_group3d.cpp:230-237), which only matters for a second build(); we start clean.
    mesh_refs_.clear(); node_refs_.clear(); material_refs_.clear(); texture_refs_.clear();
    tlas_nodes_.clear(); bvh_nodes_.clear(); blas_instances_.clear();
    triangles_.clear(); textures_.clear(); materials_.clear();

    // index 0 is always a default material (:239-247); -2 marks the built-in grey one
    material_refs_.push_back(default_material_ >= 0 ? default_material_ : -2);

    // collect_mesh_instances (:150-214) over the registered instances, in order
    for (const PendingInstance &p : scene_) {
        if (p.mesh < 0 || p.mesh >= (int)mesh_pool_.size()) continue; // mesh.is_valid()
        int slot = -1;
        for (size_t j = 0; j < mesh_refs_.size(); j++)
            if (mesh_refs_[j] == p.mesh) { slot = (int)j; break; }
        if (slot < 0) { slot = (int)mesh_refs_.size(); mesh_refs_.push_back(p.mesh); }
        const int n_surf = (int)mesh_pool_[p.mesh].indices.size(); // get_surface_override_material_count()
        NodeRef ref;
        ref.mesh_slot = slot; ref.transform = p.transform;
        if (p.material_override >= 0) {
            const unsigned id = get_material_index(p.material_override);
            for (int s = 0; s < n_surf; s++) ref.material_ids.push_back((int)id);
        } else {
            for (int s = 0; s < n_surf; s++) {
                const int h = s < (int)p.surface_overrides.size() ? p.surface_overrides[s] : -1;
                ref.material_ids.push_back((int)get_material_index(h));
            }
        }
        node_refs_.push_back(std::move(ref));
    }

    // StandardMaterial3D -> GpuMaterial (:271-292)
    for (int handle : material_refs_) {
        StandardMaterial m;
        if (handle >= 0) m = material_pool_[handle];
        else { m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.5f; m.roughness = 0.5f; m.metallic = 0.0f; }
        gdpt_material g;
        std::memset(&g, 0, sizeof(g));
        g.albedo[0] = m.albedo[0]; g.albedo[1] = m.albedo[1]; g.albedo[2] = m.albedo[2]; g.albedo[3] = 1.0f;
        g.metallic = m.metallic;
        g.roughness = m.roughness;
        g.emission[0] = m.emission[0]; g.emission[1] = m.emission[1]; g.emission[2] = m.emission[2];
        g.emission[3] = m.emission_energy_multiplier;
        g.albedo_texture_index = get_texture_index(m.albedo_texture);
        if (material_ext_) { // extension words in the reference's padding: layer + 1, 0 = none
            g.ext_roughness_texture = (uint32_t)(get_texture_index(m.roughness_texture) + 1);
            g.ext_metallic_texture = (uint32_t)(get_texture_index(m.metallic_texture) + 1);
            g.ext_flags = m.albedo_srgb ? GDPT_MATERIAL_ALBEDO_SRGB : 0u;
        }
        materials_.push_back(g);
    }
    // texture array layers (:294-303); a blank layer when the scene has no texture
    for (int handle : texture_refs_) textures_.push_back(resize_bilinear(texture_pool_[handle], texture_array_resolution_, build_threads_));
    if (textures_.empty())
        textures_.emplace_back((size_t)texture_array_resolution_ * texture_array_resolution_ * 4, (uint8_t)0);

    // one BLAS per unique mesh (:306-313)
    AccelBuilder builder;
    builder.set_threads(build_threads_);
    std::vector<uint32_t> roots;
    for (int mesh : mesh_refs_) {
        const MeshRes &m = mesh_pool_[mesh];
        std::vector<SurfaceArrays> surf(m.indices.size());
        for (size_t s = 0; s < surf.size(); s++) {
            surf[s].positions = m.positions[s].data(); surf[s].normals = m.normals[s].data(); surf[s].uvs = m.uvs[s].data();
            surf[s].vertex_count = (int64_t)m.positions[s].size() / 3;
            surf[s].indices = m.indices[s].data(); surf[s].index_count = (int64_t)m.indices[s].size();
        }
        roots.push_back(builder.build_blas(bvh_nodes_, triangles_, surf.data(), (int)surf.size()));
    }
    // one instance per node (:322-341), then the TLAS (:350-351)
    for (const NodeRef &r : node_refs_)
        blas_instances_.push_back(AccelBuilder::make_instance(roots[r.mesh_slot], r.material_ids.data(),
                                                              (int)r.material_ids.size(), r.transform, bvh_nodes_));
    AccelBuilder::build_tlas(tlas_nodes_, blas_instances_);
    surface_materials_.clear();
    if (material_ext_) { // every surface of every instance (BLASInstance.materials stops at three, bvh.h:71)
        uint32_t at = (uint32_t)node_refs_.size() + 1u;
        for (const NodeRef &r : node_refs_) { surface_materials_.push_back(at); at += (uint32_t)r.material_ids.size(); }
        surface_materials_.push_back(at);
        for (const NodeRef &r : node_refs_)
            for (int id : r.material_ids) surface_materials_.push_back((uint32_t)id);
    }

    // flatten for the GPU (:356-365)
    triangles_geometry_.resize(triangles_.size());
    triangles_data_.resize(triangles_.size());
    for (size_t i = 0; i < triangles_.size(); i++) {
        const gdpt_build_triangle &t = triangles_[i];
        std::memcpy(triangles_geometry_[i].v, t.vertices, sizeof(t.vertices));
        gdpt_triangle_data &d = triangles_data_[i];
        std::memset(&d, 0, sizeof(d));
        d.n0[0] = t.normals[0][0]; d.n0[1] = t.normals[0][1]; d.n0[2] = t.normals[0][2];
        d.material_index = t.material_index;
        std::memcpy(d.n1, t.normals[1], sizeof(d.n1));
        std::memcpy(d.n2, t.normals[2], sizeof(d.n2));
        std::memcpy(d.uvs, t.uvs, sizeof(d.uvs));
    }
    build_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace gdpt
