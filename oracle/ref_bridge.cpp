// ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).
//
// extern "C" doorway into the REFERENCE's own acceleration-structure builder,
// compiled verbatim from /root/reference/src/bvh/bvh.cpp against
// oracle/godot_shim (see oracle/Makefile; output goes to oracle/_ref/, which is
// git-ignored).  This file contains no algorithm: it only marshals flat arrays
// into the types BVHBuilder::BuildBVH / BLASInstance / TLAS::build expect
// (src/bvh/bvh.h:64-142) and copies their results out.  It mirrors what
// GeometryGroup3D::build does with them (src/path_tracing/geometry_group3d.cpp:
// 306-365).  Only tests/ and bench.py's cpu_baseline leg load the resulting .so.
#include "godot_shim.hpp"
#include "bvh/bvh.h"

#include <cstring>
#include <vector>

using namespace BVH;

namespace {
struct RefScene {
    std::vector<BVHNode> nodes;
    std::vector<Triangle> triangles;
    std::vector<BLASInstance> instances;
    std::vector<TLASNode> tlas;
    BVHBuilder builder;
};
} // namespace

extern "C" {

void *refbvh_new() { return new RefScene(); }
void refbvh_free(void *h) { delete static_cast<RefScene *>(h); }

// One ArrayMesh with n_surfaces surfaces; arrays are concatenated per surface.
// Returns the BLAS root index (BVHBuilder::BuildBVH return value).
uint32_t refbvh_add_mesh(void *h, int n_surfaces, const int32_t *vertex_counts, const int32_t *index_counts,
                         const float *positions, const float *normals, const float *uvs, const int32_t *indices)
{
    RefScene *s = static_cast<RefScene *>(h);
    godot::ArrayMesh mesh;
    int64_t voff = 0, ioff = 0;
    for (int l = 0; l < n_surfaces; l++) {
        godot::Array a;
        a.slots[godot::Mesh::ARRAY_VERTEX].v3.ptr = reinterpret_cast<const godot::Vector3 *>(positions + 3 * voff);
        a.slots[godot::Mesh::ARRAY_VERTEX].v3.n = vertex_counts[l];
        a.slots[godot::Mesh::ARRAY_NORMAL].v3.ptr = reinterpret_cast<const godot::Vector3 *>(normals + 3 * voff);
        a.slots[godot::Mesh::ARRAY_NORMAL].v3.n = vertex_counts[l];
        a.slots[godot::Mesh::ARRAY_TEX_UV].v2.ptr = reinterpret_cast<const godot::Vector2 *>(uvs + 2 * voff);
        a.slots[godot::Mesh::ARRAY_TEX_UV].v2.n = vertex_counts[l];
        a.slots[godot::Mesh::ARRAY_INDEX].i32.ptr = indices + ioff;
        a.slots[godot::Mesh::ARRAY_INDEX].i32.n = index_counts[l];
        mesh.surfaces.push_back(a);
        voff += vertex_counts[l];
        ioff += index_counts[l];
    }
    godot::Ref<godot::ArrayMesh> ref(&mesh);
    return s->builder.BuildBVH(s->nodes, s->triangles, ref);
}

// transform12 = basis rows (9 floats, row-major as godot::Basis stores them) + origin (3).
void refbvh_add_instance(void *h, uint32_t root, const int32_t *material_ids, int n_material_ids,
                         const float *transform12)
{
    RefScene *s = static_cast<RefScene *>(h);
    // Value-initialise: upstream leaves unused material[] slots indeterminate (bvh.h:73-79);
    // zeroing them first changes no decision and makes the bytes reproducible.
    BLASInstance inst = BLASInstance();
    inst.blas_index = root;
    std::vector<int> ids(material_ids, material_ids + n_material_ids);
    inst.set_materials(ids);
    godot::Transform3D t;
    for (int r = 0; r < 3; r++)
        t.basis.rows[r] = godot::Vector3(transform12[r * 3 + 0], transform12[r * 3 + 1], transform12[r * 3 + 2]);
    t.origin = godot::Vector3(transform12[9], transform12[10], transform12[11]);
    inst.set_transform(t, s->nodes);
    s->instances.push_back(inst);
}

void refbvh_build_tlas(void *h)
{
    RefScene *s = static_cast<RefScene *>(h);
    s->tlas.clear();
    TLAS tlas;
    tlas.build(s->tlas, s->instances);
}

uint64_t refbvh_node_count(void *h) { return static_cast<RefScene *>(h)->nodes.size(); }
uint64_t refbvh_triangle_count(void *h) { return static_cast<RefScene *>(h)->triangles.size(); }
uint64_t refbvh_instance_count(void *h) { return static_cast<RefScene *>(h)->instances.size(); }
uint64_t refbvh_tlas_count(void *h) { return static_cast<RefScene *>(h)->tlas.size(); }

void refbvh_copy_nodes(void *h, void *out) { RefScene *s = static_cast<RefScene *>(h); std::memcpy(out, s->nodes.data(), s->nodes.size() * sizeof(BVHNode)); }
void refbvh_copy_triangles(void *h, void *out) { RefScene *s = static_cast<RefScene *>(h); std::memcpy(out, s->triangles.data(), s->triangles.size() * sizeof(Triangle)); }
void refbvh_copy_instances(void *h, void *out) { RefScene *s = static_cast<RefScene *>(h); std::memcpy(out, s->instances.data(), s->instances.size() * sizeof(BLASInstance)); }
void refbvh_copy_tlas(void *h, void *out) { RefScene *s = static_cast<RefScene *>(h); std::memcpy(out, s->tlas.data(), s->tlas.size() * sizeof(TLASNode)); }

uint32_t refbvh_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(BVHNode);
    case 1: return sizeof(Triangle);
    case 2: return sizeof(BLASInstance);
    case 3: return sizeof(TLASNode);
    default: return 0;
    }
}

} // extern "C"
