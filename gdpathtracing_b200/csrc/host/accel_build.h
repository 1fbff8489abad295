// accel_build.h -- host-side BLAS / TLAS construction for the drop-in scene
// upload.  Output arrays are byte-for-byte what the reference's src/bvh emits
// (tests/test_accel_build.py compares against the reference bvh.cpp compiled in
// place); that byte layout is the substrate the traversal parity is defined on.
#ifndef GDPT_ACCEL_BUILD_H
#define GDPT_ACCEL_BUILD_H

#include "gdpt_wire.h"
#include "xform_math.h"

#include <cstdint>
#include <vector>

namespace gdpt {

// One mesh surface as Godot hands it out (Mesh::surface_get_arrays):
// positions/normals xyz, uvs xy, 32-bit indices, three per triangle.
struct SurfaceArrays {
    const float *positions = nullptr;
    const float *normals = nullptr;
    const float *uvs = nullptr;
    int64_t vertex_count = 0;
    const int32_t *indices = nullptr;
    int64_t index_count = 0;
};

class AccelBuilder {
public:
    // Appends the mesh's triangles to `triangles`, builds its BLAS into `nodes`
    // and returns the root's global node index.  Mirrors BVHBuilder::BuildBVH
    // (src/bvh/bvh.cpp:187-223).
    uint32_t build_blas(std::vector<gdpt_bvh_node> &nodes, std::vector<gdpt_build_triangle> &triangles,
                        const SurfaceArrays *surfaces, int n_surfaces) const;

    // BLASInstance::set_materials + set_transform (src/bvh/bvh.h:73-115).
    static gdpt_blas_instance make_instance(uint32_t root, const int *material_ids, int n_material_ids,
                                            const Xform3 &transform, const std::vector<gdpt_bvh_node> &nodes);

    // TLAS::build (src/bvh/bvh.cpp:264-317).
    static void build_tlas(std::vector<gdpt_tlas_node> &tlas, const std::vector<gdpt_blas_instance> &instances);

    // Threads used by build_blas for meshes of 32 k triangles and more: 0 = all hardware threads,
    // 1 = single-threaded like upstream.  The output bytes do not depend on it.
    void set_threads(int n) { threads_ = n; }
    int threads() const { return threads_; }

private:
    int threads_ = 0;
    void subdivide(std::vector<gdpt_bvh_node> &nodes, std::vector<gdpt_build_triangle> &tris, int first, int last) const;
};

} // namespace gdpt
#endif
