// accel_build.cpp -- see accel_build.h.
//
// The arithmetic (operation order, comparison direction, the seed values of an
// empty extent) deliberately reproduces the reference builder so that node,
// triangle, instance and TLAS arrays come out byte-identical; the control
// structure is our own (explicit work list instead of recursion).  Quirks kept
// on purpose, with their upstream location:
//   Q1  an "empty" box starts at lo=(FLT_MAX,0,0), hi=(FLT_MIN_POSITIVE,0,0)
//       (src/bvh/bvh.cpp:6-10 with vec4's default arguments, src/bvh/vec.h:49),
//       so every box also contains y=0, z=0 and x>=~0;
//   Q2  8 bins, cost = area*count with half-area xy+yz+zx, split kept only if
//       cost*0.8 < count*parent_area, leaves of <=4 triangles (bvh.cpp:39-150);
//   Q3  in-place partition swapping from the back; median split through
//       std::nth_element when the partition is one-sided (bvh.cpp:153-177);
//   Q4  centroid = (v0+v1+v2)*0.33333333f (bvh.cpp:210);
//   Q5  instance box: homogeneous accumulate starting from (0,0,0,1) and the
//       "* (2/w)" renormalisation (src/bvh/bvh.h:100-110).
#include "accel_build.h"

#include <algorithm>
#include <cfloat>
#include <cstring>

namespace gdpt {

namespace {
// std::min / std::max argument order as used by vec3/vec4::min/max (src/bvh/vec.h:78-86,158-166).
inline float keep_lo(float cur, float p) { return (p < cur) ? p : cur; }
inline float keep_hi(float cur, float p) { return (cur < p) ? p : cur; }
} // namespace

AccelBuilder::Extent AccelBuilder::seed_extent()
{
    Extent e;
    e.lo[0] = FLT_MAX; e.lo[1] = 0.0f; e.lo[2] = 0.0f;
    e.hi[0] = FLT_MIN; e.hi[1] = 0.0f; e.hi[2] = 0.0f;
    return e;
}

void AccelBuilder::grow(Extent &e, const float *p)
{
    for (int a = 0; a < 3; a++) {
        e.lo[a] = keep_lo(e.lo[a], p[a]);
        e.hi[a] = keep_hi(e.hi[a], p[a]);
    }
}

float AccelBuilder::half_area(const Extent &e)
{
    const float dx = e.hi[0] - e.lo[0], dy = e.hi[1] - e.lo[1], dz = e.hi[2] - e.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

float AccelBuilder::binned_sah(const std::vector<gdpt_build_triangle> &tris, const gdpt_bvh_node &node, int axis,
                               float &split) const
{
    enum { kBins = 8 };
    const float lo = node.aabb_min[axis];
    const float span = node.aabb_max[axis] - lo;
    if (span < 1e-6f) return 1e+30f;
    const float inv_span = 1.0f / span;

    Extent bin_box[kBins];
    int bin_n[kBins];
    for (int b = 0; b < kBins; b++) { bin_box[b] = seed_extent(); bin_n[b] = 0; }
    for (uint32_t k = 0; k < node.tri_count; k++) {
        const gdpt_build_triangle &t = tris[node.first_tri_index + k];
        int b = int(float(kBins) * (t.centroid[axis] - lo) * inv_span);
        b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
        bin_n[b]++;
        grow(bin_box[b], t.vertices[0]);
        grow(bin_box[b], t.vertices[1]);
        grow(bin_box[b], t.vertices[2]);
    }

    // prefix boxes/counts over bins 0..6
    Extent prefix_box[kBins];
    int prefix_n[kBins];
    Extent run = seed_extent();
    int run_n = 0;
    for (int b = 0; b < kBins - 1; b++) {
        grow(run, bin_box[b].lo);
        grow(run, bin_box[b].hi);
        run_n += bin_n[b];
        prefix_box[b] = run;
        prefix_n[b] = run_n;
    }
    // suffix sweep evaluates the 7 candidate planes from the right
    float best = 1e+30f;
    Extent tail = seed_extent();
    int tail_n = 0;
    for (int b = kBins - 1; b > 0; b--) {
        grow(tail, bin_box[b].lo);
        grow(tail, bin_box[b].hi);
        tail_n += bin_n[b];
        const float cost = half_area(prefix_box[b - 1]) * prefix_n[b - 1] + half_area(tail) * tail_n;
        if (cost < best) {
            best = cost;
            split = lo + (float(b) / kBins) * span;
        }
    }
    return best;
}

void AccelBuilder::subdivide(std::vector<gdpt_bvh_node> &nodes, std::vector<gdpt_build_triangle> &tris, int first,
                             int last) const
{
    struct Pending { int first, last; int64_t parent; int side; };
    std::vector<Pending> todo;
    todo.push_back({ first, last, -1, 0 });
    while (!todo.empty()) {
        const Pending job = todo.back();
        todo.pop_back();
        const uint32_t self = (uint32_t)nodes.size();
        if (job.parent >= 0) {
            if (job.side == 0) nodes[(size_t)job.parent].left_child = self;
            else nodes[(size_t)job.parent].right_child = self;
        }
        nodes.emplace_back();
        gdpt_bvh_node &node = nodes.back();

        Extent box = seed_extent();
        for (int i = job.first; i < job.last; i++)
            for (int v = 0; v < 3; v++) grow(box, tris[i].vertices[v]);
        for (int a = 0; a < 3; a++) { node.aabb_min[a] = box.lo[a]; node.aabb_max[a] = box.hi[a]; }
        node.aabb_min[3] = 1.0f; node.aabb_max[3] = 1.0f;
        node.left_child = 0; node.right_child = 0;
        node.first_tri_index = (uint32_t)job.first;
        node.tri_count = (uint32_t)(job.last - job.first);
        if (node.tri_count <= 4) continue;

        float cut = 0.0f, cut_cost = 1e30f;
        int cut_axis = -1;
        for (int axis = 0; axis < 3; axis++) {
            float s = 0.0f;
            const float c = binned_sah(tris, node, axis, s);
            if (c < cut_cost) { cut_cost = c; cut = s; cut_axis = axis; }
        }
        const float parent_cost = node.tri_count * half_area(box);
        if (cut_cost * 0.8f >= parent_cost) continue;
        // Upstream reads an unset split value when no axis produced a finite cost and the
        // test above still passes (parent cost > 8e29); nothing meaningful can follow.
        if (cut_axis < 0) continue;

        int i = job.first, j = job.last - 1;
        while (i <= j) {
            if (tris[i].centroid[cut_axis] < cut) i++;
            else std::swap(tris[i], tris[j--]);
        }
        const int n_left = i - job.first;
        if (n_left == 0 || n_left == (int)node.tri_count) {
            const int mid = job.first + (job.last - job.first) / 2;
            std::nth_element(tris.begin() + job.first, tris.begin() + mid, tris.begin() + job.last,
                             [cut_axis](const gdpt_build_triangle &a, const gdpt_build_triangle &b) {
                                 return a.centroid[cut_axis] < b.centroid[cut_axis];
                             });
            i = mid;
        }
        node.tri_count = 0; // internal from here on (`node` stays valid: no push_back since emplace)
        // LIFO: right goes in first so the whole left subtree is numbered before it (pre-order).
        todo.push_back({ i, job.last, (int64_t)self, 1 });
        todo.push_back({ job.first, i, (int64_t)self, 0 });
    }
}

uint32_t AccelBuilder::build_blas(std::vector<gdpt_bvh_node> &nodes, std::vector<gdpt_build_triangle> &triangles,
                                  const SurfaceArrays *surfaces, int n_surfaces) const
{
    const int first = (int)triangles.size();
    for (int s = 0; s < n_surfaces; s++) {
        const SurfaceArrays &a = surfaces[s];
        for (int64_t i = 0; i + 2 < a.index_count; i += 3) { // whole triangles only
            gdpt_build_triangle t;
            std::memset(&t, 0, sizeof(t));
            for (int c = 0; c < 3; c++) {
                const int64_t vi = a.indices[i + c];
                for (int k = 0; k < 3; k++) {
                    t.vertices[c][k] = a.positions[vi * 3 + k];
                    t.normals[c][k] = a.normals[vi * 3 + k];
                }
                t.vertices[c][3] = 1.0f; t.normals[c][3] = 1.0f;
                t.uvs[c][0] = a.uvs[vi * 2 + 0]; t.uvs[c][1] = a.uvs[vi * 2 + 1];
            }
            t.material_index = (uint32_t)s;
            for (int k = 0; k < 4; k++) t.centroid[k] = (t.vertices[0][k] + t.vertices[1][k] + t.vertices[2][k]) * 0.33333333f;
            triangles.push_back(t);
        }
    }
    const int last = (int)triangles.size();
    if (first >= last) return 0; // build_recursive's empty-range result (bvh.cpp:111-112)
    const uint32_t root = (uint32_t)nodes.size();
    subdivide(nodes, triangles, first, last);
    return root;
}

gdpt_blas_instance AccelBuilder::make_instance(uint32_t root, const int *material_ids, int n_material_ids,
                                               const Xform3 &transform, const std::vector<gdpt_bvh_node> &nodes)
{
    gdpt_blas_instance inst;
    std::memset(&inst, 0, sizeof(inst)); // upstream leaves unused material slots indeterminate; we zero them
    inst.root = root;
    for (int i = 0; i < n_material_ids && i < 3; i++) inst.materials[i] = (uint32_t)material_ids[i];
    transform.to_float16(inst.transform);
    transform.affine_inverse().to_float16(inst.inverse_transform);

    const gdpt_bvh_node &n = nodes[root];
    float lo[4] = { 1e34f, 1e34f, 1e34f, 1.0f }, hi[4] = { -1e34f, -1e34f, -1e34f, 1.0f };
    for (int corner = 0; corner < 8; corner++) {
        const float c[4] = { (corner & 1) ? n.aabb_max[0] : n.aabb_min[0], (corner & 2) ? n.aabb_max[1] : n.aabb_min[1],
                             (corner & 4) ? n.aabb_max[2] : n.aabb_min[2], 1.0f };
        float p[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
        for (int r = 0; r < 4; r++)
            for (int k = 0; k < 4; k++) p[r] += inst.transform[k * 4 + r] * c[k];
        const float renorm = 2.0f / p[3];
        for (int r = 0; r < 4; r++) {
            const float v = p[r] * renorm;
            lo[r] = keep_lo(lo[r], v);
            hi[r] = keep_hi(hi[r], v);
        }
    }
    std::memcpy(inst.aabb_min, lo, sizeof(lo));
    std::memcpy(inst.aabb_max, hi, sizeof(hi));
    return inst;
}

namespace {
// TLAS::FindBestMatch (src/bvh/bvh.cpp:319-340): smallest merged half-area, first wins ties.
int closest_partner(const std::vector<gdpt_tlas_node> &tlas, const std::vector<int> &live, int n, int a)
{
    float smallest = 1e30f;
    int best = -1;
    const gdpt_tlas_node &na = tlas[live[a]];
    for (int b = 0; b < n; b++) {
        if (b == a) continue;
        const gdpt_tlas_node &nb = tlas[live[b]];
        float e[3];
        for (int k = 0; k < 3; k++) e[k] = keep_hi(na.aabb_max[k], nb.aabb_max[k]) - keep_lo(na.aabb_min[k], nb.aabb_min[k]);
        const float area = e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
        if (area < smallest) { smallest = area; best = b; }
    }
    return best;
}
} // namespace

void AccelBuilder::build_tlas(std::vector<gdpt_tlas_node> &tlas, const std::vector<gdpt_blas_instance> &instances)
{
    int n = (int)instances.size();
    tlas.clear();
    if (n == 0) return; // upstream dereferences an empty list here; we emit an empty TLAS instead
    tlas.reserve((size_t)n * 2);
    gdpt_tlas_node blank;
    std::memset(&blank, 0, sizeof(blank));
    tlas.push_back(blank); // slot 0 becomes the root at the end

    std::vector<int> live;
    live.reserve(n);
    for (int i = 0; i < n; i++) {
        gdpt_tlas_node leaf = blank;
        for (int k = 0; k < 3; k++) { leaf.aabb_min[k] = instances[i].aabb_min[k]; leaf.aabb_max[k] = instances[i].aabb_max[k]; }
        leaf.blas = (uint32_t)i;
        leaf.left_right = 0;
        live.push_back((int)tlas.size());
        tlas.push_back(leaf);
    }

    // agglomerative clustering: merge mutually-closest pairs
    int a = 0, b = closest_partner(tlas, live, n, a);
    while (n > 1) {
        const int c = closest_partner(tlas, live, n, b);
        if (a == c) {
            const int ia = live[a], ib = live[b];
            gdpt_tlas_node merged = blank; // internal nodes: `blas` is indeterminate upstream, 0 here
            merged.left_right = (uint32_t)(ia + (ib << 16));
            for (int k = 0; k < 3; k++) {
                merged.aabb_min[k] = keep_lo(tlas[ia].aabb_min[k], tlas[ib].aabb_min[k]);
                merged.aabb_max[k] = keep_hi(tlas[ia].aabb_max[k], tlas[ib].aabb_max[k]);
            }
            live[a] = (int)tlas.size();
            tlas.push_back(merged);
            live[b] = live[--n];
            b = closest_partner(tlas, live, n, a);
        } else {
            a = b;
            b = c;
        }
    }
    tlas[0] = tlas[live[a]];
}

} // namespace gdpt
