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
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

namespace gdpt {

namespace {
// std::min / std::max argument order as used by vec3/vec4::min/max (src/bvh/vec.h:78-86,158-166).
inline float keep_lo(float cur, float p) { return (p < cur) ? p : cur; }
inline float keep_hi(float cur, float p) { return (cur < p) ? p : cur; }
} // namespace

// ------------------------------------------------------------------------------------------------
// BLAS subdivision.  The recursion of BVHBuilder::build_recursive (bvh.cpp:108-185) numbers nodes in
// pre-order and permutes the triangle range in place.  Here the same decisions are taken on 40-byte
// proxies (triangle bounds, centroid, source slot) instead of the 144-byte build triangles, and the
// tree is cut into a sequentially-walked top and many independent subtrees built by a thread pool:
//   * extents and bin contents are min/max/count folds -- any chunking merged in chunk order gives the
//     bits of the sequential fold (first element reaching the extreme wins, the seed wins ties);
//   * the in-place partition and the nth_element fallback depend on comparisons only, so they produce
//     the same permutation on proxies, which is applied to the triangles once at the end;
//   * a subtree is a contiguous block of the pre-order numbering, so blocks built apart are placed by
//     adding their base to the child links.
// Output bytes are those of the single-threaded reference (tests/test_accel_build.py).
namespace {

struct Proxy { float lo[3], hi[3], c[3]; uint32_t src; };
enum { kBins = 8 };

// fork-join over a fixed set of threads; the caller takes part
class WorkerPool {
public:
    explicit WorkerPool(int n_threads)
    {
        for (int i = 1; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~WorkerPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            quit_ = true;
            generation_++;
        }
        wake_.notify_all();
        for (std::thread &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }
    void run(int parts, const std::function<void(int)> &fn)
    {
        if (parts <= 0) return;
        if (workers_.empty() || parts == 1) {
            for (int p = 0; p < parts; p++) fn(p);
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; parts_ = parts; next_.store(0); pending_ = (int)workers_.size();
            generation_++;
        }
        wake_.notify_all();
        drain();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void drain()
    {
        for (;;) {
            const int p = next_.fetch_add(1);
            if (p >= parts_) break;
            (*fn_)(p);
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                wake_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                if (quit_) return;
            }
            drain();
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable wake_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{ 0 };
    int parts_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool quit_ = false;
};

struct Extent3 { float lo[3], hi[3]; };
inline Extent3 seed3()
{
    Extent3 e;
    e.lo[0] = FLT_MAX; e.lo[1] = 0.0f; e.lo[2] = 0.0f;
    e.hi[0] = FLT_MIN; e.hi[1] = 0.0f; e.hi[2] = 0.0f;
    return e;
}
// BoundingBox::extend(point) (bvh.cpp:12-16): both corners move
inline void grow_point(Extent3 &e, const float *p)
{
    for (int a = 0; a < 3; a++) { e.lo[a] = keep_lo(e.lo[a], p[a]); e.hi[a] = keep_hi(e.hi[a], p[a]); }
}
// the three vertices of one triangle, pre-folded: lo can only move the low corner, hi the high one
inline void grow_proxy(Extent3 &e, const Proxy &t)
{
    for (int a = 0; a < 3; a++) { e.lo[a] = keep_lo(e.lo[a], t.lo[a]); e.hi[a] = keep_hi(e.hi[a], t.hi[a]); }
}
// later fold into earlier one (chunk order)
inline void merge_extent(Extent3 &into, const Extent3 &next)
{
    for (int a = 0; a < 3; a++) { into.lo[a] = keep_lo(into.lo[a], next.lo[a]); into.hi[a] = keep_hi(into.hi[a], next.hi[a]); }
}
inline float half_area3(const Extent3 &e)
{
    const float dx = e.hi[0] - e.lo[0], dy = e.hi[1] - e.lo[1], dz = e.hi[2] - e.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

struct AxisBins { Extent3 box[kBins]; int n[kBins]; bool live; float lo, span, inv_span; };
struct NodeBins { AxisBins axis[3]; };

inline void bins_begin(NodeBins &nb, const Extent3 &box)
{
    for (int a = 0; a < 3; a++) {
        AxisBins &ab = nb.axis[a];
        ab.lo = box.lo[a];
        ab.span = box.hi[a] - box.lo[a];
        ab.live = !(ab.span < 1e-6f); // EvaluateSAH returns 1e30 for a flat axis (bvh.cpp:54-55)
        ab.inv_span = ab.live ? 1.0f / ab.span : 0.0f;
        for (int b = 0; b < kBins; b++) { ab.box[b] = seed3(); ab.n[b] = 0; }
    }
}
inline void bins_add(NodeBins &nb, const Proxy *px, int first, int last)
{
    for (int i = first; i < last; i++) {
        const Proxy &t = px[i];
        for (int a = 0; a < 3; a++) {
            AxisBins &ab = nb.axis[a];
            if (!ab.live) continue;
            int b = int(float(kBins) * (t.c[a] - ab.lo) * ab.inv_span);
            b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
            ab.n[b]++;
            grow_proxy(ab.box[b], t);
        }
    }
}
inline void bins_merge(NodeBins &into, const NodeBins &next)
{
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < kBins; b++) {
            merge_extent(into.axis[a].box[b], next.axis[a].box[b]);
            into.axis[a].n[b] += next.axis[a].n[b];
        }
}
// the sweep of EvaluateSAH (bvh.cpp:75-105) over filled bins
inline float sweep_bins(const AxisBins &ab, float &split)
{
    if (!ab.live) return 1e+30f;
    Extent3 prefix_box[kBins];
    int prefix_n[kBins];
    Extent3 run = seed3();
    int run_n = 0;
    for (int b = 0; b < kBins - 1; b++) {
        grow_point(run, ab.box[b].lo);
        grow_point(run, ab.box[b].hi);
        run_n += ab.n[b];
        prefix_box[b] = run;
        prefix_n[b] = run_n;
    }
    float best = 1e+30f;
    Extent3 tail = seed3();
    int tail_n = 0;
    for (int b = kBins - 1; b > 0; b--) {
        grow_point(tail, ab.box[b].lo);
        grow_point(tail, ab.box[b].hi);
        tail_n += ab.n[b];
        const float cost = half_area3(prefix_box[b - 1]) * prefix_n[b - 1] + half_area3(tail) * tail_n;
        if (cost < best) {
            best = cost;
            split = ab.lo + (float(b) / kBins) * ab.span;
        }
    }
    return best;
}

struct Split { bool leaf; int mid; };

// Everything build_recursive does at one node except creating children: fills `node`, partitions the proxies.
// `pool` (may be null) spreads the two folds of a large range over threads.
Split split_node(gdpt_bvh_node &node, Proxy *px, int first, int last, WorkerPool *pool)
{
    const int count = last - first;
    Extent3 box = seed3();
    NodeBins bins;
    const int kChunk = 16384;
    const int parts = (pool && count >= 4 * kChunk) ? (count + kChunk - 1) / kChunk : 1;
    if (parts == 1) {
        for (int i = first; i < last; i++) grow_proxy(box, px[i]);
    } else {
        std::vector<Extent3> part_box((size_t)parts);
        pool->run(parts, [&](int p) {
            Extent3 e = seed3();
            const int a = first + p * kChunk, b = std::min(a + kChunk, last);
            for (int i = a; i < b; i++) grow_proxy(e, px[i]);
            part_box[(size_t)p] = e;
        });
        for (int p = 0; p < parts; p++) merge_extent(box, part_box[(size_t)p]);
    }
    for (int a = 0; a < 3; a++) { node.aabb_min[a] = box.lo[a]; node.aabb_max[a] = box.hi[a]; }
    node.aabb_min[3] = 1.0f; node.aabb_max[3] = 1.0f;
    node.left_child = 0; node.right_child = 0;
    node.first_tri_index = (uint32_t)first;
    node.tri_count = (uint32_t)count;
    if (node.tri_count <= 4) return { true, 0 };

    bins_begin(bins, box);
    if (parts == 1) {
        bins_add(bins, px, first, last);
    } else {
        std::vector<NodeBins> part_bins((size_t)parts);
        pool->run(parts, [&](int p) {
            NodeBins nb;
            bins_begin(nb, box);
            const int a = first + p * kChunk, b = std::min(a + kChunk, last);
            bins_add(nb, px, a, b);
            part_bins[(size_t)p] = nb;
        });
        for (int p = 0; p < parts; p++) bins_merge(bins, part_bins[(size_t)p]);
    }
    float cut = 0.0f, cut_cost = 1e30f;
    int cut_axis = -1;
    for (int axis = 0; axis < 3; axis++) {
        float s = 0.0f;
        const float c = sweep_bins(bins.axis[axis], s);
        if (c < cut_cost) { cut_cost = c; cut = s; cut_axis = axis; }
    }
    const float parent_cost = node.tri_count * half_area3(box);
    if (cut_cost * 0.8f >= parent_cost) return { true, 0 };
    // Upstream reads an unset split value when no axis produced a finite cost and the
    // test above still passes (parent cost > 8e29); nothing meaningful can follow.
    if (cut_axis < 0) return { true, 0 };

    int i = first, j = last - 1;
    while (i <= j) {
        if (px[i].c[cut_axis] < cut) i++;
        else std::swap(px[i], px[j--]);
    }
    const int n_left = i - first;
    if (n_left == 0 || n_left == count) {
        const int mid = first + count / 2;
        std::nth_element(px + first, px + mid, px + last,
                         [cut_axis](const Proxy &a, const Proxy &b) { return a.c[cut_axis] < b.c[cut_axis]; });
        i = mid;
    }
    node.tri_count = 0; // internal from here on
    return { false, i };
}

// Whole subtree of [first,last) appended to `nodes` in pre-order; links are indices into `nodes`.
void subdivide_serial(std::vector<gdpt_bvh_node> &nodes, Proxy *px, int first, int last)
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
        const Split sp = split_node(nodes.back(), px, job.first, job.last, nullptr);
        if (sp.leaf) continue;
        // LIFO: right goes in first so the whole left subtree is numbered before it (pre-order).
        todo.push_back({ sp.mid, job.last, (int64_t)self, 1 });
        todo.push_back({ job.first, sp.mid, (int64_t)self, 0 });
    }
}

} // namespace

void AccelBuilder::subdivide(std::vector<gdpt_bvh_node> &nodes, std::vector<gdpt_build_triangle> &tris, int first,
                             int last) const
{
    const int count = last - first;
#ifdef GDPT_BUILD_TIMING
    const bool timing = true; // diagnostic builds only: the process environment is never consulted
#else
    const bool timing = false;
#endif
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto lap = [&](const char *what, std::chrono::steady_clock::time_point &t) {
        if (timing) std::fprintf(stderr, "[accel_build] %-12s %.3f s\n", what, std::chrono::duration<double>(now() - t).count());
        t = now();
    };
    auto t_lap = now();
    std::vector<Proxy> proxies((size_t)count);
    Proxy *px = proxies.data() - first; // indexed like the triangle array
    int threads = threads_ > 0 ? threads_ : (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, 64));
    if (count < 32768) threads = 1;
    std::unique_ptr<WorkerPool> pool(threads > 1 ? new WorkerPool(threads) : nullptr);
    const int kSlice = 8192;
    const int slices = (count + kSlice - 1) / kSlice;
    auto for_slices = [&](const std::function<void(int)> &fn) {
        if (pool) pool->run(slices, fn);
        else for (int s = 0; s < slices; s++) fn(s);
    };
    for_slices([&](int s) {
        const int a = first + s * kSlice, b = std::min(a + kSlice, last);
        for (int i = a; i < b; i++) {
            const gdpt_build_triangle &t = tris[(size_t)i];
            Proxy &p = px[i];
            for (int k = 0; k < 3; k++) {
                p.lo[k] = keep_lo(keep_lo(t.vertices[0][k], t.vertices[1][k]), t.vertices[2][k]);
                p.hi[k] = keep_hi(keep_hi(t.vertices[0][k], t.vertices[1][k]), t.vertices[2][k]);
                p.c[k] = t.centroid[k];
            }
            p.src = (uint32_t)i;
        }
    });

    lap("proxies", t_lap);
    if (!pool) {
        subdivide_serial(nodes, px, first, last);
    } else {
        // top of the tree, walked in pre-order; ranges at or below `cut_at` triangles become independent subtrees
        const int cut_at = std::max(4096, count / (threads * 8));
        struct Entry { bool is_subtree; uint32_t top; int first, last; uint32_t base; };
        struct Pending { int first, last; int64_t parent; int side; };
        std::vector<Entry> order;                      // pre-order sequence of top nodes and subtrees
        std::vector<gdpt_bvh_node> top;                // top nodes, links hold `order` positions until placed
        std::vector<Pending> todo;
        todo.push_back({ first, last, -1, 0 });
        while (!todo.empty()) {
            const Pending job = todo.back();
            todo.pop_back();
            const uint32_t pos = (uint32_t)order.size();
            if (job.parent >= 0) {
                if (job.side == 0) top[(size_t)job.parent].left_child = pos;
                else top[(size_t)job.parent].right_child = pos;
            }
            if (job.last - job.first <= cut_at && job.parent >= 0) {
                order.push_back({ true, 0u, job.first, job.last, 0u });
                continue;
            }
            const uint32_t self = (uint32_t)top.size();
            order.push_back({ false, self, job.first, job.last, 0u });
            top.emplace_back();
            const Split sp = split_node(top.back(), px, job.first, job.last, pool.get());
            if (sp.leaf) continue;
            todo.push_back({ sp.mid, job.last, (int64_t)self, 1 });
            todo.push_back({ job.first, sp.mid, (int64_t)self, 0 });
        }
        lap("top", t_lap);
        // subtrees in parallel, each numbered from 0
        std::vector<uint32_t> sub_of;
        for (uint32_t i = 0; i < order.size(); i++)
            if (order[i].is_subtree) sub_of.push_back(i);
        std::vector<std::vector<gdpt_bvh_node>> sub(sub_of.size());
        pool->run((int)sub_of.size(), [&](int k) {
            const Entry &e = order[sub_of[(size_t)k]];
            sub[(size_t)k].reserve((size_t)(e.last - e.first) / 2 + 8);
            subdivide_serial(sub[(size_t)k], px, e.first, e.last);
        });
        lap("subtrees", t_lap);
        // placement: pre-order numbering continues from nodes.size()
        uint32_t next = (uint32_t)nodes.size();
        {
            size_t k = 0;
            for (Entry &e : order) {
                e.base = next;
                if (e.is_subtree) { e.top = (uint32_t)k; next += (uint32_t)sub[k++].size(); }
                else next += 1u;
            }
        }
        const size_t base0 = nodes.size();
        nodes.resize((size_t)next);
        pool->run((int)order.size(), [&](int oi) {
            const Entry &e = order[(size_t)oi];
            if (!e.is_subtree) {
                gdpt_bvh_node n = top[e.top];
                if (n.tri_count == 0) { n.left_child = order[n.left_child].base; n.right_child = order[n.right_child].base; }
                nodes[e.base] = n;
            } else {
                const std::vector<gdpt_bvh_node> &src = sub[e.top];
                for (size_t i = 0; i < src.size(); i++) {
                    gdpt_bvh_node n = src[i];
                    if (n.tri_count == 0) { n.left_child += e.base; n.right_child += e.base; }
                    nodes[e.base + i] = n;
                }
            }
        });
        (void)base0;
    }

    lap("placement", t_lap);
    // apply the permutation to the build triangles
    std::unique_ptr<gdpt_build_triangle[]> moved(new gdpt_build_triangle[(size_t)count]);
    for_slices([&](int s) {
        const int a = first + s * kSlice, b = std::min(a + kSlice, last);
        std::memcpy(moved.get() + (a - first), tris.data() + a, (size_t)(b - a) * sizeof(gdpt_build_triangle));
    });
    for_slices([&](int s) {
        const int a = first + s * kSlice, b = std::min(a + kSlice, last);
        for (int i = a; i < b; i++) tris[(size_t)i] = moved[(size_t)(px[i].src - (uint32_t)first)];
    });
    lap("permute", t_lap);
}

uint32_t AccelBuilder::build_blas(std::vector<gdpt_bvh_node> &nodes, std::vector<gdpt_build_triangle> &triangles,
                                  const SurfaceArrays *surfaces, int n_surfaces) const
{
    const int first = (int)triangles.size();
    // BVHBuilder::BuildBVH (bvh.cpp:196-214): surfaces in order, whole triangles only
    std::vector<int64_t> surface_first((size_t)n_surfaces + 1, 0);
    for (int s = 0; s < n_surfaces; s++) surface_first[(size_t)s + 1] = surface_first[(size_t)s] + surfaces[s].index_count / 3;
    const int64_t n_new = surface_first[(size_t)n_surfaces];
    triangles.resize((size_t)first + (size_t)n_new);
    auto fill = [&](int s, int64_t a, int64_t b) {
        const SurfaceArrays &arr = surfaces[s];
        for (int64_t k = a; k < b; k++) {
            gdpt_build_triangle &t = triangles[(size_t)first + (size_t)(surface_first[(size_t)s] + k)];
            std::memset(&t, 0, sizeof(t));
            for (int c = 0; c < 3; c++) {
                const int64_t vi = arr.indices[k * 3 + c];
                for (int j = 0; j < 3; j++) {
                    t.vertices[c][j] = arr.positions[vi * 3 + j];
                    t.normals[c][j] = arr.normals[vi * 3 + j];
                }
                t.vertices[c][3] = 1.0f; t.normals[c][3] = 1.0f;
                t.uvs[c][0] = arr.uvs[vi * 2 + 0]; t.uvs[c][1] = arr.uvs[vi * 2 + 1];
            }
            t.material_index = (uint32_t)s;
            for (int j = 0; j < 4; j++) t.centroid[j] = (t.vertices[0][j] + t.vertices[1][j] + t.vertices[2][j]) * 0.33333333f;
        }
    };
    {
        int threads = threads_ > 0 ? threads_ : (int)std::thread::hardware_concurrency();
        threads = std::max(1, std::min(threads, 64));
        const int64_t kSlice = 16384;
        struct Piece { int s; int64_t a, b; };
        std::vector<Piece> pieces;
        for (int s = 0; s < n_surfaces; s++) {
            const int64_t n = surfaces[s].index_count / 3;
            for (int64_t a = 0; a < n; a += kSlice) pieces.push_back({ s, a, std::min(a + kSlice, n) });
        }
        if (threads > 1 && n_new >= 32768) {
            WorkerPool pool(threads);
            pool.run((int)pieces.size(), [&](int p) { fill(pieces[(size_t)p].s, pieces[(size_t)p].a, pieces[(size_t)p].b); });
        } else {
            for (const Piece &p : pieces) fill(p.s, p.a, p.b);
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
