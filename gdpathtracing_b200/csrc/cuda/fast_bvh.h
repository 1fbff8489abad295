// fast_bvh.h -- host-side construction of the CLOSEST-HIT tables (pt_fast.cuh).
//
// The reference traversal (main.glsl:270-350) returns, for one ray, the triangle with the smallest
// accepted t among the triangles it reaches; on equal t the one tested last wins (main.glsl:247).
// Which triangles it reaches depends on its visiting order only through ties and through box tests
// that sit exactly on the current hit distance.  The rendering kernels therefore answer a ray in two
// steps (pt_fast.cuh):
//   1. an order-free closest-hit search over OUR OWN acceleration structure -- a binned-SAH BVH over
//      the same triangles with true (inflated) bounds, built here -- that evaluates the reference's
//      Moller-Trumbore arithmetic on the reference's vertices and the reference's instance-local ray,
//      and notices ties;
//   2. a proof that the reference traversal reaches that triangle: the reference box of the triangle's
//      leaf (and of its instance's TLAS leaf) is entered strictly before the hit.  Reference boxes
//      are nested (a child's box is computed from a subset of its parent's triangles, bvh.cpp:24-37,
//      108-127; TLAS boxes are unions, bvh.cpp:299-304) and the slab arithmetic is monotone in the box,
//      so every ancestor is then entered no later, i.e. pushed while hit.t is still larger.
// A ray whose proof fails (or that saw a tie) is re-traced with the exact reference-order traversal
// (pt_trace.cuh), so results are bit-identical either way; only the amount of work differs.
// The nesting the proof relies on is verified here for every edge of the uploaded arrays; if it does
// not hold (hand-made arrays), `ok` is false and the kernels use the reference-order traversal only.
//
// Pure C++ (no CUDA) so tests/devcheck can reuse it.
#ifndef GDPT_FAST_BVH_H
#define GDPT_FAST_BVH_H

#include "derived_layout.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

namespace gdpt {

struct FastLayout {
    std::vector<FastNode> nodes;     // internal nodes of every BLAS
    std::vector<FastTri> tris;       // triangle copies in leaf order
    std::vector<FastNode> tlas;      // TLAS internal nodes: reference topology, true world boxes (appended to `nodes`
                                     // at `tlas_base` when the layout is complete: one table for both levels)
    uint32_t tlas_base = 0;
    std::vector<uint32_t> tri_leaf;  // per reference triangle index: reference node index of the leaf that holds it
    std::vector<uint32_t> inst_root; // per instance: link of the root of its BLAS in `nodes`/`tris`
    uint32_t tlas_root_link = LINK_NONE;
    uint32_t max_depth = 0;          // deepest root-to-leaf chain of our trees (stack bound)
    int build_threads = 0;           // input: threads of the tree build, 0 = all hardware threads (same tables for any value)
    // four-wide form (collapse of the trees above): one table for both levels
    std::vector<FastNode4> nodes4;
    std::vector<uint32_t> inst_root4; // per instance: link of its BLAS root in nodes4
    uint32_t root4 = LINK_NONE;       // TLAS root link in nodes4
    uint32_t need4 = 0;               // stack entries the four-wide search can need
    bool ok4 = false;
    bool ok = false;
    std::string why_not;
};

namespace fastbvh {

constexpr int kBins = 16;
constexpr uint32_t kLeafMaxDefault = 2; // A/B on the B200, C2 path kernel ms | C3 ms: round 1 (majority-phase loop) 2: 0.799, 3: 0.788, 4: 0.807, 6: 0.825, 8: 0.813; round 2 (every phase per iteration, margins / 4096) 2: 0.596 | 4.78, 3: 0.611 | 5.22, 4: 0.643 | 5.51
// triangles per leaf of our trees (1..8, the leaf link holds count-1 in 3 bits); -DGDPT_FAST_LEAF_MAX=n for A/B builds
#ifndef GDPT_FAST_LEAF_MAX
#define GDPT_FAST_LEAF_MAX kLeafMaxDefault
#endif
// A mesh of kLeafOneFrom triangles and more gets one triangle per leaf: its rays spend their time in deep, small boxes, and
// a second triangle per leaf is a second Moller-Trumbore for rays that mostly miss both (C3, 1 M-triangle soup: path kernel
// 4.78 -> 4.25 ms with 1-triangle leaves; C2's meshes, 1 k-50 k triangles: 0.596 -> 0.669, so they keep 2).
constexpr uint32_t kLeafOneFrom = 262144u;
inline uint32_t leaf_max(uint32_t mesh_triangles)
{
    static_assert((uint32_t)(GDPT_FAST_LEAF_MAX) >= 1u && (uint32_t)(GDPT_FAST_LEAF_MAX) <= 8u, "1..8 triangles per leaf");
#ifdef GDPT_FAST_LEAF_NO_RULE
    (void)mesh_triangles;
    return (uint32_t)(GDPT_FAST_LEAF_MAX);
#else
    return mesh_triangles >= kLeafOneFrom ? 1u : (uint32_t)(GDPT_FAST_LEAF_MAX);
#endif
}

struct Prim { float lo[3], hi[3], c[3]; uint32_t orig; };

inline TightBox bounds_of(const std::vector<Prim> &p, uint32_t b, uint32_t e)
{
    TightBox t = tight_empty();
    for (uint32_t i = b; i < e; i++) { tight_grow(t, p[i].lo); tight_grow(t, p[i].hi); }
    return t;
}
inline float half_area(const TightBox &b)
{
    const float x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}

// run fn(0..parts-1) on `threads` threads (the caller is one of them); parts are claimed dynamically
template <class Fn> inline void parallel_parts(int threads, int parts, const Fn &fn)
{
    if (threads <= 1 || parts <= 1) { for (int p = 0; p < parts; p++) fn(p); return; }
    std::atomic<int> next(0);
    auto work = [&] { for (;;) { const int p = next.fetch_add(1); if (p >= parts) break; fn(p); } };
    std::vector<std::thread> pool;
    const int extra = std::min(threads, parts) - 1;
    for (int i = 0; i < extra; i++) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
}

struct Builder {
    const gdpt_triangle_geometry *tris;
    std::vector<Prim> prims;
    FastLayout *out;
    float owner_extent = 0.0f;
    uint32_t leaf_cap = 1;  // triangles per leaf (leaf_max of the mesh)
    uint32_t depth_seen = 0;
    // multi-threaded form: ranges of at most `cut_at` primitives become tasks built into private tables, which are
    // placed behind one another in the depth-first order the single-threaded recursion would have produced
    enum : uint32_t { kTaskMark = 0x3F000000u }; // internal-link value space above any real node index
    struct Task { uint32_t b, e, depth, link = 0, depth_seen = 0, node_base = 0, tri_base = 0; FastLayout part; };
    std::vector<Task> tasks;
    uint32_t cut_at = 0; // 0 = single-threaded
    int threads = 1;
    enum : uint32_t { kChunk = 16384u, kParallelFrom = 65536u };
    // folds over a large range run in fixed chunks merged in chunk order: the same bits for any thread count
    template <class Fn> void for_chunks(uint32_t b, uint32_t e, const Fn &fn) const
    {
        const int parts = (int)((e - b + kChunk - 1u) / kChunk);
        parallel_parts(threads, parts, [&](int p) { fn(b + (uint32_t)p * kChunk, std::min(e, b + (uint32_t)(p + 1) * kChunk), p); });
    }
    TightBox bounds_parallel(uint32_t b, uint32_t e, bool centroids) const
    {
        const int parts = (int)((e - b + kChunk - 1u) / kChunk);
        std::vector<TightBox> part((size_t)parts);
        for_chunks(b, e, [&](uint32_t cb, uint32_t ce, int p) {
            TightBox t = tight_empty();
            for (uint32_t i = cb; i < ce; i++) {
                if (centroids) tight_grow(t, prims[i].c);
                else { tight_grow(t, prims[i].lo); tight_grow(t, prims[i].hi); }
            }
            part[(size_t)p] = t;
        });
        TightBox all = tight_empty();
        for (const TightBox &t : part) tight_merge(all, t);
        return all;
    }

    // builds [b, e) and returns its link; *box receives the true bounds
    uint32_t build(uint32_t b, uint32_t e, uint32_t depth, TightBox *box)
    {
        if (depth > depth_seen) depth_seen = depth;
        const uint32_t n = e - b;
        const bool wide_range = threads > 1 && n >= kParallelFrom;
        *box = wide_range ? bounds_parallel(b, e, false) : bounds_of(prims, b, e);
        if (cut_at && n <= cut_at) { // hand the whole range to a task; its tables are placed later (place_tasks)
            tasks.emplace_back();
            tasks.back().b = b; tasks.back().e = e; tasks.back().depth = depth;
            return kTaskMark | (uint32_t)(tasks.size() - 1);
        }
        if (n <= leaf_cap) {
            const uint32_t first = (uint32_t)out->tris.size();
            for (uint32_t i = b; i < e; i++) {
                FastTri t;
                std::memset(&t, 0, sizeof(t));
                const gdpt_triangle_geometry &g = tris[prims[i].orig];
                for (int k = 0; k < 3; k++) { t.v0[k] = g.v[0][k]; t.v1[k] = g.v[1][k]; t.v2[k] = g.v[2][k]; }
                t.orig = prims[i].orig;
                out->tris.push_back(t);
            }
            return LINK_LEAF | ((n - 1u) << FAST_LEAF_COUNT_SHIFT) | first;
        }
        // binned SAH over the centroid bounds, three axes
        TightBox cb = tight_empty();
        if (wide_range) cb = bounds_parallel(b, e, true);
        else for (uint32_t i = b; i < e; i++) tight_grow(cb, prims[i].c);
        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        for (int axis = 0; axis < 3; axis++) {
            const float lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.0f) || !std::isfinite(ext)) continue;
            TightBox bb[kBins];
            uint32_t cnt[kBins];
            for (int k = 0; k < kBins; k++) { bb[k] = tight_empty(); cnt[k] = 0; }
            const float scale = (float)kBins / ext;
            auto bin_range = [&](uint32_t rb, uint32_t re, TightBox *obb, uint32_t *ocnt) {
                for (uint32_t i = rb; i < re; i++) {
                    int k = (int)((prims[i].c[axis] - lo) * scale);
                    k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                    ocnt[k]++; tight_grow(obb[k], prims[i].lo); tight_grow(obb[k], prims[i].hi);
                }
            };
            if (wide_range) {
                const int parts = (int)((n + kChunk - 1u) / kChunk);
                std::vector<TightBox> pbb((size_t)parts * kBins, tight_empty());
                std::vector<uint32_t> pcnt((size_t)parts * kBins, 0u);
                for_chunks(b, e, [&](uint32_t rb, uint32_t re, int p) { bin_range(rb, re, &pbb[(size_t)p * kBins], &pcnt[(size_t)p * kBins]); });
                for (int p = 0; p < parts; p++)
                    for (int k = 0; k < kBins; k++) {
                        if (pcnt[(size_t)p * kBins + k]) tight_merge(bb[k], pbb[(size_t)p * kBins + k]);
                        cnt[k] += pcnt[(size_t)p * kBins + k];
                    }
            } else {
                bin_range(b, e, bb, cnt);
            }
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            TightBox acc = tight_empty();
            uint32_t c = 0;
            for (int k = kBins - 1; k > 0; k--) {
                if (cnt[k]) tight_merge(acc, bb[k]);
                c += cnt[k];
                right_area[k] = c ? half_area(acc) : 0.0f; right_cnt[k] = c;
            }
            acc = tight_empty(); c = 0;
            for (int k = 0; k < kBins - 1; k++) {
                if (cnt[k]) tight_merge(acc, bb[k]);
                c += cnt[k];
                if (c == 0 || right_cnt[k + 1] == 0) continue;
                const float cost = half_area(acc) * (float)c + right_area[k + 1] * (float)right_cnt[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = k; }
            }
        }
        uint32_t mid = b;
        if (best_axis >= 0) {
            const float lo = cb.lo[best_axis], scale = (float)kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
            const int axis = best_axis, bin = best_bin;
            auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](const Prim &p) {
                int k = (int)((p.c[axis] - lo) * scale);
                k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                return k <= bin;
            });
            mid = (uint32_t)(it - prims.begin());
        }
        if (mid == b || mid == e) { // all centroids coincide (or non-finite input): split the index range
            int axis = 0;
            for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
            mid = b + n / 2;
            std::nth_element(prims.begin() + b, prims.begin() + mid, prims.begin() + e,
                             [axis](const Prim &x, const Prim &y) { return x.c[axis] < y.c[axis]; });
        }
        const uint32_t idx = (uint32_t)out->nodes.size();
        out->nodes.push_back(FastNode());
        TightBox lb, rb;
        const uint32_t l = build(b, mid, depth + 1, &lb);
        const uint32_t r = build(mid, e, depth + 1, &rb);
        FastNode &nd = out->nodes[idx];
        std::memset(&nd, 0, sizeof(nd));
        const TightBox li = tight_inflate(lb, owner_extent), ri = tight_inflate(rb, owner_extent);
        for (int k = 0; k < 3; k++) { nd.lmin[k] = li.lo[k]; nd.lmax[k] = li.hi[k]; nd.rmin[k] = ri.lo[k]; nd.rmax[k] = ri.hi[k]; }
        nd.left = l; nd.right = r;
        return idx;
    }

    // Builds the tasks on `threads` threads and splices their tables into out->nodes / out->tris.  `first_node` /
    // `first_tri`: sizes of the shared tables when this tree's top was started.  Returns the final root link.
    uint32_t finish_tasks(uint32_t root_link, int threads, uint32_t first_node)
    {
        if (tasks.empty()) return root_link;
        parallel_parts(threads, (int)tasks.size(), [&](int k) {
            Task &t = tasks[(size_t)k];
            Builder sub;
            sub.tris = tris; sub.out = &t.part; sub.owner_extent = owner_extent; sub.leaf_cap = leaf_cap;
            sub.prims.assign(prims.begin() + t.b, prims.begin() + t.e);
            TightBox box;
            t.link = sub.build(0, t.e - t.b, t.depth, &box);
            t.depth_seen = sub.depth_seen;
        });
        // depth-first order: a top node comes before everything below it, children left to right.  Top nodes were
        // appended to out->nodes in exactly that order, with the task blocks missing: walk the top and assign bases.
        const std::vector<FastNode> top(out->nodes.begin() + first_node, out->nodes.end());
        std::vector<uint32_t> top_final(top.size(), 0u);
        uint32_t next_node = first_node, next_tri = (uint32_t)out->tris.size();
        struct Visit { uint32_t link; };
        std::vector<uint32_t> stack;
        auto is_task = [](uint32_t l) { return (l & LINK_LEAF) == 0u && (l & 0x3F000000u) == 0x3F000000u && l != LINK_NONE; };
        stack.push_back(root_link);
        while (!stack.empty()) {
            const uint32_t l = stack.back();
            stack.pop_back();
            if (is_task(l)) {
                Task &t = tasks[l & 0x00FFFFFFu];
                t.node_base = next_node; t.tri_base = next_tri;
                next_node += (uint32_t)t.part.nodes.size(); next_tri += (uint32_t)t.part.tris.size();
                if (t.depth_seen > depth_seen) depth_seen = t.depth_seen;
            } else if ((l & LINK_LEAF) == 0u) {
                const uint32_t ti = l - first_node;
                top_final[ti] = next_node++;
                stack.push_back(top[ti].right); // right below left: left is popped (numbered) first
                stack.push_back(top[ti].left);
            }
        }
        auto resolve = [&](uint32_t l) -> uint32_t {
            if (is_task(l)) {
                const Task &t = tasks[l & 0x00FFFFFFu];
                return (t.link & LINK_LEAF) ? t.link + t.tri_base : t.link + t.node_base;
            }
            return (l & LINK_LEAF) ? l : top_final[l - first_node];
        };
        out->nodes.resize(next_node);
        out->tris.resize(next_tri);
        for (size_t i = 0; i < top.size(); i++) {
            FastNode nd = top[i];
            nd.left = resolve(top[i].left); nd.right = resolve(top[i].right);
            out->nodes[top_final[i]] = nd;
        }
        parallel_parts(threads, (int)tasks.size(), [&](int k) {
            const Task &t = tasks[(size_t)k];
            for (size_t i = 0; i < t.part.nodes.size(); i++) {
                FastNode nd = t.part.nodes[i];
                nd.left += (nd.left & LINK_LEAF) ? t.tri_base : t.node_base;
                nd.right += (nd.right & LINK_LEAF) ? t.tri_base : t.node_base;
                out->nodes[t.node_base + i] = nd;
            }
            std::copy(t.part.tris.begin(), t.part.tris.end(), out->tris.begin() + t.tri_base);
        });
        const uint32_t final_root = resolve(root_link);
        tasks.clear();
        return final_root;
    }
};

// Four-wide collapse: a node takes its two children and, while it has fewer than four, replaces the internal child
// with the largest box by that child's two children.  Boxes and leaves are the two-wide tree's, so the search visits
// the same triangles; only the number of steps (and dependent loads) per ray drops.
struct Collapse {
    FastLayout *o;
    std::vector<uint32_t> done_blas; // two-wide BLAS node index -> four-wide link (0xFFFFFFFE = not yet)
    struct Child { uint32_t link; float lo[3], hi[3]; };
    const FastNode &node2(uint32_t link) const { return (link & LINK_TLAS) ? o->tlas[link & LINK_INDEX_MASK] : o->nodes[link & LINK_INDEX_MASK]; }
    static bool internal(uint32_t link) { return link != LINK_NONE && (link & LINK_LEAF) == 0u; }
    static void push2(Child *c, int &n, const FastNode &f)
    {
        if (f.left != LINK_NONE) { c[n].link = f.left; for (int k = 0; k < 3; k++) { c[n].lo[k] = f.lmin[k]; c[n].hi[k] = f.lmax[k]; } n++; }
        if (f.right != LINK_NONE) { c[n].link = f.right; for (int k = 0; k < 3; k++) { c[n].lo[k] = f.rmin[k]; c[n].hi[k] = f.rmax[k]; } n++; }
    }
    uint32_t run(uint32_t link2, uint32_t *need)
    {
        *need = 0;
        if (!internal(link2)) return link2;
        Child c[5];
        int n = 0;
        push2(c, n, node2(link2));
        while (n > 0 && n < 4) {
            int best = -1;
            float best_area = -1.0f;
            for (int i = 0; i < n; i++) {
                if (!internal(c[i].link)) continue;
                const float x = c[i].hi[0] - c[i].lo[0], y = c[i].hi[1] - c[i].lo[1], z = c[i].hi[2] - c[i].lo[2];
                const float area = x * y + y * z + z * x;
                if (best < 0 || area > best_area) { best = i; best_area = area; }
            }
            if (best < 0) break;
            const FastNode &f = node2(c[best].link);
            Child pair[2];
            int m = 0;
            push2(pair, m, f);
            if (m == 0) { c[best] = c[--n]; continue; }
            c[best] = pair[0];
            if (m == 2) c[n++] = pair[1];
        }
        const uint32_t idx = (uint32_t)o->nodes4.size();
        o->nodes4.push_back(FastNode4());
        uint32_t links[4] = { LINK_NONE, LINK_NONE, LINK_NONE, LINK_NONE }, worst = 0;
        for (int i = 0; i < n; i++) {
            uint32_t below = 0;
            links[i] = run(c[i].link, &below);
            if (below > worst) worst = below;
        }
        FastNode4 &nd = o->nodes4[idx];
        std::memset(&nd, 0, sizeof(nd));
        for (int i = 0; i < 4; i++) {
            nd.link[i] = links[i];
            if (i < n) { nd.lox[i] = c[i].lo[0]; nd.loy[i] = c[i].lo[1]; nd.loz[i] = c[i].lo[2]; nd.hix[i] = c[i].hi[0]; nd.hiy[i] = c[i].hi[1]; nd.hiz[i] = c[i].hi[2]; }
        }
        *need = (n > 0 ? (uint32_t)(n - 1) : 0u) + worst;
        return idx | (link2 & LINK_TLAS);
    }
};

inline bool box_inside(const float *cmin, const float *cmax, const float *pmin, const float *pmax)
{
    for (int k = 0; k < 3; k++)
        if (!(cmin[k] >= pmin[k]) || !(cmax[k] <= pmax[k])) return false; // NaN fails
    return true;
}

} // namespace fastbvh

// `lay` is the derived layout of the same arrays (its TLAS records carry the true world boxes).
inline void build_fast_layout(const gdpt_bvh_node *bvh, uint32_t n_nodes, const gdpt_blas_instance *blas, uint32_t n_blas,
                              const gdpt_tlas_node *tlas, uint32_t n_tlas, const gdpt_triangle_geometry *tris, uint32_t n_tris,
                              const DerivedLayout &lay, FastLayout &out)
{
#ifdef GDPT_BUILD_TIMING
    const bool timing = true; // diagnostic builds only: the process environment is never consulted
#else
    const bool timing = false;
#endif
    auto t_lap = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        const auto now = std::chrono::steady_clock::now();
        if (timing) std::fprintf(stderr, "[fast_bvh] %-10s %.3f s\n", what, std::chrono::duration<double>(now - t_lap).count());
        t_lap = now;
    };
    const int want_threads = out.build_threads;
    out = FastLayout();
    out.build_threads = want_threads;
    out.tri_leaf.assign(n_tris, 0xFFFFFFFFu);
    out.inst_root.assign(n_blas, LINK_NONE);
    if (n_tris >= (1u << FAST_LEAF_COUNT_SHIFT)) { out.why_not = "too many triangles for the leaf link encoding"; return; }

    // ---- one tree per distinct BLAS root
    std::vector<uint32_t> root_link(n_nodes, 0xFFFFFFFEu); // 0xFFFFFFFE = not built
    std::vector<uint32_t> walk;
    std::vector<uint32_t> depth_of(n_nodes, 0);
    for (uint32_t b = 0; b < n_blas; b++) {
        const uint32_t root = blas[b].root;
        if (root_link[root] != 0xFFFFFFFEu) { out.inst_root[b] = root_link[root]; continue; }
        fastbvh::Builder bld;
        bld.tris = tris; bld.out = &out;
        uint32_t ref_depth = 0;
        walk.assign(1, root);
        depth_of[root] = 0;
        while (!walk.empty()) {
            const uint32_t i = walk.back();
            walk.pop_back();
            const gdpt_bvh_node &n = bvh[i];
            if (depth_of[i] > ref_depth) ref_depth = depth_of[i];
            if (depth_of[i] > 4096u) { out.why_not = "reference BVH deeper than 4096 levels"; return; }
            if (n.tri_count > 0) {
                for (uint32_t k = 0; k < n.tri_count; k++) {
                    const uint32_t t = n.first_tri_index + k;
                    if (out.tri_leaf[t] != 0xFFFFFFFFu) { out.why_not = "a triangle belongs to two reference leaves"; return; }
                    out.tri_leaf[t] = i;
                    fastbvh::Prim p;
                    for (int a = 0; a < 3; a++) {
                        const float x = tris[t].v[0][a], y = tris[t].v[1][a], z = tris[t].v[2][a];
                        p.lo[a] = std::min(x, std::min(y, z)); p.hi[a] = std::max(x, std::max(y, z));
                        p.c[a] = (x + y + z) * (1.0f / 3.0f);
                    }
                    p.orig = t;
                    bld.prims.push_back(p);
                }
                continue;
            }
            for (int side = 0; side < 2; side++) {
                const uint32_t c = side ? n.right_child : n.left_child;
                // the proof in pt_fast.cuh needs child boxes nested in their parent's (true for bvh.cpp's output)
                if (i != root && !fastbvh::box_inside(bvh[c].aabb_min, bvh[c].aabb_max, n.aabb_min, n.aabb_max)) {
                    out.why_not = "reference BVH boxes are not nested";
                    return;
                }
                depth_of[c] = depth_of[i] + 1;
                walk.push_back(c);
            }
        }
        // the reference keeps 64 stack entries per level without a check (main.glsl:272); a tree this shallow cannot exceed them
        if (ref_depth >= 60u) { out.why_not = "reference BVH too deep to bound its stack use"; return; }
        lap("walk");
        uint32_t link = LINK_NONE;
        if (!bld.prims.empty()) {
            const TightBox all = fastbvh::bounds_of(bld.prims, 0, (uint32_t)bld.prims.size());
            bld.owner_extent = root < lay.root_extent.size() && lay.root_extent[root] > 0.0f ? lay.root_extent[root] : tight_extent(all);
            TightBox box;
            const int threads = out.build_threads > 0 ? out.build_threads : (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
            const uint32_t count = (uint32_t)bld.prims.size();
            bld.threads = threads;
            bld.leaf_cap = fastbvh::leaf_max(count);
            if (threads > 1 && count >= 65536u) bld.cut_at = std::max(4096u, count / (uint32_t)(threads * 8));
            const uint32_t first_node = (uint32_t)out.nodes.size();
            link = bld.build(0, count, 0, &box);
            lap("top");
            link = bld.finish_tasks(link, threads, first_node);
            lap("tasks");
            if (bld.depth_seen > out.max_depth) out.max_depth = bld.depth_seen;
        }
        root_link[root] = link;
        out.inst_root[b] = link;
    }
    // children of a BLAS root are tested against nothing above them, but they must still nest below the root's children:
    // handled above for every non-root parent; a root's children need no condition (the root is never box-tested).

    // ---- TLAS: reference topology with the true world boxes; nesting of the reference boxes verified per edge
    out.tlas.assign(lay.wide_tlas.size(), FastNode());
    uint32_t tlas_depth = 0;
    {
        std::vector<uint32_t> d(n_tlas, 0);
        walk.assign(1, 0u);
        uint32_t visited = 0;
        while (!walk.empty()) {
            const uint32_t i = walk.back();
            walk.pop_back();
            if (++visited > 2u * n_tlas + 2u) { out.why_not = "TLAS links do not form a tree"; return; }
            if (d[i] > tlas_depth) tlas_depth = d[i];
            if (tlas[i].left_right == 0) continue;
            const uint32_t l = tlas[i].left_right & 0xFFFFu, r = tlas[i].left_right >> 16;
            for (uint32_t c : { l, r }) {
                if (i != 0 && !fastbvh::box_inside(tlas[c].aabb_min, tlas[c].aabb_max, tlas[i].aabb_min, tlas[i].aabb_max)) {
                    out.why_not = "reference TLAS boxes are not nested";
                    return;
                }
                d[c] = d[i] + 1;
                walk.push_back(c);
            }
        }
        if (tlas_depth >= 60u) { out.why_not = "reference TLAS too deep to bound its stack use"; return; }
    }
    for (size_t i = 0; i < lay.wide_tlas.size(); i++) {
        const WideNode &w = lay.wide_tlas[i];
        FastNode &f = out.tlas[i];
        std::memset(&f, 0, sizeof(f));
        for (int k = 0; k < 3; k++) { f.lmin[k] = w.tlmin[k]; f.lmax[k] = w.tlmax[k]; f.rmin[k] = w.trmin[k]; f.rmax[k] = w.trmax[k]; }
        f.left = w.left; f.right = w.right;
    }
    out.tlas_root_link = lay.tlas_root_link;
    out.max_depth += tlas_depth + 2u;
    if (out.max_depth >= GDPT_FAST_MAX_DEPTH) { out.why_not = "closest-hit tree deeper than the traversal stack"; return; }
    lap("tlas");
    {   // four-wide form of every tree (before the TLAS nodes join the two-wide table: Collapse reads both)
        fastbvh::Collapse col;
        col.o = &out;
        out.inst_root4.assign(n_blas, LINK_NONE);
        std::vector<uint32_t> done(out.nodes.size(), 0xFFFFFFFEu);
        uint32_t need_blas = 0;
        for (uint32_t b = 0; b < n_blas; b++) {
            const uint32_t l2 = out.inst_root[b];
            if (!fastbvh::Collapse::internal(l2)) { out.inst_root4[b] = l2; continue; }
            if (done[l2 & LINK_INDEX_MASK] == 0xFFFFFFFEu) {
                uint32_t need = 0;
                done[l2 & LINK_INDEX_MASK] = col.run(l2, &need);
                if (need > need_blas) need_blas = need;
            }
            out.inst_root4[b] = done[l2 & LINK_INDEX_MASK];
        }
        uint32_t need_tlas = 0;
        out.root4 = col.run(out.tlas_root_link, &need_tlas);
        out.need4 = need_tlas + need_blas + 2u;
        out.ok4 = out.need4 < GDPT_FAST_MAX_DEPTH && out.nodes4.size() < (size_t)LINK_INDEX_MASK;
    }
    lap("collapse");
    out.tlas_base = (uint32_t)out.nodes.size();
    out.nodes.insert(out.nodes.end(), out.tlas.begin(), out.tlas.end());
    out.ok = true;
}

} // namespace gdpt
#endif
