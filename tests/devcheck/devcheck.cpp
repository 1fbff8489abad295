// devcheck.cpp -- TEST-ONLY host compile of the backend's device functions.
//
// The per-ray functions of gdpathtracing_b200/csrc/cuda/{pt_math,pt_trace,pt_shade}.cuh are
// `__host__ __device__`; this file compiles them with g++ (-ffp-contract=off, matching
// nvcc -fmad=false) and runs the same stage logic the kernels run, one pixel at a time, so
// the CPU test tier can compare them against the oracle without a GPU: derived WideNode
// layout, unified-stack traversal, visit order/hash, shading, RNG, depth and colour packing.
//
// It is NOT part of the product and is never loaded by gdpathtracing_b200: it is built into
// tests/devcheck/_build/ by tests/conftest.py and only the "not gpu" tests call it.  The
// kernel-level scheduling (queues, refill, atomics) is covered by the GPU tests.
#include "derived_layout.h"
#include "fast_bvh.h"
#include "pt_fast.cuh"
#include "pt_post.cuh"
#include "pt_shade.cuh"
#include "pt_trace.cuh"

#include <cstring>
#include <string>
#include <vector>

using namespace gdpt;

namespace {
struct HostStack {
    uint32_t slots[GDPT_MAX_STACK];
    void store(uint32_t i, uint32_t v) { slots[i] = v; }
    uint32_t load(uint32_t i) const { return slots[i]; }
};
} // namespace

static int g_stepwise = 0;
static int g_cull = 0;
static int g_fast = 0;
static uint64_t g_fast_rays = 0, g_fast_retraced = 0, g_fast_ties = 0;

template <bool CULL> static void run_ray(const SceneView &sc, RayState &r, HostStack &st, TraceCounters *tc)
{
    if (g_fast && sc.fast_ok) {
        // closest-hit search + proof; rays that fail it are re-traced in reference order (what the kernels do)
        RayState f = r; // + what fast_ray_begin adds: rays the search is not trusted with go straight to the exact traversal
        if (fast_far_origin(f.wo, sc.fast_world_reach)) f.overflow |= RAY_FAR;
        if (fast_degenerate_dir(f.wd)) f.overflow |= RAY_AXIAL;
        if (g_fast == 2 && sc.fast4_ok) fast_trace_ray4(sc, f, st); // four-wide tables
        else fast_trace_ray(sc, f, st);
        __atomic_add_fetch(&g_fast_rays, 1, __ATOMIC_RELAXED);
        if (fast_result_is_reference(sc, f)) {
            r = f; return; }
        __atomic_add_fetch(&g_fast_retraced, 1, __ATOMIC_RELAXED);
        if (f.overflow & (RAY_TIE | RAY_UNSEARCHED)) __atomic_add_fetch(&g_fast_ties, 1, __ATOMIC_RELAXED);
    }
    if (g_stepwise == 2) trace_ray_compact<true, CULL>(sc, r, st, tc);
    else if (g_stepwise) trace_ray_stepwise<true, CULL>(sc, r, st, tc);
    else trace_ray<true, CULL>(sc, r, st, tc);
}

struct devcheck_scene {
    const void *tri_geom; uint64_t n_tris;
    const void *tri_data;
    const void *materials; uint64_t n_materials;
    const void *bvh; uint64_t n_nodes;
    const void *blas; uint64_t n_blas;
    const void *tlas; uint64_t n_tlas;
    const uint8_t *textures; int32_t tex_w, tex_h, tex_layers, material_ext;
    const uint32_t *surface_materials; // material-breadth extension (gdpt_wire.h), same layout as oracle.OrcScene
};

static void make_view(const devcheck_scene *in, DerivedLayout &lay, FastLayout &fast, SceneView &sc)
{
    std::memset(&sc, 0, sizeof(sc));
    sc.tri_geom = (const gdpt_triangle_geometry *)in->tri_geom; sc.tri_data = (const gdpt_triangle_data *)in->tri_data;
    sc.materials = (const gdpt_material *)in->materials; sc.bvh = (const gdpt_bvh_node *)in->bvh;
    sc.blas = (const gdpt_blas_instance *)in->blas; sc.tlas = (const gdpt_tlas_node *)in->tlas;
    sc.textures = in->textures; sc.tex_w = in->tex_w; sc.tex_h = in->tex_h; sc.tex_layers = in->tex_layers;
    if (in->material_ext) {
        static const std::vector<float> lut = [] {
            std::vector<float> t(256);
            for (int k = 0; k < 256; k++) {
                const double c = k / 255.0;
                t[k] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
            }
            return t;
        }();
        sc.material_ext = 1u; sc.surface_materials = in->surface_materials; sc.srgb_lut = lut.data();
    }
    if (g_fast) {
        build_fast_layout(sc.bvh, (uint32_t)in->n_nodes, sc.blas, (uint32_t)in->n_blas, sc.tlas, (uint32_t)in->n_tlas, sc.tri_geom,
                          (uint32_t)in->n_tris, lay, fast);
        for (size_t b = 0; b < lay.inst_recs.size(); b++) {
            lay.inst_recs[b].fast_root = fast.inst_root[b];
            std::memcpy(&lay.inst_recs[b].tight_min[3], &fast.inst_root4[b], 4);
        }
        sc.fast4 = fast.nodes4.data(); sc.fast4_root = fast.root4; sc.fast4_ok = fast.ok4 ? 1u : 0u;
        sc.fast_nodes = fast.nodes.data(); sc.fast_tlas_base = fast.tlas_base; sc.fast_tris = fast.tris.data();
        sc.tri_leaf = fast.tri_leaf.data(); sc.fast_ok = fast.ok ? 1u : 0u;
        if (!fast.ok) std::fprintf(stderr, "devcheck: closest-hit tables unavailable: %s\n", fast.why_not.c_str());
    }
    sc.wide_nodes = lay.wide_nodes.data(); sc.leaf_recs = lay.leaf_recs.data();
    sc.wide_tlas = lay.wide_tlas.data(); sc.inst_recs = lay.inst_recs.data();
    sc.tlas_root_link = lay.tlas_root_link;
    sc.fast_world_reach = lay.world_reach;
}

extern "C" {

// 0: node-at-a-time traversal, 1: leaves split into single-triangle steps (phase-voting order),
// 2: the compact scheduler's node / instance / triangle steps
void devcheck_set_stepwise(int on) { g_stepwise = on; }
// 0: reference visit order, 1: tight-box culling (hit records must not change)
void devcheck_set_cull(int on) { g_cull = on; }
// 1: closest-hit search + proof, exact re-trace where the proof fails (pt_fast.cuh); hit records must not change
// FNV-1a digests of the closest-hit tables built with `threads` threads (0 = all): [0] two-wide nodes, [1] triangle
// copies, [2] four-wide nodes, [3] roots / depths.  Returns 0 when the tables could be built.
int devcheck_fast_layout_digest(const devcheck_scene *in, int threads, uint64_t *out4)
{
    DerivedLayout lay;
    const std::string err = derive_layout((const gdpt_bvh_node *)in->bvh, (uint32_t)in->n_nodes, (const gdpt_blas_instance *)in->blas,
                                          (uint32_t)in->n_blas, (const gdpt_tlas_node *)in->tlas, (uint32_t)in->n_tlas,
                                          (const gdpt_triangle_geometry *)in->tri_geom, (uint32_t)in->n_tris, lay);
    if (!err.empty()) return 1;
    FastLayout fast;
    fast.build_threads = threads;
    build_fast_layout((const gdpt_bvh_node *)in->bvh, (uint32_t)in->n_nodes, (const gdpt_blas_instance *)in->blas, (uint32_t)in->n_blas,
                      (const gdpt_tlas_node *)in->tlas, (uint32_t)in->n_tlas, (const gdpt_triangle_geometry *)in->tri_geom,
                      (uint32_t)in->n_tris, lay, fast);
    if (!fast.ok || !fast.ok4) return 2;
    auto fnv = [](const void *p, size_t n) {
        uint64_t h = 1469598103934665603ull;
        const uint8_t *b = static_cast<const uint8_t *>(p);
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
        return h;
    };
    out4[0] = fnv(fast.nodes.data(), fast.nodes.size() * sizeof(FastNode));
    out4[1] = fnv(fast.tris.data(), fast.tris.size() * sizeof(FastTri));
    out4[2] = fnv(fast.nodes4.data(), fast.nodes4.size() * sizeof(FastNode4));
    const uint64_t tail[4] = { fast.inst_root.empty() ? 0u : fast.inst_root[0], fast.root4, fast.max_depth, fast.need4 };
    out4[3] = fnv(tail, sizeof(tail));
    return 0;
}
// Structural check of the four-wide tables: below every BLAS root each triangle copy is reached exactly once, every
// child box contains the vertices of all triangles below it, and the stack bound covers the deepest chain.
// Returns 0 when all of that holds, otherwise a code (10 + kind of violation).
int devcheck_fast4_structure(const devcheck_scene *in, uint64_t *out_stats3)
{
    DerivedLayout lay;
    if (!derive_layout((const gdpt_bvh_node *)in->bvh, (uint32_t)in->n_nodes, (const gdpt_blas_instance *)in->blas, (uint32_t)in->n_blas,
                       (const gdpt_tlas_node *)in->tlas, (uint32_t)in->n_tlas, (const gdpt_triangle_geometry *)in->tri_geom,
                       (uint32_t)in->n_tris, lay).empty()) return 1;
    FastLayout fast;
    build_fast_layout((const gdpt_bvh_node *)in->bvh, (uint32_t)in->n_nodes, (const gdpt_blas_instance *)in->blas, (uint32_t)in->n_blas,
                      (const gdpt_tlas_node *)in->tlas, (uint32_t)in->n_tlas, (const gdpt_triangle_geometry *)in->tri_geom,
                      (uint32_t)in->n_tris, lay, fast);
    if (!fast.ok || !fast.ok4) return 2;
    std::vector<uint32_t> seen(fast.tris.size(), 0u);
    std::vector<uint32_t> roots_done; // instances of one mesh share a root
    uint64_t nodes_walked = 0, children = 0, deepest = 0;
    struct Item { uint32_t link; float lo[3], hi[3]; bool boxed; uint32_t depth; };
    for (uint32_t b = 0; b < (uint32_t)in->n_blas; b++) {
        const uint32_t root = fast.inst_root4[b];
        if (root == LINK_NONE) continue;
        if (std::find(roots_done.begin(), roots_done.end(), root) != roots_done.end()) continue;
        roots_done.push_back(root);
        std::vector<Item> todo;
        Item first; first.link = root; first.boxed = false; first.depth = 0;
        todo.push_back(first);
        while (!todo.empty()) {
            const Item it = todo.back();
            todo.pop_back();
            if (it.depth > deepest) deepest = it.depth;
            if (it.link & LINK_TLAS) return 11;
            // collect the triangles below this link and check them against the box it was given
            if (it.link & LINK_LEAF) {
                const uint32_t f = it.link & FAST_LEAF_FIRST_MASK, n = ((it.link >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u;
                if (n > 8u || f + n > fast.tris.size()) return 12;
                for (uint32_t k = f; k < f + n; k++) {
                    seen[k]++;
                    const FastTri &t = fast.tris[k];
                    const float *v[3] = { t.v0, t.v1, t.v2 };
                    if (it.boxed)
                        for (int c = 0; c < 3; c++)
                            for (int a = 0; a < 3; a++)
                                if (!(v[c][a] >= it.lo[a] && v[c][a] <= it.hi[a])) return 13;
                }
                continue;
            }
            const uint32_t idx = it.link & LINK_INDEX_MASK;
            if (idx >= fast.nodes4.size()) return 14;
            const FastNode4 &nd = fast.nodes4[idx];
            nodes_walked++;
            int live = 0;
            for (int c = 0; c < 4; c++) {
                if (nd.link[c] == LINK_NONE) continue;
                live++; children++;
                Item ch; ch.link = nd.link[c]; ch.boxed = true; ch.depth = it.depth + 1;
                ch.lo[0] = nd.lox[c]; ch.lo[1] = nd.loy[c]; ch.lo[2] = nd.loz[c]; ch.hi[0] = nd.hix[c]; ch.hi[1] = nd.hiy[c]; ch.hi[2] = nd.hiz[c];
                if (it.boxed) { // boxes are independent inflations of true bounds: triangles are checked against the
                    // intersection of all boxes on their chain
                    for (int a = 0; a < 3; a++) { ch.lo[a] = ch.lo[a] > it.lo[a] ? ch.lo[a] : it.lo[a]; ch.hi[a] = ch.hi[a] < it.hi[a] ? ch.hi[a] : it.hi[a]; }
                }
                todo.push_back(ch);
            }
            if (live < 2) return 15;
        }
    }
    for (size_t k = 0; k < seen.size(); k++)
        if (seen[k] != 1u) return 16;
    if (deepest * 3u + 2u > GDPT_FAST_MAX_DEPTH + 0u && fast.need4 >= GDPT_FAST_MAX_DEPTH) return 17;
    out_stats3[0] = nodes_walked; out_stats3[1] = children; out_stats3[2] = deepest;
    return 0;
}
void devcheck_set_fast(int on) { g_fast = on; g_fast_rays = g_fast_retraced = g_fast_ties = 0; }
void devcheck_fast_counts(uint64_t *out3) { out3[0] = g_fast_rays; out3[1] = g_fast_retraced; out3[2] = g_fast_ties; }


// Renders rows [y_begin, y_end) with the device functions.  Same outputs as orc_path_trace.
int devcheck_path_trace(const devcheck_scene *in, const gdpt_render_params *params, const gdpt_camera *cam, int max_depth,
                        int debug_steps, int y_begin, int y_end, uint8_t *out_rgba8, float *out_depth,
                        gdpt_trace_record *trace, int trace_segments, uint32_t *visits, uint32_t visits_per_ray,
                        uint64_t *out_rays)
{
    DerivedLayout lay;
    const std::string err = derive_layout((const gdpt_bvh_node *)in->bvh, (uint32_t)in->n_nodes, (const gdpt_blas_instance *)in->blas,
                                          (uint32_t)in->n_blas, (const gdpt_tlas_node *)in->tlas, (uint32_t)in->n_tlas,
                                          (const gdpt_triangle_geometry *)in->tri_geom, (uint32_t)in->n_tris, lay);
    if (!err.empty()) return -1;
    FastLayout fast;
    SceneView sc;
    make_view(in, lay, fast, sc);

    const int W = params->width, H = params->height;
    uint64_t rays = 0;
    if (trace)
        for (size_t i = 0; i < (size_t)trace_segments * W * H; i++) { std::memset(&trace[i], 0, sizeof(trace[i])); trace[i].hit = 0xFFFFFFFFu; }
    HostStack st;
    for (int y = y_begin; y < y_end && y < H; y++) {
        for (int x = 0; x < W; x++) {
            const uint32_t pixel = (uint32_t)y * W + x;
            f3 o, d;
            u2 seed = generate_primary_ray(*cam, W, H, x, y, &o, &d);
            f3 radiance = mk3(0, 0, 0), throughput = mk3(1, 1, 1);
            float depth = cam->z_far;
            const int n_seg = debug_steps ? 1 : max_depth;
            for (int i = 0; i < n_seg; i++) {
                RayState r;
                ray_begin(r, sc, o, d);
                TraceCounters tc;
                counters_init(tc, (visits && i == 0) ? visits + (size_t)pixel * visits_per_ray : nullptr, visits_per_ray);
                if (g_cull) run_ray<true>(sc, r, st, &tc);
                else run_ray<false>(sc, r, st, &tc);
                rays++;
                const bool hit = r.t < 1e9f;
                if (trace && i < trace_segments) {
                    gdpt_trace_record &rec = trace[(size_t)i * W * H + pixel];
                    rec.hit = hit ? 1u : 0u; rec.triangle = hit ? r.tri : 0u;
                    rec.blas = hit ? hit_blas(r.blas_front) : 0u; rec.front = hit ? (r.blas_front >> 31) : 0u;
                    rec.t = r.t; rec.u = hit ? r.u : 0.0f; rec.v = hit ? r.v : 0.0f;
                    rec.node_pops = tc.node_pops; rec.box_tests = tc.box_tests; rec.tri_tests = tc.tri_tests;
                    rec.tlas_leaves = tc.tlas_leaves; rec.max_stack = tc.max_stack;
                    rec.visit_hash_lo = (uint32_t)tc.hash; rec.visit_hash_hi = (uint32_t)(tc.hash >> 32);
                }
                if (debug_steps) {
                    float e = (float)tc.tri_tests / 256.0f;
                    e = e < 0.0f ? 0.0f : (e > 1.0f ? 1.0f : e);
                    radiance = mk3(e, e, e);
                    break;
                }
                if (!hit) { radiance = radiance + throughput * sample_sky(r.wd); break; }
                BounceResult br = shade_and_bounce(sc, o, d, r.t, r.u, r.v, r.tri, r.blas_front, radiance, throughput, seed);
                radiance = br.radiance;
                if (i == 0) depth = br.first_hit_distance;
                if (!br.alive) break;
                throughput = br.throughput; o = br.next_o; d = br.next_d;
            }
            const uint32_t packed = pack_rgba8(radiance);
            std::memcpy(out_rgba8 + (size_t)pixel * 4, &packed, 4);
            if (out_depth) out_depth[pixel] = encode_depth(*cam, depth);
        }
    }
    if (out_rays) *out_rays = rays;
    return 0;
}

void devcheck_rng(uint32_t px, uint32_t py, uint32_t frame, uint32_t *seed_out2, float *r_out2, uint32_t *state_out2)
{
    u2 s = prng_seed(px, py, frame);
    seed_out2[0] = s.x; seed_out2[1] = s.y;
    f2 r = pcg2d(s);
    r_out2[0] = r.x; r_out2[1] = r.y;
    state_out2[0] = s.x; state_out2[1] = s.y;
}

void devcheck_sincos(float x, float *out2) { sincos_det(x, &out2[0], &out2[1]); }

void devcheck_progressive(uint8_t *screen, float *accum, int width, int height, uint32_t frame_count)
{
    const float fc = (float)frame_count;
    for (size_t p = 0; p < (size_t)width * height; p++) {
        uint32_t in;
        std::memcpy(&in, screen + p * 4, 4);
        f3 rad = mk3((float)(in & 0xffu) / 255.0f, (float)((in >> 8) & 0xffu) / 255.0f, (float)((in >> 16) & 0xffu) / 255.0f);
        if (frame_count > 1u) rad = rad + mk3(accum[p * 4 + 0], accum[p * 4 + 1], accum[p * 4 + 2]);
        accum[p * 4 + 0] = rad.x; accum[p * 4 + 1] = rad.y; accum[p * 4 + 2] = rad.z; accum[p * 4 + 3] = 1.0f;
        const f3 avg = (rad / fc) * 1.0f;
        const uint32_t out = pack_rgba8(mk3(aces_channel(avg.x), aces_channel(avg.y), aces_channel(avg.z)));
        std::memcpy(screen + p * 4, &out, 4);
    }
}

// K3 with the kernel's per-pixel function (pt_post.cuh); same roles as orc_temporal.
void devcheck_temporal(const gdpt_temporal_params *params, uint8_t *screen, const float *depth, float *fb1, float *fb2)
{
    const bool use_first = (params->frame_count % 2u) == 0u;
    const float *history = use_first ? fb1 : fb2;
    float *next = use_first ? fb2 : fb1;
    std::vector<uint32_t> scr((size_t)params->width * params->height);
    std::memcpy(scr.data(), screen, scr.size() * 4);
    for (int y = 0; y < params->height; y++)
        for (int x = 0; x < params->width; x++) temporal_pixel(*params, x, y, scr.data(), depth, history, next);
    std::memcpy(screen, scr.data(), scr.size() * 4);
}

} // extern "C"

// Analysis hook (tools/path_cost_model.py): per-pixel step counts of the culled, compact-order traversal --
// node steps, leaf entries, triangle tests, instance steps, segments, and the largest leaf met.
extern "C" int devcheck_path_costs(const devcheck_scene *in, const gdpt_render_params *params, const gdpt_camera *cam, int max_depth,
                                   int y_begin, int y_end, uint32_t *out6)
{
    DerivedLayout lay;
    const std::string err = derive_layout((const gdpt_bvh_node *)in->bvh, (uint32_t)in->n_nodes, (const gdpt_blas_instance *)in->blas,
                                          (uint32_t)in->n_blas, (const gdpt_tlas_node *)in->tlas, (uint32_t)in->n_tlas,
                                          (const gdpt_triangle_geometry *)in->tri_geom, (uint32_t)in->n_tris, lay);
    if (!err.empty()) return -1;
    FastLayout fast;
    SceneView sc;
    make_view(in, lay, fast, sc);
    const int W = params->width, H = params->height;
    HostStack st;
    for (int y = y_begin; y < y_end && y < H; y++) {
        for (int x = 0; x < W; x++) {
            uint32_t *c = out6 + ((size_t)y * W + x) * 6;
            for (int k = 0; k < 6; k++) c[k] = 0;
            f3 o, d;
            u2 seed = generate_primary_ray(*cam, W, H, x, y, &o, &d);
            f3 radiance = mk3(0, 0, 0), throughput = mk3(1, 1, 1);
            for (int i = 0; i < max_depth; i++) {
                RayState r;
                ray_begin(r, sc, o, d);
                TraceCounters tc;
                counters_init(tc, nullptr, 0);
                uint32_t tri_next = 0, tri_end = 0;
                if (g_fast == 2 && sc.fast4_ok) { // the same over the four-wide tables
                    r.cur = sc.fast4_root;
                    while (r.cur != LINK_NONE) {
                        if (fast_link_is_leaf(r.cur)) { c[1]++; c[2] += ((r.cur >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u; fast_step_leaf(sc, r, st); }
                        else if (fast_link_is_node(r.cur, r.inst)) { c[0]++; fast_step_node4(sc, r, st); }
                        else {
                            c[3]++;
                            if (r.inst != GDPT_NO_INSTANCE) { r.o = r.wo; r.d = r.wd; r.rd = rcp3(r.wd); r.inst = GDPT_NO_INSTANCE; }
                            if (r.cur & LINK_LEAF) fast_enter_instance<true>(sc, r, st);
                        }
                    }
                    if (!fast_result_is_reference(sc, r)) { c[5]++; ray_begin(r, sc, o, d); trace_ray_compact<true, true>(sc, r, st, &tc); }
                } else if (g_fast && sc.fast_ok) { // step counts of the closest-hit search (c[5] = exact re-traces)
                    while (r.cur != LINK_NONE) {
                        if (fast_link_is_leaf(r.cur)) { c[1]++; c[2] += ((r.cur >> FAST_LEAF_COUNT_SHIFT) & 7u) + 1u; fast_step_leaf(sc, r, st); }
                        else if (fast_link_is_node(r.cur, r.inst)) { c[0]++; fast_step_node(sc, r, st); }
                        else { c[3]++; fast_step_instance(sc, r, st); }
                    }
                    if (!fast_result_is_reference(sc, r)) { c[5]++; ray_begin(r, sc, o, d); trace_ray_compact<true, true>(sc, r, st, &tc); }
                }
                while (r.cur != LINK_NONE || tri_next < tri_end) {
                    if (tri_next < tri_end || link_is_blas_leaf(r.cur)) {
                        if (tri_next == tri_end) {
                            c[1]++;
                            const q4u leaf = ldqu(sc.leaf_recs, r.cur & LINK_INDEX_MASK);
                            if (leaf.y > c[5]) c[5] = leaf.y;
                        }
                        step_blas_leaf_one<true>(sc, r, st, &tc, tri_next, tri_end);
                        c[2]++;
                    } else if (link_is_node_step(r.cur, r.inst)) { step_node<true, true>(sc, r, st, &tc); c[0]++; }
                    else { step_instance<true, true>(sc, r, st, &tc); c[3]++; }
                }
                c[4]++;
                if (!(r.t < 1e9f)) break;
                BounceResult br = shade_and_bounce(sc, o, d, r.t, r.u, r.v, r.tri, r.blas_front, radiance, throughput, seed);
                radiance = br.radiance;
                if (!br.alive) break;
                throughput = br.throughput; o = br.next_o; d = br.next_d;
            }
        }
    }
    return 0;
}
