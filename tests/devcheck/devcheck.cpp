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

template <bool CULL> static void run_ray(const SceneView &sc, RayState &r, HostStack &st, TraceCounters *tc)
{
    if (g_stepwise == 2) trace_ray_compact<true, CULL>(sc, r, st, tc);
    else if (g_stepwise) trace_ray_stepwise<true, CULL>(sc, r, st, tc);
    else trace_ray<true, CULL>(sc, r, st, tc);
}

extern "C" {

// 0: node-at-a-time traversal, 1: leaves split into single-triangle steps (phase-voting order),
// 2: the compact scheduler's node / instance / triangle steps
void devcheck_set_stepwise(int on) { g_stepwise = on; }
// 0: reference visit order, 1: tight-box culling (hit records must not change)
void devcheck_set_cull(int on) { g_cull = on; }

struct devcheck_scene {
    const void *tri_geom; uint64_t n_tris;
    const void *tri_data;
    const void *materials; uint64_t n_materials;
    const void *bvh; uint64_t n_nodes;
    const void *blas; uint64_t n_blas;
    const void *tlas; uint64_t n_tlas;
    const uint8_t *textures; int32_t tex_w, tex_h, tex_layers, _pad;
};

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
    SceneView sc;
    std::memset(&sc, 0, sizeof(sc));
    sc.tri_geom = (const gdpt_triangle_geometry *)in->tri_geom; sc.tri_data = (const gdpt_triangle_data *)in->tri_data;
    sc.materials = (const gdpt_material *)in->materials; sc.bvh = (const gdpt_bvh_node *)in->bvh;
    sc.blas = (const gdpt_blas_instance *)in->blas; sc.tlas = (const gdpt_tlas_node *)in->tlas;
    sc.textures = in->textures; sc.tex_w = in->tex_w; sc.tex_h = in->tex_h; sc.tex_layers = in->tex_layers;
    sc.wide_nodes = lay.wide_nodes.data(); sc.leaf_recs = lay.leaf_recs.data();
    sc.wide_tlas = lay.wide_tlas.data(); sc.inst_recs = lay.inst_recs.data();
    sc.tlas_root_link = lay.tlas_root_link;

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
                    rec.blas = hit ? (r.blas_front & ~GDPT_FRONT_BIT) : 0u; rec.front = hit ? (r.blas_front >> 31) : 0u;
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

} // extern "C"
