// ref_shader_bridge.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).  NEVER SHIPPED, NEVER A FALLBACK.
//
// extern "C" doorway into the REFERENCE's own shaders.  The four `#include "*.glsl"` lines below pull in the
// reference's shader text -- project/addons/jar_path_tracing/src/shaders/{main,brdfs,progressive_rendering,
// temporal_reprojection}.glsl -- as the bodies of C++ structs, after the lexical adaptation of
// oracle/glsl_shim/glsl_prep.py (rules R1-R7 there; scratch copies, deleted by oracle/Makefile after the
// compile; nothing of the reference is stored in this repository).  Types and built-ins come from
// oracle/glsl_shim/glsl.hpp.  This file contains no rendering algorithm: it binds caller memory to the
// shaders' resources, sets gl_GlobalInvocationID, calls the shaders' own `main()` once per pixel and decodes
// the buffer-read log into the parity observables of SURVEY A.5 (node-visit order, counters, hit ids).
//
// Output: oracle/_ref/libgdpt_refshader.so (git-ignored; travels to the GPU box with the snapshot).
// It pins oracle/pt_oracle.cpp -- and through it the CUDA kernels -- to the reference text:
// tests/test_ref_shader.py compares the two on frames, depth, raw radiance, traversal logs and single rays.
#include "glsl.hpp"
#include "gdpt_wire.h"

#include <atomic>
#include <thread>
#include <vector>

#undef M_PI // brdfs.glsl:1 defines its own

namespace glsl {

namespace ref_main { // main.glsl + brdfs.glsl as written
struct Shader {
    uvec3 gl_GlobalInvocationID;
    int gdpt_segments; // R7: the `5` of main.glsl:377
#include "main.glsl"
};
} // namespace ref_main

namespace ref_main_debug { // the same text with main.glsl:4 `#define DEBUG_STEPS` enabled
struct Shader {
    uvec3 gl_GlobalInvocationID;
    int gdpt_segments;
#define DEBUG_STEPS
#include "main.glsl"
#undef DEBUG_STEPS
};
} // namespace ref_main_debug

namespace ref_progressive {
struct Shader {
    uvec3 gl_GlobalInvocationID;
#include "progressive_rendering.glsl"
};
} // namespace ref_progressive

namespace ref_temporal {
struct Shader {
    uvec3 gl_GlobalInvocationID;
#include "temporal_reprojection.glsl"
};
} // namespace ref_temporal

} // namespace glsl

namespace {

using MainShader = glsl::ref_main::Shader;

// std430 sizes of the shader's own structs equal the host records; the strides the host wrote
// the arrays with are the wire sizes of gdpt_wire.h.
static_assert(sizeof(MainShader::TriangleGeometry) == 48, "TriangleGeometry");
static_assert(sizeof(MainShader::TriangleData) == 72 && sizeof(gdpt_triangle_data) == 80, "TriangleData: 72 B of fields, std430 stride 80");
static_assert(offsetof(MainShader::TriangleData, materialIndex) == 12 && offsetof(MainShader::TriangleData, uvs) == 48, "TriangleData offsets");
static_assert(sizeof(MainShader::Material) == 64 && offsetof(MainShader::Material, albedo_texture_id) == 40, "Material");
static_assert(sizeof(MainShader::BVHNode) == 48 && offsetof(MainShader::BVHNode, tri_count) == 44, "BVHNode");
static_assert(sizeof(MainShader::TLASNode) == 32 && offsetof(MainShader::TLASNode, leftRight) == 12 && offsetof(MainShader::TLASNode, blas) == 28, "TLASNode");
static_assert(sizeof(MainShader::BLASInstance) == 176 && offsetof(MainShader::BLASInstance, root) == 160, "BLASInstance");
static_assert(sizeof(MainShader::Params) == 36 && sizeof(MainShader::Camera) == 156, "Params 36 B, Camera 156 B of fields in a 160 B block");

enum : uint64_t {
    LOG_TRI_GEOM = 1ull << 32, LOG_TRI_DATA = 2ull << 32, LOG_MATERIAL = 3ull << 32,
    LOG_BVH = 4ull << 32, LOG_INSTANCE = 5ull << 32, LOG_TLAS = 6ull << 32
};

struct SceneView {
    const uint8_t *tri_geom, *tri_data, *materials, *bvh, *blas, *tlas, *textures;
    uint64_t n_tris, n_materials, n_nodes, n_blas, n_tlas;
    int tex_w, tex_h, tex_layers;
};

template <class S> void bind_scene(S &sh, const SceneView &sc, std::vector<uint64_t> *log)
{
    sh.triangles_geometry.base = sc.tri_geom; sh.triangles_geometry.stride = 48; sh.triangles_geometry.count = sc.n_tris;
    sh.triangles_data.base = sc.tri_data; sh.triangles_data.stride = 80; sh.triangles_data.count = sc.n_tris;
    sh.materials.base = sc.materials; sh.materials.stride = 64; sh.materials.count = sc.n_materials;
    sh.bvhTree.base = sc.bvh; sh.bvhTree.stride = 48; sh.bvhTree.count = sc.n_nodes;
    sh.blas_instances.base = sc.blas; sh.blas_instances.stride = 176; sh.blas_instances.count = sc.n_blas;
    sh.tlas_nodes.base = sc.tlas; sh.tlas_nodes.stride = 32; sh.tlas_nodes.count = sc.n_tlas;
    sh.triangles_geometry.tag = LOG_TRI_GEOM; sh.triangles_data.tag = LOG_TRI_DATA; sh.materials.tag = LOG_MATERIAL;
    sh.bvhTree.tag = LOG_BVH; sh.blas_instances.tag = LOG_INSTANCE; sh.tlas_nodes.tag = LOG_TLAS;
    sh.triangles_geometry.log = sh.triangles_data.log = sh.materials.log = sh.bvhTree.log = sh.blas_instances.log = sh.tlas_nodes.log = log;
    sh.textureArray.texels = sc.textures; sh.textureArray.width = sc.tex_w; sh.textureArray.height = sc.tex_h;
    sh.textureArray.layers = sc.tex_layers;
}

struct Segment {
    uint32_t hit = 0, triangle = 0, blas = 0, node_pops = 0, box_tests = 0, tri_tests = 0, tlas_leaves = 0;
    uint64_t hash = GDPT_FNV64_OFFSET;
};

// Turns the reads one pixel's main() made into per-ray-segment observables.  What is read when is fixed by the
// shader text: a popped TLAS node (main.glsl:313) is followed by its instance (:317) or by its two children
// (:333-334); a popped BVH node (:277) by its triangles (:226) or its two children (:286-287); a hit by
// triangles_data / blas_instances / materials (:196-198).  TLAS node 0 is read only as the first pop of a ray.
struct LogDecoder {
    const SceneView &sc;
    std::vector<Segment> segs;
    std::vector<uint32_t> first_visits; // pops of segment 0
    explicit LogDecoder(const SceneView &s) : sc(s) {}

    static void visit(Segment &g, uint32_t id)
    {
        g.node_pops++;
        for (int b = 0; b < 4; b++) { g.hash ^= (id >> (8 * b)) & 0xffu; g.hash *= GDPT_FNV64_PRIME; }
    }
    void decode(const std::vector<uint64_t> &log)
    {
        segs.clear(); first_visits.clear();
        int skip_tlas = 0, skip_bvh = 0;
        bool shading = false;
        for (uint64_t e : log) {
            const uint64_t buf = e & ~0xffffffffull;
            const uint32_t idx = (uint32_t)e;
            if (buf == LOG_TLAS) {
                if (skip_tlas > 0) { skip_tlas--; continue; }
                if (idx == 0) segs.emplace_back();
                Segment &g = segs.back();
                visit(g, idx | GDPT_VISIT_TLAS_TAG);
                if (segs.size() == 1) first_visits.push_back(idx | GDPT_VISIT_TLAS_TAG);
                gdpt_tlas_node n;
                memcpy(&n, sc.tlas + 32 * (size_t)idx, 32);
                if (n.left_right != 0) { skip_tlas = 2; g.box_tests += 2; }
            } else if (buf == LOG_BVH) {
                if (skip_bvh > 0) { skip_bvh--; continue; }
                Segment &g = segs.back();
                visit(g, idx);
                if (segs.size() == 1) first_visits.push_back(idx);
                gdpt_bvh_node n;
                memcpy(&n, sc.bvh + 48 * (size_t)idx, 48);
                if (n.tri_count == 0) { skip_bvh = 2; g.box_tests += 2; }
            } else if (buf == LOG_TRI_GEOM) {
                segs.back().tri_tests++;
            } else if (buf == LOG_TRI_DATA) {
                segs.back().hit = 1; segs.back().triangle = idx; shading = true;
            } else if (buf == LOG_INSTANCE) {
                if (shading) segs.back().blas = idx; else segs.back().tlas_leaves++;
            } else if (buf == LOG_MATERIAL) {
                shading = false;
            }
        }
    }
};

struct Job {
    SceneView sc;
    const void *params36, *camera160;
    int width, height, max_depth, debug_steps, y_begin, y_end, y_step, want_counters;
    uint8_t *out_rgba8; float *out_depth, *out_radiance;
    gdpt_trace_record *trace; int trace_segments;
    uint32_t *visits; uint32_t visits_per_ray;
    std::atomic<int> next_row;
    std::atomic<uint64_t> rays, primary_hits, node_pops, box_tests, tri_tests, tlas_leaves;
};

template <class S> void render_rows(Job *job)
{
    S sh;
    std::vector<uint64_t> log;
    // The read log feeds the per-segment observables; when nobody asked for them (timing runs) only the ray count is
    // kept: TLAS node 0 is read exactly once per ray_trace_tlas call (main.glsl:309-313; child reads never name it).
    const bool observe = job->trace != nullptr || job->visits != nullptr || job->want_counters;
    uint64_t tlas_root_reads = 0;
    bind_scene(sh, job->sc, observe ? &log : nullptr);
    if (!observe) sh.tlas_nodes.reads_of_record_0 = &tlas_root_reads;
    memcpy(&sh.params, job->params36, sizeof(sh.params));
    memcpy(&sh.camera, job->camera160, sizeof(sh.camera));
    sh.gdpt_segments = job->max_depth;
    sh.outputImage.data = job->out_rgba8; sh.outputImage.width = job->width; sh.outputImage.height = job->height;
    sh.outputImage.raw_store = job->out_radiance;
    sh.depthBuffer.data = job->out_depth; sh.depthBuffer.width = job->width; sh.depthBuffer.height = job->height;
    LogDecoder dec(job->sc);
    uint64_t rays = 0, phits = 0, pops = 0, boxes = 0, tris = 0, leaves = 0;
    const size_t n_pix = (size_t)job->width * job->height;
    for (;;) {
        const int y = job->next_row.fetch_add(job->y_step);
        if (y >= job->y_end) break;
        for (int x = 0; x < job->width; x++) {
            log.clear();
            sh.gl_GlobalInvocationID = glsl::uvec3((uint32_t)x, (uint32_t)y, 0u);
            sh.main();
            if (!observe) continue;
            dec.decode(log);
            const size_t pix = (size_t)y * job->width + x;
            rays += dec.segs.size();
            for (size_t i = 0; i < dec.segs.size(); i++) {
                const Segment &g = dec.segs[i];
                pops += g.node_pops; boxes += g.box_tests; tris += g.tri_tests; leaves += g.tlas_leaves;
                if (i == 0 && g.hit) phits++;
                if (job->trace && (int)i < job->trace_segments) {
                    gdpt_trace_record &tr = job->trace[i * n_pix + pix];
                    memset(&tr, 0, sizeof(tr));
                    tr.hit = g.hit; tr.triangle = g.triangle; tr.blas = g.blas;
                    tr.node_pops = g.node_pops; tr.box_tests = g.box_tests; tr.tri_tests = g.tri_tests; tr.tlas_leaves = g.tlas_leaves;
                    tr.visit_hash_lo = (uint32_t)g.hash; tr.visit_hash_hi = (uint32_t)(g.hash >> 32);
                }
            }
            if (job->visits)
                for (size_t k = 0; k < dec.first_visits.size() && k < job->visits_per_ray; k++)
                    job->visits[pix * job->visits_per_ray + k] = dec.first_visits[k];
        }
    }
    if (!observe) rays = tlas_root_reads;
    job->rays += rays; job->primary_hits += phits; job->node_pops += pops; job->box_tests += boxes;
    job->tri_tests += tris; job->tlas_leaves += leaves;
}

} // namespace

extern "C" {

typedef struct refsh_scene { // same layout as orc_scene (oracle/pt_oracle.cpp)
    const void *tri_geom; uint64_t n_tris;
    const void *tri_data;
    const void *materials; uint64_t n_materials;
    const void *bvh; uint64_t n_nodes;
    const void *blas; uint64_t n_blas;
    const void *tlas; uint64_t n_tlas;
    const uint8_t *textures; int32_t tex_w, tex_h, tex_layers, _pad;
} refsh_scene;

typedef struct refsh_stats {
    uint64_t rays, primary_hits, node_pops, box_tests, tri_tests, tlas_leaves;
    uint32_t max_stack, stack_overflow; // not observable from the shader: always 0
} refsh_stats;

static SceneView view_of(const refsh_scene *s)
{
    SceneView v;
    v.tri_geom = (const uint8_t *)s->tri_geom; v.tri_data = (const uint8_t *)s->tri_data; v.materials = (const uint8_t *)s->materials;
    v.bvh = (const uint8_t *)s->bvh; v.blas = (const uint8_t *)s->blas; v.tlas = (const uint8_t *)s->tlas; v.textures = s->textures;
    v.n_tris = s->n_tris; v.n_materials = s->n_materials; v.n_nodes = s->n_nodes; v.n_blas = s->n_blas; v.n_tlas = s->n_tlas;
    v.tex_w = s->tex_w; v.tex_h = s->tex_h; v.tex_layers = s->tex_layers;
    return v;
}

// main.glsl `main()` for every pixel of rows y_begin, y_begin + y_step, ... < y_end.  Same arguments as
// orc_path_trace; out_radiance (optional) receives the vec4 handed to imageStore(outputImage, ..) unconverted.
// trace records carry hit / triangle / blas / counters / visit hash (t, u, v, front: see refsh_trace_rays).
// stats: rays always; the other totals only when trace or visits are requested or debug_steps & 2 is set (they come
// from the read log, which costs about a quarter of the run time).
int refsh_path_trace(const refsh_scene *scene, const void *params36, const void *camera160, int max_depth, int debug_steps,
                     int n_threads, int y_begin, int y_end, int y_step, uint8_t *out_rgba8, float *out_depth,
                     gdpt_trace_record *trace, int trace_segments, uint32_t *visits, uint32_t visits_per_ray,
                     refsh_stats *stats, float *out_radiance)
{
    Job job;
    job.sc = view_of(scene);
    job.params36 = params36; job.camera160 = camera160;
    const gdpt_render_params *p = (const gdpt_render_params *)params36;
    job.width = p->width; job.height = p->height;
    job.max_depth = max_depth; job.debug_steps = debug_steps & 1; job.want_counters = (debug_steps & 2) ? 1 : 0;
    job.y_begin = y_begin < 0 ? 0 : y_begin; job.y_end = y_end > p->height ? p->height : y_end; job.y_step = y_step < 1 ? 1 : y_step;
    job.out_rgba8 = out_rgba8; job.out_depth = out_depth; job.out_radiance = out_radiance;
    job.trace = trace; job.trace_segments = trace_segments; job.visits = visits; job.visits_per_ray = visits_per_ray;
    job.next_row = job.y_begin;
    job.rays = 0; job.primary_hits = 0; job.node_pops = 0; job.box_tests = 0; job.tri_tests = 0; job.tlas_leaves = 0;
    if (trace) {
        const size_t n = (size_t)trace_segments * p->width * p->height;
        for (size_t i = 0; i < n; i++) { memset(&trace[i], 0, sizeof(trace[i])); trace[i].hit = 0xFFFFFFFFu; }
    }
    if (n_threads < 1) n_threads = 1;
    void (*worker)(Job *) = job.debug_steps ? render_rows<glsl::ref_main_debug::Shader> : render_rows<glsl::ref_main::Shader>;
    std::vector<std::thread> pool;
    for (int i = 1; i < n_threads; i++) pool.emplace_back(worker, &job);
    worker(&job);
    for (auto &t : pool) t.join();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->rays = job.rays; stats->primary_hits = job.primary_hits; stats->node_pops = job.node_pops;
        stats->box_tests = job.box_tests; stats->tri_tests = job.tri_tests; stats->tlas_leaves = job.tlas_leaves;
    }
    return 0;
}

// The shader's own ray_trace_tlas (main.glsl:305-350) on caller-given rays, with the HitInfo initialisation of
// ray_trace (main.glsl:354-356) and rD = 1.0 / d (main.glsl:421).  Fills every field of the trace record except
// max_stack.
void refsh_trace_rays(const refsh_scene *scene, uint64_t n, const float *origins, const float *directions, gdpt_trace_record *out)
{
    MainShader sh;
    std::vector<uint64_t> log;
    const SceneView sc = view_of(scene);
    bind_scene(sh, sc, &log);
    LogDecoder dec(sc);
    for (uint64_t i = 0; i < n; i++) {
        MainShader::Ray ray;
        ray.o = glsl::vec3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        ray.d = glsl::vec3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        ray.rD = 1.0f / ray.d;
        MainShader::HitInfo h;
        memset(&h, 0, sizeof(h));
        h.t = 1e9f; h.steps = 0;
        log.clear();
        const bool hit = sh.ray_trace_tlas(ray, h);
        dec.decode(log);
        const Segment &g = dec.segs.at(0);
        gdpt_trace_record &tr = out[i];
        memset(&tr, 0, sizeof(tr));
        tr.hit = hit ? 1u : 0u; tr.triangle = hit ? h.triangle : 0u; tr.blas = hit ? h.blas : 0u; tr.front = hit && h.front ? 1u : 0u;
        tr.t = h.t; tr.u = hit ? h.barycentrics.x : 0.0f; tr.v = hit ? h.barycentrics.y : 0.0f;
        tr.node_pops = g.node_pops; tr.box_tests = g.box_tests; tr.tri_tests = h.steps; tr.tlas_leaves = g.tlas_leaves;
        tr.visit_hash_lo = (uint32_t)g.hash; tr.visit_hash_hi = (uint32_t)(g.hash >> 32);
    }
}

// progressive_rendering.glsl `main()` for every pixel.  screen: rgba8 in/out, accum: rgba32f in/out.
void refsh_progressive(uint8_t *screen, float *accum, int width, int height, uint32_t frame_count)
{
    glsl::ref_progressive::Shader sh;
    sh.width = (uint32_t)width; sh.height = (uint32_t)height; sh.frame_count = frame_count;
    sh.screenTexture.data = screen; sh.screenTexture.width = width; sh.screenTexture.height = height;
    sh.frameBuffer.data = accum; sh.frameBuffer.width = width; sh.frameBuffer.height = height;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            sh.gl_GlobalInvocationID = glsl::uvec3((uint32_t)x, (uint32_t)y, 0u);
            sh.main();
        }
}

// temporal_reprojection.glsl `main()` for every pixel.  No pixel reads what another pixel of the dispatch
// writes, so the order is immaterial.
void refsh_temporal(const gdpt_temporal_params *params, uint8_t *screen, const float *depth, float *fb1, float *fb2)
{
    glsl::ref_temporal::Shader sh;
    memcpy(&sh.reprojectionMatrix, params->delta_matrix, 64);
    sh.width = (uint32_t)params->width; sh.height = (uint32_t)params->height; sh.frameCount = params->frame_count;
    sh.blendFactor = params->blend_factor; sh.nearPlane = params->near_plane; sh.farPlane = params->far_plane;
    const int W = params->width, H = params->height;
    sh.screenTexture.data = screen; sh.screenTexture.width = W; sh.screenTexture.height = H;
    sh.depthTexture.data = (void *)depth; sh.depthTexture.width = W; sh.depthTexture.height = H;
    sh.frameBuffer1.data = fb1; sh.frameBuffer1.width = W; sh.frameBuffer1.height = H;
    sh.frameBuffer2.data = fb2; sh.frameBuffer2.width = W; sh.frameBuffer2.height = H;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            sh.gl_GlobalInvocationID = glsl::uvec3((uint32_t)x, (uint32_t)y, 0u);
            sh.main();
        }
}

// The shader's RNG on its own (SURVEY A.6 known answers).
void refsh_prng_seed(uint32_t px, uint32_t py, uint32_t frame, uint32_t *out2)
{
    MainShader sh;
    glsl::uvec2 s = sh.prng_seed(glsl::vec2((float)px, (float)py), frame);
    out2[0] = s.x; out2[1] = s.y;
}
void refsh_pcg2d(uint32_t *state2, float *out2)
{
    MainShader sh;
    glsl::uvec2 s(state2[0], state2[1]);
    glsl::vec2 r = sh.pcg2d(s);
    state2[0] = s.x; state2[1] = s.y; out2[0] = r.x; out2[1] = r.y;
}

} // extern "C"
