// gdpt_capi.cu -- implementation of the C-ABI declared in include/gdpt.h:
// device / shader / RID bookkeeping that stands in for the gdcs ComputeShader +
// RenderingDevice pair, the binding-table validation, construction of the
// derived traversal layout, and the frame drivers that launch the kernels of
// pt_kernels.cu.  No CPU rendering path exists here: every entry point either
// drives the GPU or fails.
#include "gdpt.h"
#include "derived_layout.h"
#include "fast_bvh.h"
#include "pt_kernels.cuh"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace gdpt;

namespace {

thread_local std::string g_create_error;

enum ResourceKind { RES_BUFFER, RES_IMAGE, RES_LAYERED };

struct Resource {
    ResourceKind kind = RES_BUFFER;
    void *dptr = nullptr;
    uint64_t size = 0;
    int width = 0, height = 0, layers = 1;
    gdpt_data_format format = GDPT_FORMAT_R8G8B8A8_UNORM;
    std::vector<uint8_t> shadow; // host copy of storage buffers (used to derive layouts)
};

enum ShaderKind { SHADER_MAIN, SHADER_PROGRESSIVE, SHADER_TEMPORAL };

} // namespace

struct gdpt_device {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    std::map<gdpt_rid, Resource> resources;
    gdpt_rid next_rid = 1;
    void *pinned_staging = nullptr; // small H2D staging (camera, params)
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    // pipelined frames: read-backs run on their own stream; per-frame H2D blocks come from a ring
    cudaStream_t copy_stream = nullptr;
    cudaStream_t extra_streams[GDPT_MAX_FRAMES_IN_FLIGHT - 1] = {}; // frame streams of overlapped pipelined frames besides `stream`
    bool extra_streams_made = false;
    bool frame_overlap = true;          // pipelined frames run on their own streams (false: they share one)
    uint8_t *ring = nullptr;            // kRingSlots x 512 B of pinned memory
    cudaEvent_t ring_ev[8] = {};        // slot i may be rewritten once ring_ev[i] has completed
    bool ring_used[8] = {};
    unsigned ring_next = 0;
};

struct gdpt_shader {
    gdpt_device *dev = nullptr;
    ShaderKind kind = SHADER_MAIN;
    std::map<std::pair<int, int>, gdpt_rid> bindings; // (set, binding) -> rid
    std::vector<gdpt_rid> owned;
    bool initialized = false, uniforms_ready = false;
    // "#define" options
    int max_depth = 5;
    bool debug_steps = false;
    int trace_segments = 0;
    uint32_t visits_per_ray = 0;
    int variant = -1; // "#define GDPT_VARIANT n": traversal schedule, -1 = default
    int cull = -1;    // "#define GDPT_CULL n" / "#define GDPT_REFERENCE_ORDER": -1 = default
    int record_hits = 0; // "#define GDPT_RECORD_HITS n": hit records of the first n segments from the rendering kernels
    std::map<std::string, int> tuning; // "#define GDPT_TUNE_<NAME> n": scheduling knobs of the path kernels (A/B runs)
    std::string fast_why_not; // why the closest-hit tables are not in use ("" = in use)
    uint32_t fast_need4 = 0;  // stack entries the four-wide search can need (fast_bvh.h)
    bool count_work = false;  // "#define GDPT_COUNT_WORK": the path kernel also counts its own work (gdpt_frame_stats own_*)
    bool material_ext = false; // "#define GDPT_MATERIAL_EXT": extension fields of gdpt_material + set 1 binding 6 (gdpt_wire.h)
    // main-shader state built by finish_create_uniforms
    FrameArgs args;
    std::vector<void *> derived; // device allocations owned by this shader
    int shard_part = 0, shard_parts = 1, shard_band = 4;
    gdpt_frame_stats stats;
    bool stats_valid = false;
    // frames in flight of gdpt_render_frame_begin / _wait
    struct FrameSlot {
        FrameCounters *dcnt = nullptr;   // device counters of this frame
        FrameCounters *hcnt = nullptr;   // pinned host copy, valid once `done` has completed
        void *stage_rgba8 = nullptr;     // device copies the read-back streams from, so the next K1 may overwrite the images
        float *stage_depth = nullptr;
        cudaEvent_t k_done = nullptr, done = nullptr, t0 = nullptr, t1 = nullptr, t2 = nullptr, t3 = nullptr;
        bool pending = false, with_k2 = false;
        // overlapped frames: what K1 of this frame reads and writes (the other frame in flight has its own)
        gdpt_camera *cam_dev = nullptr; gdpt_progressive_params *pp_dev = nullptr;
        uint32_t *raw_rgba8 = nullptr; float *raw_depth = nullptr; uint32_t *hit_list = nullptr;
        bool finished_once = false;      // k_done has been recorded at least once
        uint32_t launches = 0;
    } slots[GDPT_MAX_FRAMES_IN_FLIGHT];
    unsigned slot_head = 0, slot_tail = 0; // begin fills slots[head % N], wait drains slots[tail % N]
    bool warp_profile = false;          // per-warp schedule profile of the path kernel
    size_t warp_prof_warps = 0;
    bool stage_timing = false;          // record an event between the K1 stage launches
    std::vector<cudaEvent_t> stage_ev;  // 2*max_depth + 1 events when enabled
    int stage_count = 0;
    std::vector<uint8_t> staged_params; // gdpt_shader_stage_params: bytes the next frame call uploads into set0 b0
    PeerScreens peers = {};             // progressive shader: images of the other GPUs K2 also writes (gdpt_shader_set_peer_screens)
};

namespace {

int fail(gdpt_device *d, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (d) d->last_error = buf; else g_create_error = buf;
    return code;
}

#define GDPT_CUDA(dev, call)                                                                                        \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return fail((dev), GDPT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int tune(const gdpt_shader *s, const char *name, int fallback)
{
    auto it = s->tuning.find(name);
    return it == s->tuning.end() ? fallback : it->second;
}

Resource *find(gdpt_device *d, gdpt_rid rid)
{
    auto it = d->resources.find(rid);
    return it == d->resources.end() ? nullptr : &it->second;
}

uint64_t texel_bytes(gdpt_data_format f)
{
    switch (f) {
    case GDPT_FORMAT_R8G8B8A8_UNORM: return 4;
    case GDPT_FORMAT_R32_SFLOAT: return 4;
    case GDPT_FORMAT_R32G32B32A32_SFLOAT: return 16;
    }
    return 0;
}

std::string basename_of(const char *path)
{
    std::string s(path ? path : "");
    size_t p = s.find_last_of("/\\");
    return p == std::string::npos ? s : s.substr(p + 1);
}

// "#define NAME [value]" -> (NAME, value)
bool parse_define(const char *line, std::string *name, long *value, bool *has_value)
{
    char n[128];
    long v = 0;
    int got = sscanf(line, " # define %127s %ld", n, &v);
    if (got < 1) got = sscanf(line, " #define %127s %ld", n, &v);
    if (got < 1) return false;
    *name = n; *value = v; *has_value = (got >= 2);
    return true;
}

Resource *bound(gdpt_shader *s, int set, int binding)
{
    auto it = s->bindings.find({ set, binding });
    if (it == s->bindings.end()) return nullptr;
    return find(s->dev, it->second);
}

template <typename T> int dev_alloc(gdpt_shader *s, T **out, size_t count)
{
    void *p = nullptr;
    GDPT_CUDA(s->dev, cudaMalloc(&p, count * sizeof(T) > 0 ? count * sizeof(T) : 16));
    s->derived.push_back(p);
    *out = static_cast<T *>(p);
    return GDPT_OK;
}

template <typename T> int dev_upload(gdpt_shader *s, const T **out, const std::vector<T> &v)
{
    T *p = nullptr;
    int rc = dev_alloc(s, &p, v.size());
    if (rc) return rc;
    if (!v.empty()) GDPT_CUDA(s->dev, cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s->dev->stream));
    *out = p;
    return GDPT_OK;
}

// Build WideNode / LeafRec / InstRec tables from the uploaded reference arrays and upload them.
int build_derived_layout(gdpt_shader *s, const Resource &bvh_r, const Resource &blas_r, const Resource &tlas_r, const Resource &tri_r)
{
    gdpt_device *d = s->dev;
    DerivedLayout lay;
    const std::string err = derive_layout(reinterpret_cast<const gdpt_bvh_node *>(bvh_r.shadow.data()),
                                          (uint32_t)(bvh_r.size / sizeof(gdpt_bvh_node)),
                                          reinterpret_cast<const gdpt_blas_instance *>(blas_r.shadow.data()),
                                          (uint32_t)(blas_r.size / sizeof(gdpt_blas_instance)),
                                          reinterpret_cast<const gdpt_tlas_node *>(tlas_r.shadow.data()),
                                          (uint32_t)(tlas_r.size / sizeof(gdpt_tlas_node)),
                                          reinterpret_cast<const gdpt_triangle_geometry *>(tri_r.shadow.data()),
                                          (uint32_t)(tri_r.size / sizeof(gdpt_triangle_geometry)), lay);
    if (!err.empty()) return fail(d, GDPT_ERR_BAD_BINDING, "%s", err.c_str());
    int rc;
    // closest-hit tables (fast_bvh.h): our own BVH over the same triangles + what the proof of pt_fast.cuh reads
    FastLayout fast;
    build_fast_layout(reinterpret_cast<const gdpt_bvh_node *>(bvh_r.shadow.data()), (uint32_t)(bvh_r.size / sizeof(gdpt_bvh_node)),
                      reinterpret_cast<const gdpt_blas_instance *>(blas_r.shadow.data()), (uint32_t)(blas_r.size / sizeof(gdpt_blas_instance)),
                      reinterpret_cast<const gdpt_tlas_node *>(tlas_r.shadow.data()), (uint32_t)(tlas_r.size / sizeof(gdpt_tlas_node)),
                      reinterpret_cast<const gdpt_triangle_geometry *>(tri_r.shadow.data()),
                      (uint32_t)(tri_r.size / sizeof(gdpt_triangle_geometry)), lay, fast);
    s->args.sc.fast_ok = fast.ok ? 1u : 0u;
    s->fast_why_not = fast.why_not;
    if (fast.ok) {
        for (size_t b = 0; b < lay.inst_recs.size(); b++) {
            lay.inst_recs[b].fast_root = fast.inst_root[b];
            std::memcpy(&lay.inst_recs[b].tight_min[3], &fast.inst_root4[b], 4); // root in the four-wide table
        }
        s->args.sc.fast4_ok = fast.ok4 ? 1u : 0u;
        s->fast_need4 = fast.need4;
        s->args.sc.fast4_root = fast.root4;
        if (fast.ok4 && (rc = dev_upload(s, &s->args.sc.fast4, fast.nodes4))) return rc;
        if ((rc = dev_upload(s, &s->args.sc.fast_nodes, fast.nodes))) return rc;
        s->args.sc.fast_tlas_base = fast.tlas_base;
        if ((rc = dev_upload(s, &s->args.sc.fast_tris, fast.tris))) return rc;
        if ((rc = dev_upload(s, &s->args.sc.tri_leaf, fast.tri_leaf))) return rc;
    }
    if ((rc = dev_upload(s, &s->args.sc.wide_nodes, lay.wide_nodes))) return rc;
    if ((rc = dev_upload(s, &s->args.sc.leaf_recs, lay.leaf_recs))) return rc;
    if ((rc = dev_upload(s, &s->args.sc.wide_tlas, lay.wide_tlas))) return rc;
    if ((rc = dev_upload(s, &s->args.sc.inst_recs, lay.inst_recs))) return rc;
    s->args.sc.tlas_root_link = lay.tlas_root_link;
    s->args.sc.fast_world_reach = lay.world_reach;
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream)); // host vectors die at return
    return GDPT_OK;
}

void compute_shard(gdpt_shader *s)
{
    FrameArgs &a = s->args;
    a.shard_part = s->shard_part; a.shard_parts = s->shard_parts; a.shard_band = s->shard_band;
    int rows = 0;
    if (a.shard_parts <= 1) rows = a.height;
    else
        for (int y = 0; y < a.height; y++)
            if ((y / a.shard_band) % a.shard_parts == a.shard_part) rows++;
    a.local_rows = rows;
    const uint32_t tiles_x = ((uint32_t)a.width + 7u) / 8u, tiles_y = ((uint32_t)rows + 3u) / 4u;
    a.n_work = tiles_x * tiles_y * 32u;
}

int ensure_slot(gdpt_shader *m, gdpt_shader::FrameSlot &sl, bool want_depth);

int alloc_warp_profile(gdpt_shader *s)
{
    FrameArgs &a = s->args;
    if (a.warp_prof) return GDPT_OK;
    s->warp_prof_warps = path_kernel_warps(a);
    int rc = dev_alloc(s, &a.warp_prof, s->warp_prof_warps * 8);
    if (rc) return rc;
    GDPT_CUDA(s->dev, cudaMemsetAsync(a.warp_prof, 0, s->warp_prof_warps * 8 * sizeof(unsigned long long), s->dev->stream));
    return GDPT_OK;
}

int finish_main(gdpt_shader *s)
{
    gdpt_device *d = s->dev;
    Resource *out = bound(s, 0, 0), *depth = bound(s, 0, 1), *params = bound(s, 0, 2), *camera = bound(s, 0, 3);
    Resource *tg = bound(s, 1, 0), *td = bound(s, 1, 1), *mat = bound(s, 1, 2), *bvh = bound(s, 1, 3), *blas = bound(s, 1, 4),
             *tlas = bound(s, 1, 5), *tex = bound(s, 2, 0);
    if (!out || !depth || !params || !camera || !tg || !td || !mat || !bvh || !blas || !tlas || !tex)
        return fail(d, GDPT_ERR_BAD_BINDING, "main.glsl needs set0 b0-3, set1 b0-5 and set2 b0 (main.glsl:98-155)");
    if (out->kind != RES_IMAGE || out->format != GDPT_FORMAT_R8G8B8A8_UNORM) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b0 must be an rgba8 image");
    if (depth->kind != RES_IMAGE || depth->format != GDPT_FORMAT_R32_SFLOAT) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b1 must be an r32f image");
    if (params->kind != RES_BUFFER || params->size < sizeof(gdpt_render_params)) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b2 must hold the 36 B Params block");
    if (camera->kind != RES_BUFFER || camera->size < 156) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b3 must hold the Camera block");
    if (tex->kind != RES_LAYERED || tex->format != GDPT_FORMAT_R8G8B8A8_UNORM) return fail(d, GDPT_ERR_BAD_BINDING, "set2 b0 must be an rgba8 texture array");
    if (tg->size % sizeof(gdpt_triangle_geometry) || td->size % sizeof(gdpt_triangle_data) || mat->size % sizeof(gdpt_material) ||
        bvh->size % sizeof(gdpt_bvh_node) || blas->size % sizeof(gdpt_blas_instance) || tlas->size % sizeof(gdpt_tlas_node))
        return fail(d, GDPT_ERR_BAD_BINDING, "a scene buffer is not a whole number of records");
    if (tg->size / sizeof(gdpt_triangle_geometry) != td->size / sizeof(gdpt_triangle_data))
        return fail(d, GDPT_ERR_BAD_BINDING, "triangle geometry/data arrays differ in length");

    gdpt_render_params rp;
    memcpy(&rp, params->shadow.data(), sizeof(rp));
    if (rp.width <= 0 || rp.height <= 0 || rp.width != out->width || rp.height != out->height || depth->width != out->width ||
        depth->height != out->height)
        return fail(d, GDPT_ERR_BAD_BINDING, "Params width/height (%d x %d) must match the bound images", rp.width, rp.height);

    // A second finish (scene rebuild: add_existing_buffer / create_*_uniform cleared uniforms_ready) frees every derived
    // buffer, so nothing may still be using them: refuse while pipelined frames are pending, drain every stream a frame
    // may have run on, and forget the frame slots' pointers into the freed set.
    for (const auto &sl : s->slots)
        if (sl.pending) return fail(d, GDPT_ERR_NOT_READY, "finish_create_uniforms while pipelined frames are in flight: call gdpt_render_frame_wait first");
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    if (d->extra_streams_made)
        for (cudaStream_t x : d->extra_streams) GDPT_CUDA(d, cudaStreamSynchronize(x));
    if (d->copy_stream) GDPT_CUDA(d, cudaStreamSynchronize(d->copy_stream));
    for (void *p : s->derived) cudaFree(p);
    s->derived.clear();
    FrameArgs &a = s->args;
    memset(&a, 0, sizeof(a));
    a.sc.tri_geom = static_cast<const gdpt_triangle_geometry *>(tg->dptr);
    a.sc.tri_data = static_cast<const gdpt_triangle_data *>(td->dptr);
    a.sc.materials = static_cast<const gdpt_material *>(mat->dptr);
    a.sc.bvh = static_cast<const gdpt_bvh_node *>(bvh->dptr);
    a.sc.blas = static_cast<const gdpt_blas_instance *>(blas->dptr);
    a.sc.tlas = static_cast<const gdpt_tlas_node *>(tlas->dptr);
    a.sc.textures = static_cast<const uint8_t *>(tex->dptr);
    a.sc.tex_w = tex->width; a.sc.tex_h = tex->height; a.sc.tex_layers = tex->layers;
    a.sc.n_tris = (uint32_t)(tg->size / sizeof(gdpt_triangle_geometry));
    a.sc.n_nodes = (uint32_t)(bvh->size / sizeof(gdpt_bvh_node));
    a.sc.n_blas = (uint32_t)(blas->size / sizeof(gdpt_blas_instance));
    a.sc.n_tlas = (uint32_t)(tlas->size / sizeof(gdpt_tlas_node));
    a.sc.n_materials = (uint32_t)(mat->size / sizeof(gdpt_material));
    if (a.sc.n_blas > 32768u || a.sc.n_tlas > 65535u)
        return fail(d, GDPT_ERR_UNSUPPORTED, "%u instances / %u TLAS nodes: the reference's TLAS addresses at most 65 535 nodes (main.glsl:329-330)", a.sc.n_blas, a.sc.n_tlas);
    int rc = build_derived_layout(s, *bvh, *blas, *tlas, *tg);
    if (rc) return rc;

    if (s->material_ext) { // the reference's padding words carry texture slots and flags; optional material tables
        const gdpt_material *mm = reinterpret_cast<const gdpt_material *>(mat->shadow.data());
        for (uint32_t i = 0; i < a.sc.n_materials; i++)
            if (mm[i].ext_roughness_texture > (uint32_t)tex->layers || mm[i].ext_metallic_texture > (uint32_t)tex->layers)
                return fail(d, GDPT_ERR_BAD_BINDING, "material %u names a texture layer beyond the %d bound ones", i, tex->layers);
        if (Resource *sm = bound(s, 1, 6)) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(sm->shadow.data());
            const uint64_t words = sm->size / 4u;
            if (sm->kind != RES_BUFFER || words < (uint64_t)a.sc.n_blas + 1u)
                return fail(d, GDPT_ERR_BAD_BINDING, "set1 b6 must hold offset[n_instances + 1] followed by the material ids");
            for (uint32_t i = 0; i <= a.sc.n_blas; i++)
                if (w[i] < a.sc.n_blas + 1u || w[i] > words || (i > 0 && w[i] < w[i - 1]))
                    return fail(d, GDPT_ERR_BAD_BINDING, "set1 b6: offset %u out of range", i);
            for (uint64_t k = a.sc.n_blas + 1u; k < words; k++)
                if (w[k] >= a.sc.n_materials) return fail(d, GDPT_ERR_BAD_BINDING, "set1 b6: material id %u of %u", w[k], a.sc.n_materials);
            a.sc.surface_materials = static_cast<const uint32_t *>(sm->dptr);
        }
        std::vector<float> lut(256);
        for (int k = 0; k < 256; k++) { // IEC 61966-2-1 decode of an 8-bit code, rounded once to binary32
            const double c = k / 255.0;
            lut[k] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        }
        if ((rc = dev_upload(s, &a.sc.srgb_lut, lut))) return rc;
        GDPT_CUDA(d, cudaStreamSynchronize(d->stream)); // `lut` dies with this block
        a.sc.material_ext = 1u;
    }
    a.camera = static_cast<const gdpt_camera *>(camera->dptr);
    a.width = rp.width; a.height = rp.height; a.max_depth = s->max_depth;
    a.out_rgba8 = static_cast<uint32_t *>(out->dptr);
    a.out_depth = static_cast<float *>(depth->dptr);
    a.queue_cap = (uint32_t)((size_t)rp.width * rp.height);
    a.heavy_cap = a.queue_cap / 4u > 1024u ? a.queue_cap / 4u : 1024u;
    if ((rc = dev_alloc(s, &a.hit_list, (size_t)a.queue_cap + (size_t)(kCostClasses - 1) * a.heavy_cap))) return rc;
    if ((rc = dev_alloc(s, &a.cost, (size_t)a.queue_cap))) return rc;
    GDPT_CUDA(d, cudaMemsetAsync(a.cost, 0, (size_t)a.queue_cap * sizeof(uint32_t), d->stream));
    if ((rc = dev_alloc(s, &a.counters, 1))) return rc;
    // Which kernels run is decided by the shader's own "#define" list (gdpt_shader_create) and by what the arrays
    // allow -- never by the process environment.
    // Default: the closest-hit search with pooled paths (6).
    a.schedule = (s->variant == 2 || s->variant == 3 || s->variant == 6) ? s->variant : 6;
    if (a.schedule == 6 && !a.sc.fast_ok) a.schedule = 3; // closest-hit tables unavailable for these arrays (fast_bvh.h)
    // culling (pt_scene.cuh) is the default for rendering; parity traces and the DEBUG_STEPS heat map
    // keep the full reference visit order (their output IS the reference's work)
    const bool observes_work = s->trace_segments > 0 || s->debug_steps;
    a.cull = s->cull >= 0 ? s->cull : (observes_work ? 0 : 1);
    // the classification kernel is a tight-box test and writes no parity records
    if (a.schedule >= 3 && (!a.cull || observes_work)) a.schedule = 2;
    init_launch_shapes(d->ordinal);
    // scheduling knobs ("#define GDPT_TUNE_<NAME> n"; results do not depend on them)
    // schedule 6: lanes without a walking ray before a pool service (8 | 12: C2 0.609 | 0.594 ms, C4 1080p 10.30 | 10.66)
    a.refill_below = tune(s, "REFILL_BELOW", a.schedule == 6 ? (a.sc.n_blas >= 64u ? 8 : 12) : 24);
    a.burst = tune(s, "BURST", a.schedule == 3 ? 4 : (a.schedule == 6 ? 8 : 16));
    a.shade_at = tune(s, "SHADE_AT", a.schedule == 6 ? 24 : 8);          // schedule 6: finished rays that justify a partial batch
    a.blocks_per_sm = tune(s, "BLOCKS_PER_SM", 0);
    a.wide_bvh = tune(s, "WIDE_BVH", 1);
    a.cost_ema = tune(s, "COST_EMA", 1);
    a.pool_alive = tune(s, "POOL_ALIVE", 0);
    a.pool_wait = tune(s, "POOL_WAIT", 16); // A/B with every phase per iteration (16 | 32): C2 0.598 | 0.594 ms, C4 1080p 10.26 | 10.66
    a.lead_min = tune(s, "LEAD_MIN", 0);
    a.sort4 = tune(s, "SORT4", 1);
    a.miss_now = tune(s, "MISS_NOW", 1);
    // A frame of a few instances is one wave of paths and bound by their latency: 128 registers, no spills, four blocks per
    // SM.  Frames with far more traversal per path are throughput-bound and take the fifth block at 96 registers
    // (A/B 4 | 5 blocks: C2 path kernel 0.606 | 0.638 ms, C4 1080p 10.22 | 9.76, C3 4.64 | 4.41).
    a.pool_dense = tune(s, "DENSE", (a.sc.n_blas >= 64u || a.sc.n_tris >= 262144u) ? 1 : 0);
    a.count_work = s->count_work ? 1 : 0;
    if (a.refill_below < 1) a.refill_below = 1;
    if (a.refill_below > 32) a.refill_below = 32;
    if (a.burst < 1) a.burst = 1;
    a.debug_steps = s->debug_steps ? 1 : 0;
    a.trace_segments = s->trace_segments;
    a.visits_per_ray = s->visits_per_ray;
    if (s->record_hits > 0 && !observes_work) a.trace_segments = s->record_hits < s->max_depth ? s->record_hits : s->max_depth;
    if (a.trace_segments > 0) {
        const size_t n = (size_t)a.trace_segments * rp.width * rp.height;
        if ((rc = dev_alloc(s, &a.trace, n))) return rc;
        if (s->visits_per_ray > 0 && (rc = dev_alloc(s, &a.visits, (size_t)rp.width * rp.height * s->visits_per_ray))) return rc;
    }
    compute_shard(s);
    // the two frame slots of gdpt_render_frame_begin are set up here, not on the first pipelined frame
    for (auto &sl : s->slots) {
        sl.dcnt = nullptr; sl.stage_rgba8 = nullptr; sl.stage_depth = nullptr; // (re)allocated with the derived buffers
        sl.cam_dev = nullptr; sl.pp_dev = nullptr; sl.raw_rgba8 = nullptr; sl.raw_depth = nullptr; sl.hit_list = nullptr;
        sl.finished_once = false;
        if ((rc = ensure_slot(s, sl, false))) return rc;
    }
    if (s->warp_profile) return alloc_warp_profile(s);
    return GDPT_OK;
}

int finish_progressive(gdpt_shader *s)
{
    gdpt_device *d = s->dev;
    Resource *params = bound(s, 0, 0), *screen = bound(s, 0, 1), *accum = bound(s, 0, 2);
    if (!params || !screen || !accum) return fail(d, GDPT_ERR_BAD_BINDING, "progressive_rendering.glsl needs set0 b0-2 (progressive_rendering.glsl:5-16)");
    if (params->kind != RES_BUFFER || params->size < sizeof(gdpt_progressive_params)) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b0 must hold the 12 B Params block");
    if (screen->kind != RES_IMAGE || screen->format != GDPT_FORMAT_R8G8B8A8_UNORM) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b1 must be the rgba8 screen image");
    if (accum->kind != RES_IMAGE || accum->format != GDPT_FORMAT_R32G32B32A32_SFLOAT) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b2 must be an rgba32f image");
    if (screen->width != accum->width || screen->height != accum->height) return fail(d, GDPT_ERR_BAD_BINDING, "screen and accumulation images differ in size");
    init_launch_shapes(d->ordinal);
    return GDPT_OK;
}

int finish_temporal(gdpt_shader *s)
{
    gdpt_device *d = s->dev;
    Resource *params = bound(s, 0, 0), *screen = bound(s, 0, 1), *depth = bound(s, 0, 2), *fb1 = bound(s, 0, 3), *fb2 = bound(s, 0, 4);
    if (!params || !screen || !depth || !fb1 || !fb2) return fail(d, GDPT_ERR_BAD_BINDING, "temporal_reprojection.glsl needs set0 b0-4 (temporal_reprojection.glsl:4-18)");
    if (params->kind != RES_BUFFER || params->size < sizeof(gdpt_temporal_params)) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b0 must hold the 88 B Params block");
    if (screen->kind != RES_IMAGE || screen->format != GDPT_FORMAT_R8G8B8A8_UNORM) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b1 must be the rgba8 screen image");
    if (depth->kind != RES_IMAGE || depth->format != GDPT_FORMAT_R32_SFLOAT) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b2 must be the r32f depth image");
    for (Resource *fb : { fb1, fb2 }) {
        if (fb->kind != RES_IMAGE || fb->format != GDPT_FORMAT_R32G32B32A32_SFLOAT) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b3/b4 must be rgba32f images");
        if (fb->width != screen->width || fb->height != screen->height) return fail(d, GDPT_ERR_BAD_BINDING, "frame buffers and screen differ in size");
    }
    if (fb1 == fb2) return fail(d, GDPT_ERR_BAD_BINDING, "set0 b3 and b4 must be two different images (ping-pong)");
    if (depth->width != screen->width || depth->height != screen->height) return fail(d, GDPT_ERR_BAD_BINDING, "depth and screen differ in size");
    init_launch_shapes(d->ordinal);
    return GDPT_OK;
}

// K3 with the Params block as last uploaded; the ping-pong roles follow frameCount's parity (temporal_reprojection.glsl:46,62,66).
int enqueue_k3(gdpt_shader *t)
{
    gdpt_device *d = t->dev;
    Resource *params = bound(t, 0, 0), *screen = bound(t, 0, 1), *depth = bound(t, 0, 2), *fb1 = bound(t, 0, 3), *fb2 = bound(t, 0, 4);
    gdpt_temporal_params host;
    memcpy(&host, params->shadow.data(), sizeof(host));
    if (host.width != screen->width || host.height != screen->height)
        return fail(d, GDPT_ERR_INVALID_ARG, "temporal Params say %dx%d, the screen image is %dx%d", host.width, host.height, screen->width, screen->height);
    const bool use_first = (host.frame_count % 2u) == 0u;
    launch_temporal(static_cast<uint32_t *>(screen->dptr), static_cast<const float *>(depth->dptr),
                    static_cast<const float *>(use_first ? fb1->dptr : fb2->dptr), static_cast<float *>(use_first ? fb2->dptr : fb1->dptr),
                    static_cast<const gdpt_temporal_params *>(params->dptr), screen->width, screen->height, d->stream);
    GDPT_CUDA(d, cudaGetLastError());
    return GDPT_OK;
}

// What the frame calls require of the post-process shader for `mode`.
int check_post(gdpt_shader *m, gdpt_shader *p, gdpt_denoising mode)
{
    gdpt_device *d = m->dev;
    if (mode == GDPT_DENOISE_PROGRESSIVE_RENDERING) {
        if (!p || p->kind != SHADER_PROGRESSIVE || p->dev != d || !gdpt_shader_check_ready(p)) return fail(d, GDPT_ERR_NOT_READY, "progressive shader missing or not ready");
        if (bound(p, 0, 1) != bound(m, 0, 0)) return fail(d, GDPT_ERR_BAD_BINDING, "progressive shader must share the main shader's output image (add_existing_buffer)");
    } else if (mode == GDPT_DENOISE_TEMPORAL_REPROJECTION) {
        if (!p || p->kind != SHADER_TEMPORAL || p->dev != d || !gdpt_shader_check_ready(p)) return fail(d, GDPT_ERR_NOT_READY, "temporal reprojection shader missing or not ready");
        if (bound(p, 0, 1) != bound(m, 0, 0) || bound(p, 0, 2) != bound(m, 0, 1))
            return fail(d, GDPT_ERR_BAD_BINDING, "temporal shader must share the main shader's output and depth images (add_existing_buffer)");
        if (m->shard_parts > 1) return fail(d, GDPT_ERR_UNSUPPORTED, "temporal reprojection reads other rows' history: not available on a row-sharded frame");
    } else if (mode != GDPT_DENOISE_NONE) {
        return fail(d, GDPT_ERR_INVALID_ARG, "unknown denoising mode %d", (int)mode);
    }
    return GDPT_OK;
}

// Stream-ordered upload of the post process's Params block from the pinned block `stage` (>= 256 B free).
int upload_post_params(gdpt_shader *m, gdpt_shader *p, gdpt_denoising mode, uint32_t frame_count, uint8_t *stage)
{
    gdpt_device *d = m->dev;
    Resource *pp = bound(p, 0, 0);
    if (mode == GDPT_DENOISE_PROGRESSIVE_RENDERING) {
        gdpt_progressive_params host_pp = { m->args.width, m->args.height, frame_count };
        memcpy(stage, &host_pp, sizeof(host_pp));
        memcpy(pp->shadow.data(), &host_pp, sizeof(host_pp));
        GDPT_CUDA(d, cudaMemcpyAsync(pp->dptr, stage, sizeof(host_pp), cudaMemcpyHostToDevice, d->stream));
    } else if (!p->staged_params.empty()) { // temporal: the block the host staged (gdpt_shader_stage_params)
        const size_t n = p->staged_params.size();
        memcpy(stage, p->staged_params.data(), n);
        memcpy(pp->shadow.data(), p->staged_params.data(), n);
        p->staged_params.clear();
        GDPT_CUDA(d, cudaMemcpyAsync(pp->dptr, stage, n, cudaMemcpyHostToDevice, d->stream));
    }
    return GDPT_OK;
}

// Enqueue one K1 dispatch on the device stream.
int enqueue_k1(gdpt_shader *s)
{
    gdpt_device *d = s->dev;
    FrameArgs &a = s->args;
    const bool trace = s->trace_segments > 0 || s->debug_steps;
    const bool rec = !trace && a.trace != nullptr; // "#define GDPT_RECORD_HITS": hit records from the rendering kernels
    GDPT_CUDA(d, cudaMemsetAsync(a.counters, 0, sizeof(FrameCounters), d->stream));
    if ((trace || rec) && a.trace) GDPT_CUDA(d, cudaMemsetAsync(a.trace, 0xFF, (size_t)a.trace_segments * a.width * a.height * sizeof(gdpt_trace_record), d->stream));
    if (trace && a.visits) GDPT_CUDA(d, cudaMemsetAsync(a.visits, 0xFF, (size_t)a.width * a.height * a.visits_per_ray * sizeof(uint32_t), d->stream));
    const bool timing = s->stage_timing && !s->stage_ev.empty();
    int ev = 0;
    if (timing) GDPT_CUDA(d, cudaEventRecord(s->stage_ev[ev++], d->stream));
    if (a.schedule >= 3) {
        launch_primary_cull(a, d->stream);
        if (timing) GDPT_CUDA(d, cudaEventRecord(s->stage_ev[ev++], d->stream));
        if (a.schedule == 6) launch_path_pool(a, rec, d->stream);
        else launch_path_list(a, rec, d->stream);
    } else {
        launch_path(a, trace, d->stream);
    }
    if (timing) GDPT_CUDA(d, cudaEventRecord(s->stage_ev[ev++], d->stream));
    s->stage_count = timing ? ev - 1 : 0;
    GDPT_CUDA(d, cudaGetLastError());
    s->stats_valid = false;
    s->stats.kernel_launches = (uint32_t)k1_launch_count(a.schedule);
    return GDPT_OK;
}

int enqueue_k2(gdpt_shader *p, int part, int parts, int band, const uint32_t *raw_in = nullptr,
               const gdpt_progressive_params *params_in = nullptr)
{
    gdpt_device *d = p->dev;
    Resource *params = bound(p, 0, 0), *screen = bound(p, 0, 1), *accum = bound(p, 0, 2);
    launch_progressive(raw_in ? raw_in : static_cast<const uint32_t *>(screen->dptr), static_cast<uint32_t *>(screen->dptr),
                       static_cast<float4 *>(accum->dptr),
                       params_in ? params_in : static_cast<const gdpt_progressive_params *>(params->dptr), 0u, screen->width, screen->height,
                       part, parts, band, p->peers, d->stream);
    GDPT_CUDA(d, cudaGetLastError());
    return GDPT_OK;
}

int collect_stats(gdpt_shader *s)
{
    gdpt_device *d = s->dev;
    FrameCounters c;
    GDPT_CUDA(d, cudaMemcpyAsync(&c, s->args.counters, sizeof(c), cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    gdpt_frame_stats &st = s->stats;
    st.rays = c.rays;
    st.primary_hits = c.primary_hits;
    st.retraced = c.retraced;
    st.own_node_steps = c.own_node_steps; st.own_box_tests = c.own_box_tests; st.own_tri_tests = c.own_tri_tests;
    st.own_inst_entries = c.own_inst_entries; st.own_proofs = c.own_proofs;
    st.node_pops = c.node_pops; st.box_tests = c.box_tests; st.tri_tests = c.tri_tests; st.tlas_leaves = c.tlas_leaves;
    st.max_stack = c.max_stack;
    if (c.overflow) return fail(d, GDPT_ERR_UNSUPPORTED, "a ray exceeded the reference's 64+64 traversal stack entries");
    s->stats_valid = true;
    return GDPT_OK;
}

} // namespace

extern "C" {

uint32_t gdpt_abi_version(void) { return 2u; }

const char *gdpt_last_error(const gdpt_device *device) { return device ? device->last_error.c_str() : g_create_error.c_str(); }

int gdpt_device_create(int cuda_ordinal, gdpt_device **out_device)
{
    if (!out_device) return fail(nullptr, GDPT_ERR_INVALID_ARG, "out_device is NULL");
    *out_device = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, GDPT_ERR_NO_DEVICE, "no CUDA device available (%s); this backend has no CPU path", cudaGetErrorString(e));
    if (cuda_ordinal < 0 || cuda_ordinal >= count) return fail(nullptr, GDPT_ERR_NO_DEVICE, "CUDA ordinal %d out of range (0..%d)", cuda_ordinal, count - 1);
    gdpt_device *d = new gdpt_device();
    d->ordinal = cuda_ordinal;
    if (cudaSetDevice(cuda_ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc(&d->pinned_staging, 4096, cudaHostAllocDefault) != cudaSuccess) {
        fail(nullptr, GDPT_ERR_CUDA, "device %d: stream/staging setup failed: %s", cuda_ordinal, cudaGetErrorString(cudaGetLastError()));
        delete d;
        return GDPT_ERR_CUDA;
    }
    for (int i = 0; i < 4; i++) cudaEventCreate(&d->ev[i]);
    if (cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void **>(&d->ring), 8 * 512, cudaHostAllocDefault) != cudaSuccess) {
        fail(nullptr, GDPT_ERR_CUDA, "device %d: copy stream / staging ring setup failed: %s", cuda_ordinal, cudaGetErrorString(cudaGetLastError()));
        gdpt_device_destroy(d);
        return GDPT_ERR_CUDA;
    }
    for (int i = 0; i < 8; i++) cudaEventCreateWithFlags(&d->ring_ev[i], cudaEventDisableTiming);
    init_launch_shapes(cuda_ordinal);
    *out_device = d;
    return GDPT_OK;
}

void gdpt_device_destroy(gdpt_device *d)
{
    if (!d) return;
    cudaSetDevice(d->ordinal);
    cudaStreamSynchronize(d->stream);
    for (auto &kv : d->resources) cudaFree(kv.second.dptr);
    for (int i = 0; i < 4; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    if (d->pinned_staging) cudaFreeHost(d->pinned_staging);
    if (d->copy_stream) { cudaStreamSynchronize(d->copy_stream); cudaStreamDestroy(d->copy_stream); }
    for (cudaStream_t &x : d->extra_streams) if (x) { cudaStreamSynchronize(x); cudaStreamDestroy(x); }
    if (d->ring) cudaFreeHost(d->ring);
    for (int i = 0; i < 8; i++) if (d->ring_ev[i]) cudaEventDestroy(d->ring_ev[i]);
    cudaStreamDestroy(d->stream);
    delete d;
}

int gdpt_device_synchronize(gdpt_device *d)
{
    if (!d) return GDPT_ERR_INVALID_ARG;
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    for (cudaStream_t x : d->extra_streams) if (x) GDPT_CUDA(d, cudaStreamSynchronize(x));
    return GDPT_OK;
}

uint64_t gdpt_device_stream(gdpt_device *d) { return d ? (uint64_t)(uintptr_t)d->stream : 0; }

void *gdpt_host_alloc(uint64_t size)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, size, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void gdpt_host_free(void *p) { if (p) cudaFreeHost(p); }

int gdpt_shader_create(gdpt_device *d, const char *shader_path, const char *const *args, int n_args, gdpt_shader **out)
{
    if (!d || !out) return fail(d, GDPT_ERR_INVALID_ARG, "device/out_shader is NULL");
    *out = nullptr;
    const std::string base = basename_of(shader_path);
    gdpt_shader *s = new gdpt_shader();
    s->dev = d;
    if (base == "main.glsl") s->kind = SHADER_MAIN;
    else if (base == "progressive_rendering.glsl") s->kind = SHADER_PROGRESSIVE;
    else if (base == "temporal_reprojection.glsl") s->kind = SHADER_TEMPORAL;
    else {
        delete s;
        return fail(d, GDPT_ERR_UNKNOWN_SHADER, "no CUDA kernel set for shader '%s'", shader_path ? shader_path : "(null)");
    }
    for (int i = 0; i < n_args; i++) {
        std::string name; long value = 0; bool has_value = false;
        if (!args[i] || !parse_define(args[i], &name, &value, &has_value)) continue;
        if (name == "DEBUG_STEPS") s->debug_steps = true;
        else if (name == "MAX_DEPTH" && has_value) s->max_depth = (int)value;
        else if (name == "GDPT_TRACE") s->trace_segments = has_value ? (int)value : 1;
        else if (name == "GDPT_TRACE_VISITS" && has_value) s->visits_per_ray = (uint32_t)value;
        else if (name == "GDPT_VARIANT" && has_value) s->variant = (int)value;
        else if (name == "GDPT_CULL" && has_value) s->cull = value != 0 ? 1 : 0;
        else if (name == "GDPT_REFERENCE_ORDER") s->cull = 0;
        else if (name == "GDPT_RECORD_HITS" && has_value) s->record_hits = (int)value;
        else if (name == "GDPT_COUNT_WORK") s->count_work = true;
        else if (name == "GDPT_MATERIAL_EXT") s->material_ext = true;
        else if (name.rfind("GDPT_TUNE_", 0) == 0 && has_value) s->tuning[name.substr(10)] = (int)value;
    }
    if (s->max_depth < 1 || s->max_depth > kMaxDepth) {
        delete s;
        return fail(d, GDPT_ERR_INVALID_ARG, "MAX_DEPTH must be in 1..%d", (int)kMaxDepth);
    }
    if (s->trace_segments > s->max_depth) s->trace_segments = s->max_depth;
    memset(&s->args, 0, sizeof(s->args));
    memset(&s->stats, 0, sizeof(s->stats));
    s->initialized = true;
    *out = s;
    return GDPT_OK;
}

void gdpt_shader_destroy(gdpt_shader *s)
{
    if (!s) return;
    gdpt_device *d = s->dev;
    cudaSetDevice(d->ordinal);
    cudaStreamSynchronize(d->stream);
    for (cudaStream_t x : d->extra_streams) if (x) cudaStreamSynchronize(x);
    if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
    for (void *p : s->derived) cudaFree(p);
    for (cudaEvent_t e : s->stage_ev) cudaEventDestroy(e);
    for (auto &sl : s->slots) {
        if (sl.hcnt) cudaFreeHost(sl.hcnt);
        for (cudaEvent_t e : { sl.k_done, sl.done, sl.t0, sl.t1, sl.t2, sl.t3 }) if (e) cudaEventDestroy(e);
    }
    for (gdpt_rid rid : s->owned) {
        Resource *r = find(d, rid);
        if (r) { cudaFree(r->dptr); d->resources.erase(rid); }
    }
    delete s;
}

static gdpt_rid new_resource(gdpt_shader *s, Resource &&r, int binding, int set)
{
    gdpt_device *d = s->dev;
    const gdpt_rid rid = d->next_rid++;
    d->resources[rid] = std::move(r);
    s->owned.push_back(rid);
    s->bindings[{ set, binding }] = rid;
    s->uniforms_ready = false;
    return rid;
}

gdpt_rid gdpt_shader_create_storage_buffer_uniform(gdpt_shader *s, const void *data, uint64_t size, int binding, int set)
{
    if (!s || !s->initialized || (size && !data)) { if (s) fail(s->dev, GDPT_ERR_INVALID_ARG, "create_storage_buffer_uniform: bad arguments"); return 0; }
    gdpt_device *d = s->dev;
    cudaSetDevice(d->ordinal);
    Resource r;
    r.kind = RES_BUFFER; r.size = size;
    if (cudaMalloc(&r.dptr, size ? size : 16) != cudaSuccess) { fail(d, GDPT_ERR_CUDA, "cudaMalloc(%llu) failed: %s", (unsigned long long)size, cudaGetErrorString(cudaGetLastError())); return 0; }
    r.shadow.assign(static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + size);
    if (size && cudaMemcpyAsync(r.dptr, r.shadow.data(), size, cudaMemcpyHostToDevice, d->stream) != cudaSuccess) {
        fail(d, GDPT_ERR_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(r.dptr);
        return 0;
    }
    cudaStreamSynchronize(d->stream);
    return new_resource(s, std::move(r), binding, set);
}

int gdpt_shader_update_storage_buffer_uniform(gdpt_shader *s, gdpt_rid rid, const void *data, uint64_t size)
{
    if (!s || !data) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    Resource *r = find(d, rid);
    if (!r || r->kind != RES_BUFFER) return fail(d, GDPT_ERR_INVALID_ARG, "update_storage_buffer_uniform: unknown buffer RID");
    if (size > r->size) return fail(d, GDPT_ERR_INVALID_ARG, "update of %llu B exceeds the %llu B buffer", (unsigned long long)size, (unsigned long long)r->size);
    cudaSetDevice(d->ordinal);
    memcpy(r->shadow.data(), data, size);
    GDPT_CUDA(d, cudaMemcpyAsync(r->dptr, r->shadow.data(), size, cudaMemcpyHostToDevice, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return GDPT_OK;
}

int gdpt_shader_stage_params(gdpt_shader *s, const void *data, uint64_t size)
{
    if (!s || !data) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    Resource *r = bound(s, 0, 0);
    if (s->kind == SHADER_MAIN || !r || r->kind != RES_BUFFER) return fail(d, GDPT_ERR_INVALID_ARG, "stage_params: not a post-process shader with a Params block at set0 b0");
    if (size > r->size || size > 256) return fail(d, GDPT_ERR_INVALID_ARG, "stage_params: %llu B exceed the Params block", (unsigned long long)size);
    s->staged_params.assign(static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + size);
    return GDPT_OK;
}

int gdpt_shader_get_storage_buffer_uniform(gdpt_shader *s, gdpt_rid rid, void *out, uint64_t capacity)
{
    if (!s || !out) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    Resource *r = find(d, rid);
    if (!r || r->kind != RES_BUFFER) return fail(d, GDPT_ERR_INVALID_ARG, "get_storage_buffer_uniform: unknown buffer RID");
    if (capacity < r->size) return fail(d, GDPT_ERR_INVALID_ARG, "output capacity too small");
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaMemcpyAsync(out, r->dptr, r->size, cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return GDPT_OK;
}

gdpt_rid gdpt_shader_create_image_uniform(gdpt_shader *s, const void *pixels, int width, int height, gdpt_data_format format,
                                          int binding, int set)
{
    if (!s || !s->initialized || width <= 0 || height <= 0 || texel_bytes(format) == 0) { if (s) fail(s->dev, GDPT_ERR_INVALID_ARG, "create_image_uniform: bad arguments"); return 0; }
    gdpt_device *d = s->dev;
    cudaSetDevice(d->ordinal);
    Resource r;
    r.kind = RES_IMAGE; r.width = width; r.height = height; r.format = format; r.layers = 1;
    r.size = (uint64_t)width * height * texel_bytes(format);
    if (cudaMalloc(&r.dptr, r.size) != cudaSuccess) { fail(d, GDPT_ERR_CUDA, "cudaMalloc(%llu) failed: %s", (unsigned long long)r.size, cudaGetErrorString(cudaGetLastError())); return 0; }
    cudaError_t e = pixels ? cudaMemcpyAsync(r.dptr, pixels, r.size, cudaMemcpyHostToDevice, d->stream) : cudaMemsetAsync(r.dptr, 0, r.size, d->stream);
    if (e != cudaSuccess) { fail(d, GDPT_ERR_CUDA, "image upload failed: %s", cudaGetErrorString(e)); cudaFree(r.dptr); return 0; }
    cudaStreamSynchronize(d->stream);
    return new_resource(s, std::move(r), binding, set);
}

gdpt_rid gdpt_shader_create_layered_image_uniform(gdpt_shader *s, const void *const *layers, int n_layers, int width, int height,
                                                  gdpt_data_format format, int binding, int set)
{
    if (!s || !s->initialized || !layers || n_layers <= 0 || width <= 0 || height <= 0 || format != GDPT_FORMAT_R8G8B8A8_UNORM) {
        if (s) fail(s->dev, GDPT_ERR_INVALID_ARG, "create_layered_image_uniform: bad arguments");
        return 0;
    }
    gdpt_device *d = s->dev;
    cudaSetDevice(d->ordinal);
    Resource r;
    r.kind = RES_LAYERED; r.width = width; r.height = height; r.format = format; r.layers = n_layers;
    const uint64_t layer_bytes = (uint64_t)width * height * 4;
    r.size = layer_bytes * n_layers;
    if (cudaMalloc(&r.dptr, r.size) != cudaSuccess) { fail(d, GDPT_ERR_CUDA, "cudaMalloc(%llu) failed: %s", (unsigned long long)r.size, cudaGetErrorString(cudaGetLastError())); return 0; }
    for (int l = 0; l < n_layers; l++) {
        cudaError_t e = layers[l] ? cudaMemcpyAsync(static_cast<uint8_t *>(r.dptr) + l * layer_bytes, layers[l], layer_bytes, cudaMemcpyHostToDevice, d->stream)
                                  : cudaMemsetAsync(static_cast<uint8_t *>(r.dptr) + l * layer_bytes, 0, layer_bytes, d->stream);
        if (e != cudaSuccess) { fail(d, GDPT_ERR_CUDA, "layer upload failed: %s", cudaGetErrorString(e)); cudaFree(r.dptr); return 0; }
    }
    cudaStreamSynchronize(d->stream);
    return new_resource(s, std::move(r), binding, set);
}

int gdpt_shader_get_image_uniform_buffer(gdpt_shader *s, gdpt_rid rid, int layer, void *out, uint64_t capacity)
{
    if (!s || !out) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    Resource *r = find(d, rid);
    if (!r || r->kind == RES_BUFFER) return fail(d, GDPT_ERR_INVALID_ARG, "get_image_uniform_buffer: unknown image RID");
    if (layer < 0 || layer >= r->layers) return fail(d, GDPT_ERR_INVALID_ARG, "layer %d out of range", layer);
    const uint64_t layer_bytes = r->size / (uint64_t)r->layers;
    if (capacity < layer_bytes) return fail(d, GDPT_ERR_INVALID_ARG, "output capacity too small");
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaMemcpyAsync(out, static_cast<uint8_t *>(r->dptr) + layer * layer_bytes, layer_bytes, cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return GDPT_OK;
}

int gdpt_shader_add_existing_buffer(gdpt_shader *s, gdpt_rid rid, gdpt_uniform_type uniform_type, int binding, int set)
{
    if (!s) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    Resource *r = find(d, rid);
    if (!r) return fail(d, GDPT_ERR_INVALID_ARG, "add_existing_buffer: RID does not belong to this device");
    if ((uniform_type == GDPT_UNIFORM_TYPE_IMAGE) != (r->kind == RES_IMAGE)) return fail(d, GDPT_ERR_INVALID_ARG, "add_existing_buffer: uniform type does not match the resource");
    s->bindings[{ set, binding }] = rid;
    s->uniforms_ready = false;
    return GDPT_OK;
}

int gdpt_shader_finish_create_uniforms(gdpt_shader *s)
{
    if (!s || !s->initialized) return GDPT_ERR_INVALID_ARG;
    cudaSetDevice(s->dev->ordinal);
    const int rc = (s->kind == SHADER_MAIN) ? finish_main(s) : (s->kind == SHADER_PROGRESSIVE ? finish_progressive(s) : finish_temporal(s));
    s->uniforms_ready = (rc == GDPT_OK);
    return rc;
}

int gdpt_shader_check_ready(const gdpt_shader *s) { return (s && s->initialized && s->uniforms_ready) ? 1 : 0; }

int gdpt_shader_set_shard(gdpt_shader *s, int part, int n_parts, int band_rows)
{
    if (!s || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    if (n_parts < 1 || part < 0 || part >= n_parts || band_rows < 4 || (band_rows & 3))
        return fail(s->dev, GDPT_ERR_INVALID_ARG, "set_shard: need 0 <= part < n_parts and band_rows a positive multiple of 4");
    s->shard_part = part; s->shard_parts = n_parts; s->shard_band = band_rows;
    if (s->uniforms_ready) compute_shard(s);
    return GDPT_OK;
}

int gdpt_shader_compute(gdpt_shader *s, int gx, int gy, int gz)
{
    if (!s) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    if (!gdpt_shader_check_ready(s)) return fail(d, GDPT_ERR_NOT_READY, "compute() before finish_create_uniforms()"); // gdcs.cpp:239-240
    cudaSetDevice(d->ordinal);
    int w, h;
    if (s->kind == SHADER_MAIN) { w = s->args.width; h = s->args.height; }
    else { Resource *screen = bound(s, 0, 1); w = screen->width; h = screen->height; }
    if (gx != (w + 31) / 32 || gy != (h + 31) / 32 || gz != 1)
        return fail(d, GDPT_ERR_INVALID_ARG, "compute(%d,%d,%d): expected ceil(W/32) x ceil(H/32) x 1 = (%d,%d,1)", gx, gy, gz, (w + 31) / 32, (h + 31) / 32);
    int rc;
    for (cudaStream_t x : d->extra_streams) if (x) GDPT_CUDA(d, cudaStreamSynchronize(x)); // overlapped pipelined frames still in flight
    if (s->kind == SHADER_MAIN) {
        GDPT_CUDA(d, cudaEventRecord(d->ev[0], d->stream));
        if ((rc = enqueue_k1(s))) return rc;
        GDPT_CUDA(d, cudaEventRecord(d->ev[1], d->stream));
        if ((rc = collect_stats(s))) return rc;
        GDPT_CUDA(d, cudaEventElapsedTime(&s->stats.k1_ms, d->ev[0], d->ev[1]));
    } else {
        GDPT_CUDA(d, cudaEventRecord(d->ev[2], d->stream));
        if ((rc = (s->kind == SHADER_PROGRESSIVE ? enqueue_k2(s, 0, 1, 4) : enqueue_k3(s)))) return rc;
        GDPT_CUDA(d, cudaEventRecord(d->ev[3], d->stream));
        GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
        GDPT_CUDA(d, cudaEventElapsedTime(&s->stats.k2_ms, d->ev[2], d->ev[3]));
    }
    return GDPT_OK;
}

static int enqueue_frame(gdpt_shader *m, gdpt_shader *p, const gdpt_camera *camera, gdpt_denoising mode, uint32_t frame_count)
{
    if (!m || m->kind != SHADER_MAIN || !camera) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = m->dev;
    if (!gdpt_shader_check_ready(m)) return fail(d, GDPT_ERR_NOT_READY, "main shader is not ready");
    int rc;
    if ((rc = check_post(m, p, mode))) return rc;
    cudaSetDevice(d->ordinal);
    // the previous frame's staging must have been consumed before we overwrite it
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    for (cudaStream_t x : d->extra_streams) if (x) GDPT_CUDA(d, cudaStreamSynchronize(x));
    uint8_t *stage = static_cast<uint8_t *>(d->pinned_staging);
    memcpy(stage, camera, sizeof(gdpt_camera));
    Resource *cam_r = bound(m, 0, 3);
    memcpy(cam_r->shadow.data(), camera, cam_r->size < sizeof(gdpt_camera) ? cam_r->size : sizeof(gdpt_camera));
    GDPT_CUDA(d, cudaMemcpyAsync(cam_r->dptr, stage, cam_r->size < sizeof(gdpt_camera) ? cam_r->size : sizeof(gdpt_camera), cudaMemcpyHostToDevice, d->stream));
    GDPT_CUDA(d, cudaEventRecord(d->ev[0], d->stream));
    if ((rc = enqueue_k1(m))) return rc;
    GDPT_CUDA(d, cudaEventRecord(d->ev[1], d->stream));
    if (mode != GDPT_DENOISE_NONE) {
        if ((rc = upload_post_params(m, p, mode, frame_count, stage + 256))) return rc;
        GDPT_CUDA(d, cudaEventRecord(d->ev[2], d->stream));
        if ((rc = (mode == GDPT_DENOISE_PROGRESSIVE_RENDERING ? enqueue_k2(p, m->shard_part, m->shard_parts, m->shard_band) : enqueue_k3(p)))) return rc;
        GDPT_CUDA(d, cudaEventRecord(d->ev[3], d->stream));
    }
    return GDPT_OK;
}

int gdpt_render_frame_async(gdpt_shader *m, gdpt_shader *p, const gdpt_camera *camera, gdpt_denoising mode, uint32_t frame_count)
{
    return enqueue_frame(m, p, camera, mode, frame_count);
}

int gdpt_render_frame(gdpt_shader *m, gdpt_shader *p, const gdpt_camera *camera, gdpt_denoising mode, uint32_t frame_count,
                      void *out_rgba8, float *out_depth)
{
    int rc = enqueue_frame(m, p, camera, mode, frame_count);
    if (rc) return rc;
    gdpt_device *d = m->dev;
    const size_t n = (size_t)m->args.width * m->args.height;
    if (out_rgba8) GDPT_CUDA(d, cudaMemcpyAsync(out_rgba8, m->args.out_rgba8, n * 4, cudaMemcpyDeviceToHost, d->stream));
    if (out_depth) GDPT_CUDA(d, cudaMemcpyAsync(out_depth, m->args.out_depth, n * 4, cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return GDPT_OK;
}

int gdpt_rid_device_pointer(gdpt_device *d, gdpt_rid rid, uint64_t *out_ptr, uint64_t *out_size)
{
    if (!d) return GDPT_ERR_INVALID_ARG;
    Resource *r = find(d, rid);
    if (!r) return fail(d, GDPT_ERR_INVALID_ARG, "unknown RID");
    if (out_ptr) *out_ptr = (uint64_t)(uintptr_t)r->dptr;
    if (out_size) *out_size = r->size;
    return GDPT_OK;
}

int gdpt_progressive_accumulate(gdpt_device *d, uint64_t raw_rgba8, uint64_t screen_rgba8, uint64_t accum_rgba32f, int width,
                                int height, uint32_t frame_count)
{
    if (!d) return GDPT_ERR_INVALID_ARG;
    if (!raw_rgba8 || !screen_rgba8 || !accum_rgba32f || width <= 0 || height <= 0 || frame_count == 0u)
        return fail(d, GDPT_ERR_INVALID_ARG, "gdpt_progressive_accumulate: null image, empty size or frame_count 0");
    if (accum_rgba32f & 15u) return fail(d, GDPT_ERR_INVALID_ARG, "the accumulation image must be 16-byte aligned");
    cudaSetDevice(d->ordinal);
    init_launch_shapes(d->ordinal);
    const PeerScreens none = {};
    launch_progressive(reinterpret_cast<const uint32_t *>(raw_rgba8), reinterpret_cast<uint32_t *>(screen_rgba8),
                       reinterpret_cast<float4 *>(accum_rgba32f), nullptr, frame_count, width, height, 0, 1, 1, none, d->stream);
    GDPT_CUDA(d, cudaGetLastError());
    return GDPT_OK;
}

gdpt_rid gdpt_device_create_buffer(gdpt_device *d, uint64_t size)
{
    if (!d || size == 0) { if (d) fail(d, GDPT_ERR_INVALID_ARG, "gdpt_device_create_buffer: empty size"); return 0; }
    cudaSetDevice(d->ordinal);
    Resource r;
    r.kind = RES_BUFFER; r.size = size; // device memory only: no host copy is kept
    if (cudaMalloc(&r.dptr, size) != cudaSuccess) { fail(d, GDPT_ERR_CUDA, "cudaMalloc(%llu) failed: %s", (unsigned long long)size, cudaGetErrorString(cudaGetLastError())); return 0; }
    if (cudaMemsetAsync(r.dptr, 0, size, d->stream) != cudaSuccess || cudaStreamSynchronize(d->stream) != cudaSuccess) {
        fail(d, GDPT_ERR_CUDA, "clearing the buffer failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(r.dptr);
        return 0;
    }
    const gdpt_rid rid = d->next_rid++;
    d->resources[rid] = std::move(r);
    return rid;
}

int gdpt_device_free_buffer(gdpt_device *d, gdpt_rid rid)
{
    if (!d) return GDPT_ERR_INVALID_ARG;
    Resource *r = find(d, rid);
    if (!r || r->kind != RES_BUFFER) return fail(d, GDPT_ERR_INVALID_ARG, "gdpt_device_free_buffer: unknown buffer RID");
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaDeviceSynchronize());
    cudaFree(r->dptr);
    d->resources.erase(rid);
    return GDPT_OK;
}

int gdpt_rid_ipc_export(gdpt_device *d, gdpt_rid rid, void *out_handle)
{
    if (!d || !out_handle) return GDPT_ERR_INVALID_ARG;
    Resource *r = find(d, rid);
    if (!r) return fail(d, GDPT_ERR_INVALID_ARG, "unknown RID");
    static_assert(sizeof(cudaIpcMemHandle_t) == GDPT_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    GDPT_CUDA(d, cudaSetDevice(d->ordinal));
    GDPT_CUDA(d, cudaIpcGetMemHandle(&h, r->dptr));
    memcpy(out_handle, &h, sizeof(h));
    return GDPT_OK;
}

int gdpt_device_ipc_open(gdpt_device *d, const void *handle, uint64_t *out_ptr)
{
    if (!d || !handle || !out_ptr) return GDPT_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    GDPT_CUDA(d, cudaSetDevice(d->ordinal));
    GDPT_CUDA(d, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out_ptr = (uint64_t)(uintptr_t)p;
    return GDPT_OK;
}

int gdpt_device_ipc_close(gdpt_device *d, uint64_t ptr)
{
    if (!d || !ptr) return GDPT_ERR_INVALID_ARG;
    GDPT_CUDA(d, cudaSetDevice(d->ordinal));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    GDPT_CUDA(d, cudaIpcCloseMemHandle(reinterpret_cast<void *>((uintptr_t)ptr)));
    return GDPT_OK;
}

int gdpt_shader_set_peer_screens(gdpt_shader *p, const uint64_t *ptrs, int n)
{
    if (!p || p->kind != SHADER_PROGRESSIVE || n < 0 || (n > 0 && !ptrs)) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = p->dev;
    if (n > kMaxPeerScreens) return fail(d, GDPT_ERR_UNSUPPORTED, "at most %d peer images", (int)kMaxPeerScreens);
    PeerScreens ps = {};
    for (int i = 0; i < n; i++) {
        if (!ptrs[i] || (ptrs[i] & 15u)) return fail(d, GDPT_ERR_INVALID_ARG, "peer image %d: null or not 16-byte aligned", i);
        ps.p[i] = reinterpret_cast<uint32_t *>((uintptr_t)ptrs[i]);
    }
    ps.n = n;
    p->peers = ps;
    return GDPT_OK;
}

int gdpt_shader_get_stats(gdpt_shader *s, gdpt_frame_stats *out)
{
    if (!s || !out || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    if (!gdpt_shader_check_ready(s)) return fail(d, GDPT_ERR_NOT_READY, "shader not ready");
    cudaSetDevice(d->ordinal);
    if (!s->stats_valid) {
        int rc = collect_stats(s);
        if (rc) return rc;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]) == cudaSuccess) s->stats.k1_ms = ms;
        if (cudaEventElapsedTime(&ms, d->ev[2], d->ev[3]) == cudaSuccess) s->stats.k2_ms = ms;
        cudaGetLastError(); // events never recorded -> not an error for the caller
    }
    *out = s->stats;
    return GDPT_OK;
}

int gdpt_shader_get_schedule(const gdpt_shader *s) { return (s && s->kind == SHADER_MAIN && s->uniforms_ready) ? s->args.schedule : -1; }

int gdpt_shader_set_stage_timing(gdpt_shader *s, int on)
{
    if (!s || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    cudaSetDevice(d->ordinal);
    if (on && s->stage_ev.empty()) {
        s->stage_ev.resize(2 * kMaxDepth + 1);
        for (auto &e : s->stage_ev) GDPT_CUDA(d, cudaEventCreate(&e));
    }
    s->stage_timing = on != 0;
    return GDPT_OK;
}

int gdpt_shader_set_warp_profile(gdpt_shader *s, int on)
{
    if (!s || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    s->warp_profile = on != 0;
    if (!on) { s->args.warp_prof = nullptr; return GDPT_OK; } // the allocation stays with the shader
    if (s->uniforms_ready) { cudaSetDevice(s->dev->ordinal); return alloc_warp_profile(s); }
    return GDPT_OK;
}

int64_t gdpt_shader_read_warp_profile(gdpt_shader *s, uint64_t *out, uint64_t capacity_words)
{
    if (!s || !out || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    if (!gdpt_shader_check_ready(s) || !s->args.warp_prof) return fail(d, GDPT_ERR_NOT_READY, "warp profile is not enabled for this shader");
    const size_t words = s->warp_prof_warps * 8;
    if (capacity_words < words) return fail(d, GDPT_ERR_INVALID_ARG, "warp profile needs %llu words", (unsigned long long)words);
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaMemcpyAsync(out, s->args.warp_prof, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return (int64_t)s->warp_prof_warps;
}

int gdpt_shader_get_stage_times(gdpt_shader *s, float *out_ms, int capacity)
{
    if (!s || !out_ms || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    const int n = s->stage_count < capacity ? s->stage_count : capacity;
    for (int i = 0; i < n; i++) GDPT_CUDA(d, cudaEventElapsedTime(&out_ms[i], s->stage_ev[i], s->stage_ev[i + 1]));
    return n;
}

int gdpt_shader_read_trace(gdpt_shader *s, int segment, gdpt_trace_record *out, uint64_t capacity)
{
    if (!s || !out || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    if (!gdpt_shader_check_ready(s) || !s->args.trace) return fail(d, GDPT_ERR_NOT_READY, "shader was not created with \"#define GDPT_TRACE\"");
    if (segment < 0 || segment >= s->args.trace_segments) return fail(d, GDPT_ERR_INVALID_ARG, "segment %d not recorded (GDPT_TRACE %d)", segment, s->args.trace_segments);
    const size_t n = (size_t)s->args.width * s->args.height;
    if (capacity < n) return fail(d, GDPT_ERR_INVALID_ARG, "trace capacity too small");
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaMemcpyAsync(out, s->args.trace + (size_t)segment * n, n * sizeof(gdpt_trace_record), cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return GDPT_OK;
}

int gdpt_shader_read_visits(gdpt_shader *s, uint32_t *out, uint32_t max_per_ray, uint64_t capacity)
{
    if (!s || !out || s->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = s->dev;
    if (!gdpt_shader_check_ready(s) || !s->args.visits) return fail(d, GDPT_ERR_NOT_READY, "shader was not created with \"#define GDPT_TRACE_VISITS n\"");
    if (max_per_ray != s->args.visits_per_ray) return fail(d, GDPT_ERR_INVALID_ARG, "max_per_ray must equal GDPT_TRACE_VISITS (%u)", s->args.visits_per_ray);
    const size_t n = (size_t)s->args.width * s->args.height * max_per_ray;
    if (capacity < n) return fail(d, GDPT_ERR_INVALID_ARG, "visits capacity too small");
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaMemcpyAsync(out, s->args.visits, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
    GDPT_CUDA(d, cudaStreamSynchronize(d->stream));
    return GDPT_OK;
}

} // extern "C"

// ---- pipelined frames -------------------------------------------------------------------------
namespace {

// 512 B of pinned memory that stays untouched until the copies enqueued from it have run
int ring_take(gdpt_device *d, uint8_t **out, unsigned *slot_out)
{
    const unsigned i = d->ring_next++ & 7u;
    if (d->ring_used[i]) GDPT_CUDA(d, cudaEventSynchronize(d->ring_ev[i]));
    *out = d->ring + (size_t)i * 512u;
    *slot_out = i;
    return GDPT_OK;
}

int ensure_slot(gdpt_shader *m, gdpt_shader::FrameSlot &sl, bool want_depth)
{
    gdpt_device *d = m->dev;
    const size_t n = (size_t)m->args.width * m->args.height;
    int rc;
    if (!sl.dcnt) {
        if ((rc = dev_alloc(m, &sl.dcnt, 1))) return rc;
        uint32_t *stage = nullptr;
        if ((rc = dev_alloc(m, &stage, n))) return rc;
        sl.stage_rgba8 = stage;
    }
    if (!sl.hcnt) {
        GDPT_CUDA(d, cudaHostAlloc(reinterpret_cast<void **>(&sl.hcnt), sizeof(FrameCounters), cudaHostAllocDefault));
        for (cudaEvent_t *e : { &sl.t0, &sl.t1, &sl.t2, &sl.t3 }) GDPT_CUDA(d, cudaEventCreate(e));
        GDPT_CUDA(d, cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming));
        GDPT_CUDA(d, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    if (want_depth && !sl.stage_depth && (rc = dev_alloc(m, &sl.stage_depth, n))) return rc;
    return GDPT_OK;
}

} // namespace

extern "C" int gdpt_render_frame_begin(gdpt_shader *m, gdpt_shader *p, const gdpt_camera *camera, gdpt_denoising mode,
                                       uint32_t frame_count, void *out_rgba8, float *out_depth)
{
    if (!m || m->kind != SHADER_MAIN || !camera || !out_rgba8) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = m->dev;
    if (!gdpt_shader_check_ready(m)) return fail(d, GDPT_ERR_NOT_READY, "main shader is not ready");
    int rc;
    if ((rc = check_post(m, p, mode))) return rc;
    gdpt_shader::FrameSlot &sl = m->slots[m->slot_head % GDPT_MAX_FRAMES_IN_FLIGHT];
    if (sl.pending) return fail(d, GDPT_ERR_NOT_READY, "%d frames are already in flight: call gdpt_render_frame_wait first", GDPT_MAX_FRAMES_IN_FLIGHT);
    cudaSetDevice(d->ordinal);
    if ((rc = ensure_slot(m, sl, out_depth != nullptr))) return rc;

    // Overlapped frames: the two frames in flight run on two streams with their own camera block, K1 images, survivor
    // lists and counters, so frame N+1's kernels fill the SMs that frame N's last paths leave idle (a whole path is
    // more than half a 1080p frame long).  What orders them: the accumulate/tone-map step (one
    // accumulation image, one presented image) waits for the other frame's.  `cost` (the scheduling hint of
    // k_primary_cull) is shared and read while the other frame writes it: any value is a valid hint.
    if (d->frame_overlap && mode != GDPT_DENOISE_TEMPORAL_REPROJECTION && m->args.trace == nullptr && !m->debug_steps &&
        !m->warp_profile && m->args.schedule >= 3) {
        const unsigned idx = m->slot_head % GDPT_MAX_FRAMES_IN_FLIGHT; // every frame in flight has its own stream
        gdpt_shader::FrameSlot &other = m->slots[(m->slot_head + GDPT_MAX_FRAMES_IN_FLIGHT - 1u) % GDPT_MAX_FRAMES_IN_FLIGHT]; // previous frame
        const size_t n = (size_t)m->args.width * m->args.height;
        if (!sl.cam_dev) {
            if ((rc = dev_alloc(m, &sl.cam_dev, 2))) return rc;
            if ((rc = dev_alloc(m, &sl.pp_dev, 16))) return rc;
            if ((rc = dev_alloc(m, &sl.raw_rgba8, n))) return rc;
            if ((rc = dev_alloc(m, &sl.raw_depth, n))) return rc;
            if ((rc = dev_alloc(m, &sl.hit_list, (size_t)m->args.queue_cap + (size_t)(kCostClasses - 1) * m->args.heavy_cap))) return rc;
        }
        if (!d->extra_streams_made) {
            for (cudaStream_t &x : d->extra_streams) GDPT_CUDA(d, cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
            d->extra_streams_made = true;
        }
        cudaStream_t const main_stream = d->stream;
        cudaStream_t const S = idx == 0 ? d->stream : d->extra_streams[idx - 1u];
        uint8_t *stage = nullptr; unsigned ring_slot = 0;
        if ((rc = ring_take(d, &stage, &ring_slot))) return rc;
        memcpy(stage, camera, sizeof(gdpt_camera));
        Resource *cam_r = bound(m, 0, 3);
        memcpy(cam_r->shadow.data(), camera, cam_r->size < sizeof(gdpt_camera) ? cam_r->size : sizeof(gdpt_camera));
        SmallCopies up = {};
        up.c[0].dst = reinterpret_cast<uint32_t *>(sl.cam_dev); up.c[0].src = reinterpret_cast<const uint32_t *>(stage);
        up.c[0].words = (uint32_t)(sizeof(gdpt_camera) / 4u);
        sl.with_k2 = (mode == GDPT_DENOISE_PROGRESSIVE_RENDERING);
        if (sl.with_k2) {
            gdpt_progressive_params host_pp = { m->args.width, m->args.height, frame_count };
            memcpy(stage + 256, &host_pp, sizeof(host_pp));
            memcpy(bound(p, 0, 0)->shadow.data(), &host_pp, sizeof(host_pp));
            up.c[1].dst = reinterpret_cast<uint32_t *>(sl.pp_dev); up.c[1].src = reinterpret_cast<const uint32_t *>(stage + 256);
            up.c[1].words = (uint32_t)(sizeof(host_pp) / 4u);
        }
        launch_small_copies(up, S);
        GDPT_CUDA(d, cudaEventRecord(d->ring_ev[ring_slot], S));
        d->ring_used[ring_slot] = true;

        const FrameArgs saved = m->args;
        m->args.counters = sl.dcnt; m->args.camera = sl.cam_dev; m->args.out_rgba8 = sl.raw_rgba8; m->args.out_depth = sl.raw_depth;
        m->args.hit_list = sl.hit_list;
        // Frames in flight share the SMs: with half a grid each (two blocks per SM) two frames' paths run side by side
        // instead of one frame's blocks waiting for the other's to retire (1080p demo frame, four in flight: 0.589 -> 0.545 ms
        // per frame; three: 0.589 -> 0.561).  Throughput-bound scenes keep their full grid.
        if (m->args.blocks_per_sm == 0 && !m->args.pool_dense) m->args.blocks_per_sm = 2;
        d->stream = S; // enqueue_k1 / enqueue_k2 launch on the device's current stream
        cudaEventRecord(sl.t0, S);
        rc = enqueue_k1(m);
        cudaEventRecord(sl.t1, S);
        m->args = saved;
        sl.launches = m->stats.kernel_launches;
        if (rc == GDPT_OK && other.finished_once) rc = cudaStreamWaitEvent(S, other.k_done, 0) == cudaSuccess ? GDPT_OK : GDPT_ERR_CUDA;
        const void *readback_rgba8 = sl.raw_rgba8;
        if (rc == GDPT_OK) {
            if (sl.with_k2) {
                cudaEventRecord(sl.t2, S);
                const PeerScreens keep = p->peers;
                if (m->shard_parts <= 1 && p->peers.n < kMaxPeerScreens) {
                    p->peers.p[p->peers.n++] = static_cast<uint32_t *>(sl.stage_rgba8);
                    readback_rgba8 = sl.stage_rgba8;
                }
                rc = enqueue_k2(p, m->shard_part, m->shard_parts, m->shard_band, sl.raw_rgba8, sl.pp_dev);
                p->peers = keep;
                cudaEventRecord(sl.t3, S);
                if (rc == GDPT_OK && readback_rgba8 == sl.raw_rgba8) { // sharded: the presented image is the bound one
                    cudaMemcpyAsync(sl.stage_rgba8, saved.out_rgba8, n * 4, cudaMemcpyDeviceToDevice, S);
                    readback_rgba8 = sl.stage_rgba8;
                }
            } else { // GDPT_DENOISE_NONE: K1's image is the frame
                cudaMemcpyAsync(saved.out_rgba8, sl.raw_rgba8, n * 4, cudaMemcpyDeviceToDevice, S);
            }
            cudaMemcpyAsync(saved.out_depth, sl.raw_depth, n * 4, cudaMemcpyDeviceToDevice, S);
            SmallCopies down = {};
            down.c[0].dst = reinterpret_cast<uint32_t *>(sl.hcnt); down.c[0].src = reinterpret_cast<const uint32_t *>(sl.dcnt);
            down.c[0].words = (uint32_t)(sizeof(FrameCounters) / 4u);
            launch_small_copies(down, S);
            cudaEventRecord(sl.k_done, S);
            sl.finished_once = true;
        }
        d->stream = main_stream;
        if (rc) return rc;
        GDPT_CUDA(d, cudaGetLastError());
        GDPT_CUDA(d, cudaStreamWaitEvent(d->copy_stream, sl.k_done, 0));
        GDPT_CUDA(d, cudaMemcpyAsync(out_rgba8, readback_rgba8, n * 4, cudaMemcpyDeviceToHost, d->copy_stream));
        if (out_depth) GDPT_CUDA(d, cudaMemcpyAsync(out_depth, sl.raw_depth, n * 4, cudaMemcpyDeviceToHost, d->copy_stream));
        GDPT_CUDA(d, cudaEventRecord(sl.done, d->copy_stream));
        sl.pending = true;
        m->slot_head++;
        m->stats_valid = false;
        return GDPT_OK;
    }

    {   // single-stream form; an overlapped frame of the other slot may still be running on the second stream
        gdpt_shader::FrameSlot &other = m->slots[(m->slot_head + GDPT_MAX_FRAMES_IN_FLIGHT - 1u) % GDPT_MAX_FRAMES_IN_FLIGHT];
        if (d->extra_streams_made && other.finished_once) GDPT_CUDA(d, cudaStreamWaitEvent(d->stream, other.k_done, 0));
    }
    // per-frame H2D blocks come from the pinned ring: nothing here waits for the GPU
    uint8_t *stage = nullptr; unsigned ring_slot = 0;
    if ((rc = ring_take(d, &stage, &ring_slot))) return rc;
    memcpy(stage, camera, sizeof(gdpt_camera));
    Resource *cam_r = bound(m, 0, 3);
    const size_t cam_bytes = cam_r->size < sizeof(gdpt_camera) ? cam_r->size : sizeof(gdpt_camera);
    memcpy(cam_r->shadow.data(), camera, cam_bytes);
    SmallCopies up = {};
    up.c[0].dst = static_cast<uint32_t *>(cam_r->dptr); up.c[0].src = reinterpret_cast<const uint32_t *>(stage);
    up.c[0].words = (uint32_t)(cam_bytes / 4u);
    sl.with_k2 = (mode != GDPT_DENOISE_NONE);
    if (sl.with_k2) { // post-process parameter block (upload_post_params), staged behind the camera block
        Resource *pp = bound(p, 0, 0);
        size_t n = 0;
        if (mode == GDPT_DENOISE_PROGRESSIVE_RENDERING) {
            gdpt_progressive_params host_pp = { m->args.width, m->args.height, frame_count };
            n = sizeof(host_pp);
            memcpy(stage + 256, &host_pp, n);
        } else if (!p->staged_params.empty()) {
            n = p->staged_params.size();
            memcpy(stage + 256, p->staged_params.data(), n);
            p->staged_params.clear();
        }
        if (n) {
            memcpy(pp->shadow.data(), stage + 256, n);
            up.c[1].dst = static_cast<uint32_t *>(pp->dptr); up.c[1].src = reinterpret_cast<const uint32_t *>(stage + 256);
            up.c[1].words = (uint32_t)(n / 4u);
        }
    }
    launch_small_copies(up, d->stream); // reads the page-locked ring directly: no copy-engine operation in the frame
    GDPT_CUDA(d, cudaEventRecord(d->ring_ev[ring_slot], d->stream));
    d->ring_used[ring_slot] = true;

    FrameCounters *const classic = m->args.counters;
    m->args.counters = sl.dcnt; // this frame counts into its own block
    GDPT_CUDA(d, cudaEventRecord(sl.t0, d->stream));
    rc = enqueue_k1(m);
    m->args.counters = classic;
    if (rc) return rc;
    sl.launches = m->stats.kernel_launches;
    GDPT_CUDA(d, cudaEventRecord(sl.t1, d->stream));
    // the read-back streams from a staging copy so the next K1 may overwrite the image.  An unsharded progressive
    // frame gets that copy for free: K2 stores every quad into the staging image as well (the peer-screen store)
    bool staged_by_k2 = false;
    if (sl.with_k2) {
        GDPT_CUDA(d, cudaEventRecord(sl.t2, d->stream));
        if (mode == GDPT_DENOISE_PROGRESSIVE_RENDERING) {
            const PeerScreens keep = p->peers;
            if (m->shard_parts <= 1 && p->peers.n < kMaxPeerScreens && (((size_t)m->args.width * m->args.height) & 3u) == 0u) {
                p->peers.p[p->peers.n++] = static_cast<uint32_t *>(sl.stage_rgba8);
                staged_by_k2 = true;
            }
            rc = enqueue_k2(p, m->shard_part, m->shard_parts, m->shard_band);
            p->peers = keep;
        } else {
            rc = enqueue_k3(p);
        }
        if (rc) return rc;
        GDPT_CUDA(d, cudaEventRecord(sl.t3, d->stream));
    }
    const size_t n = (size_t)m->args.width * m->args.height;
    if (!staged_by_k2) GDPT_CUDA(d, cudaMemcpyAsync(sl.stage_rgba8, m->args.out_rgba8, n * 4, cudaMemcpyDeviceToDevice, d->stream));
    if (out_depth) GDPT_CUDA(d, cudaMemcpyAsync(sl.stage_depth, m->args.out_depth, n * 4, cudaMemcpyDeviceToDevice, d->stream));
    SmallCopies down = {};
    down.c[0].dst = reinterpret_cast<uint32_t *>(sl.hcnt); down.c[0].src = reinterpret_cast<const uint32_t *>(sl.dcnt);
    down.c[0].words = (uint32_t)(sizeof(FrameCounters) / 4u);
    launch_small_copies(down, d->stream); // counters into the page-locked block the host reads after `done`
    GDPT_CUDA(d, cudaEventRecord(sl.k_done, d->stream));
    sl.finished_once = true; // an overlapped frame that follows orders its post-process step after this one
    // the read-back leaves on the copy stream while the compute stream starts the next frame
    GDPT_CUDA(d, cudaStreamWaitEvent(d->copy_stream, sl.k_done, 0));
    GDPT_CUDA(d, cudaMemcpyAsync(out_rgba8, sl.stage_rgba8, n * 4, cudaMemcpyDeviceToHost, d->copy_stream));
    if (out_depth) GDPT_CUDA(d, cudaMemcpyAsync(out_depth, sl.stage_depth, n * 4, cudaMemcpyDeviceToHost, d->copy_stream));
    GDPT_CUDA(d, cudaEventRecord(sl.done, d->copy_stream));
    sl.pending = true;
    m->slot_head++;
    m->stats_valid = false;
    return GDPT_OK;
}

extern "C" int gdpt_render_frame_wait(gdpt_shader *m, gdpt_frame_stats *out_stats)
{
    if (!m || m->kind != SHADER_MAIN) return GDPT_ERR_INVALID_ARG;
    gdpt_device *d = m->dev;
    gdpt_shader::FrameSlot &sl = m->slots[m->slot_tail % GDPT_MAX_FRAMES_IN_FLIGHT];
    if (!sl.pending) return fail(d, GDPT_ERR_NOT_READY, "no frame in flight");
    cudaSetDevice(d->ordinal);
    GDPT_CUDA(d, cudaEventSynchronize(sl.done));
    sl.pending = false;
    m->slot_tail++;
    const FrameCounters &c = *sl.hcnt;
    if (c.overflow) return fail(d, GDPT_ERR_UNSUPPORTED, "a ray exceeded the reference's 64+64 traversal stack entries");
    if (out_stats) {
        memset(out_stats, 0, sizeof(*out_stats));
        out_stats->rays = c.rays;
        out_stats->primary_hits = c.primary_hits;
        out_stats->retraced = c.retraced;
        out_stats->own_node_steps = c.own_node_steps; out_stats->own_box_tests = c.own_box_tests; out_stats->own_tri_tests = c.own_tri_tests;
        out_stats->own_inst_entries = c.own_inst_entries; out_stats->own_proofs = c.own_proofs;
        out_stats->node_pops = c.node_pops; out_stats->box_tests = c.box_tests; out_stats->tri_tests = c.tri_tests;
        out_stats->tlas_leaves = c.tlas_leaves; out_stats->max_stack = c.max_stack;
        out_stats->kernel_launches = sl.launches;
        cudaEventElapsedTime(&out_stats->k1_ms, sl.t0, sl.t1);
        if (sl.with_k2) cudaEventElapsedTime(&out_stats->k2_ms, sl.t2, sl.t3);
    }
    return GDPT_OK;
}
