// pt_oracle.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).  NEVER SHIPPED, NEVER A FALLBACK.
//
// CPU restatement of the reference's GPU hot path, used as the parity checker by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  Nothing under gdpathtracing_b200/ may include, link or load this file.
//
// PARITY PINNING: the reference ships no tests, golden vectors or KATs for this
// path, and its GPU half (GLSL on Godot's Vulkan
// RenderingDevice) cannot be executed in this container, so the shader
// restatement below is "parity unpinned" by the reference itself.  What IS
// pinned: the RNG against the integer KATs of SURVEY A.6 (tests/golden/rng_kat.json)
// and every acceleration-structure byte against the reference's own bvh.cpp
// compiled in place (oracle/_ref, ref_bridge.cpp).
//
// Each function cites the reference lines it follows.  Paths:
//   main.glsl  = project/addons/jar_path_tracing/src/shaders/main.glsl
//   brdfs.glsl = .../shaders/brdfs.glsl
//   prog.glsl  = .../shaders/progressive_rendering.glsl
//   temp.glsl  = .../shaders/temporal_reprojection.glsl
//
// GLSL leaves built-in precision and NaN behaviour implementation-defined.  The
// choices made here ARE the parity contract the CUDA kernels implement too
// (DESIGN.md "Arithmetic contract"):
//   * every operation is a single IEEE-754 binary32 op, round-to-nearest-even,
//     evaluated left to right as written in the GLSL; no FMA contraction
//     (compile with -ffp-contract=off); '/' and sqrt are correctly rounded.
//   * dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z;  m*v = ((m0*v.x + m1*v.y) + m2*v.z) + m3*v.w
//   * normalize(v) = v / sqrt(dot(v,v));  mix(a,b,t) = a*(1-t) + b*t
//   * min/max inside intersectAABB are IEEE minNum/maxNum (what FMNMX / v_min_f32
//     do): a NaN operand is dropped.  Everywhere else min/max against a constant
//     is the select (x < c) ? .. : ..
//   * sin/cos: the Cody-Waite + minimax polynomial below (orc_sincos).
//   * imageStore to rgba8 = round-half-even(clamp(x,0,1)*255), NaN -> 0;
//     imageLoad = byte / 255.0f;  texture() = nearest, clamp-to-edge, LOD 0.
#include "gdpt_wire.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct V3 { float x, y, z; };
struct V2 { float x, y; };
static inline V3 mk(float x, float y, float z) { V3 r = { x, y, z }; return r; }
static inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 scl(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline V3 mulv(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 divs(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
static inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length3(V3 a) { return sqrtf(dot(a, a)); }
static inline V3 normalize3(V3 a) { return divs(a, length3(a)); }
static inline V3 mix3(V3 a, V3 b, float t) { return add(scl(a, 1.0f - t), scl(b, t)); }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float min_num(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((b < a) ? b : a)); }
static inline float max_num(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((a < b) ? b : a)); }
static inline float max_c(float c, float x) { return (c < x) ? x : c; } // GLSL max(c, x) as a select
static inline float min_c(float c, float x) { return (x < c) ? x : c; } // GLSL min(c, x) as a select

// column-major mat4 (src/utils.h:15-49) times (v, w)
static inline V3 xform(const float *m, V3 v, float w)
{
    return mk(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * w,
              ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * w,
              ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * w);
}

// sin and cos of x >= 0 (every argument on the path is in [0, 2*pi]).
// k = nearest multiple of pi/2, three-constant Cody-Waite reduction, then the
// classic single-precision minimax polynomials on [-pi/4, pi/4].
static inline void orc_sincos(float x, float *s_out, float *c_out)
{
    int k = (int)(x * 0.636619772f + 0.5f);
    float kf = (float)k;
    float r = ((x - kf * 1.5703125f) - kf * 4.837512969970703125e-4f) - kf * 7.54978995489188216e-8f;
    float z = r * r;
    float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    switch (k & 3) {
    case 0: *s_out = sp; *c_out = cp; break;
    case 1: *s_out = cp; *c_out = -sp; break;
    case 2: *s_out = -sp; *c_out = -cp; break;
    default: *s_out = -cp; *c_out = sp; break;
    }
}

// ---------------------------------------------------------------- RNG

// main.glsl:176-181
static inline void prng_seed(uint32_t px, uint32_t py, uint32_t frame, uint32_t *sx, uint32_t *sy)
{
    uint32_t x = px * 0x9e3779b9u + frame;
    uint32_t y = py * 0x9e3779b9u + frame;
    x ^= x >> 16; y ^= y >> 16;
    *sx = x * 0x9e3779b9u; *sy = y * 0x9e3779b9u;
}

// main.glsl:163-174
static inline V2 pcg2d(uint32_t *sx, uint32_t *sy)
{
    uint32_t x = 1664525u * *sx + 1013904223u;
    uint32_t y = 1664525u * *sy + 1013904223u;
    x += 1664525u * y; y += 1664525u * x;
    x ^= x >> 16; y ^= y >> 16;
    x += 1664525u * y; y += 1664525u * x;
    x ^= x >> 16; y ^= y >> 16;
    *sx = x; *sy = y;
    V2 r = { (float)x * 2.32830643654e-10f, (float)y * 2.32830643654e-10f };
    return r;
}

// ---------------------------------------------------------------- scene view

struct Scene {
    const gdpt_triangle_geometry *tri_geom; uint64_t n_tris;
    const gdpt_triangle_data *tri_data;
    const gdpt_material *materials; uint64_t n_materials;
    const gdpt_bvh_node *bvh; uint64_t n_nodes;
    const gdpt_blas_instance *blas; uint64_t n_blas;
    const gdpt_tlas_node *tlas; uint64_t n_tlas;
    const uint8_t *textures; int32_t tex_w, tex_h, tex_layers;
    // material-breadth extension (include/gdpt_wire.h, "#define GDPT_MATERIAL_EXT"); NOT part of the reference, hence not
    // pinned by its shader text: this restatement is the definition the CUDA path is held to.  material_ext == 0 = reference
    int32_t material_ext = 0;
    const uint32_t *surface_materials = nullptr;
    float srgb_lut[256];
};

struct Ray { V3 d, o, rD; };

struct Hit { // main.glsl:62-71 HitInfo
    V3 position; float t; uint32_t blas, triangle, steps; float bu, bv; bool front; V3 out_dir;
};

struct Shading { // main.glsl:73-82 ShadingInfo
    V3 position, normal, out_dir; float lambert_out; V3 emission, diffuse_albedo, fresnel_0; float roughness;
};

struct Counters {
    uint32_t node_pops, box_tests, tri_tests, tlas_leaves, max_stack;
    uint64_t hash;
    uint32_t *visits; uint32_t visits_cap; uint32_t visits_n;
    bool overflow;
};

static inline void visit(Counters *c, uint32_t id)
{
    c->node_pops++;
    for (int b = 0; b < 4; b++) { c->hash ^= (id >> (8 * b)) & 0xffu; c->hash *= GDPT_FNV64_PRIME; }
    if (c->visits && c->visits_n < c->visits_cap) c->visits[c->visits_n] = id;
    c->visits_n++;
}

// main.glsl:224-257
static inline bool intersect_triangle(const Scene &sc, const Ray &ray, uint32_t tri_index, Hit &h)
{
    h.steps++;
    const gdpt_triangle_geometry &tri = sc.tri_geom[tri_index];
    V3 v0 = mk(tri.v[0][0], tri.v[0][1], tri.v[0][2]);
    V3 v1 = mk(tri.v[1][0], tri.v[1][1], tri.v[1][2]);
    V3 v2 = mk(tri.v[2][0], tri.v[2][1], tri.v[2][2]);
    V3 edge1 = sub(v1, v0), edge2 = sub(v2, v0);
    V3 pvec = cross(ray.d, edge2);
    float det = dot(edge1, pvec);
    if (fabsf(det) < 1e-5f) return false;
    float inv_det = 1.0f / det;
    V3 tvec = sub(ray.o, v0);
    float u = dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    V3 qvec = cross(tvec, edge1);
    float v = dot(ray.d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = dot(edge2, qvec) * inv_det;
    if (t < 0.0f || t > h.t) return false;
    h.position = add(ray.o, scl(ray.d, t));
    h.t = t;
    h.triangle = tri_index;
    h.bu = u; h.bv = v;
    h.out_dir = neg(ray.d);
    h.front = dot(cross(edge1, edge2), ray.d) > 0.0f;
    return true;
}

// main.glsl:259-268
static inline float intersect_aabb(const Ray &ray, const float *bmin, const float *bmax)
{
    float tx1 = (bmin[0] - ray.o.x) * ray.rD.x, tx2 = (bmax[0] - ray.o.x) * ray.rD.x;
    float tmin = min_num(tx1, tx2), tmax = max_num(tx1, tx2);
    float ty1 = (bmin[1] - ray.o.y) * ray.rD.y, ty2 = (bmax[1] - ray.o.y) * ray.rD.y;
    tmin = max_num(tmin, min_num(ty1, ty2)); tmax = min_num(tmax, max_num(ty1, ty2));
    float tz1 = (bmin[2] - ray.o.z) * ray.rD.z, tz2 = (bmax[2] - ray.o.z) * ray.rD.z;
    tmin = max_num(tmin, min_num(tz1, tz2)); tmax = min_num(tmax, max_num(tz1, tz2));
    if (tmax >= tmin && tmax > 0.0f) return tmin;
    return 1e30f;
}

#define ORC_STACK 64 // main.glsl:272,307 (no overflow check upstream; we flag it)

// main.glsl:270-303
static inline bool trace_blas(const Scene &sc, uint32_t root, const Ray &ray, Hit &h, Counters *c)
{
    uint32_t stack[ORC_STACK];
    uint32_t sp = 0;
    stack[sp++] = root;
    if (c->max_stack < sp) c->max_stack = sp;
    while (sp > 0) {
        uint32_t id = stack[--sp];
        const gdpt_bvh_node &node = sc.bvh[id];
        visit(c, id);
        if (node.tri_count > 0) {
            for (uint32_t i = 0; i < node.tri_count; i++) intersect_triangle(sc, ray, node.first_tri_index + i, h);
            continue;
        }
        const gdpt_bvh_node &L = sc.bvh[node.left_child];
        const gdpt_bvh_node &R = sc.bvh[node.right_child];
        float d1 = intersect_aabb(ray, L.aabb_min, L.aabb_max);
        float d2 = intersect_aabb(ray, R.aabb_min, R.aabb_max);
        c->box_tests += 2;
        bool lv = d1 < h.t, rv = d2 < h.t;
        if (sp + 2 > ORC_STACK) { c->overflow = true; return h.t < 1e9f; }
        if (d1 < d2) {
            if (rv) stack[sp++] = node.right_child;
            if (lv) stack[sp++] = node.left_child;
        } else {
            if (lv) stack[sp++] = node.left_child;
            if (rv) stack[sp++] = node.right_child;
        }
        if (c->max_stack < sp) c->max_stack = sp;
    }
    return h.t < 1e9f;
}

// main.glsl:305-350
static inline bool trace_tlas(const Scene &sc, const Ray &ray, Hit &h, Counters *c)
{
    uint32_t stack[ORC_STACK];
    int sp = 0;
    stack[sp++] = 0;
    float min_t = 1e9f;
    if (c->max_stack < (uint32_t)sp) c->max_stack = sp;
    while (sp > 0) {
        uint32_t id = stack[--sp];
        const gdpt_tlas_node &node = sc.tlas[id];
        visit(c, id | GDPT_VISIT_TLAS_TAG);
        if (node.left_right == 0) {
            const gdpt_blas_instance &b = sc.blas[node.blas];
            Ray br;
            br.o = xform(b.inverse_transform, ray.o, 1.0f);
            br.d = xform(b.inverse_transform, ray.d, 0.0f);
            br.rD = mk(1.0f / br.d.x, 1.0f / br.d.y, 1.0f / br.d.z);
            c->tlas_leaves++;
            trace_blas(sc, b.root, br, h, c);
            if (h.t < min_t) { h.blas = node.blas; min_t = h.t; }
            continue;
        }
        uint32_t left = node.left_right & 0xFFFFu, right = node.left_right >> 16;
        const gdpt_tlas_node &L = sc.tlas[left];
        const gdpt_tlas_node &R = sc.tlas[right];
        float d1 = intersect_aabb(ray, L.aabb_min, L.aabb_max);
        float d2 = intersect_aabb(ray, R.aabb_min, R.aabb_max);
        c->box_tests += 2;
        bool lv = d1 < h.t, rv = d2 < h.t;
        if (sp + 2 > ORC_STACK) { c->overflow = true; return h.t < 1e9f; }
        if (d1 < d2) {
            if (rv) stack[sp++] = right;
            if (lv) stack[sp++] = left;
        } else {
            if (lv) stack[sp++] = left;
            if (rv) stack[sp++] = right;
        }
        if (c->max_stack < (uint32_t)sp) c->max_stack = sp;
    }
    return h.t < 1e9f;
}

// texture(textureArray, vec3(uv, layer)).rgb with the default RDSamplerState
// (gdcs.cpp:183-187): nearest, clamp-to-edge, single mip.
static inline V3 sample_texture(const Scene &sc, float u, float v, int layer)
{
    int ix = (int)floorf(u * (float)sc.tex_w), iy = (int)floorf(v * (float)sc.tex_h);
    ix = ix < 0 ? 0 : (ix > sc.tex_w - 1 ? sc.tex_w - 1 : ix);
    iy = iy < 0 ? 0 : (iy > sc.tex_h - 1 ? sc.tex_h - 1 : iy);
    if (layer > sc.tex_layers - 1) layer = sc.tex_layers - 1; // array layer index is clamped by the sampler
    const uint8_t *p = sc.textures + (((size_t)layer * sc.tex_h + iy) * sc.tex_w + ix) * 4;
    return mk((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f);
}
static inline const uint8_t *texel_bytes(const Scene &sc, float u, float v, int layer)
{
    int ix = (int)floorf(u * (float)sc.tex_w), iy = (int)floorf(v * (float)sc.tex_h);
    ix = ix < 0 ? 0 : (ix > sc.tex_w - 1 ? sc.tex_w - 1 : ix);
    iy = iy < 0 ? 0 : (iy > sc.tex_h - 1 ? sc.tex_h - 1 : iy);
    if (layer > sc.tex_layers - 1) layer = sc.tex_layers - 1;
    return sc.textures + (((size_t)layer * sc.tex_h + iy) * sc.tex_w + ix) * 4;
}
static void fill_srgb_lut(Scene &sc)
{
    for (int k = 0; k < 256; k++) { // IEC 61966-2-1 decode of an 8-bit code, rounded once to binary32
        const double c = k / 255.0;
        sc.srgb_lut[k] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
}

// main.glsl:194-222
static inline Shading get_shading_data(const Scene &sc, const Hit &h)
{
    Shading s;
    const gdpt_triangle_data &tri = sc.tri_data[h.triangle];
    const gdpt_blas_instance &b = sc.blas[h.blas];
    const uint32_t material_id = (sc.material_ext && sc.surface_materials)
                                     ? sc.surface_materials[sc.surface_materials[h.blas] + tri.material_index]
                                     : b.materials[tri.material_index];
    const gdpt_material &m = sc.materials[material_id];
    s.position = xform(b.transform, h.position, 1.0f);
    s.out_dir = normalize3(xform(b.transform, h.out_dir, 0.0f));
    float u = h.bu, v = h.bv;
    float w = 1.0f - u - v;
    float tu = (tri.uvs[0][0] * w + tri.uvs[1][0] * u) + tri.uvs[2][0] * v;
    float tv = (tri.uvs[0][1] * w + tri.uvs[1][1] * u) + tri.uvs[2][1] * v;
    V3 n = add(add(scl(mk(tri.n0[0], tri.n0[1], tri.n0[2]), w), scl(mk(tri.n1[0], tri.n1[1], tri.n1[2]), u)),
               scl(mk(tri.n2[0], tri.n2[1], tri.n2[2]), v));
    n = normalize3(xform(b.transform, n, 0.0f));
    s.normal = h.front ? n : neg(n);
    s.lambert_out = dot(s.normal, s.out_dir);
    s.emission = scl(mk(m.emission[0], m.emission[1], m.emission[2]), max_c(0.0f, m.emission[3]));
    V3 albedo = mk(m.albedo[0], m.albedo[1], m.albedo[2]);
    float metal = m.metallic, rough = m.roughness;
    if (!sc.material_ext) {
        if (m.albedo_texture_index >= 0) albedo = mulv(albedo, sample_texture(sc, tu, tv, m.albedo_texture_index));
    } else { // the extension: sRGB albedo layers, roughness / metallic scaled by the red channel of their layers
        if (m.albedo_texture_index >= 0) {
            const uint8_t *p = texel_bytes(sc, tu, tv, m.albedo_texture_index);
            const V3 c = (m.ext_flags & GDPT_MATERIAL_ALBEDO_SRGB) ? mk(sc.srgb_lut[p[0]], sc.srgb_lut[p[1]], sc.srgb_lut[p[2]])
                                                                   : mk((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f);
            albedo = mulv(albedo, c);
        }
        if (m.ext_roughness_texture) rough = rough * ((float)texel_bytes(sc, tu, tv, (int)m.ext_roughness_texture - 1)[0] / 255.0f);
        if (m.ext_metallic_texture) metal = metal * ((float)texel_bytes(sc, tu, tv, (int)m.ext_metallic_texture - 1)[0] / 255.0f);
    }
    s.fresnel_0 = mix3(mk(0.02f, 0.02f, 0.02f), albedo, metal);
    s.diffuse_albedo = sub(albedo, scl(albedo, metal));
    s.roughness = max_c(0.006f, rough);
    return s;
}

// ---------------------------------------------------------------- brdfs.glsl

#define ORC_PI 3.14159274f          // float(M_PI), brdfs.glsl:1
#define ORC_TWO_PI 6.28318548f      // 2.0 * M_PI folded
#define ORC_TWO_OVER_PI 0.636619772f // 2.0 / M_PI folded

// brdfs.glsl:3-8 specialised: scalar f0/f90
static inline float fresnel_schlick_f(float f0, float f90, float cosine)
{
    float f = 1.0f - cosine, f2 = f * f, f5 = f2 * f2 * f;
    return mixf(f0, f90, f5);
}

// brdfs.glsl:10-38
static inline V3 brdf(const Shading &s, V3 light)
{
    float ndl = dot(s.normal, light), ndv = s.lambert_out;
    if (((ndv < ndl) ? ndv : ndl) < 0.0f) return mk(0, 0, 0);
    V3 half = normalize3(add(light, s.out_dir));
    float hdv = dot(half, s.out_dir);
    float f90 = (hdv * hdv) * (2.0f * s.roughness) + 0.5f;
    float diffuse_fresnel = fresnel_schlick_f(1.0f, f90, ndv) * fresnel_schlick_f(1.0f, f90, ndl);
    V3 r = scl(s.diffuse_albedo, diffuse_fresnel);
    float hdn = dot(half, s.normal);
    float r2 = s.roughness * s.roughness;
    float denom = hdn * (r2 - 1.0f) + 1.0f;
    float distribution = r2 / (denom * denom);
    float masking = ndl * sqrtf((ndv - r2 * ndv) * ndv + r2);
    float shadowing = ndv * sqrtf((ndl - r2 * ndl) * ndl + r2);
    float geometry = 0.5f / (masking + shadowing);
    float c = max_c(0.0f, hdv);
    V3 spec = mk(fresnel_schlick_f(s.fresnel_0.x, 1.0f, c), fresnel_schlick_f(s.fresnel_0.y, 1.0f, c),
                 fresnel_schlick_f(s.fresnel_0.z, 1.0f, c));
    r = add(r, scl(spec, distribution * geometry));
    return divs(r, ORC_PI);
}

// brdfs.glsl:40-54 with roughness.x == roughness.y
static inline V3 sample_ggx_vndf(V3 view, float rough, V2 rnd)
{
    V3 tv = normalize3(mk(view.x * rough, view.y * rough, view.z));
    float phi = ORC_TWO_PI * rnd.x;
    float z = 1.0f - rnd.y * (1.0f + tv.z);
    float sin_theta = sqrtf(max_c(0.0f, 1.0f - z * z));
    float sphi, cphi;
    orc_sincos(phi, &sphi, &cphi);
    V3 hs = mk(sin_theta * cphi, sin_theta * sphi, z);
    V3 sum = add(hs, tv);
    return normalize3(mk(sum.x * rough, sum.y * rough, sum.z));
}

// brdfs.glsl:56-67
static inline float ggx_vndf_density(float ndv, float hdn, float hdv, float rough)
{
    if (hdn < 0.0f) return 0.0f;
    float r2 = rough * rough, inv = 1.0f - r2;
    float denom = ndv + sqrtf(r2 + inv * ndv * ndv);
    float d_vis = max_c(0.0f, hdv) * ORC_TWO_OVER_PI / denom;
    float m = 1.0f - inv * hdn * hdn;
    return d_vis * r2 / (m * m);
}

// brdfs.glsl:83-93
static inline void shading_space(V3 n, V3 *c0, V3 *c1, V3 *c2)
{
    float sign = n.z > 0.0f ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    *c0 = mk(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    *c1 = mk(b, sign + n.y * n.y * a, -n.y);
    *c2 = n;
}
static inline V3 mat3_mul(V3 c0, V3 c1, V3 c2, V3 v) { return add(add(scl(c0, v.x), scl(c1, v.y)), scl(c2, v.z)); }

// brdfs.glsl:107-110
static inline float diffuse_probability(const Shading &s)
{
    float lum = dot(s.diffuse_albedo, mk(0.2126f, 0.7152f, 0.0722f));
    return min_c(0.5f, lum);
}

// brdfs.glsl:112-128 (+ :95-101 sample_hemisphere_psa, :69-72 sample_ggx_in_dir)
static inline V3 sample_brdf(const Shading &s, V2 rnd)
{
    V3 c0, c1, c2;
    shading_space(s.normal, &c0, &c1, &c2);
    float p = diffuse_probability(s);
    if (rnd.x < p) {
        rnd.x = rnd.x / p;
        float phi = ORC_TWO_PI * rnd.x, radius = sqrtf(rnd.y), z = sqrtf(1.0f - radius * radius);
        float sphi, cphi;
        orc_sincos(phi, &sphi, &cphi);
        return mat3_mul(c0, c1, c2, mk(radius * cphi, radius * sphi, z));
    }
    rnd.x = (rnd.x - p) / (1.0f - p);
    V3 local_view = mk(dot(c0, s.out_dir), dot(c1, s.out_dir), dot(c2, s.out_dir)); // transpose(M) * v
    V3 half = sample_ggx_vndf(local_view, s.roughness, rnd);
    float k = 2.0f * dot(half, local_view);                                      // reflect(I,N) = I - 2*dot(N,I)*N
    V3 local_light = neg(sub(local_view, scl(half, k)));
    return mat3_mul(c0, c1, c2, local_light);
}

// brdfs.glsl:130-138 (+ :74-81 get_ggx_in_dir_density, :103-105)
static inline float brdf_density(const Shading &s, V3 dir)
{
    float p = diffuse_probability(s);
    V3 half = normalize3(add(dir, s.out_dir));
    float hdv = dot(half, s.out_dir), hdn = dot(half, s.normal);
    float spec = ggx_vndf_density(s.lambert_out, hdn, hdv, s.roughness) / (4.0f * hdv);
    float diff = max_c(0.0f, dot(s.normal, dir)) / ORC_PI;
    return mixf(spec, diff, p);
}

// ---------------------------------------------------------------- main.glsl driver

static inline uint8_t to_unorm8(float x)
{
    if (x != x) return 0;
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return (uint8_t)nearbyintf(x * 255.0f); // default rounding mode: nearest-even
}

struct RenderJob {
    Scene sc;
    const gdpt_render_params *params;
    const gdpt_camera *cam;
    int max_depth, debug_steps;
    int y_begin, y_end, y_step;
    uint8_t *out_rgba8; float *out_depth; float *out_radiance;
    gdpt_trace_record *trace; int trace_segments;
    uint32_t *visits; uint32_t visits_per_ray;
    std::atomic<int> next_row;
    std::atomic<uint64_t> rays, primary_hits, node_pops, box_tests, tri_tests, tlas_leaves;
    std::atomic<uint32_t> max_stack, overflow;
};

static void render_rows(RenderJob *job)
{
    const Scene &sc = job->sc;
    const gdpt_camera &cam = *job->cam;
    const int W = job->params->width, H = job->params->height;
    uint64_t rays = 0, phits = 0, pops = 0, boxes = 0, tris = 0, leaves = 0;
    uint32_t max_stack = 0; bool overflow = false;
    for (;;) {
        int y = job->next_row.fetch_add(job->y_step);
        if (y >= job->y_end) break;
        for (int x = 0; x < W; x++) {
            const size_t pix = (size_t)y * W + x;
            // main.glsl:405-421
            uint32_t sx, sy;
            prng_seed((uint32_t)x, (uint32_t)y, cam.frame_index, &sx, &sy);
            V2 r0 = pcg2d(&sx, &sy);
            float theta = 6.2831853f * (r0.y * 0.25f); // box_muller, main.glsl:183-187 (R is dead code)
            float jc, js;
            orc_sincos(theta, &js, &jc);
            float scx = ((float)x + jc) / (float)W * 2.0f - 1.0f;
            float scy = ((float)y + js) / (float)H * 2.0f - 1.0f;
            const float *m = cam.ivp;
            float nx = scx, ny = -scy;
            float wx = ((m[0] * nx + m[4] * ny) + m[8] * 1.0f) + m[12] * 1.0f;
            float wy = ((m[1] * nx + m[5] * ny) + m[9] * 1.0f) + m[13] * 1.0f;
            float wz = ((m[2] * nx + m[6] * ny) + m[10] * 1.0f) + m[14] * 1.0f;
            float ww = ((m[3] * nx + m[7] * ny) + m[11] * 1.0f) + m[15] * 1.0f;
            wx = wx / ww; wy = wy / ww; wz = wz / ww;
            Ray ray;
            ray.o = mk(cam.position[0], cam.position[1], cam.position[2]);
            ray.d = normalize3(sub(mk(wx, wy, wz), ray.o));
            ray.rD = mk(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);

            // path_trace, main.glsl:372-401 (loop bound 5 upstream; parameter here)
            float depth = cam.z_far;
            V3 radiance = mk(0, 0, 0), throughput = mk(1, 1, 1);
            const int n_seg = job->debug_steps ? 1 : job->max_depth;
            for (int i = 0; i < n_seg; i++) {
                // ray_trace, main.glsl:352-370
                Hit h;
                memset(&h, 0, sizeof(h));
                h.t = 1e9f; h.steps = 0;
                Counters c;
                memset(&c, 0, sizeof(c));
                c.hash = GDPT_FNV64_OFFSET;
                if (job->visits && i == 0) { c.visits = job->visits + pix * job->visits_per_ray; c.visits_cap = job->visits_per_ray; }
                bool hit = trace_tlas(sc, ray, h, &c);
                rays++; pops += c.node_pops; boxes += c.box_tests; tris += h.steps; leaves += c.tlas_leaves;
                if (c.max_stack > max_stack) max_stack = c.max_stack;
                overflow |= c.overflow;
                if (i == 0 && hit) phits++;
                if (job->trace && i < job->trace_segments) {
                    gdpt_trace_record &tr = job->trace[(size_t)i * W * H + pix];
                    tr.hit = hit ? 1u : 0u; tr.triangle = hit ? h.triangle : 0u; tr.blas = hit ? h.blas : 0u;
                    tr.front = hit ? (h.front ? 1u : 0u) : 0u;
                    tr.t = h.t; tr.u = hit ? h.bu : 0.0f; tr.v = hit ? h.bv : 0.0f;
                    tr.node_pops = c.node_pops; tr.box_tests = c.box_tests; tr.tri_tests = h.steps;
                    tr.tlas_leaves = c.tlas_leaves; tr.max_stack = c.max_stack;
                    tr.visit_hash_lo = (uint32_t)c.hash; tr.visit_hash_hi = (uint32_t)(c.hash >> 32);
                }
                if (job->debug_steps) { // main.glsl:358-361,423-427
                    float e = (float)h.steps / 256.0f;
                    e = e < 0.0f ? 0.0f : (e > 1.0f ? 1.0f : e);
                    radiance = mk(e, e, e);
                    break;
                }
                Shading s;
                if (hit) s = get_shading_data(sc, h);
                else { // sampleSky, main.glsl:189-192
                    float t = 0.5f * (ray.d.y + 1.0f);
                    s.emission = scl(mix3(mk(0.95f, 0.95f, 0.95f), mk(0.9f, 0.94f, 1.0f), t), 1.0f);
                }
                radiance = add(radiance, mulv(throughput, s.emission));
                if (!hit) break;
                if (i == 0) depth = length3(sub(s.position, ray.o));
                ray.o = add(s.position, scl(s.normal, 0.001f));
                ray.d = sample_brdf(s, pcg2d(&sx, &sy));
                ray.rD = mk(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
                float density = brdf_density(s, ray.d);
                float lambert_in = dot(s.normal, ray.d);
                if (lambert_in <= 0.0f) break;
                V3 f = brdf(s, ray.d);
                throughput = mulv(throughput, divs(scl(f, lambert_in), density));
            }
            // main.glsl:430-435
            depth = cam.z_far / (cam.z_far - cam.z_near) * (1.0f - cam.z_near / depth);
            uint8_t *px = job->out_rgba8 + pix * 4;
            px[0] = to_unorm8(radiance.x); px[1] = to_unorm8(radiance.y); px[2] = to_unorm8(radiance.z); px[3] = 255;
            if (job->out_depth) job->out_depth[pix] = depth;
            if (job->out_radiance) { float *r = job->out_radiance + pix * 4; r[0] = radiance.x; r[1] = radiance.y; r[2] = radiance.z; r[3] = 1.0f; }
        }
    }
    job->rays += rays; job->primary_hits += phits; job->node_pops += pops; job->box_tests += boxes;
    job->tri_tests += tris; job->tlas_leaves += leaves;
    uint32_t prev = job->max_stack.load();
    while (prev < max_stack && !job->max_stack.compare_exchange_weak(prev, max_stack)) {}
    if (overflow) job->overflow.store(1);
}

} // namespace

extern "C" {

typedef struct orc_scene {
    const void *tri_geom; uint64_t n_tris;
    const void *tri_data;
    const void *materials; uint64_t n_materials;
    const void *bvh; uint64_t n_nodes;
    const void *blas; uint64_t n_blas;
    const void *tlas; uint64_t n_tlas;
    const uint8_t *textures; int32_t tex_w, tex_h, tex_layers, material_ext;
    const uint32_t *surface_materials; // extension only (gdpt_wire.h): offset[n_blas + 1], ids
} orc_scene;

typedef struct orc_stats {
    uint64_t rays, primary_hits, node_pops, box_tests, tri_tests, tlas_leaves;
    uint32_t max_stack, stack_overflow;
} orc_stats;

// K1 on rows y_begin, y_begin + y_step, ... < y_end of the frame (y_step <= 1: every row).  trace: [trace_segments][W*H] or NULL.
// visits: [W*H][visits_per_ray] (primary rays) or NULL.  out_radiance: [W*H][4] or NULL, the vec4 of main.glsl:434 before
// the rgba8 conversion.
int orc_path_trace(const orc_scene *scene, const gdpt_render_params *params, const gdpt_camera *camera,
                   int max_depth, int debug_steps, int n_threads, int y_begin, int y_end, int y_step,
                   uint8_t *out_rgba8, float *out_depth, gdpt_trace_record *trace, int trace_segments,
                   uint32_t *visits, uint32_t visits_per_ray, orc_stats *stats, float *out_radiance)
{
    RenderJob job;
    job.sc.tri_geom = (const gdpt_triangle_geometry *)scene->tri_geom; job.sc.n_tris = scene->n_tris;
    job.sc.tri_data = (const gdpt_triangle_data *)scene->tri_data;
    job.sc.materials = (const gdpt_material *)scene->materials; job.sc.n_materials = scene->n_materials;
    job.sc.bvh = (const gdpt_bvh_node *)scene->bvh; job.sc.n_nodes = scene->n_nodes;
    job.sc.blas = (const gdpt_blas_instance *)scene->blas; job.sc.n_blas = scene->n_blas;
    job.sc.tlas = (const gdpt_tlas_node *)scene->tlas; job.sc.n_tlas = scene->n_tlas;
    job.sc.textures = scene->textures; job.sc.tex_w = scene->tex_w; job.sc.tex_h = scene->tex_h; job.sc.tex_layers = scene->tex_layers;
    job.sc.material_ext = scene->material_ext; job.sc.surface_materials = scene->material_ext ? scene->surface_materials : nullptr;
    fill_srgb_lut(job.sc);
    job.params = params; job.cam = camera;
    job.max_depth = max_depth; job.debug_steps = debug_steps;
    if (y_begin < 0) y_begin = 0;
    if (y_end > params->height) y_end = params->height;
    job.y_begin = y_begin; job.y_end = y_end; job.y_step = y_step < 1 ? 1 : y_step;
    job.out_rgba8 = out_rgba8; job.out_depth = out_depth; job.out_radiance = out_radiance;
    job.trace = trace; job.trace_segments = trace_segments;
    job.visits = visits; job.visits_per_ray = visits_per_ray;
    job.next_row = y_begin;
    job.rays = 0; job.primary_hits = 0; job.node_pops = 0; job.box_tests = 0; job.tri_tests = 0; job.tlas_leaves = 0;
    job.max_stack = 0; job.overflow = 0;
    if (trace) {
        const size_t n = (size_t)trace_segments * params->width * params->height;
        for (size_t i = 0; i < n; i++) { memset(&trace[i], 0, sizeof(trace[i])); trace[i].hit = 0xFFFFFFFFu; }
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    for (int i = 1; i < n_threads; i++) pool.emplace_back(render_rows, &job);
    render_rows(&job);
    for (auto &t : pool) t.join();
    if (stats) {
        stats->rays = job.rays; stats->primary_hits = job.primary_hits; stats->node_pops = job.node_pops;
        stats->box_tests = job.box_tests; stats->tri_tests = job.tri_tests; stats->tlas_leaves = job.tlas_leaves;
        stats->max_stack = job.max_stack; stats->stack_overflow = job.overflow;
    }
    return 0;
}

// ray_trace_tlas (main.glsl:305-350) on caller-given rays: HitInfo initialised as in ray_trace (main.glsl:354-356),
// rD = 1.0 / d (main.glsl:421).  One trace record per ray.
void orc_trace_rays(const orc_scene *scene, uint64_t n, const float *origins, const float *directions, gdpt_trace_record *out)
{
    Scene sc;
    sc.tri_geom = (const gdpt_triangle_geometry *)scene->tri_geom; sc.n_tris = scene->n_tris;
    sc.tri_data = (const gdpt_triangle_data *)scene->tri_data;
    sc.materials = (const gdpt_material *)scene->materials; sc.n_materials = scene->n_materials;
    sc.bvh = (const gdpt_bvh_node *)scene->bvh; sc.n_nodes = scene->n_nodes;
    sc.blas = (const gdpt_blas_instance *)scene->blas; sc.n_blas = scene->n_blas;
    sc.tlas = (const gdpt_tlas_node *)scene->tlas; sc.n_tlas = scene->n_tlas;
    sc.textures = scene->textures; sc.tex_w = scene->tex_w; sc.tex_h = scene->tex_h; sc.tex_layers = scene->tex_layers;
    for (uint64_t i = 0; i < n; i++) {
        Ray ray;
        ray.o = mk(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        ray.d = mk(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        ray.rD = mk(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        Hit h;
        memset(&h, 0, sizeof(h));
        h.t = 1e9f; h.steps = 0;
        Counters c;
        memset(&c, 0, sizeof(c));
        c.hash = GDPT_FNV64_OFFSET;
        const bool hit = trace_tlas(sc, ray, h, &c);
        gdpt_trace_record &tr = out[i];
        memset(&tr, 0, sizeof(tr));
        tr.hit = hit ? 1u : 0u; tr.triangle = hit ? h.triangle : 0u; tr.blas = hit ? h.blas : 0u;
        tr.front = hit ? (h.front ? 1u : 0u) : 0u;
        tr.t = h.t; tr.u = hit ? h.bu : 0.0f; tr.v = hit ? h.bv : 0.0f;
        tr.node_pops = c.node_pops; tr.box_tests = c.box_tests; tr.tri_tests = h.steps;
        tr.tlas_leaves = c.tlas_leaves; tr.max_stack = c.max_stack;
        tr.visit_hash_lo = (uint32_t)c.hash; tr.visit_hash_hi = (uint32_t)(c.hash >> 32);
    }
}

// K2: prog.glsl:19-46.  screen is RGBA8 in/out, accum is RGBA32F in/out.
void orc_progressive(uint8_t *screen, float *accum, int width, int height, uint32_t frame_count)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float fc = (float)frame_count;
    for (size_t p = 0; p < (size_t)width * height; p++) {
        float rad[3];
        for (int k = 0; k < 3; k++) {
            rad[k] = (float)screen[p * 4 + k] / 255.0f;
            if (frame_count > 1) rad[k] = rad[k] + accum[p * 4 + k];
            accum[p * 4 + k] = rad[k];
        }
        accum[p * 4 + 3] = 1.0f;
        for (int k = 0; k < 3; k++) {
            float x = rad[k] / fc * 1.0f;
            float y = (x * (a * x + b)) / (x * (c * x + d) + e);
            screen[p * 4 + k] = to_unorm8(y);
        }
        screen[p * 4 + 3] = 255;
    }
}

// K3: temp.glsl:31-71.  One dispatch over the whole image, pixel by pixel in row-major order.  No pixel
// reads what another pixel of the same dispatch writes (screen is read and written at `pos` only; depth
// and the history buffer are read-only), so the order is immaterial -- as on the GPU.
//   params     88 B block of temp.glsl:4-12 (delta matrix column-major, width, height, frameCount, ...)
//   screen     rgba8 in/out;  depth: r32f of the current frame;  fb1/fb2: rgba32f ping-pong buffers
// ivec2(vec2): truncation toward zero; NaN and values outside the int range become INT_MIN.
static inline int to_int_trunc(float v)
{
    if (v != v) return INT32_MIN;
    if (v <= -2147483648.0f || v >= 2147483648.0f) return INT32_MIN;
    return (int)v;
}

void orc_temporal(const gdpt_temporal_params *params, uint8_t *screen, const float *depth, float *fb1, float *fb2)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f; // temp.glsl:23-29
    const int W = params->width, H = params->height;
    const float *M = params->delta_matrix;
    const float fW = (float)(uint32_t)W, fH = (float)(uint32_t)H;
    const bool use_first = (params->frame_count % 2u) == 0u;            // temp.glsl:46
    const float *history = use_first ? fb1 : fb2;                       // temp.glsl:62
    float *target = use_first ? fb2 : fb1;                              // temp.glsl:66
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t pos = (size_t)y * W + x;
            float cur[3], rep[3];
            for (int k = 0; k < 3; k++) { cur[k] = (float)screen[pos * 4 + k] / 255.0f; rep[k] = cur[k]; } // :36, :48
            const float z = depth[pos];                                                                    // :37
            float ndc[4];
            ndc[0] = ((float)x + 0.5f) / fW * 2.0f - 1.0f;   // :40
            ndc[1] = ((float)y + 0.5f) / fH * -2.0f + 1.0f;  // :41
            ndc[2] = z; ndc[3] = 1.0f;
            if (params->frame_count > 0u) {                  // :49
                float clip[4];
                for (int r = 0; r < 4; r++)                  // :50, m*v = ((m0*v.x + m1*v.y) + m2*v.z) + m3*v.w
                    clip[r] = ((M[0 + r] * ndc[0] + M[4 + r] * ndc[1]) + M[8 + r] * ndc[2]) + M[12 + r] * ndc[3];
                for (int r = 0; r < 3; r++) clip[r] = clip[r] / clip[3]; // :51
                const float pu = (clip[0] + 1.0f) * 0.5f;    // :54
                const float pv = (1.0f - clip[1]) * 0.5f;    // :55
                const int qx = to_int_trunc(pu * fW), qy = to_int_trunc(pv * fH); // :57
                if (qx >= 0 && qx < W && qy >= 0 && qy < H) { // :59
                    const size_t q = (size_t)qy * W + qx;
                    float dz = depth[q] - clip[2];
                    if (dz < 0.0f) dz = -dz;
                    if (dz < 0.1f)                           // :59
                        for (int k = 0; k < 3; k++) rep[k] = history[q * 4 + k]; // :60
                }
            }
            for (int k = 0; k < 3; k++) {
                const float blended = cur[k] * (1.0f - 0.75f) + rep[k] * 0.75f; // :64 mix(a,b,t) = a*(1-t) + b*t
                target[pos * 4 + k] = blended;                                  // :66
                const float tm = (blended * (a * blended + b)) / (blended * (c * blended + d) + e); // :68
                screen[pos * 4 + k] = to_unorm8(tm);                            // :70
            }
            target[pos * 4 + 3] = 1.0f;
            screen[pos * 4 + 3] = 255;
        }
    }
}

// Host arithmetic the camera and the temporal post process rely on, restated from godot-cpp @56571dc (real_t = float).
// Matrices are 16 floats, column-major, m[col*4 + row], like godot::Projection::columns.
// Projection::operator* (godot-cpp/src/variant/projection.cpp:709-723)
void orc_mat4_mul(const float *A, const float *B, float *out)
{
    float r[16];
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) {
            float ab = 0;
            for (int k = 0; k < 4; k++) ab += A[k * 4 + i] * B[j * 4 + k];
            r[j * 4 + i] = ab;
        }
    memcpy(out, r, sizeof(r));
}

// Projection::invert (godot-cpp/src/variant/projection.cpp:601-698): Gauss-Jordan with full pivoting.
void orc_mat4_inverse(const float *in, float *out)
{
    float m[4][4];
    memcpy(m, in, sizeof(m));
    int pvt_i[4], pvt_j[4];
    float det = 1.0f;
    for (int k = 0; k < 4; k++) {
        float pvt_val = m[k][k];
        pvt_i[k] = k; pvt_j[k] = k;
        for (int i = k; i < 4; i++)
            for (int j = k; j < 4; j++)
                if (fabsf(m[i][j]) > fabsf(pvt_val)) { pvt_i[k] = i; pvt_j[k] = j; pvt_val = m[i][j]; }
        det *= pvt_val;
        if (fabsf(det) < 0.00001f) { memcpy(out, m, sizeof(m)); return; }
        int i = pvt_i[k];
        if (i != k)
            for (int j = 0; j < 4; j++) { const float hold = -m[k][j]; m[k][j] = m[i][j]; m[i][j] = hold; }
        int j = pvt_j[k];
        if (j != k)
            for (i = 0; i < 4; i++) { const float hold = -m[i][k]; m[i][k] = m[i][j]; m[i][j] = hold; }
        for (i = 0; i < 4; i++)
            if (i != k) m[i][k] /= (-pvt_val);
        for (i = 0; i < 4; i++) {
            const float hold = m[i][k];
            for (j = 0; j < 4; j++)
                if (i != k && j != k) m[i][j] += hold * m[k][j];
        }
        for (j = 0; j < 4; j++)
            if (j != k) m[k][j] /= pvt_val;
        m[k][k] = 1.0f / pvt_val;
    }
    for (int k = 4 - 2; k >= 0; k--) {
        int i = pvt_j[k];
        if (i != k)
            for (int j = 0; j < 4; j++) { const float hold = m[k][j]; m[k][j] = -m[i][j]; m[i][j] = hold; }
        int j = pvt_i[k];
        if (j != k)
            for (i = 0; i < 4; i++) { const float hold = m[i][k]; m[i][k] = -m[i][j]; m[i][j] = hold; }
    }
    memcpy(out, m, sizeof(m));
}

// KAT helpers.
void orc_prng_seed(uint32_t px, uint32_t py, uint32_t frame, uint32_t *out2) { prng_seed(px, py, frame, &out2[0], &out2[1]); }
void orc_pcg2d(uint32_t *state2, float *out2) { V2 r = pcg2d(&state2[0], &state2[1]); out2[0] = r.x; out2[1] = r.y; }
void orc_sincosf(float x, float *out2) { orc_sincos(x, &out2[0], &out2[1]); }
unsigned orc_hardware_threads(void) { unsigned n = std::thread::hardware_concurrency(); return n ? n : 1; }

} // extern "C"
