// glsl.hpp -- TEST INFRASTRUCTURE ONLY (oracle/).  NEVER SHIPPED.
//
// The slice of GLSL 4.60 the reference's shaders use, as C++ types and functions, so that the shader TEXT
// (main.glsl, brdfs.glsl, progressive_rendering.glsl, temporal_reprojection.glsl of the reference) compiles
// as the body of a C++ struct and runs on the CPU (oracle/ref_shader_bridge.cpp, oracle/Makefile).  The
// algorithm is whatever the shader text says; this header only supplies what a GLSL compiler and a GPU would:
//
//   * vector / matrix types with the swizzles and operators the shaders use.  Every operator is
//     component-wise IEEE binary32, one rounding per operation, no FMA contraction (-ffp-contract=off).
//   * the built-ins.  GLSL leaves their precision implementation-defined; the definitions here are the
//     arithmetic contract of DESIGN.md section 2, restated independently of pt_oracle.cpp:
//       dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z        m*v = ((m0*v.x + m1*v.y) + m2*v.z) + m3*v.w
//       cross per the GLSL specification                 normalize(v) = v / sqrt(dot(v,v))
//       length(v) = sqrt(dot(v,v))                       mix(a,b,t) = a*(1-t) + b*t
//       reflect(I,N) = I - 2*dot(N,I)*N                  clamp(x,lo,hi) = min(max(x,lo),hi)
//       min / max drop a NaN operand (what FMNMX and v_min_f32 do), otherwise (b<a)?b:a and (a<b)?b:a
//       sqrt and '/' correctly rounded; sin / cos = the Cody-Waite + minimax polynomial of the contract
//       imageStore to rgba8 = round-half-even(clamp(x,0,1)*255), NaN -> 0;  imageLoad = byte / 255.0f
//       texture() = nearest, clamp-to-edge, level 0 (the default RDSamplerState, gdcs.cpp:183-187)
//   * resources (ssbo<T>, image2D<format>, sampler2DArray) as views over caller memory.  ssbo<T> reads
//     records at the std430 stride the host wrote them with and can log the index of every read: the
//     node-visit order, triangle-test count and hit ids of the reference traversal are observed from those
//     reads without touching the shader text.
#ifndef GDPT_ORACLE_GLSL_HPP
#define GDPT_ORACLE_GLSL_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

namespace glsl {

typedef uint32_t uint;

struct vec2; struct vec3; struct vec4; struct uvec2; struct uvec3; struct ivec2;

// GLSL implicit conversions between vector types (4.1.10): int -> uint -> float.
template <class From, class To> struct implicit_to : std::false_type {};
template <> struct implicit_to<uvec2, vec2> : std::true_type {};
template <> struct implicit_to<ivec2, vec2> : std::true_type {};
template <> struct implicit_to<ivec2, uvec2> : std::true_type {};

// Swizzles alias the storage of the vector they belong to (they are members of its union).
template <class V, class T, int N, int A, int B> struct swz2 {
    T v[N];
    operator V() const { return V(v[A], v[B]); }
    template <class W, class = typename std::enable_if<implicit_to<V, W>::value>::type> operator W() const { return W(V(v[A], v[B])); }
    void operator=(const V &o) { v[A] = o.x; v[B] = o.y; }
};
template <class V, class T, int N, int A, int B, int C> struct swz3 {
    T v[N];
    operator V() const { return V(v[A], v[B], v[C]); }
    void operator=(const V &o) { v[A] = o.x; v[B] = o.y; v[C] = o.z; }
    void operator/=(T s) { v[A] = v[A] / s; v[B] = v[B] / s; v[C] = v[C] / s; }
};

struct vec2 {
    union {
        struct { float x, y; };
        swz2<vec2, float, 2, 0, 1> xy;
    };
    vec2() = default;
    explicit vec2(float s) { x = s; y = s; }
    vec2(float a, float b) { x = a; y = b; }
    vec2(const uvec2 &u);
    vec2(const ivec2 &i);
};
struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz2<vec2, float, 3, 0, 1> xy;
        swz3<vec3, float, 3, 0, 1, 2> xyz;
        swz3<vec3, float, 3, 0, 1, 2> rgb;
    };
    vec3() = default;
    explicit vec3(float s) { x = s; y = s; z = s; }
    vec3(float a, float b, float c) { x = a; y = b; z = c; }
    vec3(const vec2 &a, float c) { x = a.x; y = a.y; z = c; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<vec2, float, 4, 0, 1> xy;
        swz3<vec3, float, 4, 0, 1, 2> xyz;
        swz3<vec3, float, 4, 0, 1, 2> rgb;
    };
    vec4() = default;
    explicit vec4(float s) { x = s; y = s; z = s; w = s; }
    vec4(float a, float b, float c, float d) { x = a; y = b; z = c; w = d; }
    vec4(const vec3 &v, float d) { x = v.x; y = v.y; z = v.z; w = d; }
};
struct uvec2 {
    union {
        struct { uint x, y; };
        swz2<uvec2, uint, 2, 0, 1> xy;
    };
    uvec2() = default;
    explicit uvec2(uint s) { x = s; y = s; }
    uvec2(uint a, uint b) { x = a; y = b; }
    explicit uvec2(const vec2 &f) { x = (uint)f.x; y = (uint)f.y; }
};
struct uvec3 {
    union {
        struct { uint x, y, z; };
        swz2<uvec2, uint, 3, 0, 1> xy;
    };
    uvec3() = default;
    uvec3(uint a, uint b, uint c) { x = a; y = b; z = c; }
};
// float -> int: truncation toward zero; NaN and out-of-range values give INT_MIN (what cvttss2si returns;
// GLSL leaves them undefined).  Only temporal_reprojection.glsl:57 can reach that case.
inline int trunc_to_int(float v)
{
    if (v != v || v <= -2147483648.0f || v >= 2147483648.0f) return INT32_MIN;
    return (int)v;
}
struct ivec2 {
    union {
        struct { int x, y; };
        swz2<ivec2, int, 2, 0, 1> xy;
    };
    ivec2() = default;
    ivec2(int a, int b) { x = a; y = b; }
    explicit ivec2(const uvec2 &u) { x = (int)u.x; y = (int)u.y; }
    explicit ivec2(const vec2 &f) { x = trunc_to_int(f.x); y = trunc_to_int(f.y); }
    template <class T, int N, int A, int B> explicit ivec2(const swz2<uvec2, T, N, A, B> &s) { uvec2 u = s; x = (int)u.x; y = (int)u.y; }
};
inline vec2::vec2(const uvec2 &u) { x = (float)u.x; y = (float)u.y; }
inline vec2::vec2(const ivec2 &i) { x = (float)i.x; y = (float)i.y; }

// ---- float vectors: component-wise, one rounding per operation
#define GLSL_BINOP(V, OP, BODY_VV, BODY_VS, BODY_SV)                  \
    inline V operator OP(const V &a, const V &b) { return BODY_VV; }  \
    inline V operator OP(const V &a, float b) { return BODY_VS; }     \
    inline V operator OP(float a, const V &b) { return BODY_SV; }     \
    inline V &operator OP##=(V &a, const V &b) { a = a OP b; return a; } \
    inline V &operator OP##=(V &a, float b) { a = a OP b; return a; }
#define GLSL_OPS2(OP) GLSL_BINOP(vec2, OP, vec2(a.x OP b.x, a.y OP b.y), vec2(a.x OP b, a.y OP b), vec2(a OP b.x, a OP b.y))
#define GLSL_OPS3(OP) GLSL_BINOP(vec3, OP, vec3(a.x OP b.x, a.y OP b.y, a.z OP b.z), vec3(a.x OP b, a.y OP b, a.z OP b), vec3(a OP b.x, a OP b.y, a OP b.z))
#define GLSL_OPS4(OP) GLSL_BINOP(vec4, OP, vec4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w), vec4(a.x OP b, a.y OP b, a.z OP b, a.w OP b), vec4(a OP b.x, a OP b.y, a OP b.z, a OP b.w))
GLSL_OPS2(+) GLSL_OPS2(-) GLSL_OPS2(*) GLSL_OPS2(/)
GLSL_OPS3(+) GLSL_OPS3(-) GLSL_OPS3(*) GLSL_OPS3(/)
GLSL_OPS4(+) GLSL_OPS4(-) GLSL_OPS4(*) GLSL_OPS4(/)
inline vec2 operator-(const vec2 &a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4 &a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

// ---- uvec2: modulo-2^32 arithmetic
inline uvec2 operator*(uint a, const uvec2 &b) { return uvec2(a * b.x, a * b.y); }
inline uvec2 operator*(const uvec2 &a, uint b) { return uvec2(a.x * b, a.y * b); }
inline uvec2 operator+(const uvec2 &a, uint b) { return uvec2(a.x + b, a.y + b); }
inline uvec2 operator+(const uvec2 &a, const uvec2 &b) { return uvec2(a.x + b.x, a.y + b.y); }
inline uvec2 operator>>(const uvec2 &a, uint s) { return uvec2(a.x >> s, a.y >> s); }
inline uvec2 operator^(const uvec2 &a, const uvec2 &b) { return uvec2(a.x ^ b.x, a.y ^ b.y); }
inline uvec2 &operator^=(uvec2 &a, const uvec2 &b) { a = a ^ b; return a; }

// ---- matrices: column-major, as in GLSL and in the host blobs (src/utils.h:15-49)
struct mat3 {
    vec3 c[3];
    mat3() = default;
    mat3(const vec3 &c0, const vec3 &c1, const vec3 &c2) { c[0] = c0; c[1] = c1; c[2] = c2; }
};
struct mat4 { vec4 c[4]; };
inline vec3 operator*(const mat3 &m, const vec3 &v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
inline vec4 operator*(const mat4 &m, const vec4 &v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }
inline mat3 transpose(const mat3 &m)
{
    return mat3(vec3(m.c[0].x, m.c[1].x, m.c[2].x), vec3(m.c[0].y, m.c[1].y, m.c[2].y), vec3(m.c[0].z, m.c[1].z, m.c[2].z));
}

// ---- scalar built-ins
inline float min(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((b < a) ? b : a)); }
inline float max(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((a < b) ? b : a)); }
inline float max(int a, float b) { return max((float)a, b); } // main.glsl:213 max(0, material.emission.w)
inline float abs(float a) { return std::fabs(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float log(float a) { return std::log(a); } // feeds only the dead `R` of box_muller (main.glsl:184)
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }

// sin and cos of x >= 0 (every argument the shaders form is in [0, 2*pi]): k = nearest multiple of pi/2,
// three-constant Cody-Waite reduction, single-precision minimax polynomials on [-pi/4, pi/4].
inline void sincos_contract(float x, float *s_out, float *c_out)
{
    const int k = (int)(x * 0.636619772f + 0.5f);
    const float kf = (float)k;
    const float r = ((x - kf * 1.5703125f) - kf * 4.837512969970703125e-4f) - kf * 7.54978995489188216e-8f;
    const float z = r * r;
    const float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    const float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    const float s[4] = { sp, cp, -sp, -cp }, c[4] = { cp, -sp, -cp, sp };
    *s_out = s[k & 3];
    *c_out = c[k & 3];
}
inline float sin(float x) { float s, c; sincos_contract(x, &s, &c); return s; }
inline float cos(float x) { float s, c; sincos_contract(x, &s, &c); return c; }

// ---- vector built-ins
inline float dot(const vec3 &a, const vec3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec3 &a) { return sqrt(dot(a, a)); }
inline vec3 normalize(const vec3 &a) { return a / sqrt(dot(a, a)); }
inline vec3 mix(const vec3 &a, const vec3 &b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 reflect(const vec3 &I, const vec3 &N) { return I - 2.0f * dot(N, I) * N; }
inline vec3 clamp(const vec3 &v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }

// ---- resources
// std430 runtime array in a storage buffer.  `tag` marks which buffer a logged read came from.
template <class T> struct ssbo {
    const uint8_t *base = nullptr;
    size_t stride = 0, count = 0;
    std::vector<uint64_t> *log = nullptr;
    uint64_t tag = 0;
    uint64_t *reads_of_record_0 = nullptr; // optional cheap observer (ray count = reads of TLAS node 0)
    T operator[](uint i) const
    {
        if (log) log->push_back(tag | i);
        if (reads_of_record_0 && i == 0u) ++*reads_of_record_0;
        T t;
        std::memcpy(&t, base + stride * (size_t)i, sizeof(T));
        return t;
    }
};

struct rgba8 {}; struct r32f {}; struct rgba32f {};
template <class Format> struct image2D {
    void *data = nullptr;
    int width = 0, height = 0;
    float *raw_store = nullptr; // optional: the unconverted vec4 of every imageStore (observer, rgba8 only)
};
inline uint8_t to_unorm8(float x)
{
    if (x != x) return 0;
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return (uint8_t)std::nearbyint(x * 255.0f); // default rounding mode: nearest, ties to even
}
inline void imageStore(image2D<rgba8> &img, const ivec2 &p, const vec4 &v)
{
    const size_t i = (size_t)p.y * img.width + p.x;
    uint8_t *q = (uint8_t *)img.data + i * 4;
    q[0] = to_unorm8(v.x); q[1] = to_unorm8(v.y); q[2] = to_unorm8(v.z); q[3] = to_unorm8(v.w);
    if (img.raw_store) std::memcpy(img.raw_store + i * 4, &v, 16);
}
inline void imageStore(image2D<r32f> &img, const ivec2 &p, const vec4 &v) { ((float *)img.data)[(size_t)p.y * img.width + p.x] = v.x; }
inline void imageStore(image2D<rgba32f> &img, const ivec2 &p, const vec4 &v) { std::memcpy((float *)img.data + ((size_t)p.y * img.width + p.x) * 4, &v, 16); }
inline vec4 imageLoad(const image2D<rgba8> &img, const ivec2 &p)
{
    const uint8_t *q = (const uint8_t *)img.data + ((size_t)p.y * img.width + p.x) * 4;
    return vec4((float)q[0] / 255.0f, (float)q[1] / 255.0f, (float)q[2] / 255.0f, (float)q[3] / 255.0f);
}
inline vec4 imageLoad(const image2D<r32f> &img, const ivec2 &p) { return vec4(((const float *)img.data)[(size_t)p.y * img.width + p.x], 0.0f, 0.0f, 1.0f); }
inline vec4 imageLoad(const image2D<rgba32f> &img, const ivec2 &p)
{
    vec4 v;
    std::memcpy(&v, (const float *)img.data + ((size_t)p.y * img.width + p.x) * 4, 16);
    return v;
}

struct sampler2DArray {
    const uint8_t *texels = nullptr; // RGBA8 UNORM, [layer][y][x][4] (path_tracing_camera.cpp:182)
    int width = 0, height = 0, layers = 0;
};
inline vec4 texture(const sampler2DArray &s, const vec3 &p)
{
    int ix = (int)std::floor(p.x * (float)s.width), iy = (int)std::floor(p.y * (float)s.height);
    ix = ix < 0 ? 0 : (ix > s.width - 1 ? s.width - 1 : ix);
    iy = iy < 0 ? 0 : (iy > s.height - 1 ? s.height - 1 : iy);
    int layer = (int)std::floor(p.z + 0.5f);
    layer = layer < 0 ? 0 : (layer > s.layers - 1 ? s.layers - 1 : layer);
    const uint8_t *q = s.texels + (((size_t)layer * s.height + iy) * s.width + ix) * 4;
    return vec4((float)q[0] / 255.0f, (float)q[1] / 255.0f, (float)q[2] / 255.0f, (float)q[3] / 255.0f);
}

} // namespace glsl
#endif
