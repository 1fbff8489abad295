// pt_math.cuh -- scalar/vector helpers that implement the arithmetic contract of
// DESIGN.md ("every GLSL operation is one IEEE binary32 op, evaluated left to
// right, no FMA contraction").  The translation unit MUST be compiled with
// -fmad=false (nvcc) and without --use_fast_math; division and sqrt are the
// correctly rounded forms (-prec-div=true -prec-sqrt=true are nvcc defaults).
//
// GDPT_HD lets the same functions be compiled for the host by the CPU-side
// unit check of the device functions (tests/devcheck); the shipped library only
// ever runs them on the GPU.
#ifndef GDPT_PT_MATH_CUH
#define GDPT_PT_MATH_CUH

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define GDPT_HD __host__ __device__ __forceinline__
#else
#define GDPT_HD inline
#endif

namespace gdpt {

struct f3 { float x, y, z; };

GDPT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
GDPT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
GDPT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
GDPT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
GDPT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
GDPT_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
GDPT_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
GDPT_HD float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
GDPT_HD f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GDPT_HD float length3(f3 a) { return sqrtf(dot3(a, a)); }
GDPT_HD f3 normalize3(f3 a) { return a / length3(a); }
GDPT_HD f3 mix3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
GDPT_HD float mix1(float a, float b, float t) { return a * (1.0f - t) + b * t; }
GDPT_HD f3 rcp3(f3 a) { return mk3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }

// GLSL max(c, x) / min(c, x) against a constant, as selects (defined NaN flow).
GDPT_HD float max_c(float c, float x) { return (c < x) ? x : c; }
GDPT_HD float min_c(float c, float x) { return (x < c) ? x : c; }

// IEEE minNum/maxNum for the slab test: one FMNMX each on the GPU.
GDPT_HD float min_num(float a, float b) { return fminf(a, b); }
GDPT_HD float max_num(float a, float b) { return fmaxf(a, b); }

// Column-major mat4 (16 floats) times (v, w); rows 0..2.
GDPT_HD f3 xform_point(const float *m, f3 v, float w)
{
    return mk3(((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * w,
               ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * w,
               ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * w);
}

// sin/cos for x >= 0 (all call sites pass [0, 2*pi]): nearest multiple of pi/2,
// three-constant Cody-Waite reduction, minimax polynomials on [-pi/4, pi/4].
GDPT_HD void sincos_det(float x, float *s_out, float *c_out)
{
    const int k = (int)(x * 0.636619772f + 0.5f);
    const float kf = (float)k;
    const float r = ((x - kf * 1.5703125f) - kf * 4.837512969970703125e-4f) - kf * 7.54978995489188216e-8f;
    const float z = r * r;
    const float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    const float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    const int q = k & 3;
    const float s = (q & 1) ? cp : sp;
    const float c = (q & 1) ? sp : cp;
    *s_out = (q & 2) ? -s : s;
    *c_out = (q == 1 || q == 2) ? -c : c;
}

// imageStore to an rgba8 image: round-half-even(clamp(x,0,1)*255), NaN -> 0.
GDPT_HD uint32_t to_unorm8(float x)
{
#if defined(__CUDA_ARCH__)
    return __float2uint_rn(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f);
#else
    if (x != x) return 0u;
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    return (uint32_t)nearbyintf(x * 255.0f);
#endif
}
GDPT_HD uint32_t pack_rgba8(f3 c)
{
    return to_unorm8(c.x) | (to_unorm8(c.y) << 8) | (to_unorm8(c.z) << 16) | 0xFF000000u;
}

// ---- RNG (main.glsl:163-181) -------------------------------------------------
struct u2 { uint32_t x, y; };
struct f2 { float x, y; };

GDPT_HD u2 prng_seed(uint32_t px, uint32_t py, uint32_t frame)
{
    u2 s;
    s.x = px * 0x9e3779b9u + frame;
    s.y = py * 0x9e3779b9u + frame;
    s.x ^= s.x >> 16; s.y ^= s.y >> 16;
    s.x *= 0x9e3779b9u; s.y *= 0x9e3779b9u;
    return s;
}

GDPT_HD f2 pcg2d(u2 &s)
{
    s.x = 1664525u * s.x + 1013904223u;
    s.y = 1664525u * s.y + 1013904223u;
    s.x += 1664525u * s.y; s.y += 1664525u * s.x;
    s.x ^= s.x >> 16; s.y ^= s.y >> 16;
    s.x += 1664525u * s.y; s.y += 1664525u * s.x;
    s.x ^= s.x >> 16; s.y ^= s.y >> 16;
    f2 r;
    r.x = (float)s.x * 2.32830643654e-10f;
    r.y = (float)s.y * 2.32830643654e-10f;
    return r;
}

} // namespace gdpt
#endif
