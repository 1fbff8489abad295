// pt_post.cuh -- per-pixel body of K3, the temporal-reprojection post process
// (temporal_reprojection.glsl:31-71), as one host/device function so that the CUDA kernel and the
// CPU test tier (tests/devcheck) execute the same source.
//
// Arithmetic contract as everywhere else: every GLSL operation is one binary32 operation,
// left to right, no FMA contraction; mat4 * vec4 is expanded column by column like `xform`
// (pt_math.cuh); mix(x, y, a) = x*(1-a) + y*a; imageLoad of rgba8 = k / 255; ivec2(vec2) truncates
// toward zero, and a value that does not fit an int (or NaN) becomes INT_MIN, i.e. "outside the image"
// (GLSL leaves it undefined).
#ifndef GDPT_PT_POST_CUH
#define GDPT_PT_POST_CUH

#include "pt_shade.cuh"

namespace gdpt {

GDPT_HD int float_to_int_trunc(float v)
{
    if (!(v > -2147483648.0f) || !(v < 2147483648.0f)) return (int)0x80000000;
    return (int)v;
}

// One pixel of temporal_reprojection.glsl:31-71.
//   screen   rgba8, read at `pos`, rewritten at `pos` (:36, :70)
//   depth    r32f of the CURRENT frame, read at `pos` and at the reprojected position (:37, :61)
//   history  the rgba32f frame buffer the previous dispatch wrote (frameBuffer1 when frameCount is even, :46, :62)
//   next     the other frame buffer, written at `pos` (:66)
// `unorm` (device only, may be null): table of the 256 quotients k / 255.0f, the same IEEE divisions done once per block
GDPT_HD void temporal_pixel(const gdpt_temporal_params &p, int x, int y, uint32_t *screen, const float *depth,
                            const float *history, float *next, const float *unorm = nullptr)
{
    const int width = p.width, height = p.height;
    const size_t at = (size_t)y * (size_t)width + (size_t)x;
    const uint32_t in = screen[at];
    const f3 current = unorm ? mk3(unorm[in & 0xffu], unorm[(in >> 8) & 0xffu], unorm[(in >> 16) & 0xffu])
                             : mk3((float)(in & 0xffu) / 255.0f, (float)((in >> 8) & 0xffu) / 255.0f, (float)((in >> 16) & 0xffu) / 255.0f);
    const float d = depth[at];
    const float fw = (float)(uint32_t)width, fh = (float)(uint32_t)height;
    const float nx = ((float)x + 0.5f) / fw * 2.0f - 1.0f;
    const float ny = ((float)y + 0.5f) / fh * -2.0f + 1.0f;
    f3 reprojected = current;
    if (p.frame_count > 0u) {
        const float *m = p.delta_matrix;
        float cx = ((m[0] * nx + m[4] * ny) + m[8] * d) + m[12] * 1.0f;
        float cy = ((m[1] * nx + m[5] * ny) + m[9] * d) + m[13] * 1.0f;
        float cz = ((m[2] * nx + m[6] * ny) + m[10] * d) + m[14] * 1.0f;
        const float cw = ((m[3] * nx + m[7] * ny) + m[11] * d) + m[15] * 1.0f;
        cx = cx / cw; cy = cy / cw; cz = cz / cw;
        const float u = (cx + 1.0f) * 0.5f, v = (1.0f - cy) * 0.5f;
        const int px = float_to_int_trunc(u * fw), py = float_to_int_trunc(v * fh);
        if (px >= 0 && px < width && py >= 0 && py < height) {
            const size_t prev = (size_t)py * (size_t)width + (size_t)px;
            if (fabsf(depth[prev] - cz) < 0.1f) {
#if defined(__CUDA_ARCH__)
                const float4 h = reinterpret_cast<const float4 *>(history)[prev]; // one 128-bit load
                reprojected = mk3(h.x, h.y, h.z);
#else
                reprojected = mk3(history[prev * 4 + 0], history[prev * 4 + 1], history[prev * 4 + 2]);
#endif
            }
        }
    }
    const f3 blended = mix3(current, reprojected, 0.75f);
#if defined(__CUDA_ARCH__)
    reinterpret_cast<float4 *>(next)[at] = make_float4(blended.x, blended.y, blended.z, 1.0f);
#else
    next[at * 4 + 0] = blended.x; next[at * 4 + 1] = blended.y; next[at * 4 + 2] = blended.z; next[at * 4 + 3] = 1.0f;
#endif
    screen[at] = tone_map_rgba8(blended);
}

} // namespace gdpt
#endif
