// um_device.cuh -- per-pixel FarnebackUpdateMatrices as a device function, shared by
// the stand-alone kernel (first UpdateMatrices of a scale, fused with the inter-scale flow up-sample)
// and by the fused iteration kernel (blur -> solve -> UpdateMatrices in one launch).
// All f32, uncontracted (-fmad=false), in the upstream order of operations.
#pragma once
#include "common.cuh"

namespace ofb {

struct M5 { float v[5]; };

__device__ __forceinline__ M5 um_pixel(int x, int y, float dx, float dy, const Planes5& R0, const Planes5& R1, int W, int H)
{
    float fx = x + dx, fy = y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= x1; fy -= y1;
    const unsigned o0 = (unsigned)y * (unsigned)R0.pitch + (unsigned)x;
    const float q0 = R0.ch(0)[o0], q1 = R0.ch(1)[o0], q2 = R0.ch(2)[o0], q3 = R0.ch(3)[o0], q4 = R0.ch(4)[o0];
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1)) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const unsigned o1 = (unsigned)y1 * (unsigned)R1.pitch + (unsigned)x1;
        const int p = R1.pitch;
        const float* c0 = R1.ch(0) + o1; const float* c1 = R1.ch(1) + o1; const float* c2 = R1.ch(2) + o1;
        const float* c3 = R1.ch(3) + o1; const float* c4 = R1.ch(4) + o1;
        // issue all 20 gathers before the arithmetic
        float t00 = c0[0], t01 = c0[1], t02 = c0[p], t03 = c0[p + 1];
        float t10 = c1[0], t11 = c1[1], t12 = c1[p], t13 = c1[p + 1];
        float t20 = c2[0], t21 = c2[1], t22 = c2[p], t23 = c2[p + 1];
        float t30 = c3[0], t31 = c3[1], t32 = c3[p], t33 = c3[p + 1];
        float t40 = c4[0], t41 = c4[1], t42 = c4[p], t43 = c4[p + 1];
        r2 = a00 * t00 + a01 * t01 + a10 * t02 + a11 * t03;
        r3 = a00 * t10 + a01 * t11 + a10 * t12 + a11 * t13;
        r4 = a00 * t20 + a01 * t21 + a10 * t22 + a11 * t23;
        r5 = a00 * t30 + a01 * t31 + a10 * t32 + a11 * t33;
        r6 = a00 * t40 + a01 * t41 + a10 * t42 + a11 * t43;
        r4 = (q2 + r4) * 0.5f;
        r5 = (q3 + r5) * 0.5f;
        r6 = (q4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = q2; r5 = q3; r6 = q4 * 0.5f;
    }
    r2 = (q0 - r2) * 0.5f;
    r3 = (q1 - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(W - 10) || (unsigned)(y - 5) >= (unsigned)(H - 10)) {
        const float b0 = 0.14f, b2 = 0.4472f;
        float sx0 = x < 5 ? (x < 2 ? b0 : b2) : 1.f;
        float sx1 = x >= W - 5 ? ((W - x - 1) < 2 ? b0 : b2) : 1.f;
        float sy0 = y < 5 ? (y < 2 ? b0 : b2) : 1.f;
        float sy1 = y >= H - 5 ? ((H - y - 1) < 2 ? b0 : b2) : 1.f;
        float scale = sx0 * sx1 * sy0 * sy1;
        r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
    }
    M5 m;
    m.v[0] = r4 * r4 + r6 * r6;
    m.v[1] = (r4 + r5) * r6;
    m.v[2] = r5 * r5 + r6 * r6;
    m.v[3] = r4 * r2 + r6 * r3;
    m.v[4] = r6 * r2 + r5 * r3;
    return m;
}

}  // namespace ofb
