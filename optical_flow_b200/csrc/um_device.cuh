// um_device.cuh -- per-pixel FarnebackUpdateMatrices as a device function, shared by
// the stand-alone kernel (first UpdateMatrices of a scale, fused with the inter-scale flow up-sample)
// and by the fused iteration kernel (blur -> solve -> UpdateMatrices in one launch).
// All f32, uncontracted (-fmad=false), in the upstream order of operations.
#pragma once
#include "common.cuh"

namespace ofb {

struct M5 { float v[5]; };

__device__ __forceinline__ M5 um_pixel(int x, int y, float dx, float dy, const RView& R0, const RView& R1, int W, int H)
{
    float fx = x + dx, fy = y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= x1; fy -= y1;
    const unsigned o0 = (unsigned)y * (unsigned)R0.pitch + (unsigned)x;
    const float4 q = R0.a[o0];
    const float q4 = R0.b[o0];
    const float q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1)) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const unsigned o1 = (unsigned)y1 * (unsigned)R1.pitch + (unsigned)x1;
        const int p = R1.pitch;
        const float4* pa = R1.a + o1;
        const float* pb = R1.b + o1;
        // 8 gathers, all issued before the arithmetic
        const float4 A00 = pa[0], A01 = pa[1], A10 = pa[p], A11 = pa[p + 1];
        const float B00 = pb[0], B01 = pb[1], B10 = pb[p], B11 = pb[p + 1];
        r2 = a00 * A00.x + a01 * A01.x + a10 * A10.x + a11 * A11.x;
        r3 = a00 * A00.y + a01 * A01.y + a10 * A10.y + a11 * A11.y;
        r4 = a00 * A00.z + a01 * A01.z + a10 * A10.z + a11 * A11.z;
        r5 = a00 * A00.w + a01 * A01.w + a10 * A10.w + a11 * A11.w;
        r6 = a00 * B00 + a01 * B01 + a10 * B10 + a11 * B11;
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
