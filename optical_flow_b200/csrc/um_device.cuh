// um_device.cuh -- per-pixel FarnebackUpdateMatrices as a device function, shared by
// the stand-alone kernel (first UpdateMatrices of a scale, fused with the inter-scale flow up-sample)
// and by the fused iteration kernel (blur -> solve -> UpdateMatrices in one launch).
// All f32, uncontracted (-fmad=false), in the upstream order of operations.
#pragma once
#include "common.cuh"

namespace ofb {

struct M5 { float v[5]; };

// Loads of one pixel's UpdateMatrices, separated from the arithmetic so that a thread can keep the gathers
// of several pixels in flight before it consumes any of them.
struct UmLoads {
    float4 q; float q4;                 // R0 at the pixel
    float4 A00, A01, A10, A11;          // R1 channels 0..3 at the four neighbours
    float B00, B01, B10, B11;           // R1 channel 4
    float fx, fy, dx, dy;               // fractional position, flow
    int x, y;
    bool inside;
};

__device__ __forceinline__ UmLoads um_load(int x, int y, float dx, float dy, const RView& R0, const RView& R1, int W, int H)
{
    UmLoads L;
    L.x = x; L.y = y; L.dx = dx; L.dy = dy;
    float fx = x + dx, fy = y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    L.fx = fx - x1; L.fy = fy - y1;
    const unsigned o0 = (unsigned)y * (unsigned)R0.pitch + (unsigned)x;
#ifdef OFB_R0_NOALLOC                   // R0 at the pixel itself is used once: keep it out of L1, which the R1 gathers re-use
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(L.q.x), "=f"(L.q.y), "=f"(L.q.z), "=f"(L.q.w) : "l"(R0.a + o0));
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(L.q4) : "l"(R0.b + o0));
#else
    L.q = R0.a[o0];
    L.q4 = R0.b[o0];
#endif
    L.inside = (unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1);
    if (L.inside) {
        const unsigned o1 = (unsigned)y1 * (unsigned)R1.pitch + (unsigned)x1;
        const int p = R1.pitch;
        const float4* pa = R1.a + o1;
        const float* pb = R1.b + o1;
        L.A00 = pa[0]; L.A01 = pa[1]; L.A10 = pa[p]; L.A11 = pa[p + 1];
        L.B00 = pb[0]; L.B01 = pb[1]; L.B10 = pb[p]; L.B11 = pb[p + 1];
    }
    return L;
}

__device__ __forceinline__ M5 um_compute(const UmLoads& L, int W, int H)
{
    const int x = L.x, y = L.y;
    const float dx = L.dx, dy = L.dy, fx = L.fx, fy = L.fy;
    const float q0 = L.q.x, q1 = L.q.y, q2 = L.q.z, q3 = L.q.w, q4 = L.q4;
    float r2, r3, r4, r5, r6;
    if (L.inside) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * L.A00.x + a01 * L.A01.x + a10 * L.A10.x + a11 * L.A11.x;
        r3 = a00 * L.A00.y + a01 * L.A01.y + a10 * L.A10.y + a11 * L.A11.y;
        r4 = a00 * L.A00.z + a01 * L.A01.z + a10 * L.A10.z + a11 * L.A11.z;
        r5 = a00 * L.A00.w + a01 * L.A01.w + a10 * L.A10.w + a11 * L.A11.w;
        r6 = a00 * L.B00 + a01 * L.B01 + a10 * L.B10 + a11 * L.B11;
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

__device__ __forceinline__ M5 um_pixel(int x, int y, float dx, float dy, const RView& R0, const RView& R1, int W, int H)
{
    UmLoads L = um_load(x, y, dx, dy, R0, R1, W, H);
    return um_compute(L, W, H);
}

}  // namespace ofb
