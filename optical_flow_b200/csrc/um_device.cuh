// um_device.cuh -- per-pixel FarnebackUpdateMatrices as a device function, shared by
// the stand-alone kernel (first UpdateMatrices of a scale, fused with the inter-scale flow up-sample)
// and by the fused iteration kernel (blur -> solve -> UpdateMatrices in one launch).
// All f32, uncontracted (-fmad=false), in the upstream order of operations.
//
// Issue slots: k_um0 runs at 72 % issue utilisation and UpdateMatrices is two thirds of k_iter's instructions, so the
// arithmetic is written with Blackwell's packed f32x2 instructions wherever that cannot change a bit:
//   * a packed MULTIPLY (FMUL2) is two IEEE products -- identical to two FMULs; the float4 layout of R puts channel pairs in
//     aligned register pairs already, and the second operand may be one register broadcast to both lanes;
//   * a packed ADD (FADD2) is used only where neither operand is a product: ptxas contracts mul.rn.f32x2 feeding
//     add.rn.f32x2 into FFMA2 even under .rn and -fmad=false (seen in SASS), which would round once instead of twice.
//     A scalar FADD fed by a packed product is left alone (also checked in SASS: tools/check_sass.sh greps for FFMA2).
// 51 floating-point instructions per interior pixel instead of 77; results bit-identical (stage tests against the C oracle).
#pragma once
#include "common.cuh"

namespace ofb {

struct M5 { float v[5]; };

// ---- packed f32x2 helpers ----
// (Carrying a pair as one 64-bit value between the packed instructions was tried: ptxas then needs MORE registers -- 62 instead
// of 56 in k_um0, spills in k_iter64 -- for the same instruction count.)
__device__ __forceinline__ float2 mul2(float2 a, float2 b)                 // (a.x*b.x, a.y*b.y)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2s(float2 a, float s)                 // (a.x*s, a.y*s): FMUL2 with a broadcast operand
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
// packed add / subtract: ONLY for operands that are not products (see the header)
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

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

// Per-CTA view of the two R slots of a pair: only the slot bases.  Channel 4 (xy) of a slot starts plane4 = 4 * plane floats
// behind its float4 part (SlotRing::slot), so every address of a pixel is a slot base plus a 32-bit element offset -- one
// IMAD.WIDE each, no 64-bit adds (a slot is far below 2^32 floats).  o0 = y * pitch + x is shared with the caller's M address.
struct UmBase {
    const float4* r0; const float4* r1;
    unsigned plane4;
    int pitch;
};
__device__ __forceinline__ UmBase um_base(const SlotRing& R, int slot0, int z)
{
    const int s0 = R.first(slot0, z), s1 = R.wrap(s0 + 1);
    return UmBase{reinterpret_cast<const float4*>(R.base + (size_t)s0 * R.slot_stride),
                  reinterpret_cast<const float4*>(R.base + (size_t)s1 * R.slot_stride), 4u * (unsigned)R.plane, R.pitch};
}

__device__ __forceinline__ UmLoads um_load(int x, int y, unsigned o0, float dx, float dy, const UmBase& b, int W, int H)
{
    UmLoads L;
    L.x = x; L.y = y; L.dx = dx; L.dy = dy;
    float fx = x + dx, fy = y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    L.fx = fx - x1; L.fy = fy - y1;
    L.q = b.r0[o0];
    L.q4 = reinterpret_cast<const float*>(b.r0)[b.plane4 + o0];
    L.inside = (unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1);
    if (L.inside) {
        const unsigned o1 = (unsigned)y1 * (unsigned)b.pitch + (unsigned)x1;
        const int p = b.pitch;
        const float4* pa = b.r1 + o1;
        const float* pb = reinterpret_cast<const float*>(b.r1) + (b.plane4 + o1);
        L.A00 = pa[0]; L.A01 = pa[1]; L.A10 = pa[p]; L.A11 = pa[p + 1];
        L.B00 = pb[0]; L.B01 = pb[1]; L.B10 = pb[p]; L.B11 = pb[p + 1];
    }
    return L;
}

// INTERIOR = true: the caller knows the pixel is at least 5 px from every border (the border attenuation of A.8 is skipped
// without the per-pixel test)
template <bool INTERIOR = false>
__device__ __forceinline__ M5 um_compute(const UmLoads& L, int W, int H)
{
    const int x = L.x, y = L.y;
    const float dx = L.dx, dy = L.dy, fx = L.fx, fy = L.fy;
    // r2..r6 of SURVEY.md A.8 as the pairs the loads deliver them in: r23 = (r2, r3), r45 = (r4, r5); r6 alone.  Every packed
    // product below pairs (r2,r3), (r4,r5), (dx,dy) or their swaps, so no pair has to be assembled with MOVs.
    float2 r23, r45;
    float r6;
    if (L.inside) {
        // a00 = (1-fx)(1-fy), a01 = fx(1-fy), a10 = (1-fx)fy, a11 = fx fy
        const float2 wx = make_float2(1.f - fx, fx);
        const float2 a0 = mul2s(wx, 1.f - fy), a1 = mul2s(wx, fy);
        // r_c = a00*R1[y1][x1] + a01*R1[y1][x1+1] + a10*R1[y1+1][x1] + a11*R1[y1+1][x1+1], summed left to right
        const float2 p0 = mul2s(make_float2(L.A00.x, L.A00.y), a0.x), p1 = mul2s(make_float2(L.A01.x, L.A01.y), a0.y);
        const float2 p2 = mul2s(make_float2(L.A10.x, L.A10.y), a1.x), p3 = mul2s(make_float2(L.A11.x, L.A11.y), a1.y);
        const float2 s0 = mul2s(make_float2(L.A00.z, L.A00.w), a0.x), s1 = mul2s(make_float2(L.A01.z, L.A01.w), a0.y);
        const float2 s2 = mul2s(make_float2(L.A10.z, L.A10.w), a1.x), s3 = mul2s(make_float2(L.A11.z, L.A11.w), a1.y);
        const float2 b0 = mul2(make_float2(L.B00, L.B01), a0), b1 = mul2(make_float2(L.B10, L.B11), a1);
        const float r2 = p0.x + p1.x + p2.x + p3.x;
        const float r3 = p0.y + p1.y + p2.y + p3.y;
        const float r4 = s0.x + s1.x + s2.x + s3.x;
        const float r5 = s0.y + s1.y + s2.y + s3.y;
        r6 = (L.q4 + (b0.x + b0.y + b1.x + b1.y)) * 0.25f;
        r45 = mul2s(add2(make_float2(L.q.z, L.q.w), make_float2(r4, r5)), 0.5f);
        r23 = mul2s(sub2(make_float2(L.q.x, L.q.y), make_float2(r2, r3)), 0.5f);
    } else {
        r6 = L.q4 * 0.5f;
        r45 = make_float2(L.q.z, L.q.w);
        r23 = mul2s(sub2(make_float2(L.q.x, L.q.y), make_float2(0.f, 0.f)), 0.5f);
    }
    // r2 += r4*dy + r6*dx;  r3 += r6*dy + r5*dx
    {
        const float2 t = mul2(r45, make_float2(dy, dx));                // (r4*dy, r5*dx)
        const float2 u = mul2s(make_float2(dx, dy), r6);                // (r6*dx, r6*dy)
        r23.x += t.x + u.x;
        r23.y += u.y + t.y;
    }
    if (!INTERIOR && ((unsigned)(x - 5) >= (unsigned)(W - 10) || (unsigned)(y - 5) >= (unsigned)(H - 10))) {
        const float b0 = 0.14f, b2 = 0.4472f;
        float sx0 = x < 5 ? (x < 2 ? b0 : b2) : 1.f;
        float sx1 = x >= W - 5 ? ((W - x - 1) < 2 ? b0 : b2) : 1.f;
        float sy0 = y < 5 ? (y < 2 ? b0 : b2) : 1.f;
        float sy1 = y >= H - 5 ? ((H - y - 1) < 2 ? b0 : b2) : 1.f;
        float scale = sx0 * sx1 * sy0 * sy1;
        r23 = mul2s(r23, scale); r45 = mul2s(r45, scale); r6 *= scale;
    }
    const float2 sq = mul2(r45, r45);                                   // r4*r4, r5*r5
    const float2 d = mul2(r45, r23);                                    // r4*r2, r5*r3
    const float2 g = mul2s(r23, r6);                                    // r6*r2, r6*r3
    const float r66 = r6 * r6;
    M5 m;
    m.v[0] = sq.x + r66;                                                // r4*r4 + r6*r6
    m.v[1] = (r45.x + r45.y) * r6;                                      // (r4 + r5) * r6
    m.v[2] = sq.y + r66;                                                // r5*r5 + r6*r6
    m.v[3] = d.x + g.y;                                                 // r4*r2 + r6*r3
    m.v[4] = g.x + d.y;                                                 // r6*r2 + r5*r3
    return m;
}

__device__ __forceinline__ M5 um_pixel(int x, int y, float dx, float dy, const RView& R0, const RView& R1, int W, int H)
{
    UmLoads L = um_load(x, y, dx, dy, R0, R1, W, H);
    return um_compute(L, W, H);
}
template <bool INTERIOR = false>
__device__ __forceinline__ M5 um_pixel(int x, int y, unsigned o0, float dx, float dy, const UmBase& b, int W, int H)
{
    UmLoads L = um_load(x, y, o0, dx, dy, b, W, H);
    return um_compute<INTERIOR>(L, W, H);
}

}  // namespace ofb
