// viz.cu -- the flow -> picture / feature tail of the reference, fused:
//     mag, ang = cv2.cartToPolar(flow[...,0], flow[...,1])          optical_flow.py:61, visualize_optical_flow.py:48
//     hsv[...,0] = ang*180/np.pi   (f32, truncated, wrapped mod 256 by the uint8 store)     :53
//     hsv[...,1] = 255                                                                      :52
//     hsv[...,2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)  (truncated by the uint8 store)   :54
//     bgr = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)                                            :55
//     np.sum(mag)                                                   optical_flow.py:64
// Two launches per picture: k_minmax_mag (per-frame min/max of the magnitude: warp-shuffle
// reduction, one atomic pair per CTA) and k_flow_to_bgr (recomputes mag/angle, quantises, converts).
// Every expression is written so that it rounds exactly like cv2 4.13 / NumPy do (bit-exact hue,
// value, magnitude and angle on cv2's own flow; HSV->BGR follows cv2's vector body, which truncates).
// The file is compiled with -fmad=false; the FMAs that cv2 uses are explicit fmaf calls.
// Roofline: HBM; algorithmic bytes 19 B/px (flow read twice, 3 B written).
#include "common.cuh"
#include "launch.cuh"
#include <float.h>

namespace ofb {

struct Polar { float mag, ang_deg; };

// cv::cartToPolar's kernel: magnitude = sqrt(fma(x,x,y*y)); angle = fastAtan2 (7th-order odd polynomial, degrees)
__device__ __forceinline__ Polar polar_of(float x, float y)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    Polar o;
    o.mag = sqrtf(fmaf(x, x, y * y));
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float c = mn / (mx + (float)DBL_EPSILON);
    float c2 = c * c;
    float a = fmaf(fmaf(fmaf(c2, p7, p5), c2, p3), c2, p1) * c;
    if (ax < ay) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    o.ang_deg = a;
    return o;
}

__device__ __forceinline__ float deg_to_rad_cv(float a) { return a * (float)(3.14159265358979323846 / 180.0); }

// ---- per-frame min / max of the magnitude ------------------------------------------------------
// Magnitudes are >= 0, so their IEEE bit patterns order like unsigned integers.
__global__ void k_minmax_reset(unsigned* mm) { mm[0] = 0x7f800000u; mm[1] = 0u; }
__global__ void k_minmax_reset_batch(unsigned* mm, int batch)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) { mm[2 * i] = 0x7f800000u; mm[2 * i + 1] = 0u; }
}

__global__ void __launch_bounds__(256)
k_minmax_mag(const float2* __restrict__ flow, size_t n, unsigned* __restrict__ mm, size_t flow_item = 0)
{
    flow += (size_t)blockIdx.y * flow_item; mm += 2 * blockIdx.y;
    float lo = __int_as_float(0x7f800000), hi = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 f = flow[i];
        float m = sqrtf(fmaf(f.x, f.x, f.y * f.y));
        lo = fminf(lo, m); hi = fmaxf(hi, m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float slo[8], shi[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { slo[w] = lo; shi[w] = hi; }
    __syncthreads();
    if (w == 0) {
        lo = l < (blockDim.x >> 5) ? slo[l] : __int_as_float(0x7f800000);
        hi = l < (blockDim.x >> 5) ? shi[l] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (l == 0) { atomicMin(mm, __float_as_uint(lo)); atomicMax(mm + 1, __float_as_uint(hi)); }
    }
}

// ---- quantise + HSV -> BGR ----------------------------------------------------------------------
struct NormCoef { float fs, fsh; };
// cv::normalize(NORM_MINMAX, dtype f32): scale rounded to f32 before the shift is formed.
__device__ __forceinline__ NormCoef norm_coef(const unsigned* mm)
{
    double smin = (double)__uint_as_float(mm[0]), smax = (double)__uint_as_float(mm[1]);
    double scale = (255.0 - 0.0) * ((smax - smin) > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0);
    NormCoef c;
    c.fs = (float)scale;
    c.fsh = 0.f - (float)(smin * (double)c.fs);
    return c;
}

// f32 flow -> the two bytes NumPy stores into hsv[...,0] and hsv[...,2] (visualize_optical_flow.py:53-54)
__device__ __forceinline__ void quantise_hv(float2 f, NormCoef nc, int& hq, int& vq)
{
    Polar p = polar_of(f.x, f.y);
    float ang = deg_to_rad_cv(p.ang_deg);
    float hue = (ang * 180.f) / (float)3.14159265358979323846;   // NumPy: f32 * 180 then / f32(pi)
    hq = ((int)hue) & 255;
    float nv = fmaf(p.mag, nc.fs, nc.fsh);
    vq = ((int)nv) & 255;
}

// cv::cvtColor(COLOR_HSV2BGR), 8-bit, S = 255, vector-body rounding (truncate): (H byte, V byte) -> B | G<<8 | R<<16
__device__ __forceinline__ unsigned hsv_s255_to_bgr(int hq, int vq)
{
    float h = (float)hq * (6.f / 180.f);
    const float s = 255.f * (1.f / 255.f);
    float v = (float)vq * (1.f / 255.f);
    while (h >= 6.f) h -= 6.f;
    int sector = (int)floorf(h);
    h -= (float)sector;
    if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
    float t0 = v, t1 = v * (1.f - s), t2 = v * (1.f - s * h), t3 = v * (1.f - s * (1.f - h));
    float b, g, r;
    switch (sector) {
        case 0: b = t1; g = t3; r = t0; break;
        case 1: b = t1; g = t0; r = t2; break;
        case 2: b = t3; g = t0; r = t1; break;
        case 3: b = t0; g = t2; r = t1; break;
        case 4: b = t0; g = t1; r = t3; break;
        default: b = t2; g = t1; r = t0; break;
    }
    int bi = (int)(b * 255.f), gi = (int)(g * 255.f), ri = (int)(r * 255.f);
    return (unsigned)min(max(bi, 0), 255) | ((unsigned)min(max(gi, 0), 255) << 8) | ((unsigned)min(max(ri, 0), 255) << 16);
}

// The colour conversion has a 16-bit domain, so the batched picture kernel looks it up instead of spending ~8 of its
// ~11 XU-pipe operations per pixel (I2F / F2I / floor) on it: table[h << 8 | v] is filled once per context by the SAME
// function, so table and arithmetic agree by construction (and both are pinned by tests/golden/hsv2bgr_table.npz).
__global__ void k_build_hsv_table(unsigned* __restrict__ table) { table[blockIdx.x * 256 + threadIdx.x] = hsv_s255_to_bgr(blockIdx.x, threadIdx.x); }

void launch_build_hsv_table(Launch& L, unsigned* table)
{
    L.run("hsv_table", [&](cudaStream_t s) { k_build_hsv_table<<<256, 256, 0, s>>>(table); });
}

template <bool LUT>
__device__ __forceinline__ unsigned pixel_bgr(float2 f, NormCoef nc, const unsigned* __restrict__ table)
{
    int hq, vq;
    quantise_hv(f, nc, hq, vq);
    return LUT ? __ldg(table + ((hq << 8) | vq)) : hsv_s255_to_bgr(hq, vq);
}

// 4 pixels per thread: two 16-byte flow loads, three 4-byte picture stores.
template <bool LUT>
__global__ void __launch_bounds__(256)
k_flow_to_bgr_v4(const float4* __restrict__ flow4, size_t n4, const unsigned* __restrict__ mm, uint32_t* __restrict__ bgr,
                 const unsigned* __restrict__ table, size_t flow_item4 = 0, size_t bgr_item4 = 0)
{
    flow4 += (size_t)blockIdx.y * flow_item4; bgr += (size_t)blockIdx.y * bgr_item4; mm += 2 * blockIdx.y;
    NormCoef nc = norm_coef(mm);
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (size_t)gridDim.x * blockDim.x) {
        float4 a = flow4[2 * g], b = flow4[2 * g + 1];
        const unsigned p0 = pixel_bgr<LUT>(make_float2(a.x, a.y), nc, table), p1 = pixel_bgr<LUT>(make_float2(a.z, a.w), nc, table);
        const unsigned p2 = pixel_bgr<LUT>(make_float2(b.x, b.y), nc, table), p3 = pixel_bgr<LUT>(make_float2(b.z, b.w), nc, table);
        bgr[3 * g] = p0 | (p1 << 24);
        bgr[3 * g + 1] = (p1 >> 8) | (p2 << 16);
        bgr[3 * g + 2] = (p2 >> 16) | (p3 << 8);
    }
}

__global__ void __launch_bounds__(256)
k_flow_to_bgr_scalar(const float2* __restrict__ flow, size_t n, const unsigned* __restrict__ mm, uint8_t* __restrict__ bgr,
                     size_t flow_item = 0, size_t bgr_item = 0)
{
    flow += (size_t)blockIdx.y * flow_item; bgr += (size_t)blockIdx.y * bgr_item; mm += 2 * blockIdx.y;
    NormCoef nc = norm_coef(mm);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned p = pixel_bgr<false>(flow[i], nc, nullptr);
        bgr[3 * i] = (uint8_t)p; bgr[3 * i + 1] = (uint8_t)(p >> 8); bgr[3 * i + 2] = (uint8_t)(p >> 16);
    }
}

__global__ void __launch_bounds__(256)
k_cart_to_polar(const float2* __restrict__ flow, size_t n, float* __restrict__ mag, float* __restrict__ ang, int degrees)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 f = flow[i];
        Polar p = polar_of(f.x, f.y);
        mag[i] = p.mag;
        ang[i] = degrees ? p.ang_deg : deg_to_rad_cv(p.ang_deg);      // cv2: fastAtan2 gives degrees; radians = degrees * (pi/180)f
    }
}

// ---- np.sum(mag): f64 accumulation (NumPy's pairwise f32 sum agrees to ~1e-7 relative) ----------
__global__ void __launch_bounds__(256)
k_sum_magnitude(const float2* __restrict__ flow, size_t n, double* __restrict__ acc, size_t flow_item = 0)
{
    flow += (size_t)blockIdx.y * flow_item; acc += blockIdx.y;
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 f = flow[i];
        s += (double)sqrtf(fmaf(f.x, f.x, f.y * f.y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ss[8];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) ss[w] = s;
    __syncthreads();
    if (w == 0) {
        s = l < (blockDim.x >> 5) ? ss[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (l == 0) atomicAdd(acc, s);
    }
}
__global__ void k_sum_finish(const double* acc, float* out) { *out = (float)*acc; }
__global__ void k_zero_double(double* p) { *p = 0.0; }
__global__ void k_sum_finish_batch(const double* acc, float* out, int batch)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) out[i] = (float)acc[i];
}
__global__ void k_zero_double_batch(double* p, int batch)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) p[i] = 0.0;
}

static inline unsigned reduce_grid(size_t n, int per_thread)
{
    size_t b = (n + 256 * (size_t)per_thread - 1) / (256 * (size_t)per_thread);
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;
    return (unsigned)b;
}

void launch_minmax_reset(Launch& L, unsigned* minmax)
{
    L.run("minmax_reset", [&](cudaStream_t s) { k_minmax_reset<<<1, 1, 0, s>>>(minmax); });
}

void launch_minmax_mag(Launch& L, const float2* flow, size_t n, unsigned* minmax)
{
    L.run("minmax_mag", [&](cudaStream_t s) { k_minmax_mag<<<reduce_grid(n, 8), 256, 0, s>>>(flow, n, minmax); });
}

void launch_flow_to_bgr(Launch& L, const float2* flow, size_t n, const unsigned* minmax, uint8_t* bgr, const unsigned* table)
{
    bool vec = (n % 4 == 0) && ((uintptr_t)flow % 16 == 0) && ((uintptr_t)bgr % 4 == 0);
    if (vec) {
        size_t n4 = n / 4;
        L.run("flow_to_bgr_v4", [&](cudaStream_t s) {
            if (table) k_flow_to_bgr_v4<true><<<reduce_grid(n4, 2), 256, 0, s>>>((const float4*)flow, n4, minmax, (uint32_t*)bgr, table);
            else k_flow_to_bgr_v4<false><<<reduce_grid(n4, 2), 256, 0, s>>>((const float4*)flow, n4, minmax, (uint32_t*)bgr, nullptr);
        });
    } else {
        L.run("flow_to_bgr_scalar", [&](cudaStream_t s) {
            k_flow_to_bgr_scalar<<<reduce_grid(n, 4), 256, 0, s>>>(flow, n, minmax, bgr);
        });
    }
}

void launch_minmax_reset_batch(Launch& L, unsigned* minmax, int batch)
{
    L.run("minmax_reset", [&](cudaStream_t s) { k_minmax_reset_batch<<<divup(batch, 64), 64, 0, s>>>(minmax, batch); });
}

void launch_picture_batch(Launch& L, const float2* flow, size_t flow_item, size_t n, unsigned* minmax,
                          uint8_t* bgr, size_t bgr_item, int batch, bool minmax_done, const unsigned* table)
{
    if (!minmax_done) {
        launch_minmax_reset_batch(L, minmax, batch);
        dim3 g1(reduce_grid(n, 8), batch);
        L.run("minmax_mag", [&](cudaStream_t s) { k_minmax_mag<<<g1, 256, 0, s>>>(flow, n, minmax, flow_item); });
    }
    bool vec = (n % 4 == 0) && ((uintptr_t)flow % 16 == 0) && ((uintptr_t)bgr % 4 == 0) && (flow_item % 2 == 0) && (bgr_item % 4 == 0);
    if (vec) {
        size_t n4 = n / 4;
        dim3 g2(reduce_grid(n4, 2), batch);
        L.run("flow_to_bgr_v4", [&](cudaStream_t s) {
            if (table) k_flow_to_bgr_v4<true><<<g2, 256, 0, s>>>((const float4*)flow, n4, minmax, (uint32_t*)bgr, table, flow_item / 2, bgr_item / 4);
            else k_flow_to_bgr_v4<false><<<g2, 256, 0, s>>>((const float4*)flow, n4, minmax, (uint32_t*)bgr, nullptr, flow_item / 2, bgr_item / 4);
        });
    } else {
        dim3 g2(reduce_grid(n, 4), batch);
        L.run("flow_to_bgr_scalar", [&](cudaStream_t s) {
            k_flow_to_bgr_scalar<<<g2, 256, 0, s>>>(flow, n, minmax, bgr, flow_item, bgr_item);
        });
    }
}

void launch_sum_magnitude_batch(Launch& L, const float2* flow, size_t flow_item, size_t n, double* acc, float* out, int batch)
{
    L.run("sum_zero", [&](cudaStream_t s) { k_zero_double_batch<<<divup(batch, 64), 64, 0, s>>>(acc, batch); });
    dim3 g(reduce_grid(n, 8), batch);
    L.run("sum_magnitude", [&](cudaStream_t s) { k_sum_magnitude<<<g, 256, 0, s>>>(flow, n, acc, flow_item); });
    L.run("sum_finish", [&](cudaStream_t s) { k_sum_finish_batch<<<divup(batch, 64), 64, 0, s>>>(acc, out, batch); });
}

void launch_cart_to_polar(Launch& L, const float2* flow, size_t n, float* mag, float* ang, bool degrees)
{
    L.run("cart_to_polar", [&](cudaStream_t s) { k_cart_to_polar<<<reduce_grid(n, 4), 256, 0, s>>>(flow, n, mag, ang, degrees ? 1 : 0); });
}

void launch_sum_magnitude(Launch& L, const float2* flow, size_t n, double* acc, float* out)
{
    L.run("sum_zero", [&](cudaStream_t s) { k_zero_double<<<1, 1, 0, s>>>(acc); });
    L.run("sum_magnitude", [&](cudaStream_t s) { k_sum_magnitude<<<reduce_grid(n, 8), 256, 0, s>>>(flow, n, acc); });
    L.run("sum_finish", [&](cudaStream_t s) { k_sum_finish<<<1, 1, 0, s>>>(acc, out); });
}

}  // namespace ofb
