// matrices.cu -- inter-scale flow initialisation and FarnebackUpdateMatrices (A.8),
// the parts of cv2.calcOpticalFlowFarneback (/root/reference/optical_flow.py:51,
// visualize_optical_flow.py:38) that warp the second frame's coefficients by the current flow.
//
// k_update_matrices is element-wise plus one bilinear gather: per pixel it reads R0 (5 floats),
// R1 at 4 neighbours x 5 channels (planar: each of the 20 loads is coalesced across the warp as long
// as the flow is smooth), the flow (8 B) and writes M (5 floats).  All f32, uncontracted, in the
// upstream order of operations.  Roofline: HBM; algorithmic bytes 68 B/px (20+20+8 read, 20 written).
#include "common.cuh"
#include "launch.cuh"
#include "um_device.cuh"

namespace ofb {

// A.2: resize(prevFlow, INTER_LINEAR) then flow *= 1/pyr_scale
__global__ void __launch_bounds__(256)
k_upsample_flow(const float2* __restrict__ prev, int Wp, int Hp, float2* __restrict__ flow, int W, int H,
                double sx_scale, double sy_scale, float mul)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    float a1, b1;
    int sx = linear_coord(x, sx_scale, Wp, &a1, true);
    int sy = linear_coord(y, sy_scale, Hp, &b1, true);
    float a0 = 1.f - a1, b0 = 1.f - b1;
    int sx1 = min(sx + 1, Wp - 1), sy1 = min(sy + 1, Hp - 1);
    float2 p00 = prev[(size_t)sy * Wp + sx], p01 = prev[(size_t)sy * Wp + sx1];
    float2 p10 = prev[(size_t)sy1 * Wp + sx], p11 = prev[(size_t)sy1 * Wp + sx1];
    float hx0 = p00.x * a0 + p01.x * a1, hx1 = p10.x * a0 + p11.x * a1;
    float hy0 = p00.y * a0 + p01.y * a1, hy1 = p10.y * a0 + p11.y * a1;
    float2 o;
    o.x = (hx0 * b0 + hx1 * b1) * mul;
    o.y = (hy0 * b0 + hy1 * b1) * mul;
    flow[(size_t)y * W + x] = o;
}

// A.2 with OPTFLOW_USE_INITIAL_FLOW: resize(flow0, INTER_AREA) * scale.  Same accumulation order as
// cv::resize's area path: per source row a horizontal weighted sum, rows combined with beta.
struct AreaSpan { int s_first; int n; float a_first, a_mid, a_last; bool has_first, has_last; int s_mid0, n_mid, s_last; };
__device__ inline AreaSpan area_span(int d, int ssize, double scale)
{
    AreaSpan sp;
    double fsx1 = d * scale, fsx2 = fsx1 + scale;
    double cell = fmin(scale, ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = min(sx2, ssize - 1); sx1 = min(sx1, sx2);
    sp.has_first = (sx1 - fsx1 > 1e-3);
    sp.s_first = sx1 - 1; sp.a_first = (float)((sx1 - fsx1) / cell);
    sp.s_mid0 = sx1; sp.n_mid = sx2 - sx1; sp.a_mid = (float)(1.0 / cell);
    sp.has_last = (fsx2 - sx2 > 1e-3);
    sp.s_last = sx2; sp.a_last = (float)(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
    sp.n = 0;
    return sp;
}
__device__ inline float2 area_row(const float2* row, const AreaSpan& sp)
{
    float2 acc = make_float2(0.f, 0.f);
    if (sp.has_first) { float2 v = row[sp.s_first]; acc.x += v.x * sp.a_first; acc.y += v.y * sp.a_first; }
    for (int i = 0; i < sp.n_mid; i++) { float2 v = row[sp.s_mid0 + i]; acc.x += v.x * sp.a_mid; acc.y += v.y * sp.a_mid; }
    if (sp.has_last) { float2 v = row[sp.s_last]; acc.x += v.x * sp.a_last; acc.y += v.y * sp.a_last; }
    return acc;
}
__global__ void __launch_bounds__(256)
k_area_flow(const float2* __restrict__ src, int Ws, int Hs, float2* __restrict__ dst, int Wd, int Hd,
            double sx, double sy, int isx, int isy, int integer_ratio, float mul)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Wd || y >= Hd) return;
    float2 o;
    if (integer_ratio) {
        float sc = 1.f / (isx * isy);
        float ax = 0.f, ay = 0.f;
        for (int j = 0; j < isy; j++)
            for (int i = 0; i < isx; i++) {
                float2 v = src[(size_t)(y * isy + j) * Ws + (x * isx + i)];
                ax += v.x; ay += v.y;
            }
        o.x = ax * sc; o.y = ay * sc;
    } else {
        AreaSpan cx = area_span(x, Ws, sx), cy = area_span(y, Hs, sy);
        bool first = true;
        float2 sum = make_float2(0.f, 0.f);
        auto add_row = [&](int srow, float beta) {
            float2 b = area_row(src + (size_t)srow * Ws, cx);
            if (first) { sum.x = beta * b.x; sum.y = beta * b.y; first = false; }
            else { sum.x += beta * b.x; sum.y += beta * b.y; }
        };
        if (cy.has_first) add_row(cy.s_first, cy.a_first);
        for (int j = 0; j < cy.n_mid; j++) add_row(cy.s_mid0 + j, cy.a_mid);
        if (cy.has_last) add_row(cy.s_last, cy.a_last);
        o = sum;
    }
    o.x *= mul; o.y *= mul;
    dst[(size_t)y * Wd + x] = o;
}

__global__ void k_scale_flow(float2* flow, size_t n, float mul)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float2 v = flow[i]; v.x *= mul; v.y *= mul; flow[i] = v; }
}

// A.8 (per-pixel body in um_device.cuh)
__global__ void __launch_bounds__(256)
k_update_matrices(RView R0, RView R1, const float2* __restrict__ flow, int W, int H, Planes5 M)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float2 d = flow[(size_t)y * W + x];
    M5 m = um_pixel(x, y, d.x, d.y, R0, R1, W, H);
    const size_t om = (size_t)y * M.pitch + x;
#pragma unroll
    for (int c = 0; c < 5; c++) M.ch(c)[om] = m.v[c];
}

__global__ void k_r_interleave(RView src, int W, int H, float* __restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    size_t o = (size_t)y * src.pitch + x;
    float4 a = src.a[o];
    float* d = dst + ((size_t)y * W + x) * 5;
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = src.b[o];
}
__global__ void k_r_deinterleave(const float* __restrict__ src, int W, int H, RView dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* s = src + ((size_t)y * W + x) * 5;
    size_t o = (size_t)y * dst.pitch + x;
    dst.a[o] = make_float4(s[0], s[1], s[2], s[3]);
    dst.b[o] = s[4];
}

__global__ void k_interleave5(Planes5 src, int W, int H, float* __restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    for (int c = 0; c < 5; c++) dst[((size_t)y * W + x) * 5 + c] = src.ch(c)[(size_t)y * src.pitch + x];
}
__global__ void k_deinterleave5(const float* __restrict__ src, int W, int H, Planes5 dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    for (int c = 0; c < 5; c++) dst.ch(c)[(size_t)y * dst.pitch + x] = src[((size_t)y * W + x) * 5 + c];
}

static inline dim3 grid2d(int W, int H, dim3 b) { return dim3(divup(W, b.x), divup(H, b.y)); }

void launch_upsample_flow(Launch& L, const float2* prev, int Wp, int Hp, float2* flow, int W, int H, float mul)
{
    dim3 b(64, 4);
    double sx = 1.0 / ((double)W / Wp), sy = 1.0 / ((double)H / Hp);
    L.run("upsample_flow", [&](cudaStream_t s) {
        k_upsample_flow<<<grid2d(W, H, b), b, 0, s>>>(prev, Wp, Hp, flow, W, H, sx, sy, mul);
    });
}

void launch_area_flow(Launch& L, const float2* src, int Ws, int Hs, float2* dst, int Wd, int Hd, float mul)
{
    dim3 b(64, 4);
    double sx = (double)Ws / Wd, sy = (double)Hs / Hd;
    int isx = (int)floor(sx + 0.5), isy = (int)floor(sy + 0.5);
    int integer_ratio = (fabs(sx - isx) < 2.220446049250313e-16 && fabs(sy - isy) < 2.220446049250313e-16) ? 1 : 0;
    L.run("area_flow", [&](cudaStream_t s) {
        k_area_flow<<<grid2d(Wd, Hd, b), b, 0, s>>>(src, Ws, Hs, dst, Wd, Hd, sx, sy, isx, isy, integer_ratio, mul);
    });
}

void launch_scale_flow(Launch& L, float2* flow, size_t n, float mul)
{
    L.run("scale_flow", [&](cudaStream_t s) {
        k_scale_flow<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flow, n, mul);
    });
}

void launch_update_matrices(Launch& L, RView R0, RView R1, const float2* flow, int W, int H, Planes5 M)
{
    dim3 b(64, 4);
    L.run("update_matrices", [&](cudaStream_t s) {
        k_update_matrices<<<grid2d(W, H, b), b, 0, s>>>(R0, R1, flow, W, H, M);
    });
}

void launch_r_interleave(Launch& L, RView src, int W, int H, float* dst)
{
    dim3 b(64, 4);
    L.run("r_interleave", [&](cudaStream_t s) { k_r_interleave<<<grid2d(W, H, b), b, 0, s>>>(src, W, H, dst); });
}
void launch_r_deinterleave(Launch& L, const float* src, int W, int H, RView dst)
{
    dim3 b(64, 4);
    L.run("r_deinterleave", [&](cudaStream_t s) { k_r_deinterleave<<<grid2d(W, H, b), b, 0, s>>>(src, W, H, dst); });
}

void launch_interleave5(Launch& L, Planes5 src, int W, int H, float* dst)
{
    dim3 b(64, 4);
    L.run("interleave5", [&](cudaStream_t s) { k_interleave5<<<grid2d(W, H, b), b, 0, s>>>(src, W, H, dst); });
}
void launch_deinterleave5(Launch& L, const float* src, int W, int H, Planes5 dst)
{
    dim3 b(64, 4);
    L.run("deinterleave5", [&](cudaStream_t s) { k_deinterleave5<<<grid2d(W, H, b), b, 0, s>>>(src, W, H, dst); });
}

}  // namespace ofb
