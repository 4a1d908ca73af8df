// polyexp.cu -- FarnebackPolyExp as cv2 computes it (call sites
// /root/reference/optical_flow.py:51, visualize_optical_flow.py:38).
//
// Per pixel: a (2n+1)^2 Gaussian-weighted least-squares fit of a quadratic, done separably:
//   vertical pass   (f32)  r0 = sum g[k] I(y+k), r1 = sum xg[k] I(y+k), r2 = sum xxg[k] I(y+k)
//   horizontal pass (f64 accumulators over the f32 rows) -> b1..b6 -> five coefficients.
// The float/double mix of every expression below is the one of the upstream C++ (float*float
// products stay float, `double tg = a + b` adds in float first) -- the file is compiled with
// -fmad=false so nothing is contracted.  Border: replicate in both directions.
//
// k_polyexp_tiled: one CTA = TW x TH output pixels; the (TH+2n) x (TW+2n) input patch is staged in
// shared memory once (replicate-clamped on load, so the border costs nothing later), the vertical
// pass writes three shared arrays, the horizontal pass reads them.  Output is 5 planar planes.
// Roofline: reads 4 B/px (+halo from L2), writes 20 B/px -> HBM-bound; algorithmic bytes 24 B/px.
#include "common.cuh"
#include "launch.cuh"

namespace ofb {

struct PolyDev { const float* g; const float* xg; const float* xxg; int n; double ig11, ig03, ig33, ig55; };

__device__ __forceinline__ void polyexp_store(Planes5 R, size_t o, double b1, double b2, double b3, double b4,
                                              double b5, double b6, const PolyDev& pc)
{
    R.ch(1)[o] = (float)(b2 * pc.ig11);
    R.ch(0)[o] = (float)(b3 * pc.ig11);
    R.ch(3)[o] = (float)(b1 * pc.ig03 + b4 * pc.ig33);
    R.ch(2)[o] = (float)(b1 * pc.ig03 + b5 * pc.ig33);
    R.ch(4)[o] = (float)(b6 * pc.ig55);
}

template <int TW, int TH>
__global__ void __launch_bounds__(256)
k_polyexp_tiled(const float* __restrict__ I, int W, int H, int pitch, PolyDev pc, Planes5 R)
{
    extern __shared__ float sm[];
    const int n = pc.n, PW = TW + 2 * n, PH = TH + 2 * n;
    float* sg = sm;                     // g[0..n]
    float* sxg = sg + (n + 1);          // xg[0..n]
    float* sxxg = sxg + (n + 1);        // xxg[0..n]
    float* sI = sxxg + (n + 1);         // PH x PW
    float* sR0 = sI + PH * PW;          // TH x PW each
    float* sR1 = sR0 + TH * PW;
    float* sR2 = sR1 + TH * PW;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

    for (int i = tid; i <= n; i += NT) { sg[i] = pc.g[n + i]; sxg[i] = pc.xg[n + i]; sxxg[i] = pc.xxg[n + i]; }
    for (int i = tid; i < PH * PW; i += NT) {
        int py = i / PW, px = i - py * PW;
        int gy = min(max(y0 + py - n, 0), H - 1), gx = min(max(x0 + px - n, 0), W - 1);
        sI[i] = I[(size_t)gy * pitch + gx];
    }
    __syncthreads();

    for (int i = tid; i < TH * PW; i += NT) {
        int ty = i / PW, px = i - ty * PW;
        const float* col = sI + (ty + n) * PW + px;
        float r0 = col[0] * sg[0], r1 = 0.f, r2 = 0.f;
        for (int k = 1; k <= n; k++) {
            float a = col[-k * PW], b = col[k * PW];
            float p = a + b;
            r0 = r0 + sg[k] * p;
            r1 = r1 + sxg[k] * (b - a);
            r2 = r2 + sxxg[k] * p;
        }
        sR0[i] = r0; sR1[i] = r1; sR2[i] = r2;
    }
    __syncthreads();

    for (int i = tid; i < TH * TW; i += NT) {
        int ly = i / TW, lx = i - ly * TW;
        int gx = x0 + lx, gy = y0 + ly;
        if (gx >= W || gy >= H) continue;
        int base = ly * PW + lx + n;
        float g0 = sg[0];
        double b1 = sR0[base] * g0, b2 = 0, b3 = sR1[base] * g0, b4 = 0, b5 = sR2[base] * g0, b6 = 0;
        for (int k = 1; k <= n; k++) {
            float r0p = sR0[base + k], r0m = sR0[base - k];
            float r1p = sR1[base + k], r1m = sR1[base - k];
            float r2p = sR2[base + k], r2m = sR2[base - k];
            float gk = sg[k], xgk = sxg[k], xxgk = sxxg[k];
            double tg = r0p + r0m;
            b1 += tg * gk; b4 += tg * xxgk;
            b2 += (r0p - r0m) * xgk;
            b3 += (r1p + r1m) * gk;
            b6 += (r1p - r1m) * xgk;
            b5 += (r2p + r2m) * gk;
        }
        polyexp_store(R, (size_t)gy * R.pitch + gx, b1, b2, b3, b4, b5, b6, pc);
    }
}

// Generic fallback (any n): two global-memory passes through three temporary planes.
__global__ void __launch_bounds__(256)
k_polyexp_v_generic(const float* __restrict__ I, int W, int H, int pitch, PolyDev pc, float* __restrict__ tmp3, size_t plane)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int n = pc.n;
    float r0 = I[(size_t)y * pitch + x] * pc.g[n], r1 = 0.f, r2 = 0.f;
    for (int k = 1; k <= n; k++) {
        float a = I[(size_t)max(y - k, 0) * pitch + x], b = I[(size_t)min(y + k, H - 1) * pitch + x];
        float p = a + b;
        r0 = r0 + pc.g[n + k] * p;
        r1 = r1 + pc.xg[n + k] * (b - a);
        r2 = r2 + pc.xxg[n + k] * p;
    }
    size_t o = (size_t)y * pitch + x;
    tmp3[o] = r0; tmp3[plane + o] = r1; tmp3[2 * plane + o] = r2;
}

__global__ void __launch_bounds__(256)
k_polyexp_h_generic(const float* __restrict__ tmp3, size_t plane, int W, int H, int pitch, PolyDev pc, Planes5 R)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int n = pc.n;
    const float* q0 = tmp3 + (size_t)y * pitch;
    const float* q1 = q0 + plane;
    const float* q2 = q1 + plane;
    float g0 = pc.g[n];
    double b1 = q0[x] * g0, b2 = 0, b3 = q1[x] * g0, b4 = 0, b5 = q2[x] * g0, b6 = 0;
    for (int k = 1; k <= n; k++) {
        int xp = min(x + k, W - 1), xm = max(x - k, 0);
        float r0p = q0[xp], r0m = q0[xm], r1p = q1[xp], r1m = q1[xm], r2p = q2[xp], r2m = q2[xm];
        float gk = pc.g[n + k], xgk = pc.xg[n + k], xxgk = pc.xxg[n + k];
        double tg = r0p + r0m;
        b1 += tg * gk; b4 += tg * xxgk;
        b2 += (r0p - r0m) * xgk;
        b3 += (r1p + r1m) * gk;
        b6 += (r1p - r1m) * xgk;
        b5 += (r2p + r2m) * gk;
    }
    polyexp_store(R, (size_t)y * R.pitch + x, b1, b2, b3, b4, b5, b6, pc);
}

void launch_polyexp(Launch& L, const float* I, int W, int H, int pitch, const PolyConst& c,
                    float* tmp3, Planes5 R, bool generic)
{
    PolyDev pc{c.g, c.xg, c.xxg, c.n, c.ig11, c.ig03, c.ig33, c.ig55};
    constexpr int TW = 64, TH = 16;
    const int n = c.n;
    size_t smem = sizeof(float) * (3 * (n + 1) + (size_t)(TH + 2 * n) * (TW + 2 * n) + 3 * (size_t)TH * (TW + 2 * n));
    if (!generic && smem <= 48 * 1024) {
        dim3 grid(divup(W, TW), divup(H, TH));
        L.run("polyexp_tiled", [&](cudaStream_t s) {
            k_polyexp_tiled<TW, TH><<<grid, 256, smem, s>>>(I, W, H, pitch, pc, R);
        });
    } else {
        dim3 block(64, 4), grid(divup(W, 64), divup(H, 4));
        size_t plane = (size_t)H * pitch;
        L.run("polyexp_v_generic", [&](cudaStream_t s) {
            k_polyexp_v_generic<<<grid, block, 0, s>>>(I, W, H, pitch, pc, tmp3, plane);
        });
        L.run("polyexp_h_generic", [&](cudaStream_t s) {
            k_polyexp_h_generic<<<grid, block, 0, s>>>(tmp3, plane, W, H, pitch, pc, R);
        });
    }
}

}  // namespace ofb
