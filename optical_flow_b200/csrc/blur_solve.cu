// blur_solve.cu -- FarnebackUpdateFlow_Blur / FarnebackUpdateFlow_GaussianBlur without their trailing
// UpdateMatrices (the iteration body of cv2.calcOpticalFlowFarneback called at
// /root/reference/optical_flow.py:51 and visualize_optical_flow.py:38):
//     B = blur(M) over a (2m+1)^2 window, replicate border, m = winsize/2
//     flow = solve2x2(B)   with  idet = 1/(B0*B2 - B1^2 + 1e-3)   in double
//
// k_box_strip<M>  (box window, the reference's flags=0 path; the dominant kernel of the pipeline)
//   One CTA owns TW = CW-2M output columns (CW = 96 loaded columns incl. halo) and walks DOWN a strip
//   of rows in steps of R = 2M+1 rows:
//     V phase  thread = (channel, column): running vertical window sum in an f64 register; the R
//              rows that enter the window this step are loaded coalesced into registers, the R rows
//              that leave it are last step's registers -> each M element is read from L2/HBM once
//              (plus the horizontal halo), never re-read.  Sums go to shared memory as f32.
//     H phase  thread = (row, channel, segment): horizontal running sum (f64) along a segment of the
//              row; one direct (2M+1)-sum to start, then +new -old.
//     S phase  thread = pixel: 2x2 solve in f64, coalesced float2 store of the flow.
//   cv2 itself keeps running sums in double down the whole column; per-strip restarts differ from
//   that by ~1e-16 relative.  Shared-memory staging is f32 (2^-24 relative, the precision cv2's own
//   Gaussian variant uses for the same quantity).
//   Roofline: HBM.  Algorithmic bytes 28 B/px (20 read M, 8 written flow); fp64 pipe ~35 op/px.
//
// Generic kernels (any winsize, also the Gaussian window): two global-memory passes, direct sums.
#include "common.cuh"
#include "launch.cuh"
#include <algorithm>

namespace ofb {

__device__ __forceinline__ float2 solve_flow(double g11, double g12, double g22, double h1, double h2)
{
    double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
    float2 f;
    f.x = (float)((g11 * h2 - g12 * h1) * idet);
    f.y = (float)((g22 * h1 - g12 * h2) * idet);
    return f;
}

// ------------------------------------------------------------------------------------------------
// generic box: vertical pass (f64 out), horizontal pass + solve
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_box_v_generic(Planes5 M, int W, int H, int m, double* __restrict__ tmp, size_t tplane, int tpitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.z;
    if (x >= W || y >= H) return;
    const float* p = M.ch(c);
    double s = 0;
    for (int i = -m; i <= m; i++) s += (double)p[(size_t)min(max(y + i, 0), H - 1) * M.pitch + x];
    // winsize == 1: cv2 seeds its running sum with row0*(m+2) and then adds row[y+m]-row[y-m-1]; for
    // m == 0 that telescopes to M[y] + M[0], not M[y] (a quirk to reproduce).
    if (m == 0) s += (double)p[x];
    tmp[(size_t)c * tplane + (size_t)y * tpitch + x] = s;
}

__global__ void __launch_bounds__(256)
k_box_h_solve_generic(const double* __restrict__ tmp, size_t tplane, int tpitch, int W, int H, int m, double scale,
                      float2* __restrict__ flow)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    double b[5];
    for (int c = 0; c < 5; c++) {
        const double* row = tmp + (size_t)c * tplane + (size_t)y * tpitch;
        double s = 0;
        for (int i = -m; i <= m; i++) s += row[min(max(x + i, 0), W - 1)];
        if (m == 0) s += row[0];          // same quirk horizontally: vsum[x] + vsum[0]
        b[c] = s * scale;
    }
    flow[(size_t)y * W + x] = solve_flow(b[0], b[1], b[2], b[3], b[4]);
}

// ------------------------------------------------------------------------------------------------
// generic Gaussian: f32 taps and f32 accumulation in cv2's folded order, solve in f64
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gauss_v_generic(Planes5 M, int W, int H, int m, const float* __restrict__ kern, float* __restrict__ tmp,
                  size_t tplane, int tpitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.z;
    if (x >= W || y >= H) return;
    const float* p = M.ch(c);
    float s0 = p[(size_t)y * M.pitch + x] * kern[0];
    for (int i = 1; i <= m; i++) {
        float dn = p[(size_t)min(y + i, H - 1) * M.pitch + x], up = p[(size_t)max(y - i, 0) * M.pitch + x];
        s0 += (dn + up) * kern[i];
    }
    tmp[(size_t)c * tplane + (size_t)y * tpitch + x] = s0;
}

__global__ void __launch_bounds__(256)
k_gauss_h_solve_generic(const float* __restrict__ tmp, size_t tplane, int tpitch, int W, int H, int m,
                        const float* __restrict__ kern, float2* __restrict__ flow)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    double b[5];
    for (int c = 0; c < 5; c++) {
        const float* row = tmp + (size_t)c * tplane + (size_t)y * tpitch;
        float s = row[x] * kern[0];
        for (int i = 1; i <= m; i++) s += kern[i] * (row[max(x - i, 0)] + row[min(x + i, W - 1)]);
        b[c] = s;
    }
    flow[(size_t)y * W + x] = solve_flow(b[0], b[1], b[2], b[3], b[4]);
}

// ------------------------------------------------------------------------------------------------
// k_box_strip<M>
// ------------------------------------------------------------------------------------------------
constexpr int BS_CW = 96;              // columns loaded per CTA (3 warps per channel)
constexpr int BS_THREADS = 5 * BS_CW;  // 480
constexpr int BS_VPITCH = BS_CW + 1;   // odd pitch: H-phase lanes walk rows, keep them on distinct banks

template <int M>
__global__ void __launch_bounds__(BS_THREADS)
k_box_strip(Planes5 Min, int W, int H, int strip_rows, double scale, float2* __restrict__ flow)
{
    constexpr int R = 2 * M + 1;               // rows per step == window height
    constexpr int TW = BS_CW - 2 * M;          // output columns per CTA
    constexpr int HP = TW + 1;                 // pitch of the H buffer
    constexpr int SEG = (BS_THREADS / (5 * R)) < 1 ? 1 : (BS_THREADS / (5 * R));   // segments per row in H phase
    constexpr int SEGLEN = (TW + SEG - 1) / SEG;
    extern __shared__ float bs_smem[];
    float* sV = bs_smem;                       // 5 * R * BS_VPITCH
    float* sH = bs_smem + 5 * R * BS_VPITCH;   // 5 * R * HP

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;                      // first output column
    const int ybeg = blockIdx.y * strip_rows;
    const int yend = min(ybeg + strip_rows, H);
    if (ybeg >= H) return;

    // ---- V-phase identity: (channel, column) ----
    const int vc = tid / BS_CW;
    const int vcol = tid - vc * BS_CW;
    const int gx = min(max(x0 - M + vcol, 0), W - 1);    // replicate border in x
    const float* __restrict__ src = Min.ch(vc) + gx;
    const int pitch = Min.pitch;

    float oldv[R];
    double vsum = 0;
#pragma unroll
    for (int i = 0; i < R; i++) {
        int yy = min(max(ybeg - M - 1 + i, 0), H - 1);
        oldv[i] = src[(size_t)yy * pitch];
    }
#pragma unroll
    for (int i = 0; i < R; i++) vsum += (double)oldv[i];   // window of row ybeg-1

    for (int ys = ybeg; ys < yend; ys += R) {
        // ---- V phase ----
        float newv[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int yy = min(ys + M + r, H - 1);               // replicate border in y (ys+M+r >= 0 always)
            newv[r] = src[(size_t)yy * pitch];
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            vsum += (double)newv[r];
            vsum -= (double)oldv[r];
            sV[(vc * R + r) * BS_VPITCH + vcol] = (float)vsum;
            oldv[r] = newv[r];
        }
        __syncthreads();

        // ---- H phase: item = (channel, row, segment) ----
        if (tid < 5 * R * SEG) {
            int seg = tid % SEG;
            int rc = tid / SEG;                 // = c*R + r
            const float* v = sV + rc * BS_VPITCH;
            float* h = sH + rc * HP;
            int xa = seg * SEGLEN, xb = min(xa + SEGLEN, TW);
            if (xa < xb) {
                double s = 0;
#pragma unroll
                for (int i = 0; i < R; i++) s += (double)v[xa + i];     // window of output xa: cols xa..xa+2M
                h[xa] = (float)s;
                for (int x = xa + 1; x < xb; x++) {
                    s += (double)v[x + 2 * M];
                    s -= (double)v[x - 1];
                    h[x] = (float)s;
                }
            }
        }
        __syncthreads();

        // ---- S phase: item = pixel ----
        for (int i = tid; i < R * TW; i += BS_THREADS) {
            int r = i / TW, lx = i - r * TW;
            int y = ys + r, x = x0 + lx;
            if (y < yend && x < W) {
                const float* h = sH + r * HP + lx;
                double g11 = (double)h[0 * R * HP] * scale, g12 = (double)h[1 * R * HP] * scale,
                       g22 = (double)h[2 * R * HP] * scale, h1 = (double)h[3 * R * HP] * scale,
                       h2 = (double)h[4 * R * HP] * scale;
                flow[(size_t)y * W + x] = solve_flow(g11, g12, g22, h1, h2);
            }
        }
        // no barrier needed here: the next V phase only writes sV (last read before the barrier above),
        // and sH is next written after the barrier that follows it.
    }
}

template <int M>
static void run_box_strip(Launch& L, Planes5 Min, int W, int H, double scale, float2* flow, int sm_count)
{
    constexpr int R = 2 * M + 1, TW = BS_CW - 2 * M, HP = TW + 1;
    const size_t smem = sizeof(float) * (5 * R * BS_VPITCH + 5 * R * HP);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_box_strip<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    int xt = divup(W, TW);
    // aim at ~2 CTAs per SM over the grid; strips are whole steps of R rows
    int want = std::max(1, (2 * sm_count + xt - 1) / xt);
    int strip = divup(divup(H, want), R) * R;
    int ns = divup(H, strip);
    dim3 grid(xt, ns);
    L.run("box_strip", [&](cudaStream_t s) {
        k_box_strip<M><<<grid, BS_THREADS, smem, s>>>(Min, W, H, strip, scale, flow);
    });
}

static int g_sm_count = 148;
void set_sm_count(int n) { g_sm_count = n > 0 ? n : 148; }

void launch_blur_solve_box(Launch& L, Planes5 M, int W, int H, int winsize, double* tmp, float2* flow, bool generic)
{
    int m = winsize / 2;
    double scale = 1. / ((double)winsize * winsize);
    if (!generic) {
        switch (m) {
#define OFB_CASE(MM) case MM: run_box_strip<MM>(L, M, W, H, scale, flow, g_sm_count); return;
            OFB_CASE(1) OFB_CASE(2) OFB_CASE(3) OFB_CASE(4) OFB_CASE(5) OFB_CASE(6) OFB_CASE(7) OFB_CASE(8)
            OFB_CASE(9) OFB_CASE(10) OFB_CASE(11) OFB_CASE(12) OFB_CASE(13) OFB_CASE(14) OFB_CASE(15) OFB_CASE(16)
#undef OFB_CASE
            default: break;
        }
    }
    dim3 b(64, 4), g(divup(W, 64), divup(H, 4), 5);
    size_t tplane = (size_t)H * M.pitch;
    int tpitch = M.pitch;
    L.run("box_v_generic", [&](cudaStream_t s) { k_box_v_generic<<<g, b, 0, s>>>(M, W, H, m, tmp, tplane, tpitch); });
    dim3 g2(divup(W, 64), divup(H, 4));
    L.run("box_h_solve_generic", [&](cudaStream_t s) {
        k_box_h_solve_generic<<<g2, b, 0, s>>>(tmp, tplane, tpitch, W, H, m, scale, flow);
    });
}

void launch_blur_solve_gauss(Launch& L, Planes5 M, int W, int H, int winsize, const float* half_taps,
                             float* tmp, float2* flow, bool /*generic*/)
{
    int m = winsize / 2;
    dim3 b(64, 4), g(divup(W, 64), divup(H, 4), 5);
    size_t tplane = (size_t)H * M.pitch;
    int tpitch = M.pitch;
    L.run("gauss_v_generic", [&](cudaStream_t s) {
        k_gauss_v_generic<<<g, b, 0, s>>>(M, W, H, m, half_taps, tmp, tplane, tpitch);
    });
    dim3 g2(divup(W, 64), divup(H, 4));
    L.run("gauss_h_solve_generic", [&](cudaStream_t s) {
        k_gauss_h_solve_generic<<<g2, b, 0, s>>>(tmp, tplane, tpitch, W, H, m, half_taps, flow);
    });
}

}  // namespace ofb
