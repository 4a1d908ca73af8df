// blur_solve.cu -- GENERIC (any winsize) FarnebackUpdateFlow_Blur / FarnebackUpdateFlow_GaussianBlur without their
// trailing UpdateMatrices (the iteration body of cv2.calcOpticalFlowFarneback called at
// /root/reference/optical_flow.py:51 and visualize_optical_flow.py:38):
//     B = blur(M) over a (2m+1)^2 window, replicate border, m = winsize/2
//     flow = solve2x2(B)   with  idet = 1/(B0*B2 - B1^2 + 1e-3)   in double
// Two global-memory passes with direct sums: f64 for the box window (cv2 keeps f64 running sums), f32 taps in cv2's
// folded order for the Gaussian window, f64 solve -- bit-exact against the CPU oracle for the Gaussian window.
// These kernels serve winsize = 1 (with cv2's m = 0 quirk), winsize > 33, iterations = 0 and option
// "generic_kernels"; every other case runs the fused strip kernel k_iter in iter.cu.
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

void launch_blur_solve_box(Launch& L, Planes5 M, int W, int H, int winsize, double* tmp, float2* flow, bool /*generic*/)
{
    int m = winsize / 2;
    double scale = 1. / ((double)winsize * winsize);
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
