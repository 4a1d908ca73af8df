// iter.cu -- the per-scale iteration of cv2.calcOpticalFlowFarneback (call
// sites /root/reference/optical_flow.py:51, visualize_optical_flow.py:38), batched over frame pairs.
//
//   k_um0<SRC>      first UpdateMatrices of a scale, fused with the inter-scale flow initialisation
//                   (zeros | an existing flow | bilinear up-sample of the coarser flow * 1/pyr_scale).
//   k_iter<M,FUSE>  FarnebackUpdateFlow_Blur for a box window of half-width M:
//                     blur(M_in) -> 2x2 solve -> flow                       (FUSE = false, last iteration)
//                     blur(M_in) -> 2x2 solve -> UpdateMatrices -> M_out    (FUSE = true; the flow of a middle
//                                                                           iteration never leaves registers)
//
// k_iter: one CTA owns TW = 96-2M output columns and walks down a strip of rows in steps of R = 2M+1.
//   V phase  thread = (channel, column).  The R rows entering the window are loaded once (coalesced) into
//            registers; window sums for the R rows of the step come from suffix sums of the previous block
//            and prefix sums of the new one (van Herk / Gil-Werman): 3 FADD per row, no re-read of M.
//   H phase  thread = (channel, row, segment of R outputs): the same trick along x out of shared memory.
//   S phase  thread = pixel: determinant and numerators by Kahan's FMA-compensated a*d-b*c in f32
//            (error <= 1.5 ulp of the exact f32-input result), one IEEE reciprocal (cv2: idet = 1/det), then (FUSE) the
//            per-pixel UpdateMatrices with its bilinear gather of R1.
// Precision: cv2 keeps f64 running sums and an f64 solve.  Here every window sum is a <= 2R-term f32 sum
// (no running sum over the image, so no drift) and the solve is compensated; measured endpoint difference
// vs cv2 stays at the 1e-6 px level (tests/test_gpu_parity.py), four orders inside the tolerance.
// The f32->f64 conversions this avoids ran on the 16-lane XU pipe and were the bottleneck of the first
// version (profiles/r1_ncu_first.md).
// Roofline: HBM.  Algorithmic bytes per pixel: 28 (FUSE=false), 96 (FUSE=true), 68 (k_um0).
#include "common.cuh"
#include "launch.cuh"
#include "um_device.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace ofb {

// ------------------------------------------------------------------------------------------------
// k_um0
// ------------------------------------------------------------------------------------------------
// CTA = UM0_BX x UM0_BY threads, UM0_ROWS consecutive rows per thread.  Block shape measured on B200 with one row per thread (1080p,
// ms per 300-pair step): 32x8 9.1, 32x16 9.6, 64x4 9.0, 32x4 8.6, 128x1 8.7, 64x1 8.7, 256x1 9.2, 64x2 8.5 -- small CTAs of two long
// rows hide the gathers best; rows per thread: 1 / 2 / 4 / 8 -> 8.74 / 8.62 / 7.93 / 7.81 (profiles/r2x_ab_um_packed_rows.log).
#ifndef OFB_UM0_BX
#define OFB_UM0_BX 64
#endif
#ifndef OFB_UM0_BY
#define OFB_UM0_BY 2
#endif
constexpr int UM0_BX = OFB_UM0_BX, UM0_BY = OFB_UM0_BY;

#ifndef OFB_UM0_ROWS
#define OFB_UM0_ROWS 4
#endif
#ifndef OFB_UM0_MINB
#define OFB_UM0_MINB 8
#endif
constexpr int UM0_ROWS = OFB_UM0_ROWS;    // consecutive rows per thread (a sequential loop, not ILP): the slot / base-pointer
                                          // arithmetic and the per-column table loads -- a quarter of the instructions of the
                                          // one-pixel-per-thread version, which ran at 72 % issue utilisation -- are paid once

template <int SRC>   // 0 zero flow, 1 read flow, 2 up-sample coarse flow
__global__ void __launch_bounds__(UM0_BX * UM0_BY, OFB_UM0_MINB)
k_um0(Um0Args a)
{
    // blockIdx.x = batch item (fastest-varying in dispatch order): the CTAs of consecutive pairs for the same tile
    // run together, so the frame slot pair z reads as R1 and pair z+1 reads as R0 comes from HBM once.
    const int x = blockIdx.y * UM0_BX + threadIdx.x, z = blockIdx.x;
    const int ybeg = (blockIdx.z * UM0_BY + threadIdx.y) * UM0_ROWS;
    // The two per-item bases go through a warp shuffle: ptxas otherwise folds z * item into EVERY address of the loop (a 64-bit
    // multiply-add, LEA and LEA.HI.X per load / store: 50 instructions per pixel); a shuffled pointer is just a register pair.
    float* const mout = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(a.M + (size_t)z * a.m_item), 0));
    const float2* const fl = reinterpret_cast<const float2*>(      // SRC 1: this scale's flow; SRC 2: the coarser one
        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(a.flow + (size_t)z * a.flow_item), 0));
    __builtin_assume(__isGlobal(mout));
    if (SRC != 0) __builtin_assume(__isGlobal(fl));            // SRC 0 has no flow (a.flow may be null) and never touches fl
    if (x >= a.W || ybeg >= a.H) return;
    const int yend = min(ybeg + UM0_ROWS, a.H);
    const UmBase ub = um_base(a.R, a.slot0, z);                 // M and R of a level share pitch and plane size (launch_um0)
    const int W = a.W, H = a.H;
    const unsigned plane = (unsigned)a.plane;
    unsigned o0 = (unsigned)ybeg * (unsigned)a.pitch + (unsigned)x;
    unsigned sx = 0, sx1 = 0;
    float a1 = 0.f, a0 = 1.f;
    if (SRC == 2) {
        sx = a.ux[x]; a1 = a.uax[x];
        a0 = 1.f - a1;
        sx1 = min(sx + 1, (unsigned)a.Wp - 1);
    }
#pragma unroll 1
    for (int y = ybeg; y < yend; y++, o0 += a.pitch) {
        float dx = 0.f, dy = 0.f;
        if (SRC == 1) {
            float2 d = fl[(unsigned)y * (unsigned)W + (unsigned)x];
            dx = d.x; dy = d.y;
        } else if (SRC == 2) {
            const int sy = a.uy[y];
            const float b1 = a.uay[y], b0 = 1.f - b1;
            const unsigned c0 = (unsigned)sy * (unsigned)a.Wp;                  // one item of the coarse flow is far below 2^32 pixels
            const unsigned c1 = (unsigned)min(sy + 1, a.Hp - 1) * (unsigned)a.Wp;
            // hx = p0.x*a0 + p1.x*a1 per row, d = (h0*b0 + h1*b1) * mul: products packed over (x, y), adds scalar (um_device.cuh)
            const float2 t00 = mul2s(fl[c0 + sx], a0), t01 = mul2s(fl[c0 + sx1], a1);
            const float2 t10 = mul2s(fl[c1 + sx], a0), t11 = mul2s(fl[c1 + sx1], a1);
            const float2 h0 = mul2s(make_float2(t00.x + t01.x, t00.y + t01.y), b0);
            const float2 h1 = mul2s(make_float2(t10.x + t11.x, t10.y + t11.y), b1);
            const float2 d = mul2s(make_float2(h0.x + h1.x, h0.y + h1.y), a.mul);
            dx = d.x; dy = d.y;
        }
        M5 m = um_pixel(x, y, o0, dx, dy, ub, W, H);
#pragma unroll
        for (int c = 0; c < 5; c++) mout[o0 + c * plane] = m.v[c];          // 5 * plane < 2^32 floats
    }
}

// The kernels address M and R of a level with ONE pixel offset y * pitch + x (32 bits): both must share pitch and plane size, and
// a batch item must stay below 2^32 floats.  Every call site in engine.cu allocates them that way; anything else is a bug here.
static void check_shared_layout(const SlotRing& R, size_t plane, int pitch)
{
    if (R.pitch != pitch || R.plane != plane || 5 * plane >= (1ull << 32)) {
        fprintf(stderr, "ofb: internal error: M / R layout mismatch (pitch %d vs %d, plane %zu vs %zu)\n", pitch, R.pitch, plane, R.plane);
        abort();
    }
}

void launch_um0(Launch& L, int src, const Um0Args& a, int batch)
{
    check_shared_layout(a.R, a.plane, a.pitch);
    dim3 block(UM0_BX, UM0_BY), grid(batch, divup(a.W, UM0_BX), divup(a.H, UM0_BY * UM0_ROWS));
    const char* names[3] = {"um0_zero", "um0_flow", "um0_upsample"};
    L.run(names[src], [&](cudaStream_t s) {
        if (src == 0) k_um0<0><<<grid, block, 0, s>>>(a);
        else if (src == 1) k_um0<1><<<grid, block, 0, s>>>(a);
        else k_um0<2><<<grid, block, 0, s>>>(a);
    });
}

// ------------------------------------------------------------------------------------------------
// k_iter
// ------------------------------------------------------------------------------------------------
constexpr int IT_CW = 96;                 // columns loaded per CTA (3 warps per channel)
constexpr int IT_THREADS = 5 * IT_CW;     // 480
constexpr int IT_VP = IT_CW + 1;          // odd pitch of the V buffer

__device__ __forceinline__ float kahan_det(float a, float d, float b, float c)   // a*d - b*c
{
    float w = __fmul_rn(b, c);
    float e = __fmaf_rn(-b, c, w);
    float f = __fmaf_rn(a, d, -w);
    return __fadd_rn(f, e);
}

// ILP = number of pixels whose UpdateMatrices gathers one thread keeps in flight in the S phase.
// ILP 1 fits two CTAs per SM (<= 68 registers); ILP > 1 trades the second CTA for deeper memory parallelism.
// GAUSS = FarnebackUpdateFlow_GaussianBlur (flags & 256): the same strip walk, but the window is a
// separable Gaussian applied tap by tap in cv2's folded order (k0*c + sum k_i*(a_{+i} + a_{-i}), f32, uncontracted),
// so the blurred field is bit-identical to cv2's; only the solve differs (compensated f32 instead of f64).
template <int M, bool FUSE, int ILP, bool GAUSS>
__global__ void __launch_bounds__(IT_THREADS, (M <= 8 && ILP == 1) ? 2 : 1)
k_iter(IterArgs a)
{
    constexpr int R = 2 * M + 1;
    constexpr int TW = IT_CW - 2 * M;
    constexpr int HP = TW + 1;
    constexpr int NSEG = (TW + R - 1) / R;
    static_assert(TW >= 1 && 5 * R * NSEG <= IT_THREADS, "H phase: one thread per (channel, row, segment)");
    extern __shared__ float it_smem[];
    float* sV = it_smem;                      // 5 * R * IT_VP
    float* sH = it_smem + 5 * R * IT_VP;      // 5 * R * HP

    // batch item fastest-varying (see k_um0); optionally the whole grid order reversed (IterArgs::reverse)
    const int tid = threadIdx.x;
    const int z = a.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const int bx = a.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    const int bs = a.reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int W = a.W, H = a.H;
    // first column of the tile; through a shuffle so that it stays in a register (see k_iter64)
    const int x0 = (FUSE && ILP == 1) ? __shfl_sync(0xffffffffu, bx * TW, 0) : bx * TW;
    const int ybeg = bs * a.strip_rows;
    const int yend = min(ybeg + a.strip_rows, H);
    if (ybeg >= H) return;

    const int vc = tid / IT_CW, vcol = tid - vc * IT_CW;
    const int gx = min(max(x0 - M + vcol, 0), W - 1);                 // replicate border in x
    const float* __restrict__ src = a.Min + (size_t)z * a.m_item + (size_t)vc * a.plane + gx;
    const int pitch = a.pitch;

    float blkA[R];                                                     // rows ys-M .. ys+M of this column
#pragma unroll
    for (int i = 0; i < R; i++) blkA[i] = src[(size_t)min(max(ybeg - M + i, 0), H - 1) * pitch];

    // Output base of this batch item, pinned in a register pair through a warp shuffle (see k_um0), and the two R slots.
    // M and R of a level share pitch and plane size, so one pixel offset o0 = y * pitch + x (32 bits) addresses all of them.
    UmBase ub{nullptr, nullptr, 0u, 0};
    RView R0{nullptr, nullptr, 0}, R1{nullptr, nullptr, 0};            // for the prefetch addresses and the ILP > 1 path
    float* mout = nullptr;
    float2* fout = nullptr;
    if (FUSE) {
        ub = um_base(a.R, a.slot0, z);
        R0 = RView{const_cast<float4*>(ub.r0), const_cast<float*>(reinterpret_cast<const float*>(ub.r0)) + ub.plane4, pitch};
        R1 = RView{const_cast<float4*>(ub.r1), const_cast<float*>(reinterpret_cast<const float*>(ub.r1)) + ub.plane4, pitch};
        mout = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(a.Mout + (size_t)z * a.m_item), 0));
        __builtin_assume(__isGlobal(mout));
    } else {
        fout = a.flow + (size_t)z * a.flow_item;
    }
    const unsigned mplane = (unsigned)a.plane;

    float mag_lo = __int_as_float(0x7f800000), mag_hi = 0.f;      // per-thread min / max of |flow| (a.minmax)

    for (int ys = ybeg; ys < yend; ys += R) {
        // ---- V phase ----
        float blkB[R];                                                 // rows ys+M+1 .. ys+3M+1
        if (ys + 3 * M + 1 < H) {                                      // block-uniform: no row of the block is clamped
            const float* pb = src + (size_t)(ys + M + 1) * pitch;
#pragma unroll
            for (int r = 0; r < R; r++) blkB[r] = pb[r * pitch];
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) blkB[r] = src[(size_t)min(ys + M + 1 + r, H - 1) * pitch];
        }
        if (a.prefetch && ys + R < yend) {
            // Software prefetch into L2, one step (R rows) ahead: the M rows the next V phase will load and the
            // R0 / R1 rows the next S phase will read (R1 at the un-displaced position; flow displacements are
            // small next to a step).  One or two prefetch instructions per thread.
            const int yn = ys + R;                                   // first output row of the next step
            if (tid < 5 * R * 4) {                                   // M: 5 channels x R rows x 4 lines (96 floats + slack)
                const int c = tid / (R * 4), rem = tid - c * (R * 4), r = rem >> 2, l = rem & 3;
                const int row = min(yn + M + 1 + r, H - 1);
                const int col = max(x0 - M, 0) + 32 * l;
                if (col < pitch) {
                    const float* p = a.Min + (size_t)z * a.m_item + (size_t)c * a.plane + (size_t)row * pitch + col;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
            if (FUSE) {
                constexpr int LA = (TW * 16 + 127) / 128 + 1;        // 128-byte lines per row of the float4 part
                if (tid < R * LA) {
                    const int r = tid / LA, l = tid - r * LA;
                    const int row = min(yn + r, H - 1), col = x0 + 8 * l;
                    if (col < pitch) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R0.a + (size_t)row * pitch + col));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R1.a + (size_t)row * pitch + col));
                    }
                } else if (tid < R * LA + R * 4) {
                    const int j = tid - R * LA, r = j >> 2, l = j & 3;
                    const int row = min(yn + r, H - 1), col = x0 + 32 * l;
                    if (col < pitch) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R0.b + (size_t)row * pitch + col));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R1.b + (size_t)row * pitch + col));
                    }
                }
            }
        }
        if (GAUSS) {
            float* v = sV + vc * R * IT_VP + vcol;
#pragma unroll
            for (int r = 0; r < R; r++) {
                // window rows r .. r+2M of the concatenation [blkA | blkB]; centre at r+M
                const int cidx = r + M;
                float s0 = (cidx < R ? blkA[cidx] : blkB[cidx - R]) * a.gk[0];
#pragma unroll
                for (int i = 1; i <= M; i++) {
                    const int dn = cidx + i, up = cidx - i;
                    const float vdn = dn < R ? blkA[dn] : blkB[dn - R];
                    const float vup = up < R ? blkA[up] : blkB[up - R];
                    s0 = s0 + (vdn + vup) * a.gk[i];
                }
                v[r * IT_VP] = s0;
            }
        } else {
#pragma unroll
            for (int i = R - 2; i >= 0; i--) blkA[i] = __fadd_rn(blkA[i], blkA[i + 1]);      // suffix sums
            float* v = sV + vc * R * IT_VP + vcol;
            v[0] = blkA[0];
            float p = blkB[0];
#pragma unroll
            for (int r = 1; r < R; r++) {
                v[r * IT_VP] = __fadd_rn(blkA[r], p);
                p = __fadd_rn(p, blkB[r]);
            }
        }
#pragma unroll
        for (int i = 0; i < R; i++) blkA[i] = blkB[i];
        __syncthreads();

        // ---- H phase: item = (channel*R + row, segment) ----
        if (tid < 5 * R * NSEG) {
            // lanes of a warp = consecutive (channel,row) lines of ONE segment: both pitches are odd, so the
            // lanes' reads of sV and writes of sH fall in distinct banks
            const int seg = tid / (5 * R), rc = tid - seg * (5 * R);
            const int xa = seg * R;
            const float* v = sV + rc * IT_VP + xa;
            float* h = sH + rc * HP + xa;
            if (GAUSS) {
                float w[2 * R - 1];
#pragma unroll
                for (int i = 0; i < 2 * R - 1; i++) w[i] = (xa + i < IT_CW) ? v[i] : 0.f;
#pragma unroll
                for (int j = 0; j < R; j++) {
                    if (xa + j < TW) {
                        float sum = w[j + M] * a.gk[0];
#pragma unroll
                        for (int i = 1; i <= M; i++) sum = sum + a.gk[i] * (w[j + M - i] + w[j + M + i]);
                        h[j] = sum;
                    }
                }
            } else {
                float sa[R];
#pragma unroll
                for (int i = 0; i < R; i++) sa[i] = (xa + i < IT_CW) ? v[i] : 0.f;
#pragma unroll
                for (int i = R - 2; i >= 0; i--) sa[i] = __fadd_rn(sa[i], sa[i + 1]);
                if (xa < TW) h[0] = sa[0];
                float p = 0.f;
#pragma unroll
                for (int r = 1; r < R; r++) {
                    if (xa + r < TW) {
                        float nb = v[R + r - 1];                           // column xa+r+2M <= CW-1
                        p = (r == 1) ? nb : __fadd_rn(p, nb);
                        h[r] = __fadd_rn(sa[r], p);
                    }
                }
            }
        }
        __syncthreads();

        // ---- S phase: item = pixel ----
        if (ILP == 1) {
            // A thread's items are tid, tid + T, ...: (r, lx) advance by (T / TW, T % TW) with a carry instead of a division per
            // pixel.  `interior` (block-uniform): no pixel of this step is within 5 px of a border, so UpdateMatrices skips the
            // per-pixel border test.
            auto s_phase = [&](auto interior_tag) {
                constexpr bool INTERIOR = decltype(interior_tag)::value;
                constexpr int DR = IT_THREADS / TW, DL = IT_THREADS % TW;
                int r = tid / TW, lx = tid - r * TW;
                while (r < R) {
                    const int y = ys + r, x = x0 + lx;
                    if (y < yend && x < W) {
                        const float* h = sH + r * HP + lx;
                        const float g11 = h[0], g12 = h[R * HP], g22 = h[2 * R * HP], h1 = h[3 * R * HP], h2 = h[4 * R * HP];
                        // flow = [g11*h2 - g12*h1, g22*h1 - g12*h2] * scale^2 / ((g11*g22 - g12^2) * scale^2 + 1e-3)
                        //      = [ ... ] / (g11*g22 - g12^2 + 1e-3 / scale^2)
                        const float det = __fadd_rn(kahan_det(g11, g22, g12, g12), a.c);
                        const float idet = __frcp_rn(det);
                        const float fx = __fmul_rn(kahan_det(g11, h2, g12, h1), idet);
                        const float fy = __fmul_rn(kahan_det(g22, h1, g12, h2), idet);
                        if (FUSE) {
                            const unsigned o0 = (unsigned)y * (unsigned)pitch + (unsigned)x;
                            M5 m = um_pixel<INTERIOR>(x, y, o0, fx, fy, ub, W, H);
#pragma unroll
                            for (int c = 0; c < 5; c++) mout[o0 + c * mplane] = m.v[c];
                        } else {
                            fout[(unsigned)y * (unsigned)W + (unsigned)x] = make_float2(fx, fy);
                            if (a.minmax) {                        // cv::cartToPolar's magnitude, as in viz.cu
                                const float mg = sqrtf(fmaf(fx, fx, fy * fy));
                                mag_lo = fminf(mag_lo, mg); mag_hi = fmaxf(mag_hi, mg);
                            }
                        }
                    }
                    r += DR; lx += DL;
                    if (lx >= TW) { lx -= TW; ++r; }
                }
            };
            if (FUSE && x0 >= 5 && x0 + TW <= W - 5 && ys >= 5 && ys + R <= H - 5) s_phase(std::true_type{});
            else s_phase(std::false_type{});
        } else {
            for (int i0 = tid; i0 < R * TW; i0 += ILP * IT_THREADS) {
                UmLoads L[ILP];
                bool ok[ILP];
#pragma unroll
                for (int j = 0; j < ILP; j++) {
                    const int i = i0 + j * IT_THREADS;
                    const int r = i / TW, lx = i - r * TW;
                    const int y = ys + r, x = x0 + lx;
                    ok[j] = (i < R * TW) && y < yend && x < W;
                    if (ok[j]) {
                        const float* h = sH + r * HP + lx;
                        const float g11 = h[0], g12 = h[R * HP], g22 = h[2 * R * HP], h1 = h[3 * R * HP], h2 = h[4 * R * HP];
                        const float det = __fadd_rn(kahan_det(g11, g22, g12, g12), a.c);
                        const float idet = __frcp_rn(det);
                        const float fx = __fmul_rn(kahan_det(g11, h2, g12, h1), idet);
                        const float fy = __fmul_rn(kahan_det(g22, h1, g12, h2), idet);
                        if (FUSE) L[j] = um_load(x, y, fx, fy, R0, R1, W, H);
                        else fout[(size_t)y * W + x] = make_float2(fx, fy);
                    }
                }
                if (FUSE) {
#pragma unroll
                    for (int j = 0; j < ILP; j++) {
                        if (ok[j]) {
                            M5 m = um_compute(L[j], W, H);
                            const unsigned o0 = (unsigned)L[j].y * (unsigned)pitch + (unsigned)L[j].x;
#pragma unroll
                            for (int c = 0; c < 5; c++) mout[o0 + c * mplane] = m.v[c];
                        }
                    }
                }
            }
        }
        // no barrier here: the next V phase writes only sV (last read before the barrier above); sH is
        // next written after the barrier that follows that V phase.
    }
    if (!FUSE && a.minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mag_lo = fminf(mag_lo, __shfl_xor_sync(0xffffffffu, mag_lo, o));
            mag_hi = fmaxf(mag_hi, __shfl_xor_sync(0xffffffffu, mag_hi, o));
        }
        if ((tid & 31) == 0) {                                 // magnitudes are >= 0: their bit patterns order like integers
            atomicMin(a.minmax + 2 * z, __float_as_uint(mag_lo));
            atomicMax(a.minmax + 2 * z + 1, __float_as_uint(mag_hi));
        }
    }
}

template <int M, bool FUSE, int ILP, bool GAUSS>
static void run_iter(Launch& L, IterArgs a, int batch)
{
    constexpr int R = 2 * M + 1, TW = IT_CW - 2 * M, HP = TW + 1;
    const size_t smem = sizeof(float) * (5 * R * IT_VP + 5 * R * HP);
    static unsigned long long configured = 0;             // bit d: attribute set on device d
    L.dyn_smem(k_iter<M, FUSE, ILP, GAUSS>, smem, configured);
    const int sm_count = L.sm_count;
    const int xt = divup(a.W, TW);
    // aim at ~1 CTA per SM for ONE pair; strips are whole steps of R rows.  The partition must not depend
    // on the batch size: the van Herk blocks restart at strip boundaries, so it fixes the f32 rounding, and a
    // pair must give bit-identical results whether it is processed alone or inside a batch.
    int want = std::max(1, (sm_count + xt - 1) / xt);       // ~one CTA per SM for a single pair; batches bring the rest
    int strip = divup(divup(a.H, want), R) * R;
    strip = std::max(strip, std::min(a.H, 4 * R));          // the block-A preload of a strip costs one step: keep it <= 25 %
    strip = divup(strip, R) * R;
    a.strip_rows = strip;
    a.prefetch = L.opt.iter_prefetch;
    dim3 grid(batch, xt, divup(a.H, strip));
    L.run(GAUSS ? (FUSE ? "iter_fused_gauss" : "iter_last_gauss") : (FUSE ? "iter_fused" : "iter_last"), [&](cudaStream_t s) {
        k_iter<M, FUSE, ILP, GAUSS><<<grid, IT_THREADS, smem, s>>>(a);
    });
}


// M is streamed (every element is used by one thread, once): under OFB_I64_NOALLOC its loads bypass L1 allocation so that the
// small L1 left beside k_iter64's shared memory keeps the R1 lines the UpdateMatrices gathers re-use.
__device__ __forceinline__ float ld_stream(const float* p)
{
#ifdef OFB_I64_NOALLOC
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}

// ------------------------------------------------------------------------------------------------
// k_iter64<M, FUSE>: the box-window iteration with cv2's OWN arithmetic for the window sums and the solve (option
// "exact_window_sums"): FarnebackUpdateFlow_Blur keeps, per column, ONE double running sum down the whole image,
//     vsum = float(row0 * (m+2)) + rows 1..m-1;   for every row y:  vsum += float(row[min(y+m, H-1)] - row[max(y-m-1, 0)])
// (the difference of the entering and the leaving row is formed in FLOAT before it is added to the double), double running
// sums along x, and a double solve (restated in oracle/farneback_oracle.c orc_blur_solve_box).
//
// Why it exists.  Where a window is rank-deficient (one edge direction, or the replicated rows at the bottom of a periodic
// pattern) det = g11*g22 - g12^2 cancels completely and the flow cv2 returns is decided by the rounding history of ITS sums:
// the float differences above leave a drift of ~1e-7 relative that grows down the column.  Any other summation -- the f32
// van Herk sums of k_iter, or exact f64 sums -- is equally "right" and lands up to 0.2 px elsewhere on a 0/255 checkerboard
// (19 000 pixels in the bottom 10 rows of a 1080p frame; cv2's own optimised and plain builds differ by 3e-3 there;
// profiles/r2e_diag_stage_swap.log).  Reproducing cv2 on such input needs its arithmetic, including the drift, and the drift
// needs the whole column: in this mode a CTA walks the FULL height of its 82-column tile (no strips).
// Same phases as k_iter.  V phase: thread = (channel, column), 1 FADD + 1 F2F + 1 DADD per element.  H phase: thread =
// (channel, row, segment), van Herk in f64 out of shared memory (order differs from cv2's running sum by 1e-16 relative).
// S phase: thread = pixel, determinant and numerators with DFMA, rounded to f32 only for the final quotient, then (FUSE)
// UpdateMatrices exactly as in k_iter.  Shared memory: 8 * 5R * (97 + TW + 1) bytes = 108 KB for winsize 15, two CTAs per SM.
// Measured on B200 against k_iter: iter_fused +26 %, iter_last +42 % (f64 shared-memory traffic), whole step -13 %.
// ------------------------------------------------------------------------------------------------
// Tile width of k_iter64 (columns loaded per CTA; 5 channels x I64_CW threads).  Measured on B200, 1080p, ms per 300-pair step
// (iter_fused / iter_last): see DESIGN.md section 4d.
#ifndef OFB_I64_CW
#define OFB_I64_CW 96
#endif
constexpr int I64_CW = OFB_I64_CW, I64_THREADS = 5 * I64_CW, I64_VP = I64_CW + 1;
constexpr int i64_min_blocks(int M)
{
    // CTAs per SM that shared memory (227 KB) and the register file (64 K / threads, <= 64 registers each... the cap the bound sets) allow
    const int R = 2 * M + 1, TW = I64_CW - 2 * M;
    const long smem = 8L * 5 * R * (I64_VP + TW + 1) + 1024;
    int by_smem = (int)(232448L / smem);
    int by_regs = 65536 / (I64_THREADS * 64);
    int n = by_smem < by_regs ? by_smem : by_regs;
    return n < 1 ? 1 : (n > 3 ? 3 : n);
}

template <int M, bool FUSE>
__global__ void __launch_bounds__(I64_THREADS, i64_min_blocks(M))
k_iter64(IterArgs a)
{
    constexpr int R = 2 * M + 1;
    constexpr int TW = I64_CW - 2 * M;
    constexpr int HP = TW + 1;
    constexpr int NSEG = (TW + R - 1) / R;
    static_assert(TW >= 1 && 5 * R * NSEG <= I64_THREADS, "H phase: one thread per (channel, row, segment)");
    extern __shared__ double it_smem64[];
    double* sV = it_smem64;                     // 5 * R * I64_VP
    double* sH = it_smem64 + 5 * R * I64_VP;     // 5 * R * HP

    const int tid = threadIdx.x;
    const int z = blockIdx.x, bx = blockIdx.y, bs = blockIdx.z;
    const int W = a.W, H = a.H;
    // first column of the tile; through a shuffle so that it stays in a register (ptxas otherwise re-reads %ctaid.y and
    // multiplies again at every use inside the S phase)
    const int x0 = FUSE ? __shfl_sync(0xffffffffu, bx * TW, 0) : bx * TW;
    const int ybeg = bs * a.strip_rows;
    const int yend = min(ybeg + a.strip_rows, H);
    if (ybeg >= H) return;

    const int vc = tid / I64_CW, vcol = tid - vc * I64_CW;
    const int gx = min(max(x0 - M + vcol, 0), W - 1);                 // replicate border in x
    const float* __restrict__ src = a.Min + (size_t)z * a.m_item + (size_t)vc * a.plane + gx;
    const int pitch = a.pitch;

    // old[r] = the row that LEAVES the window when it moves down to row ys + r:  max(ys + r - M - 1, 0)
#ifdef OFB_I64_RELOAD_OLD
    // The leaving rows are loaded again (they entered one step ago: L2 hits) instead of being carried in R registers across the
    // H and S phases, whose UpdateMatrices working set otherwise makes ptxas re-derive tile coordinates and shuffle pairs.
    float old0 = src[(size_t)min(max(ybeg - M - 1, 0), H - 1) * pitch];
#else
    float old[R];
#pragma unroll
    for (int i = 0; i < R; i++) old[i] = ld_stream(src + (size_t)min(max(ybeg - M - 1 + i, 0), H - 1) * pitch);
#endif
    // S = cv2's vsum BEFORE row ybeg is processed.  At the top of the image that is float(row0 * (m+2)) + rows 1..m-1;
    // a strip that starts lower (not used by the exact mode) starts from the plain sum of the window of row ybeg - 1.
    double S;
    if (ybeg == 0) {
#ifdef OFB_I64_RELOAD_OLD
        S = (double)__fmul_rn(old0, (float)(M + 2));
#else
        S = (double)__fmul_rn(old[0], (float)(M + 2));
#endif
        for (int y = 1; y < M; y++) S += (double)src[(size_t)min(y, H - 1) * pitch];
    } else {
        S = 0.0;
        for (int j = ybeg - 1 - M; j <= ybeg - 1 + M; j++) S += (double)src[(size_t)min(max(j, 0), H - 1) * pitch];
    }

    // Output base of this batch item, pinned in a register pair through a warp shuffle (see k_um0), and the two R slots.
    // M and R of a level share pitch and plane size, so one pixel offset o0 = y * pitch + x (32 bits) addresses all of them.
    UmBase ub{nullptr, nullptr, 0u, 0};
    RView R0{nullptr, nullptr, 0}, R1{nullptr, nullptr, 0};            // only for the prefetch addresses
    float* mout = nullptr;
    float2* fout = nullptr;
    if (FUSE) {
        ub = um_base(a.R, a.slot0, z);
        R0 = RView{const_cast<float4*>(ub.r0), const_cast<float*>(reinterpret_cast<const float*>(ub.r0)) + ub.plane4, pitch};
        R1 = RView{const_cast<float4*>(ub.r1), const_cast<float*>(reinterpret_cast<const float*>(ub.r1)) + ub.plane4, pitch};
        mout = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(a.Mout + (size_t)z * a.m_item), 0));
        __builtin_assume(__isGlobal(mout));
    } else {
        fout = a.flow + (size_t)z * a.flow_item;
    }
    const unsigned mplane = (unsigned)a.plane;
    const double c64 = a.c64;
    float mag_lo = __int_as_float(0x7f800000), mag_hi = 0.f;

    for (int ys = ybeg; ys < yend; ys += R) {
        // ---- V phase ----
        float nb[R];                                                   // rows ys+M .. ys+3M enter the window during this step
        if (ys + 3 * M < H) {
            const float* pb = src + (size_t)(ys + M) * pitch;
#pragma unroll
            for (int r = 0; r < R; r++) nb[r] = ld_stream(pb + r * pitch);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) nb[r] = ld_stream(src + (size_t)min(ys + M + r, H - 1) * pitch);
        }
        if (a.prefetch && ys + R < yend) {
            const int yn = ys + R;
            if (tid < 5 * R * 4) {
                const int c = tid / (R * 4), rem = tid - c * (R * 4), r = rem >> 2, l = rem & 3;
                const int row = min(yn + M + r, H - 1);
                const int col = max(x0 - M, 0) + 32 * l;
                if (col < pitch) {
                    const float* p = a.Min + (size_t)z * a.m_item + (size_t)c * a.plane + (size_t)row * pitch + col;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
            if (FUSE) {
                constexpr int LA = (TW * 16 + 127) / 128 + 1;
                if (tid < R * LA) {
                    const int r = tid / LA, l = tid - r * LA;
                    const int row = min(yn + r, H - 1), col = x0 + 8 * l;
                    if (col < pitch) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R0.a + (size_t)row * pitch + col));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R1.a + (size_t)row * pitch + col));
                    }
                } else if (tid < R * LA + R * 4) {
                    const int j = tid - R * LA, r = j >> 2, l = j & 3;
                    const int row = min(yn + r, H - 1), col = x0 + 32 * l;
                    if (col < pitch) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R0.b + (size_t)row * pitch + col));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(R1.b + (size_t)row * pitch + col));
                    }
                }
            }
        }
        {
            double* v = sV + vc * R * I64_VP + vcol;
#ifdef OFB_I64_RELOAD_OLD
            float old[R];
            if (ys - M - 1 >= 0) {
                const float* po = src + (size_t)(ys - M - 1) * pitch;      // rows ys-M-1 .. ys+M-1 < H: no clamp
#pragma unroll
                for (int r = 0; r < R; r++) old[r] = (ys - M - 1 + r < H) ? po[r * pitch] : 0.f;
#pragma unroll
                for (int r = 0; r < R; r++) if (ys - M - 1 + r >= H) old[r] = src[(size_t)(H - 1) * pitch];
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) old[r] = src[(size_t)min(max(ys - M - 1 + r, 0), H - 1) * pitch];
            }
#endif
#pragma unroll
            for (int r = 0; r < R; r++) {
                S += (double)__fsub_rn(nb[r], old[r]);                 // cv2: vsum[x] += srow1[x] - srow0[x]  (float difference)
                v[r * I64_VP] = S;                                      // window of row ys + r
#ifndef OFB_I64_RELOAD_OLD
                old[r] = nb[r];                                        // the rows that entered now leave during the next step
#endif
            }
        }
        __syncthreads();

        // ---- H phase: item = (channel*R + row, segment), van Herk along x in f64 ----
        if (tid < 5 * R * NSEG) {
            const int seg = tid / (5 * R), rc = tid - seg * (5 * R);
            const int xa = seg * R;
            const double* v = sV + rc * I64_VP + xa;
            double* h = sH + rc * HP + xa;
            double sa[R];
#pragma unroll
            for (int i = 0; i < R; i++) sa[i] = (xa + i < I64_CW) ? v[i] : 0.0;
#pragma unroll
            for (int i = R - 2; i >= 0; i--) sa[i] = sa[i] + sa[i + 1];
            if (xa < TW) h[0] = sa[0];
            double p = 0.0;
#pragma unroll
            for (int r = 1; r < R; r++) {
                if (xa + r < TW) {
                    const double q = v[R + r - 1];                     // column xa+r+2M <= CW-1
                    p = (r == 1) ? q : p + q;
                    h[r] = sa[r] + p;
                }
            }
        }
        __syncthreads();

        // ---- S phase: item = pixel.  A thread's items are tid, tid + T, ...: (r, lx) advance by (T / TW, T % TW) with a carry
        // instead of a division per pixel.  `interior` (block-uniform): no pixel of this step is within 5 px of a border, so
        // UpdateMatrices skips the per-pixel border test (22 of 24 tiles, 70 of 72 steps at 1080p). ----
        auto s_phase = [&](auto interior_tag) {
            constexpr bool INTERIOR = decltype(interior_tag)::value;
            constexpr int DR = I64_THREADS / TW, DL = I64_THREADS % TW;
            int r = tid / TW, lx = tid - r * TW;
            while (r < R) {
                const int y = ys + r, x = x0 + lx;
                if (y < yend && x < W) {
                    const double* h = sH + r * HP + lx;
                    const double g11 = h[0], g12 = h[R * HP], g22 = h[2 * R * HP], h1 = h[3 * R * HP], h2 = h[4 * R * HP];
                    // cv2: idet = 1 / (g11*g22 - g12^2 + 1e-3) on sums scaled by 1/w^2; here unscaled sums, c64 = 1e-3 * w^4.  The
                    // three cancelling differences stay in f64 (DFMA); only the well-conditioned quotient is formed in f32.
                    const double det = fma(g11, g22, -(g12 * g12)) + c64;
                    const double nx = fma(g11, h2, -(g12 * h1));
                    const double ny = fma(g22, h1, -(g12 * h2));
                    const float idet = __frcp_rn((float)det);
                    const float fx = __fmul_rn((float)nx, idet);
                    const float fy = __fmul_rn((float)ny, idet);
                    if (FUSE) {
                        const unsigned o0 = (unsigned)y * (unsigned)pitch + (unsigned)x;
                        M5 m = um_pixel<INTERIOR>(x, y, o0, fx, fy, ub, W, H);
#pragma unroll
                        for (int c = 0; c < 5; c++) mout[o0 + c * mplane] = m.v[c];
                    } else {
                        fout[(unsigned)y * (unsigned)W + (unsigned)x] = make_float2(fx, fy);
                        if (a.minmax) {
                            const float mg = sqrtf(fmaf(fx, fx, fy * fy));
                            mag_lo = fminf(mag_lo, mg); mag_hi = fmaxf(mag_hi, mg);
                        }
                    }
                }
                r += DR; lx += DL;
                if (lx >= TW) { lx -= TW; ++r; }
            }
        };
        if (FUSE && x0 >= 5 && x0 + TW <= W - 5 && ys >= 5 && ys + R <= H - 5) s_phase(std::true_type{});
        else s_phase(std::false_type{});
        // no barrier here: the next V phase writes only sV (last read before the barrier above); sH is next written after
        // the barrier that follows that V phase.
    }
    if (!FUSE && a.minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mag_lo = fminf(mag_lo, __shfl_xor_sync(0xffffffffu, mag_lo, o));
            mag_hi = fmaxf(mag_hi, __shfl_xor_sync(0xffffffffu, mag_hi, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(a.minmax + 2 * z, __float_as_uint(mag_lo));
            atomicMax(a.minmax + 2 * z + 1, __float_as_uint(mag_hi));
        }
    }
}

template <int M, bool FUSE>
static void run_iter64(Launch& L, IterArgs a, int batch)
{
    constexpr int R = 2 * M + 1, TW = I64_CW - 2 * M, HP = TW + 1;
    const size_t smem = sizeof(double) * (5 * R * I64_VP + 5 * R * HP);
    static unsigned long long configured = 0;
    L.dyn_smem(k_iter64<M, FUSE>, smem, configured);
    const int xt = divup(a.W, TW);
    a.strip_rows = divup(a.H, R) * R;                        // ONE strip: cv2's running sums (and their drift) run down the whole column
    a.prefetch = L.opt.iter_prefetch;
    dim3 grid(batch, xt, 1);
    L.run(FUSE ? "iter_fused" : "iter_last", [&](cudaStream_t s) { k_iter64<M, FUSE><<<grid, I64_THREADS, smem, s>>>(a); });
}

bool iter_supported(int winsize) { int m = winsize / 2; return m >= 1 && m <= 16; }


void launch_iter(Launch& L, const IterArgs& a, int winsize, bool fuse_um, int batch)
{
    if (fuse_um) check_shared_layout(a.R, a.plane, a.pitch);
    const int m = winsize / 2;
    if (a.gauss) {
        switch (m) {
#define OFB_CASE(MM)                                                                          \
    case MM:                                                                                  \
        if (!fuse_um) run_iter<MM, false, 1, true>(L, a, batch);                    \
        else run_iter<MM, true, 1, true>(L, a, batch);                              \
        return;
            OFB_CASE(1) OFB_CASE(2) OFB_CASE(3) OFB_CASE(4) OFB_CASE(5) OFB_CASE(6) OFB_CASE(7) OFB_CASE(8)
            OFB_CASE(9) OFB_CASE(10) OFB_CASE(11) OFB_CASE(12) OFB_CASE(13) OFB_CASE(14) OFB_CASE(15) OFB_CASE(16)
#undef OFB_CASE
            default: return;
        }
    }
    if (L.opt.exact_window_sums) {
        switch (m) {
#define OFB_CASE(MM)                                                                          \
    case MM:                                                                                  \
        if (!fuse_um) run_iter64<MM, false>(L, a, batch);                                     \
        else run_iter64<MM, true>(L, a, batch);                                               \
        return;
            OFB_CASE(1) OFB_CASE(2) OFB_CASE(3) OFB_CASE(4) OFB_CASE(5) OFB_CASE(6) OFB_CASE(7) OFB_CASE(8)
            OFB_CASE(9) OFB_CASE(10) OFB_CASE(11) OFB_CASE(12) OFB_CASE(13) OFB_CASE(14) OFB_CASE(15) OFB_CASE(16)
#undef OFB_CASE
            default: return;
        }
    }
    switch (m) {
#define OFB_CASE(MM)                                                                          \
    case MM:                                                                                  \
        if (!fuse_um) run_iter<MM, false, 1, false>(L, a, batch);                   \
        else if (L.opt.iter_ilp >= 3) run_iter<MM, true, 3, false>(L, a, batch);        \
        else if (L.opt.iter_ilp == 2) run_iter<MM, true, 2, false>(L, a, batch);        \
        else run_iter<MM, true, 1, false>(L, a, batch);                             \
        return;
        OFB_CASE(1) OFB_CASE(2) OFB_CASE(3) OFB_CASE(4) OFB_CASE(5) OFB_CASE(6) OFB_CASE(7) OFB_CASE(8)
        OFB_CASE(9) OFB_CASE(10) OFB_CASE(11) OFB_CASE(12) OFB_CASE(13) OFB_CASE(14) OFB_CASE(15) OFB_CASE(16)
#undef OFB_CASE
        default: break;
    }
}

}  // namespace ofb
