// pyramid.cu -- level image I_k of SURVEY.md A.3, the per-frame front of cv2.calcOpticalFlowFarneback
// (call sites /root/reference/optical_flow.py:51, visualize_optical_flow.py:38):
//     convertTo(CV_32F) -> GaussianBlur(ksize_k, sigma_k, reflect-101) -> resize(INTER_LINEAR)
// always taken from the FULL-RESOLUTION frame.
//
// cv2 blurs the whole frame and then samples it; only the two blurred columns / rows that the
// bilinear sample touches are needed, and the four linear, separable operators commute
// (Vlerp o Hlerp o Vblur o Hblur == [Vlerp o Vblur] o [Hlerp o Hblur]).  So:
//   k_pyr_h : T(r, x)   = (1-ax) * Hblur(r, sx) + ax * Hblur(r, sx+1)      for all H source rows
//   k_pyr_v : I(y, x)   = (1-ay) * Vblur_T(sy, x) + ay * Vblur_T(sy+1, x)
// Work is O(H*W_k*ksize) instead of O(H*W*ksize) per level.  f32 taps, f32 accumulation, taps
// applied left-to-right (the order of the CPU oracle, oracle/farneback_oracle.c gaussian_blur_f32).
#include "common.cuh"
#include "launch.cuh"

namespace ofb {

template <typename T>
__global__ void __launch_bounds__(256)
k_pyr_h(const T* __restrict__ src, int W, int H, size_t pitch_bytes, const float* __restrict__ taps, int ksize,
        double scale_x, float* __restrict__ dst, int Wk, int dst_pitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Wk || r >= H) return;
    float a1;
    int sx = linear_coord(x, scale_x, W, &a1);
    float a0 = 1.f - a1;
    const T* row = (const T*)((const char*)src + (size_t)r * pitch_bytes);
    int c = ksize / 2;
    float b0 = 0.f, b1 = 0.f;
    if (sx - c >= 0 && sx + 1 + c < W) {            // interior: no border arithmetic
        const T* p = row + sx - c;
        if (a1 != 0.f) {
            float prev = (float)p[0];
            for (int j = 0; j < ksize; j++) {
                float cur = (float)p[j + 1];
                float t = __ldg(taps + j);
                b0 += t * prev;
                b1 += t * cur;
                prev = cur;
            }
        } else {
            for (int j = 0; j < ksize; j++) b0 += __ldg(taps + j) * (float)p[j];
        }
    } else {
        int sx1 = min(sx + 1, W - 1);
        for (int j = 0; j < ksize; j++) {
            float t = __ldg(taps + j);
            b0 += t * (float)row[reflect101(sx + j - c, W)];
            if (a1 != 0.f) b1 += t * (float)row[reflect101(sx1 + j - c, W)];
        }
    }
    dst[(size_t)r * dst_pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}

__global__ void __launch_bounds__(256)
k_pyr_v(const float* __restrict__ T, int H, int t_pitch, const float* __restrict__ taps, int ksize,
        double scale_y, float* __restrict__ dst, int Wk, int Hk, int dst_pitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Wk || y >= Hk) return;
    float a1;
    int sy = linear_coord(y, scale_y, H, &a1);
    float a0 = 1.f - a1;
    int sy1 = min(sy + 1, H - 1);
    int c = ksize / 2;
    float b0 = 0.f, b1 = 0.f;
    for (int j = 0; j < ksize; j++) {
        float t = __ldg(taps + j);
        b0 += t * T[(size_t)reflect101(sy + j - c, H) * t_pitch + x];
        if (a1 != 0.f) b1 += t * T[(size_t)reflect101(sy1 + j - c, H) * t_pitch + x];
    }
    dst[(size_t)y * dst_pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}

void launch_pyr_h(Launch& L, const void* frame, int dtype, int W, int H, size_t pitch_bytes,
                  const float* taps, int ksize, float* T, int Wk, int t_pitch)
{
    dim3 block(64, 4), grid(divup(Wk, 64), divup(H, 4));
    double scale_x = 1.0 / ((double)Wk / W);
    if (dtype == 0)
        L.run("pyr_h_u8", [&](cudaStream_t s) {
            k_pyr_h<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)frame, W, H, pitch_bytes, taps, ksize, scale_x, T, Wk, t_pitch);
        });
    else
        L.run("pyr_h_f32", [&](cudaStream_t s) {
            k_pyr_h<float><<<grid, block, 0, s>>>((const float*)frame, W, H, pitch_bytes, taps, ksize, scale_x, T, Wk, t_pitch);
        });
}

void launch_pyr_v(Launch& L, const float* T, int H, int t_pitch, const float* taps, int ksize,
                  float* I, int Wk, int Hk, int i_pitch)
{
    dim3 block(64, 4), grid(divup(Wk, 64), divup(Hk, 4));
    double scale_y = 1.0 / ((double)Hk / H);
    L.run("pyr_v", [&](cudaStream_t s) {
        k_pyr_v<<<grid, block, 0, s>>>(T, H, t_pitch, taps, ksize, scale_y, I, Wk, Hk, i_pitch);
    });
}


// ------------------------------------------------------------------------------------------------
// Batched variants (blockIdx.z = frame of the batch).  The bilinear source index / weight of every
// destination column and row is a property of the level, so the engine builds it once per plan on
// the host (same double-precision rule as linear_coord) instead of per thread.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_pyr_h2(PyrArgs a)
{
    const int x = blockIdx.x * 64 + threadIdx.x, r = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= a.Wk || r >= a.H) return;
    const int W = a.W, ksize = a.ksize, c = ksize / 2;
    const int sx = a.sx[x];
    const float a1 = a.ax[x], a0 = 1.f - a1;
    const T* row = (const T*)((const char*)a.src + (size_t)z * a.src_item + (size_t)r * a.src_pitch);
    const float* __restrict__ taps = a.taps;
    float b0 = 0.f, b1 = 0.f;
    if (sx - c >= 0 && sx + 1 + c < W) {
        const T* p = row + sx - c;
        if (a1 != 0.f) {
            float prev = px_to_f32(p[0]);
            for (int j = 0; j < ksize; j++) {
                float cur = px_to_f32(p[j + 1]);
                float t = __ldg(taps + j);
                b0 += t * prev;
                b1 += t * cur;
                prev = cur;
            }
        } else {
            for (int j = 0; j < ksize; j++) b0 += __ldg(taps + j) * px_to_f32(p[j]);
        }
    } else {
        int sx1 = min(sx + 1, W - 1);
        for (int j = 0; j < ksize; j++) {
            float t = __ldg(taps + j);
            b0 += t * px_to_f32(row[reflect101(sx + j - c, W)]);
            if (a1 != 0.f) b1 += t * px_to_f32(row[reflect101(sx1 + j - c, W)]);
        }
    }
    a.T[(size_t)z * a.t_item + (size_t)r * a.pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}

__global__ void __launch_bounds__(256)
k_pyr_v2(PyrArgs a)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= a.Wk || y >= a.Hk) return;
    const int H = a.H, ksize = a.ksize, c = ksize / 2;
    const int sy = a.sy[y];
    const float a1 = a.ay[y], a0 = 1.f - a1;
    const int sy1 = min(sy + 1, H - 1);
    const float* T = a.T + (size_t)z * a.t_item + x;
    const float* __restrict__ taps = a.taps;
    float b0 = 0.f, b1 = 0.f;
    if (sy - c >= 0 && sy + 1 + c < H) {
        const float* p = T + (size_t)(sy - c) * a.pitch;
        if (a1 != 0.f) {
            float prev = p[0];
            for (int j = 0; j < ksize; j++) {
                float cur = p[(size_t)(j + 1) * a.pitch];
                float t = __ldg(taps + j);
                b0 += t * prev;
                b1 += t * cur;
                prev = cur;
            }
        } else {
            for (int j = 0; j < ksize; j++) b0 += __ldg(taps + j) * p[(size_t)j * a.pitch];
        }
    } else {
        for (int j = 0; j < ksize; j++) {
            float t = __ldg(taps + j);
            b0 += t * T[(size_t)reflect101(sy + j - c, H) * a.pitch];
            if (a1 != 0.f) b1 += t * T[(size_t)reflect101(sy1 + j - c, H) * a.pitch];
        }
    }
    a.I[(size_t)z * a.i_item + (size_t)y * a.pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}

// Unrolled row / column passes for the kernel sizes of pyr_scale = 0.5 (3, 9, 19, 39, 79 and 5): taps come from the
// constant bank, no loop or tap-load overhead.  Same arithmetic and order as k_pyr_h2 / k_pyr_v2 (bit-identical).
template <typename T, int KS>
__global__ void __launch_bounds__(256)
k_pyr_h4(PyrArgs a)
{
    const int x = blockIdx.x * 64 + threadIdx.x, r = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= a.Wk || r >= a.H) return;
    constexpr int c = KS / 2;
    const int W = a.W;
    const int sx = a.sx[x];
    const float a1 = a.ax[x], a0 = 1.f - a1;
    const T* row = (const T*)((const char*)a.src + (size_t)z * a.src_item + (size_t)r * a.src_pitch);
    float b0 = 0.f, b1 = 0.f;
    if (sx - c >= 0 && sx + 1 + c < W) {
        const T* p = row + sx - c;
        if (a1 != 0.f) {
            float prev = px_to_f32(p[0]);
#pragma unroll
            for (int j = 0; j < KS; j++) {
                float cur = px_to_f32(p[j + 1]);
                b0 += a.tapsv[j] * prev;
                b1 += a.tapsv[j] * cur;
                prev = cur;
            }
        } else {
#pragma unroll
            for (int j = 0; j < KS; j++) b0 += a.tapsv[j] * px_to_f32(p[j]);
        }
    } else {
        int sx1 = min(sx + 1, W - 1);
#pragma unroll 1
        for (int j = 0; j < KS; j++) {
            float t = a.taps[j];
            b0 += t * px_to_f32(row[reflect101(sx + j - c, W)]);
            if (a1 != 0.f) b1 += t * px_to_f32(row[reflect101(sx1 + j - c, W)]);
        }
    }
    a.T[(size_t)z * a.t_item + (size_t)r * a.pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}

template <int KS>
__global__ void __launch_bounds__(256)
k_pyr_v4(PyrArgs a)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= a.Wk || y >= a.Hk) return;
    constexpr int c = KS / 2;
    const int H = a.H;
    const int sy = a.sy[y];
    const float a1 = a.ay[y], a0 = 1.f - a1;
    const int sy1 = min(sy + 1, H - 1);
    const float* T = a.T + (size_t)z * a.t_item + x;
    float b0 = 0.f, b1 = 0.f;
    if (sy - c >= 0 && sy + 1 + c < H) {
        const float* p = T + (size_t)(sy - c) * a.pitch;
        if (a1 != 0.f) {
            float prev = p[0];
#pragma unroll
            for (int j = 0; j < KS; j++) {
                float cur = p[(size_t)(j + 1) * a.pitch];
                b0 += a.tapsv[j] * prev;
                b1 += a.tapsv[j] * cur;
                prev = cur;
            }
        } else {
#pragma unroll
            for (int j = 0; j < KS; j++) b0 += a.tapsv[j] * p[(size_t)j * a.pitch];
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < KS; j++) {
            float t = a.taps[j];
            b0 += t * T[(size_t)reflect101(sy + j - c, H) * a.pitch];
            if (a1 != 0.f) b1 += t * T[(size_t)reflect101(sy1 + j - c, H) * a.pitch];
        }
    }
    a.I[(size_t)z * a.i_item + (size_t)y * a.pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}


// ------------------------------------------------------------------------------------------------
// Column-first variant of the unrolled passes (the fast path when rows are 4-pixel aligned).  The row-first
// kernels above spend ~140 instructions per frame pixel, almost all of it single-byte loads: every output column
// re-reads KS+1 neighbouring bytes.  Running the COLUMN pass first turns those into aligned 4-pixel loads shared by
// four outputs, and it shrinks the image to Hk rows before the (strided) row pass runs:
//   k_pyr_vf : T'(y, x) = (1-ay) * Vblur(sy, x) + ay * Vblur(sy+1, x)      for all W source columns, Hk rows
//   k_pyr_hf : I(y, x)  = (1-ax) * Hblur_T'(y, sx) + ax * Hblur_T'(y, sx+1)
// The four operators are linear and separable, so this is the same level image up to f32 rounding order
// (tests/test_gpu_parity.py::test_stage_level_image, 2e-4 on a 0..255 scale).  Because the order already differs from
// cv2's, the taps are applied with explicit FMAs here (half the FP32 instructions of these FMA-pipe-bound kernels,
// and one rounding less per tap); the stages that ARE bit-exact against the oracle stay uncontracted.
// ------------------------------------------------------------------------------------------------
template <typename T> struct Px4;
template <> struct Px4<unsigned char> {
    static __device__ __forceinline__ void load(const unsigned char* p, float v[4])
    {
        const unsigned w = *reinterpret_cast<const unsigned*>(p);
        // byte i of w into the mantissa of 2^23 (one PRMT), minus 2^23: u8 -> f32 without the XU pipe
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | i)) - 8388608.f;
    }
};
template <> struct Px4<float> {
    static __device__ __forceinline__ void load(const float* p, float v[4])
    {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
};

template <typename T, int KS>
__global__ void __launch_bounds__(256)
k_pyr_vf(PyrArgs a)
{
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x4 >= a.W || y >= a.Hk) return;
    constexpr int c = KS / 2;
    const int H = a.H;
    const int sy = a.sy[y];
    const float a1 = a.ay[y], a0 = 1.f - a1;
    const char* base = (const char*)a.src + (size_t)z * a.src_item + (size_t)x4 * sizeof(T);
    float b0[4] = {0.f, 0.f, 0.f, 0.f}, b1[4] = {0.f, 0.f, 0.f, 0.f};
    if (sy - c >= 0 && sy + 1 + c < H) {
        const char* p = base + (size_t)(sy - c) * a.src_pitch;
        if (a1 != 0.f) {
            float prev[4], cur[4];
            Px4<T>::load((const T*)p, prev);
#pragma unroll
            for (int j = 0; j < KS; j++) {
                Px4<T>::load((const T*)(p + (size_t)(j + 1) * a.src_pitch), cur);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    b0[i] = fmaf(a.tapsv[j], prev[i], b0[i]);
                    b1[i] = fmaf(a.tapsv[j], cur[i], b1[i]);
                    prev[i] = cur[i];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < KS; j++) {
                float v[4];
                Px4<T>::load((const T*)(p + (size_t)j * a.src_pitch), v);
#pragma unroll
                for (int i = 0; i < 4; i++) b0[i] = fmaf(a.tapsv[j], v[i], b0[i]);
            }
        }
    } else {
        const int sy1 = min(sy + 1, H - 1);
#pragma unroll 1
        for (int j = 0; j < KS; j++) {
            const float t = a.taps[j];
            float v[4];
            Px4<T>::load((const T*)(base + (size_t)reflect101(sy + j - c, H) * a.src_pitch), v);
#pragma unroll
            for (int i = 0; i < 4; i++) b0[i] = fmaf(t, v[i], b0[i]);
            if (a1 != 0.f) {
                Px4<T>::load((const T*)(base + (size_t)reflect101(sy1 + j - c, H) * a.src_pitch), v);
#pragma unroll
                for (int i = 0; i < 4; i++) b1[i] = fmaf(t, v[i], b1[i]);
            }
        }
    }
    float4 o;
    if (a1 != 0.f) o = make_float4(fmaf(b1[0], a1, b0[0] * a0), fmaf(b1[1], a1, b0[1] * a0), fmaf(b1[2], a1, b0[2] * a0), fmaf(b1[3], a1, b0[3] * a0));
    else o = make_float4(b0[0], b0[1], b0[2], b0[3]);
    *reinterpret_cast<float4*>(a.T + (size_t)z * a.t_item + (size_t)y * a.W + x4) = o;       // T': Hk rows of W floats
}

template <int KS>
__global__ void __launch_bounds__(256)
k_pyr_hf(PyrArgs a)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= a.Wk || y >= a.Hk) return;
    constexpr int c = KS / 2;
    const int W = a.W;
    const int sx = a.sx[x];
    const float a1 = a.ax[x], a0 = 1.f - a1;
    const float* row = a.T + (size_t)z * a.t_item + (size_t)y * W;
    float b0 = 0.f, b1 = 0.f;
    if (sx - c >= 0 && sx + 1 + c < W) {
        const float* p = row + sx - c;
        if (a1 != 0.f) {
            float prev = p[0];
#pragma unroll
            for (int j = 0; j < KS; j++) {
                float cur = p[j + 1];
                b0 = fmaf(a.tapsv[j], prev, b0);
                b1 = fmaf(a.tapsv[j], cur, b1);
                prev = cur;
            }
        } else {
#pragma unroll
            for (int j = 0; j < KS; j++) b0 = fmaf(a.tapsv[j], p[j], b0);
        }
    } else {
        const int sx1 = min(sx + 1, W - 1);
#pragma unroll 1
        for (int j = 0; j < KS; j++) {
            const float t = a.taps[j];
            b0 = fmaf(t, row[reflect101(sx + j - c, W)], b0);
            if (a1 != 0.f) b1 = fmaf(t, row[reflect101(sx1 + j - c, W)], b1);
        }
    }
    a.I[(size_t)z * a.i_item + (size_t)y * a.pitch + x] = (a1 != 0.f) ? b0 * a0 + b1 * a1 : b0;
}

static bool pyr_colfirst_ok(int dtype, const PyrArgs& a)
{
    const size_t al = dtype == 0 ? 4 : 16;            // bytes of one 4-pixel load
    return a.W % 4 == 0 && a.src_pitch % al == 0 && a.src_item % al == 0 && ((uintptr_t)a.src % al) == 0 &&
           (size_t)a.Hk * a.W <= a.t_cap && a.t_item % 4 == 0 && ((uintptr_t)a.T % 16) == 0;
}

template <int KS>
static void run_pyr4(Launch& L, int dtype, const PyrArgs& a, int batch)
{
    dim3 block(64, 4);
    if (pyr_colfirst_ok(dtype, a)) {
        dim3 gv(divup(a.W / 4, 64), divup(a.Hk, 4), batch), gh(divup(a.Wk, 64), divup(a.Hk, 4), batch);
        L.run(dtype == 0 ? "pyr_v_u8" : "pyr_v_f32", [&](cudaStream_t s) {
            if (dtype == 0) k_pyr_vf<uint8_t, KS><<<gv, block, 0, s>>>(a);
            else k_pyr_vf<float, KS><<<gv, block, 0, s>>>(a);
        });
        L.run("pyr_h", [&](cudaStream_t s) { k_pyr_hf<KS><<<gh, block, 0, s>>>(a); });
        return;
    }
    dim3 gh(divup(a.Wk, 64), divup(a.H, 4), batch), gv(divup(a.Wk, 64), divup(a.Hk, 4), batch);
    L.run(dtype == 0 ? "pyr_h_u8" : "pyr_h_f32", [&](cudaStream_t s) {
        if (dtype == 0) k_pyr_h4<uint8_t, KS><<<gh, block, 0, s>>>(a);
        else k_pyr_h4<float, KS><<<gh, block, 0, s>>>(a);
    });
    L.run("pyr_v", [&](cudaStream_t s) { k_pyr_v4<KS><<<gv, block, 0, s>>>(a); });
}

void launch_pyr2(Launch& L, int dtype, const PyrArgs& a, int batch)
{
    switch (a.ksize) {
        case 3: run_pyr4<3>(L, dtype, a, batch); return;
        case 5: run_pyr4<5>(L, dtype, a, batch); return;
        case 9: run_pyr4<9>(L, dtype, a, batch); return;
        case 19: run_pyr4<19>(L, dtype, a, batch); return;
        case 39: run_pyr4<39>(L, dtype, a, batch); return;
        case 79: run_pyr4<79>(L, dtype, a, batch); return;
        default: break;
    }
    dim3 block(64, 4);
    dim3 gh(divup(a.Wk, 64), divup(a.H, 4), batch), gv(divup(a.Wk, 64), divup(a.Hk, 4), batch);
    L.run(dtype == 0 ? "pyr_h_u8" : "pyr_h_f32", [&](cudaStream_t s) {
        if (dtype == 0) k_pyr_h2<uint8_t><<<gh, block, 0, s>>>(a);
        else k_pyr_h2<float><<<gh, block, 0, s>>>(a);
    });
    L.run("pyr_v", [&](cudaStream_t s) { k_pyr_v2<<<gv, block, 0, s>>>(a); });
}


// ------------------------------------------------------------------------------------------------
// k_pyr_fused: the level images of scales 1..3 of ONE frame tile in one pass over the frame (pyr_scale = 0.5, frame sizes that
// are multiples of 8, u8 rows 4-byte aligned).  The per-level kernels above read the whole u8 frame once per level
// (ncu: 3 x 2.07 MB per 1080p frame, plus the Hk x W intermediate written and read back) and were latency-bound on small
// grids; here a CTA stages a 144 x 76 pixel patch of the frame in shared memory ONCE (as f32, reflect-101 applied while
// staging, so the passes below have no border arithmetic) and produces from it the 64 x 32 tile of I_1, the 32 x 16 tile of
// I_2 and the 16 x 8 tile of I_3 that depend on it.  Per level: the column pass at the sampled rows for every patch column
// (4 columns per thread, 16-byte shared loads), a barrier, the row pass at the sampled columns.
// Arithmetic and order are those of k_pyr_vf / k_pyr_hf (FMA per tap, the two bilinear weights at the end): bit-identical.
// ------------------------------------------------------------------------------------------------
constexpr int PF_TX = 128, PF_TY = 64;           // frame pixels per CTA
constexpr int PF_HX = 8, PF_HY = 6;              // halo: level 3 needs c + 1 = 10 columns beyond its sampled column 8x+3 -> 6 (+2 to stay 4-aligned)
constexpr int PF_SW = PF_TX + 2 * PF_HX, PF_SH = PF_TY + 2 * PF_HY + 0;     // 144 x 76 staged pixels
constexpr int PF_TP = PF_SW + 1;                 // pitch of the column-pass result (odd)
constexpr int PF_THREADS = 256;

struct PyrFusedArgs {
    const void* src; size_t src_item, src_pitch;
    int W, H, nlev;
    int Wk[3], Hk[3], pitch[3];
    const int* sx[3]; const float* ax[3]; const int* sy[3]; const float* ay[3];
    float* I[3]; size_t i_item[3];
    float taps1[3], taps2[9], taps3[19];
};

template <int KS>
__device__ __forceinline__ void pf_level(const PyrFusedArgs& a, const float (&taps)[KS], const int lv, const float* sS, float* sT,
                                         const int cx0, const int cy0, const int bx, const int by, const int z)
{
    constexpr int c = KS / 2;
    const int tw = PF_TX >> (lv + 1), th = PF_TY >> (lv + 1);           // output tile of this level
    const int X0 = bx * tw, Y0 = by * th;
    const int tid = threadIdx.x;
    // column pass: item = (output row, group of 4 patch columns)
#pragma unroll 2
    for (int i = tid; i < th * (PF_SW / 4); i += PF_THREADS) {
        const int yl = i / (PF_SW / 4), xg = i - yl * (PF_SW / 4);
        const int Y = Y0 + yl;
        if (Y >= a.Hk[lv]) continue;
        const int sy = a.sy[lv][Y];
        const float a1 = a.ay[lv][Y], a0 = 1.f - a1;
        const float* p = sS + (sy - c - cy0) * PF_SW + 4 * xg;
        float b0[4] = {0.f, 0.f, 0.f, 0.f}, b1[4] = {0.f, 0.f, 0.f, 0.f};
        float4 o;
        if (a1 != 0.f) {
            float4 prev = *reinterpret_cast<const float4*>(p);
#pragma unroll
            for (int j = 0; j < KS; j++) {
                const float4 cur = *reinterpret_cast<const float4*>(p + (j + 1) * PF_SW);
                b0[0] = fmaf(taps[j], prev.x, b0[0]); b0[1] = fmaf(taps[j], prev.y, b0[1]);
                b0[2] = fmaf(taps[j], prev.z, b0[2]); b0[3] = fmaf(taps[j], prev.w, b0[3]);
                b1[0] = fmaf(taps[j], cur.x, b1[0]); b1[1] = fmaf(taps[j], cur.y, b1[1]);
                b1[2] = fmaf(taps[j], cur.z, b1[2]); b1[3] = fmaf(taps[j], cur.w, b1[3]);
                prev = cur;
            }
            o = make_float4(fmaf(b1[0], a1, b0[0] * a0), fmaf(b1[1], a1, b0[1] * a0), fmaf(b1[2], a1, b0[2] * a0), fmaf(b1[3], a1, b0[3] * a0));
        } else {
#pragma unroll
            for (int j = 0; j < KS; j++) {
                const float4 v = *reinterpret_cast<const float4*>(p + j * PF_SW);
                b0[0] = fmaf(taps[j], v.x, b0[0]); b0[1] = fmaf(taps[j], v.y, b0[1]);
                b0[2] = fmaf(taps[j], v.z, b0[2]); b0[3] = fmaf(taps[j], v.w, b0[3]);
            }
            o = make_float4(b0[0], b0[1], b0[2], b0[3]);
        }
        float* t = sT + yl * PF_TP + 4 * xg;
        t[0] = o.x; t[1] = o.y; t[2] = o.z; t[3] = o.w;
    }
    __syncthreads();
    // row pass: item = output pixel, lanes along x (coalesced stores)
    float* out = a.I[lv] + (size_t)z * a.i_item[lv];
#pragma unroll 2
    for (int i = tid; i < th * tw; i += PF_THREADS) {
        const int yl = i / tw, xl = i - yl * tw;
        const int X = X0 + xl, Y = Y0 + yl;
        if (X >= a.Wk[lv] || Y >= a.Hk[lv]) continue;
        const int sx = a.sx[lv][X];
        const float a1 = a.ax[lv][X], a0 = 1.f - a1;
        const float* p = sT + yl * PF_TP + (sx - c - cx0);
        float b0 = 0.f, b1 = 0.f, r;
        if (a1 != 0.f) {
            float prev = p[0];
#pragma unroll
            for (int j = 0; j < KS; j++) {
                const float cur = p[j + 1];
                b0 = fmaf(taps[j], prev, b0);
                b1 = fmaf(taps[j], cur, b1);
                prev = cur;
            }
            r = __fadd_rn(__fmul_rn(b0, a0), __fmul_rn(b1, a1));
        } else {
#pragma unroll
            for (int j = 0; j < KS; j++) b0 = fmaf(taps[j], p[j], b0);
            r = b0;
        }
        out[(size_t)Y * a.pitch[lv] + X] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PF_THREADS)
k_pyr_fused(PyrFusedArgs a)
{
    extern __shared__ __align__(16) float pf_smem[];
    float* sS = pf_smem;                          // PF_SH x PF_SW staged frame patch (f32)
    float* sT = pf_smem + PF_SH * PF_SW;          // column-pass result of the current level: <= 32 x PF_TP
    const int bx = blockIdx.x, by = blockIdx.y, z = blockIdx.z, tid = threadIdx.x;
    const int cx0 = bx * PF_TX - PF_HX, cy0 = by * PF_TY - PF_HY;
    const int W = a.W, H = a.H;
    const unsigned char* src = (const unsigned char*)a.src + (size_t)z * a.src_item;
    const bool inside = cx0 >= 0 && cx0 + PF_SW <= W && cy0 >= 0 && cy0 + PF_SH <= H;
    if (inside) {
        constexpr int NITEM = PF_SH * (PF_SW / 4), NIT = (NITEM + PF_THREADS - 1) / PF_THREADS;
        unsigned wv[NIT];
#pragma unroll
        for (int k = 0; k < NIT; k++) {                     // every load of the thread in flight before the first use
            const int i = tid + k * PF_THREADS;
            if (i < NITEM) {
                const int r = i / (PF_SW / 4), v = i - r * (PF_SW / 4);
                wv[k] = *reinterpret_cast<const unsigned*>(src + (size_t)(cy0 + r) * a.src_pitch + cx0 + 4 * v);
            }
        }
#pragma unroll
        for (int k = 0; k < NIT; k++) {
            const int i = tid + k * PF_THREADS;
            if (i < NITEM) {
                const int r = i / (PF_SW / 4), v = i - r * (PF_SW / 4);
                const unsigned w = wv[k];
                float4 o;
                o.x = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u)) - 8388608.f;
                o.y = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441u)) - 8388608.f;
                o.z = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442u)) - 8388608.f;
                o.w = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7443u)) - 8388608.f;
                *reinterpret_cast<float4*>(sS + r * PF_SW + 4 * v) = o;
            }
        }
    } else {
        for (int i = tid; i < PF_SH * PF_SW; i += PF_THREADS) {
            const int r = i / PF_SW, cc = i - r * PF_SW;
            const int fy = reflect101(cy0 + r, H), fx = reflect101(cx0 + cc, W);
            sS[i] = u8_to_f32(src[(size_t)fy * a.src_pitch + fx]);
        }
    }
    __syncthreads();
    pf_level<3>(a, a.taps1, 0, sS, sT, cx0, cy0, bx, by, z);
    if (a.nlev > 1) pf_level<9>(a, a.taps2, 1, sS, sT, cx0, cy0, bx, by, z);
    if (a.nlev > 2) pf_level<19>(a, a.taps3, 2, sS, sT, cx0, cy0, bx, by, z);
}

// Conditions under which the sampled rows / columns of every tile stay inside the staged patch: pyr_scale 0.5 on frame sizes
// that are multiples of 8 gives the affine tables sx = 2x (a 1/2), 4x+1, 8x+3 that the halo constants above are derived from.
bool pyr_fused_supported(int dtype, int W, int H, double pyr_scale, const void* src, size_t src_pitch, size_t src_item)
{
    return dtype == 0 && pyr_scale == 0.5 && W % 8 == 0 && H % 8 == 0 && W >= 64 && H >= 64 && src_pitch % 4 == 0 && src_item % 4 == 0 &&
           ((uintptr_t)src % 4) == 0;
}

void launch_pyr_fused(Launch& L, const PyrFusedLaunch& f, int batch)
{
    PyrFusedArgs a{};
    a.src = f.src; a.src_item = f.src_item; a.src_pitch = f.src_pitch; a.W = f.W; a.H = f.H; a.nlev = f.nlev;
    for (int k = 0; k < f.nlev; k++) {
        a.Wk[k] = f.Wk[k]; a.Hk[k] = f.Hk[k]; a.pitch[k] = f.pitch[k];
        a.sx[k] = f.sx[k]; a.ax[k] = f.ax[k]; a.sy[k] = f.sy[k]; a.ay[k] = f.ay[k];
        a.I[k] = f.I[k]; a.i_item[k] = f.i_item[k];
    }
    for (int j = 0; j < 3; j++) a.taps1[j] = f.taps[0][j];
    if (f.nlev > 1) for (int j = 0; j < 9; j++) a.taps2[j] = f.taps[1][j];
    if (f.nlev > 2) for (int j = 0; j < 19; j++) a.taps3[j] = f.taps[2][j];
    const size_t smem = sizeof(float) * (PF_SH * PF_SW + 32 * PF_TP);
    static unsigned long long configured = 0;
    L.dyn_smem(k_pyr_fused, smem, configured);
    dim3 grid(divup(f.W, PF_TX), divup(f.H, PF_TY), batch);
    L.run("pyr_fused", [&](cudaStream_t s) { k_pyr_fused<<<grid, PF_THREADS, smem, s>>>(a); });
}

}  // namespace ofb
