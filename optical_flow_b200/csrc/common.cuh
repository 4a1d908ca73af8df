// common.cuh -- shared declarations of the sm_100a Farneback engine (internal; the public ABI is
// include/optflow_b200.h).
//
// Data layout in HBM:
//   frames        u8 or f32, row pitch in bytes
//   level images  f32, row pitch `pitch` floats (multiple of 32 -> 128-byte aligned rows)
//   R, M          5 PLANAR f32 planes (plane stride = H*pitch floats).  cv2 interleaves 5 floats per
//                 pixel (20-byte pixels, hostile to vector loads); planes keep every warp access a
//                 run of consecutive floats.  Channel order is cv2's:
//                   R: 0 = d/dy, 1 = d/dx, 2 = yy, 3 = xx, 4 = xy     M: G11, G12, G22, h1, h2
//   flow          float2 interleaved, tightly packed (H, W, 2) -- exactly the array cv2 returns
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace ofb {

struct Planes5 {          // five planar f32 images
    float* base;
    size_t plane;         // floats between planes
    int pitch;            // floats between rows
    __host__ __device__ float* ch(int c) const { return base + (size_t)c * plane; }
};

static inline int divup(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return divup(a, b) * b; }

// BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba); loops because the pre-blur kernel can be wider
// than a tiny frame.
__host__ __device__ inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

// Source coordinate of cv::resize(INTER_LINEAR), evaluated in double like the
// installed wheel does: returns the left/top sample index, writes the f32 weight of the next one.
__device__ inline int linear_coord(int d, double scale, int src_len, float* w1)
{
    double f = (d + 0.5) * scale - 0.5;
    int s = (int)floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0; }
    if (s >= src_len - 1) { s = src_len - 1; f = 0; }
    *w1 = (float)f;
    return s;
}

// ---- kernel launchers (one translation unit per kernel family) -----------------------------
struct Launch;   // profiler hook, engine.cu

// pyramid.cu -- A.3: convertTo(f32) -> GaussianBlur(reflect-101) -> resize(INTER_LINEAR)
void launch_pyr_h(Launch& L, const void* frame, int dtype, int W, int H, size_t pitch_bytes,
                  const float* taps, int ksize, float* T, int Wk, int t_pitch);
void launch_pyr_v(Launch& L, const float* T, int H, int t_pitch, const float* taps, int ksize,
                  float* I, int Wk, int Hk, int i_pitch);

// polyexp.cu -- A.5/A.6
struct PolyConst { const float* g; const float* xg; const float* xxg; int n; double ig11, ig03, ig33, ig55; };
void launch_polyexp(Launch& L, const float* I, int W, int H, int pitch, const PolyConst& pc,
                    float* tmp3 /* 3 planes */, Planes5 R, bool generic);

// matrices.cu -- A.2 and A.8
void launch_upsample_flow(Launch& L, const float2* prev, int Wp, int Hp, float2* flow, int W, int H, float mul);
void launch_area_flow(Launch& L, const float2* src, int Ws, int Hs, float2* dst, int Wd, int Hd, float mul);
void launch_scale_flow(Launch& L, float2* flow, size_t n, float mul);
void launch_update_matrices(Launch& L, Planes5 R0, Planes5 R1, const float2* flow, int W, int H, Planes5 M);
void launch_interleave5(Launch& L, Planes5 src, int W, int H, float* dst /* (H,W,5) */);
void launch_deinterleave5(Launch& L, const float* src /* (H,W,5) */, int W, int H, Planes5 dst);

// blur_solve.cu -- A.9 / A.10
void set_sm_count(int n);
void launch_blur_solve_box(Launch& L, Planes5 M, int W, int H, int winsize, double* tmp /* 5 planes f64 */,
                           float2* flow, bool generic);
void launch_blur_solve_gauss(Launch& L, Planes5 M, int W, int H, int winsize, const float* half_taps,
                             float* tmp /* 5 planes f32 */, float2* flow, bool generic);

// viz.cu -- Appendix B
void launch_minmax_mag(Launch& L, const float2* flow, size_t n, unsigned* minmax /* [2], pre-set */);
void launch_minmax_reset(Launch& L, unsigned* minmax);
void launch_flow_to_bgr(Launch& L, const float2* flow, size_t n, const unsigned* minmax, uint8_t* bgr);
void launch_cart_to_polar(Launch& L, const float2* flow, size_t n, float* mag, float* ang);
void launch_sum_magnitude(Launch& L, const float2* flow, size_t n, double* acc /* pre-zeroed */, float* out);

}  // namespace ofb
