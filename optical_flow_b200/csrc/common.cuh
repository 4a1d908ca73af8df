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

// Polynomial-expansion coefficients R of one image: channels 0..3 (d/dy, d/dx, yy, xx) interleaved as one
// float4 per pixel, channel 4 (xy) as a separate plane.  UpdateMatrices gathers R1 at 4 neighbours: with this
// layout that is 4 x (LDG.128 + LDG.32) instead of 20 scalar loads, and a warp still reads whole 128-byte lines.
struct RView {
    float4* a;                  // H x pitch float4
    float* b;                   // H x pitch float
    int pitch;                  // pixels between rows (multiple of 32)
};

// Batched views: R of one pyramid level for a ring of frame slots; pair z of a batch reads the slots
// (slot0+z) % nslots and (slot0+z+1) % nslots.  A slot is 5*plane floats: 4*plane of float4 data, then plane of ch 4.
struct SlotRing {
    float* base;                // slot s at base + s * slot_stride
    size_t slot_stride;         // floats (= 5 * plane)
    size_t plane;               // pixels per plane (H * pitch)
    int pitch;
    int nslots;
    int step;                   // slot distance between consecutive batch items: 1 inside a shot (pair z = frames z, z+1),
                                // 2 for independent pairs (pair z = slots 2z, 2z+1)
    __host__ __device__ RView slot(int s) const
    {
        float* p = base + (size_t)s * slot_stride;
        return RView{reinterpret_cast<float4*>(p), p + 4 * plane, pitch};
    }
    __host__ __device__ int wrap(int s) const { return s >= nslots ? s - nslots : s; }   // s < 2 * nslots
    __host__ __device__ int first(int slot0, int z) const { return wrap(slot0 + z * step); }
};

// first UpdateMatrices of a scale (iter.cu k_um0); batch item z
struct Um0Args {
    const float2* flow; size_t flow_item;   // SRC 1: flow of this scale; SRC 2: coarser flow (Wp x Hp); per-item stride in float2
    int Wp, Hp; float mul;
    const int* ux; const float* uax;        // SRC 2: bilinear tables coarser -> this scale (per column / per row)
    const int* uy; const float* uay;
    SlotRing R; int slot0;
    float* M; size_t m_item, plane; int pitch;
    int W, H;
};

// blur + solve (+ UpdateMatrices) (iter.cu k_iter); batch item z
struct IterArgs {
    const float* Min; float* Mout; size_t m_item, plane; int pitch;
    SlotRing R; int slot0;
    float2* flow; size_t flow_item;
    int W, H, strip_rows; float c;          // c = 1e-3 * winsize^4
    double c64;                             // the same in f64 (k_iter64)
    int prefetch;                           // 1 = software-prefetch the next step's lines into L2
    int reverse;                            // 1 = walk the grid backwards: the tiles the previous launch wrote last (still in L2) are read first
    unsigned* minmax;                       // non-null (FUSE = false only): fold min / max of |flow| of item z into minmax[2z..]
    int gauss;                              // 1 = Gaussian window (flags & 256): taps gk[0..M], c = 1e-3
    float gk[17];
};

// polynomial expansion (polyexp.cu); batch item z = frame
struct PolyArgs {
    const void* src; size_t src_item;       // bytes between batch items
    size_t src_pitch;                       // bytes between rows
    int W, H;
    SlotRing R; int slot0;                  // output slot of frame z: (slot0 + z * R.step) % nslots
    float g[9], xg[9], xxg[9];              // taps k = 0..n (f32, as cv2 builds them)
    double gd[9], xgd[9], xxgd[9];          // the same values widened (exact)
    double ig11, ig03, ig33, ig55;
    int n;
};

// level image of scale k >= 1 (pyramid.cu, batched); T is the row-pass intermediate (H x Wk), I the result
struct PyrArgs {
    const void* src; size_t src_item, src_pitch;    // frames: bytes between batch items / rows
    int W, H, Wk, Hk, ksize;
    const float* taps;
    const int* sx; const float* ax;                 // per destination column: source index, weight of the next one
    const int* sy; const float* ay;                 // per destination row
    float* T; size_t t_item;                        // floats
    size_t t_cap;                                   // floats available per item of T (the column-first path stores Hk x W there)
    float* I; size_t i_item;
    int pitch;                                      // row pitch of T and I (floats)
    float tapsv[80];                                // the same taps by value (unrolled kernels read them from the constant bank)
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

// u8 -> f32 without an I2F (which runs on the 16-lane XU pipe): place the byte in the mantissa of 2^23.
__device__ __forceinline__ float u8_to_f32(unsigned int b) { return __uint_as_float(0x4B000000u | b) - 8388608.f; }
template <typename T> __device__ __forceinline__ float px_to_f32(T v) { return (float)v; }
template <> __device__ __forceinline__ float px_to_f32<unsigned char>(unsigned char v) { return u8_to_f32(v); }

// Source coordinate of cv::resize(INTER_LINEAR): returns the left/top sample index, writes the f32 weight of
// the next one.  f32_coord = false: coordinate kept in double (cv2's one-channel f32 resize, the level images); true: rounded
// to f32 before the floor (cv2's generic path, which the two-channel flow up-sample takes) -- see linear_table in engine.cu.
__device__ inline int linear_coord(int d, double scale, int src_len, float* w1, bool f32_coord = false)
{
    double f = (d + 0.5) * scale - 0.5;
    if (f32_coord) f = (double)(float)f;
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

void launch_pyr2(Launch& L, int dtype, const PyrArgs& a, int batch);
// scales 1..nlev (<= 3) of a batch of u8 frames in one pass over each frame (pyr_scale 0.5, sizes multiple of 8)
struct PyrFusedLaunch {
    const void* src; size_t src_item, src_pitch;
    int W, H, nlev;
    int Wk[3], Hk[3], pitch[3];
    const int* sx[3]; const float* ax[3]; const int* sy[3]; const float* ay[3];
    float* I[3]; size_t i_item[3];
    const float* taps[3];               // host pointers: 3, 9 and 19 taps
};
bool pyr_fused_supported(int dtype, int W, int H, double pyr_scale, const void* src, size_t src_pitch, size_t src_item);
void launch_pyr_fused(Launch& L, const PyrFusedLaunch& f, int batch);

// polyexp.cu -- A.5/A.6
struct PolyConst { const float* g; const float* xg; const float* xxg; int n; double ig11, ig03, ig33, ig55; };
void launch_polyexp(Launch& L, const float* I, int W, int H, int pitch, const PolyConst& pc,
                    float* tmp3 /* 3 planes */, RView R, bool generic);
// batched, unrolled variant: src_kind 0 = level image (f32), 1 = u8 frame + fused [1/4 1/2 1/4]^2 pre-blur
// (scale 0), 2 = f32 frame + the same pre-blur.  Returns false when poly_n has no unrolled instance.
bool polyexp2_supported(int n);
void launch_polyexp2(Launch& L, int src_kind, const PolyArgs& a, int batch);

// matrices.cu -- A.2 and A.8
void launch_upsample_flow(Launch& L, const float2* prev, int Wp, int Hp, float2* flow, int W, int H, float mul);
void launch_area_flow(Launch& L, const float2* src, int Ws, int Hs, float2* dst, int Wd, int Hd, float mul);
void launch_scale_flow(Launch& L, float2* flow, size_t n, float mul);
void launch_update_matrices(Launch& L, RView R0, RView R1, const float2* flow, int W, int H, Planes5 M);
void launch_r_interleave(Launch& L, RView src, int W, int H, float* dst /* (H,W,5) */);
void launch_r_deinterleave(Launch& L, const float* src /* (H,W,5) */, int W, int H, RView dst);
void launch_interleave5(Launch& L, Planes5 src, int W, int H, float* dst /* (H,W,5) */);
void launch_deinterleave5(Launch& L, const float* src /* (H,W,5) */, int W, int H, Planes5 dst);

// iter.cu -- A.8 / A.9 / A.11 batched
void launch_um0(Launch& L, int src, const Um0Args& a, int batch);
bool iter_supported(int winsize);
void launch_iter(Launch& L, const IterArgs& a, int winsize, bool fuse_um, int batch);

// blur_solve.cu -- A.9 / A.10
void launch_blur_solve_box(Launch& L, Planes5 M, int W, int H, int winsize, double* tmp /* 5 planes f64 */,
                           float2* flow, bool generic);
void launch_blur_solve_gauss(Launch& L, Planes5 M, int W, int H, int winsize, const float* half_taps,
                             float* tmp /* 5 planes f32 */, float2* flow, bool generic);

// preprocess.cu -- SURVEY.md 8f row N2: BGR->gray and the 8-bit bilinear resize, batched over frames
struct ResizeTab {             // per destination column / row: the two source indices and their 11-bit weights
    const int* x0; const int* x1; const short* ax;     // ax[2x], ax[2x+1]
    const int* y0; const int* y1; const short* ay;
};
void launch_bgr2gray(Launch& L, const uint8_t* src, size_t src_item, size_t src_pitch, uint8_t* dst, size_t dst_item,
                     size_t dst_pitch, int W, int H, int batch);
void launch_resize_u8(Launch& L, const uint8_t* src, size_t src_item, size_t src_pitch, int cn, bool to_gray, uint8_t* dst,
                      size_t dst_item, size_t dst_pitch, int dW, int dH, const ResizeTab& t, int batch);

// viz.cu -- Appendix B
// batched over pairs: flow / bgr / minmax / sums advance by *_item per batch element
// table: 65536-entry (H byte << 8 | V byte) -> B | G<<8 | R<<16 lookup built by launch_build_hsv_table, or nullptr (arithmetic)
void launch_build_hsv_table(Launch& L, unsigned* table);
void launch_picture_batch(Launch& L, const float2* flow, size_t flow_item, size_t n, unsigned* minmax /* 2 per item */,
                          uint8_t* bgr, size_t bgr_item, int batch, bool minmax_done = false, const unsigned* table = nullptr);
void launch_minmax_reset_batch(Launch& L, unsigned* minmax, int batch);
void launch_sum_magnitude_batch(Launch& L, const float2* flow, size_t flow_item, size_t n, double* acc /* 1 per item */,
                                float* out /* 1 per item */, int batch);
void launch_minmax_mag(Launch& L, const float2* flow, size_t n, unsigned* minmax /* [2], pre-set */);
void launch_minmax_reset(Launch& L, unsigned* minmax);
void launch_flow_to_bgr(Launch& L, const float2* flow, size_t n, const unsigned* minmax, uint8_t* bgr, const unsigned* table = nullptr);
void launch_cart_to_polar(Launch& L, const float2* flow, size_t n, float* mag, float* ang, bool degrees = false);
void launch_sum_magnitude(Launch& L, const float2* flow, size_t n, double* acc /* pre-zeroed */, float* out);

}  // namespace ofb
