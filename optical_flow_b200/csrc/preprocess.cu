// preprocess.cu -- the frame preprocessing either side of the hot path, so that the decoded
// BGR frame can be uploaded as it is and never makes a second pass over host memory:
//   k_bgr2gray     cv2.cvtColor(frame, COLOR_BGR2GRAY)      /root/reference/optical_flow.py:44,
//                                                           /root/reference/visualize_optical_flow.py:31,35
//   k_resize_u8    cv2.resize(frame, (w, h)) [INTER_LINEAR] /root/reference/optical_flow.py:25-31
//                  (<3, true>: resize of the BGR frame fused with the gray conversion that follows it, :42-44)
// Both are integer / fixed-point algorithms and are reproduced BIT-EXACTLY (oracle/preprocess_oracle.c,
// tests/golden/preprocess.npz):
//   gray  = (3735*B + 19235*G + 9798*R + 16384) >> 15
//   resize: 11-bit weights from an f32 coordinate (host tables, engine.cu resize_table), row pass in 32-bit ints,
//           column pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
// Roofline: HBM (3 B/px read, 1 B/px written); batched over frames (blockIdx.z).
#include "common.cuh"
#include "launch.cuh"

namespace ofb {

__device__ __forceinline__ unsigned gray_of(unsigned b, unsigned g, unsigned r)
{
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

// 4 pixels per thread: three aligned 32-bit loads (12 bytes of BGR), one 32-bit store.
__global__ void __launch_bounds__(256)
k_bgr2gray_v4(const uint8_t* __restrict__ src, size_t src_item, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_item,
              size_t dst_pitch, int W4, int H)
{
    const int x4 = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x4 >= W4 || y >= H) return;
    const unsigned* p = reinterpret_cast<const unsigned*>(src + (size_t)z * src_item + (size_t)y * src_pitch) + 3 * x4;
    const unsigned w0 = p[0], w1 = p[1], w2 = p[2];       // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
    const unsigned g0 = gray_of(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
    const unsigned g1 = gray_of(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
    const unsigned g2 = gray_of((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
    const unsigned g3 = gray_of((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
    *reinterpret_cast<unsigned*>(dst + (size_t)z * dst_item + (size_t)y * dst_pitch + 4 * x4) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

__global__ void __launch_bounds__(256)
k_bgr2gray(const uint8_t* __restrict__ src, size_t src_item, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_item,
           size_t dst_pitch, int W, int H)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= W || y >= H) return;
    const uint8_t* p = src + (size_t)z * src_item + (size_t)y * src_pitch + 3 * x;
    dst[(size_t)z * dst_item + (size_t)y * dst_pitch + x] = (uint8_t)gray_of(p[0], p[1], p[2]);
}

void launch_bgr2gray(Launch& L, const uint8_t* src, size_t src_item, size_t src_pitch, uint8_t* dst, size_t dst_item,
                     size_t dst_pitch, int W, int H, int batch)
{
    dim3 block(64, 4);
    const bool v4 = (W & 3) == 0 && (src_pitch & 3) == 0 && (src_item & 3) == 0 && (dst_pitch & 3) == 0 && (dst_item & 3) == 0 &&
                    ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0;
    L.run("bgr2gray", [&](cudaStream_t s) {
        if (v4) k_bgr2gray_v4<<<dim3(divup(W / 4, 64), divup(H, 4), batch), block, 0, s>>>(src, src_item, src_pitch, dst, dst_item, dst_pitch, W / 4, H);
        else k_bgr2gray<<<dim3(divup(W, 64), divup(H, 4), batch), block, 0, s>>>(src, src_item, src_pitch, dst, dst_item, dst_pitch, W, H);
    });
}

// One destination pixel per thread, all CN channels; GRAY (CN = 3): the three resized channels go straight into the
// gray formula and one byte is written.
template <int CN, bool GRAY>
__global__ void __launch_bounds__(256)
k_resize_u8(const uint8_t* __restrict__ src, size_t src_item, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_item,
            size_t dst_pitch, int dW, int dH, ResizeTab t)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
    if (x >= dW || y >= dH) return;
    const int xa = t.x0[x] * CN, xb = t.x1[x] * CN;
    const int a0 = t.ax[2 * x], a1 = t.ax[2 * x + 1], b0 = t.ay[2 * y], b1 = t.ay[2 * y + 1];
    const uint8_t* r0 = src + (size_t)z * src_item + (size_t)t.y0[y] * src_pitch;
    const uint8_t* r1 = src + (size_t)z * src_item + (size_t)t.y1[y] * src_pitch;
    unsigned v[CN];
#pragma unroll
    for (int c = 0; c < CN; c++) {
        const int S0 = r0[xa + c] * a0 + r0[xb + c] * a1;
        const int S1 = r1[xa + c] * a0 + r1[xb + c] * a1;
        const int o = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v[c] = (unsigned)min(max(o, 0), 255);
    }
    uint8_t* o = dst + (size_t)z * dst_item + (size_t)y * dst_pitch;
    if (GRAY) {
        o[x] = (uint8_t)gray_of(v[0], v[CN > 1 ? 1 : 0], v[CN > 2 ? 2 : 0]);
    } else {
#pragma unroll
        for (int c = 0; c < CN; c++) o[x * CN + c] = (uint8_t)v[c];
    }
}

void launch_resize_u8(Launch& L, const uint8_t* src, size_t src_item, size_t src_pitch, int cn, bool to_gray, uint8_t* dst,
                      size_t dst_item, size_t dst_pitch, int dW, int dH, const ResizeTab& t, int batch)
{
    dim3 block(64, 4), grid(divup(dW, 64), divup(dH, 4), batch);
    L.run(to_gray ? "resize_bgr_gray" : "resize_u8", [&](cudaStream_t s) {
        if (cn == 3 && to_gray) k_resize_u8<3, true><<<grid, block, 0, s>>>(src, src_item, src_pitch, dst, dst_item, dst_pitch, dW, dH, t);
        else if (cn == 3) k_resize_u8<3, false><<<grid, block, 0, s>>>(src, src_item, src_pitch, dst, dst_item, dst_pitch, dW, dH, t);
        else k_resize_u8<1, false><<<grid, block, 0, s>>>(src, src_item, src_pitch, dst, dst_item, dst_pitch, dW, dH, t);
    });
}

}  // namespace ofb
