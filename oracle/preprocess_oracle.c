/*
 * preprocess_oracle.c -- CPU restatement of the frame preprocessing either side of the hot path
 * (SURVEY.md section 8f, row N2).  TEST INFRASTRUCTURE ONLY: nothing under optical_flow_b200/ may link
 * or call this; tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg are its only users.
 *
 * Reference call sites (the arithmetic itself lives in the un-vendored dependency opencv-python,
 * pinned 4.2.0.32 in /root/reference/requirements_optical_flow.txt:3; restated from OpenCV's published
 * imgproc algorithms and pinned against cv2 4.13.0 by tests/golden/preprocess_*.npz):
 *   orc_bgr2gray        cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)      optical_flow.py:44,
 *                                                                    visualize_optical_flow.py:31,35
 *   orc_resize_u8       cv2.resize(frame, (w, h))  [INTER_LINEAR]    optical_flow.py:25-31
 *
 * Both are integer / fixed-point algorithms: the bar is bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define ORC_API __attribute__((visibility("default")))

/* 8-bit BGR -> gray: 15-bit fixed-point weights B 3735, G 19235, R 9798 (sum 32768), round to nearest. */
ORC_API void orc_bgr2gray(const uint8_t* bgr, int W, int H, uint8_t* gray)
{
    for (size_t i = 0, n = (size_t)W * H; i < n; i++)
        gray[i] = (uint8_t)((bgr[3 * i] * 3735 + bgr[3 * i + 1] * 19235 + bgr[3 * i + 2] * 9798 + 16384) >> 15);
}

/* Source index and 11-bit weights of one destination coordinate of the 8-bit bilinear resize:
 * the coordinate (d + 0.5) * scale - 0.5 is formed in double and ROUNDED TO f32 before the floor and the
 * fraction (this is what distinguishes the u8 path from the f32 resize of SURVEY.md A.4);
 * weights = round-half-even((1 - frac) * 2048), round-half-even(frac * 2048) in f32.
 * Columns (clamp_weights = 1): an index left of 0 or at / right of the last column is moved inside AND its
 * fraction is zeroed.  Rows (clamp_weights = 0): only the two row indices are clipped; the weights keep the
 * fraction, so a clipped row is blended with itself through both truncating terms (differs by 1 now and then). */
static void coord_u8(int d, double scale, int slen, int clamp_weights, int* s0, int* s1, int* w0, int* w1)
{
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_weights) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= slen - 1) { s = slen - 1; f = 0.f; }
    }
    *s0 = s < 0 ? 0 : s > slen - 1 ? slen - 1 : s;
    *s1 = s + 1 < 0 ? 0 : s + 1 > slen - 1 ? slen - 1 : s + 1;
    *w0 = (int)lrintf((1.f - f) * 2048.f);
    *w1 = (int)lrintf(f * 2048.f);
}

/* cn interleaved channels, tightly packed.  Row pass: S = a*w0 + b*w1 (22-bit ints); column pass:
 * (((v0 * (S0 >> 4)) >> 16) + ((v1 * (S1 >> 4)) >> 16) + 2) >> 2, saturated to 8 bits. */
ORC_API void orc_resize_u8(const uint8_t* src, int W, int H, int cn, uint8_t* dst, int dW, int dH)
{
    const double sx = 1.0 / ((double)dW / W), sy = 1.0 / ((double)dH / H);
    int* x0 = (int*)malloc(sizeof(int) * 4 * (size_t)dW);
    int *x1 = x0 + dW, *a0 = x1 + dW, *a1 = a0 + dW;
    for (int x = 0; x < dW; x++) coord_u8(x, sx, W, 1, &x0[x], &x1[x], &a0[x], &a1[x]);
    for (int y = 0; y < dH; y++) {
        int y0, y1, b0, b1;
        coord_u8(y, sy, H, 0, &y0, &y1, &b0, &b1);
        const uint8_t *r0 = src + (size_t)y0 * W * cn, *r1 = src + (size_t)y1 * W * cn;
        uint8_t* o = dst + (size_t)y * dW * cn;
        for (int x = 0; x < dW; x++)
            for (int c = 0; c < cn; c++) {
                int S0 = r0[x0[x] * cn + c] * a0[x] + r0[x1[x] * cn + c] * a1[x];
                int S1 = r1[x0[x] * cn + c] * a0[x] + r1[x1[x] * cn + c] * a1[x];
                int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
                o[x * cn + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
    }
    free(x0);
}
