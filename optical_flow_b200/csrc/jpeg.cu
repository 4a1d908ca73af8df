// jpeg.cu -- baseline JPEG encoding of the flow picture ON THE GPU: the artefact the reference writes with
// cv2.imwrite("flow_<ms>.jpeg", rgb) (/root/reference/visualize_optical_flow.py:57-58).  Why it is here: the raw BGR picture
// is 6.2 MB per 1080p pair and is what saturates the host side of the box at 4-8 GPUs; the JPEG of the same picture is
// ~0.2 MB, and it is the only form of the picture the reference keeps.
//
// Byte-for-byte the stream libjpeg(-turbo) produces under cv2's defaults (quality 95, 4:2:0, baseline SOF0, the Annex-K
// Huffman tables, no restart markers), checked against cv2.imencode itself and against oracle/jpeg_oracle.c:
//   k_jpeg_dct     4 MCUs (64 x 16 px) per CTA: BGR -> YCbCr (16-bit fixed point), 2x2 chroma box with the 1,2,1,2 bias,
//                  replicated edges (columns before, rows after the down-sample), the "islow" integer 8x8 DCT as two
//                  passes through shared memory, quantisation by 8q with an exact multiply-high reciprocal, zigzag;
//                  coefficients leave as int16 in scan order (MCU-major, Y00 Y01 Y10 Y11 Cb Cr)
//   k_jpeg_count   one thread per 8x8 block: length in bits of its Huffman code (DC difference against the previous block of
//                  the same component, AC run/size pairs, ZRL, EOB); the loop runs over the set bits of the block's
//                  non-zero mask (built by k_jpeg_dct), i.e. once per non-zero coefficient
//   k_jpeg_scan    one CTA per picture: exclusive prefix sum of the block lengths -> the bit offset of every block
//   k_jpeg_emit    128 blocks per CTA: every thread writes its block's code at its bit offset into a shared-memory window
//                  (atomicOr on the two words it shares with its neighbours), the window goes out as whole words
//   k_jpeg_ffcount / k_jpeg_layout / k_jpeg_stuff   0xFF byte stuffing: count the 0xFF bytes per 4 KB segment, prefix-sum
//                  them (and the pictures of the chunk: the streams are packed back to back), then copy every segment to its
//                  final place with a 0x00 after each 0xFF, the 623-byte header in front and EOI behind
// Everything is integer / byte work: no tensor cores, bound by shared-memory and issue rate, negligible next to the flow
// kernels (measured per picture in profiles/).  The Huffman tables are built on the host from the Annex-K BITS / HUFFVAL lists.
#include "common.cuh"
#include "launch.cuh"
#include "jpeg.cuh"

#include <cstring>
#include <vector>

namespace ofb {

// ------------------------------------------------------------------------------------------------
// host: tables and header (ITU-T T.81 Annex K; libjpeg jcparam.c quality scaling, jcmarker.c segment order)
// ------------------------------------------------------------------------------------------------
namespace {

const uint8_t STD_LUMA_Q[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t STD_CHROMA_Q[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t ZIGZAG[64] = {   // zigzag position -> natural (row-major) index
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t DC_LUMA_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t DC_CHROMA_BITS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t AC_LUMA_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t AC_LUMA_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
const uint8_t AC_CHROMA_BITS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t AC_CHROMA_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};

void quant_table(int quality, bool chroma, uint8_t out[64])
{
    quality = quality <= 0 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    const uint8_t* base = chroma ? STD_CHROMA_Q : STD_LUMA_Q;
    for (int i = 0; i < 64; i++) {
        long t = ((long)base[i] * scale + 50L) / 100L;
        out[i] = (uint8_t)(t <= 0 ? 1 : (t > 255 ? 255 : t));
    }
}

// canonical Huffman codes (T.81 Annex C) packed as (code << 8) | length, indexed by symbol
void build_codes(const uint8_t* bits, const uint8_t* vals, int nvals, uint32_t* out /* 256 */)
{
    memset(out, 0, 256 * sizeof(uint32_t));
    int k = 0; unsigned code = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1] && k < nvals; i++, k++) out[vals[k]] = (code++ << 8) | (unsigned)l;
        code <<= 1;
    }
}

size_t put_seg(uint8_t* o, size_t n, int marker, const uint8_t* payload, int len)
{
    o[n++] = 0xff; o[n++] = (uint8_t)marker; o[n++] = (uint8_t)((len + 2) >> 8); o[n++] = (uint8_t)((len + 2) & 0xff);
    memcpy(o + n, payload, (size_t)len);
    return n + (size_t)len;
}

}  // namespace

void jpeg_build_tables(int W, int H, int quality, JpegTables& t)
{
    memset(&t, 0, sizeof(t));
    uint8_t q[2][64];
    quant_table(quality, false, q[0]);
    quant_table(quality, true, q[1]);
    for (int c = 0; c < 2; c++)
        for (int nat = 0; nat < 64; nat++) {
            const unsigned d = 8u * q[c][nat];                                   // the islow DCT leaves its output scaled by 8
            t.recip[c][nat] = (uint32_t)(((1ull << 32) + d - 1) / d);            // floor(n / d) == umulhi(n, recip) for n < 2^16
            t.half[c][nat] = (uint16_t)(d / 2);
        }
    for (int k = 0; k < 64; k++) t.zz_of_nat[ZIGZAG[k]] = (uint8_t)k;
    build_codes(DC_LUMA_BITS, DC_VALS, 12, t.dc[0]);
    build_codes(DC_CHROMA_BITS, DC_VALS, 12, t.dc[1]);
    build_codes(AC_LUMA_BITS, AC_LUMA_VALS, 162, t.ac[0]);
    build_codes(AC_CHROMA_BITS, AC_CHROMA_VALS, 162, t.ac[1]);
    // SOI, APP0 (JFIF 1.01, density 1:1), DQT x2 (zigzag order), SOF0 (Y 2x2, Cb 1x1, Cr 1x1), DHT x4, SOS
    uint8_t* o = t.header;
    size_t n = 0;
    o[n++] = 0xff; o[n++] = 0xd8;
    const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    n = put_seg(o, n, 0xe0, app0, 14);
    for (int c = 0; c < 2; c++) {
        uint8_t seg[65];
        seg[0] = (uint8_t)c;
        for (int i = 0; i < 64; i++) seg[1 + i] = q[c][ZIGZAG[i]];
        n = put_seg(o, n, 0xdb, seg, 65);
    }
    const uint8_t sof[15] = {8, (uint8_t)(H >> 8), (uint8_t)(H & 255), (uint8_t)(W >> 8), (uint8_t)(W & 255), 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    n = put_seg(o, n, 0xc0, sof, 15);
    const uint8_t* bits[4] = {DC_LUMA_BITS, AC_LUMA_BITS, DC_CHROMA_BITS, AC_CHROMA_BITS};
    const uint8_t* vals[4] = {DC_VALS, AC_LUMA_VALS, DC_VALS, AC_CHROMA_VALS};
    const int nv[4] = {12, 162, 12, 162}, id[4] = {0x00, 0x10, 0x01, 0x11};
    for (int i = 0; i < 4; i++) {
        uint8_t seg[1 + 16 + 162];
        seg[0] = (uint8_t)id[i];
        memcpy(seg + 1, bits[i], 16);
        memcpy(seg + 17, vals[i], (size_t)nv[i]);
        n = put_seg(o, n, 0xc4, seg, 17 + nv[i]);
    }
    const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    n = put_seg(o, n, 0xda, sos, 10);
    t.header_len = (int)n;
}

JpegGeom jpeg_geometry(int W, int H)
{
    JpegGeom g{};
    g.W = W; g.H = H;
    g.mcux = (W + 15) / 16; g.mcuy = (H + 15) / 16;
    g.nblk = g.mcux * g.mcuy * 6;
    g.ywb = (W + 7) / 8; g.yhb = (H + 7) / 8;
    g.ch_real = (H + 1) / 2;
    g.bits_cap = ((size_t)g.nblk * JPEG_MAX_BLOCK_BITS / 8 + 4 + 15) & ~(size_t)15;     // unstuffed stream, worst case, bytes
    g.nseg_cap = (int)((g.bits_cap + JPEG_SEG - 1) / JPEG_SEG);
    g.out_cap = ((size_t)W * H * 3 + 4096 + 15) & ~(size_t)15;                          // stuffed stream + header, per picture
    return g;
}

// ------------------------------------------------------------------------------------------------
// k_jpeg_dct
// ------------------------------------------------------------------------------------------------
constexpr int DCT_THREADS = 256;
constexpr int BLK_PITCH = 72;            // ints per 8x8 block in shared memory: 4 blocks of a warp land in distinct banks

// one 1-D pass of jfdctint.c on 8 values; FIRST = row pass (results scaled up by 4), else column pass
template <bool FIRST>
__device__ __forceinline__ void fdct8(int (&d)[8])
{
    constexpr int CB = 13, P1 = 2;
    constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137,
                  F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
    constexpr int SH = FIRST ? CB - P1 : CB + P1, RND = 1 << (SH - 1);
    int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    if (FIRST) { d[0] = (t10 + t11) << P1; d[4] = (t10 - t11) << P1; }
    else { d[0] = (t10 + t11 + (1 << (P1 - 1))) >> P1; d[4] = (t10 - t11 + (1 << (P1 - 1))) >> P1; }
    int z1 = (t12 + t13) * F0541;
    d[2] = (z1 + t13 * F0765 + RND) >> SH;
    d[6] = (z1 + t12 * (-F1847) + RND) >> SH;
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * F1175;
    t4 *= F0298; t5 *= F2053; t6 *= F3072; t7 *= F1501;
    z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
    z3 += z5; z4 += z5;
    d[7] = (t4 + z1 + z3 + RND) >> SH;
    d[5] = (t5 + z2 + z4 + RND) >> SH;
    d[3] = (t6 + z2 + z3 + RND) >> SH;
    d[1] = (t7 + z1 + z4 + RND) >> SH;
}

// Picture z of the batch: (H, W, 3) uint8 BGR at bgr + z * bgr_item.  Output: coef + z * nblk * 64.
__global__ void __launch_bounds__(DCT_THREADS)
k_jpeg_dct(const uint8_t* __restrict__ bgr, size_t bgr_item, JpegGeom g, const JpegTables* __restrict__ tab, int16_t* __restrict__ coef,
           unsigned long long* __restrict__ blk_mask)
{
    __shared__ int sblk[24 * BLK_PITCH];                 // 4 MCUs x 6 blocks
    __shared__ __align__(16) int16_t sout[24 * 64];
    __shared__ unsigned smask[24 * 2];                   // per block: bit k set <=> quantised coefficient at zigzag position k is non-zero
    __shared__ uint32_t srecip[2][64];                   // the quantisation tables, once per CTA (3 table reads per coefficient)
    __shared__ uint16_t shalf[2][64];
    __shared__ uint8_t szz[64];
    if (threadIdx.x < 48) smask[threadIdx.x] = 0u;
    if (threadIdx.x < 128) {
        srecip[threadIdx.x >> 6][threadIdx.x & 63] = tab->recip[threadIdx.x >> 6][threadIdx.x & 63];
        shalf[threadIdx.x >> 6][threadIdx.x & 63] = tab->half[threadIdx.x >> 6][threadIdx.x & 63];
        if (threadIdx.x < 64) szz[threadIdx.x] = tab->zz_of_nat[threadIdx.x];
    }
    const int tid = threadIdx.x;
    const int mx0 = blockIdx.x * 4, my = blockIdx.y, z = blockIdx.z;
    const int W = g.W, H = g.H;
    const uint8_t* src = bgr + (size_t)z * bgr_item;

    // ---- colour conversion: one 2x2 quad per thread (32 x 8 quads = 64 x 16 pixels) ----
    {
        const int qx = tid & 31, qy = tid >> 5;              // quad coordinates inside the tile
        const int cx = mx0 * 8 + qx, cy = my * 8 + qy;       // absolute chroma sample
        // Y samples use the replicated frame (column min(x, W-1), row min(y, H-1)).  The chroma sample replicates COLUMNS before
        // the 2x2 box and ROWS after it: chroma row cy >= ch_real is a copy of chroma row ch_real - 1.
        const int cyc = min(cy, g.ch_real - 1);
        int yv[4], cb = 0, cr = 0;
        if (2 * cx + 1 < W && 2 * cy + 1 < H && (W & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 1) == 0) {
            // interior quad of an even-width picture: its two pixel pairs are 6 contiguous, 2-byte aligned bytes per row
#pragma unroll
            for (int dy = 0; dy < 2; dy++) {
                const unsigned short* p = reinterpret_cast<const unsigned short*>(src + ((size_t)(2 * cy + dy) * W + 2 * cx) * 3);
                const unsigned h0 = p[0], h1 = p[1], h2 = p[2];                 // B0 G0 | R0 B1 | G1 R1
                const int b0 = h0 & 255, g0 = h0 >> 8, r0 = h1 & 255, b1 = h1 >> 8, g1 = h2 & 255, r1 = h2 >> 8;
                yv[2 * dy] = ((19595 * r0 + 38470 * g0 + 7471 * b0 + 32768) >> 16) - 128;
                yv[2 * dy + 1] = ((19595 * r1 + 38470 * g1 + 7471 * b1 + 32768) >> 16) - 128;
                cb += ((-11059 * r0 - 21709 * g0 + 32768 * b0 + (128 << 16) + 32767) >> 16) + ((-11059 * r1 - 21709 * g1 + 32768 * b1 + (128 << 16) + 32767) >> 16);
                cr += ((32768 * r0 - 27439 * g0 - 5329 * b0 + (128 << 16) + 32767) >> 16) + ((32768 * r1 - 27439 * g1 - 5329 * b1 + (128 << 16) + 32767) >> 16);
            }
        } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int dx = i & 1, dy = i >> 1;
            const int xs = min(2 * cx + dx, W - 1);
            {
                const int ys = min(2 * cy + dy, H - 1);
                const uint8_t* p = src + ((size_t)ys * W + xs) * 3;
                const int b = p[0], gg = p[1], r = p[2];
                yv[i] = ((19595 * r + 38470 * gg + 7471 * b + 32768) >> 16) - 128;
                if (cyc == cy) {
                    cb += (-11059 * r - 21709 * gg + 32768 * b + (128 << 16) + 32767) >> 16;
                    cr += (32768 * r - 27439 * gg - 5329 * b + (128 << 16) + 32767) >> 16;
                }
            }
            if (cyc != cy) {                                  // bottom padding rows of the chroma planes
                const int ys = min(2 * cyc + dy, H - 1);
                const uint8_t* p = src + ((size_t)ys * W + xs) * 3;
                const int b = p[0], gg = p[1], r = p[2];
                cb += (-11059 * r - 21709 * gg + 32768 * b + (128 << 16) + 32767) >> 16;
                cr += (32768 * r - 27439 * gg - 5329 * b + (128 << 16) + 32767) >> 16;
            }
        }
        }
        const int bias = 1 + (cx & 1);
        cb = ((cb + bias) >> 2) - 128;
        cr = ((cr + bias) >> 2) - 128;
        const int m = qx >> 3;                               // MCU of the tile
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int lx = 2 * (qx & 7) + (i & 1), ly = 2 * qy + (i >> 1);      // pixel inside the 16 x 16 MCU
            const int blk = m * 6 + (ly >> 3) * 2 + (lx >> 3);
            sblk[blk * BLK_PITCH + (ly & 7) * 8 + (lx & 7)] = yv[i];
        }
        sblk[(m * 6 + 4) * BLK_PITCH + qy * 8 + (qx & 7)] = cb;
        sblk[(m * 6 + 5) * BLK_PITCH + qy * 8 + (qx & 7)] = cr;
    }
    __syncthreads();
    // ---- row pass: thread = (block, row) ----
    if (tid < 192) {
        int* p = sblk + (tid >> 3) * BLK_PITCH + (tid & 7) * 8;
        int d[8];
        const int4 a = *reinterpret_cast<const int4*>(p), b = *reinterpret_cast<const int4*>(p + 4);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
        fdct8<true>(d);
        *reinterpret_cast<int4*>(p) = make_int4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<int4*>(p + 4) = make_int4(d[4], d[5], d[6], d[7]);
    }
    __syncthreads();
    // ---- column pass + quantisation + zigzag: thread = (block, column) ----
    if (tid < 192) {
        const int blk = tid >> 3, col = tid & 7;
        const int* p = sblk + blk * BLK_PITCH + col;
        int d[8];
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = p[i * 8];
        fdct8<false>(d);
        const int c = (blk % 6) >= 4 ? 1 : 0;
        unsigned m0 = 0u, m1 = 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int nat = i * 8 + col;
            const int v = d[i];
            const unsigned a = (unsigned)abs(v) + shalf[c][nat];
            const int qv = (int)__umulhi(a, srecip[c][nat]);                    // (|v| + d/2) / d, exact
            const int zz = szz[nat];
            sout[blk * 64 + zz] = (int16_t)(v < 0 ? -qv : qv);
            if (qv) { if (zz < 32) m0 |= 1u << zz; else m1 |= 1u << (zz - 32); }
        }
        if (m0) atomicOr(&smask[blk * 2], m0);
        if (m1) atomicOr(&smask[blk * 2 + 1], m1);
    }
    __syncthreads();
    // ---- dummy blocks (jccoefct.c): a luma block beyond the picture is all-zero AC with the DC of the preceding block ----
    if (tid < 4) {
        const int mx = mx0 + tid;
        int16_t* mc = sout + tid * 6 * 64;
        for (int b = 1; b < 4; b++) {
            const bool real = (2 * mx + (b & 1)) < g.ywb && (2 * my + (b >> 1)) < g.yhb;
            if (!real) mc[b * 64] = mc[(b - 1) * 64];
        }
    }
    __syncthreads();
    // ---- store: the MCUs of the tile are consecutive in scan order ----
    const int nm = min(4, g.mcux - mx0);
    int16_t* dst = coef + ((size_t)z * g.nblk + ((size_t)my * g.mcux + mx0) * 6) * 64;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sout);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    for (int i = tid; i < nm * 6 * 32; i += DCT_THREADS) {
        const int blk = i >> 5, m = blk / 6, b = blk - m * 6;
        const bool real = b >= 4 || b == 0 || ((2 * (mx0 + m) + (b & 1)) < g.ywb && (2 * my + (b >> 1)) < g.yhb);
        uint32_t v = s32[i];
        if (!real) v = (i & 31) == 0 ? (v & 0xffffu) : 0u;               // keep the propagated DC, zero every AC
        d32[i] = v;
    }
    if (tid < nm * 6) {
        const int m = tid / 6, b = tid - m * 6;
        const bool real = b >= 4 || b == 0 || ((2 * (mx0 + m) + (b & 1)) < g.ywb && (2 * my + (b >> 1)) < g.yhb);
        const unsigned long long mk = real ? ((unsigned long long)smask[tid * 2 + 1] << 32) | smask[tid * 2] : 0ull;
        blk_mask[(size_t)z * g.nblk + ((size_t)my * g.mcux + mx0) * 6 + tid] = mk & ~1ull;      // AC positions only
    }
}

// ------------------------------------------------------------------------------------------------
// Huffman: count, scan, emit
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nbits_of(int v) { return 32 - __clz(abs(v)); }

// Previous block of the same component in scan order: Y blocks chain through the MCU and into the next one; Cb / Cr chain
// from MCU to MCU.  Returns its DC, or 0 at the start of the scan.
__device__ __forceinline__ int dc_predictor(const int16_t* __restrict__ coef, int blk)
{
    const int m = blk / 6, b = blk - m * 6;
    int prev;
    if (b >= 4) prev = m > 0 ? blk - 6 : -1;
    else if (b > 0) prev = blk - 1;
    else prev = m > 0 ? blk - 3 : -1;                    // Y00 of MCU m follows Y11 of MCU m-1
    return prev < 0 ? 0 : (int)coef[(size_t)prev * 64];
}

constexpr int HUF_THREADS = 128;

// One thread per block.  The non-zero AC positions come as a 64-bit mask from k_jpeg_dct, so the loop runs once per NON-ZERO
// coefficient (a dozen for a flow picture at quality 95) instead of 63 times with a divergent zero test; each coefficient is a
// 2-byte load from the block's own 128-byte line.
__global__ void __launch_bounds__(HUF_THREADS)
k_jpeg_count(const int16_t* __restrict__ coef, const unsigned long long* __restrict__ blk_mask, JpegGeom g,
             const JpegTables* __restrict__ tab, uint32_t* __restrict__ blk_bits)
{
    __shared__ uint8_t slen[2][256];                     // AC code lengths
    for (int i = threadIdx.x; i < 512; i += HUF_THREADS) slen[i >> 8][i & 255] = (uint8_t)(tab->ac[i >> 8][i & 255] & 0xff);
    __syncthreads();
    const int z = blockIdx.y;
    const int blk = blockIdx.x * HUF_THREADS + threadIdx.x;
    if (blk >= g.nblk) return;
    const int16_t* pc = coef + (size_t)z * g.nblk * 64;
    const int16_t* mine = pc + (size_t)blk * 64;
    const int c = (blk % 6) >= 4 ? 1 : 0;
    unsigned long long m = blk_mask[(size_t)z * g.nblk + blk];
    const int diff = (int)mine[0] - dc_predictor(pc, blk);
    int n = nbits_of(diff);
    unsigned bits = (tab->dc[c][n] & 0xff) + n;
    int prev = 0;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int run = k - prev - 1;
        prev = k;
        bits += (run >> 4) * slen[c][0xf0];
        n = nbits_of((int)mine[k]);
        bits += slen[c][((run & 15) << 4) | n] + n;
    }
    if (prev < 63) bits += slen[c][0];
    blk_bits[(size_t)z * g.nblk + blk] = bits;
}

// Sum of the code lengths of the 128 blocks of every emit CTA (grid = CTAs x pictures): the input of k_jpeg_scan.
__global__ void __launch_bounds__(HUF_THREADS)
k_jpeg_ctasum(const uint32_t* __restrict__ blk_bits, JpegGeom g, int ncta, uint32_t* __restrict__ cta_bits)
{
    __shared__ uint32_t sw[HUF_THREADS / 32];
    const int z = blockIdx.y, blk = blockIdx.x * HUF_THREADS + threadIdx.x;
    uint32_t v = blk < g.nblk ? blk_bits[(size_t)z * g.nblk + blk] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) cta_bits[(size_t)z * ncta + blockIdx.x] = sw[0] + sw[1] + sw[2] + sw[3];
}

// One CTA per picture: exclusive prefix sum of the per-CTA code lengths (a few hundred values) -> the bit offset of every emit
// CTA and the total of the picture; also zeroes the words of the unstuffed stream that two emit CTAs share (they are written
// with atomicOr) and the last, partly used word.
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
k_jpeg_scan(uint32_t* __restrict__ cta_bits, int ncta, JpegGeom g, uint32_t* __restrict__ bits32, uint32_t* __restrict__ total_bits)
{
    __shared__ uint32_t swarp[32];
    __shared__ uint32_t stot;
    const int z = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t* a = cta_bits + (size_t)z * ncta;
    uint32_t* stream = bits32 + (size_t)z * (g.bits_cap / 4);
    uint32_t carry = 0;
    for (int s0 = 0; s0 < ncta; s0 += SCAN_THREADS) {
        const int i = s0 + tid;
        const uint32_t v = i < ncta ? a[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        __syncthreads();
        if (lane == 31) swarp[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t x = swarp[lane], ix = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, ix, o); if (lane >= o) ix += u; }
            swarp[lane] = ix - x;
            if (lane == 31) stot = ix;
        }
        __syncthreads();
        if (i < ncta) {
            const uint32_t off = carry + swarp[w] + inc - v;
            a[i] = off;
            stream[off >> 5] = 0u;                                     // first word of emit CTA i (shared with CTA i-1 when off % 32 != 0)
        }
        carry += stot;
    }
    if (tid == 0) { total_bits[z] = carry; stream[carry >> 5] = 0u; }
}

// 128 consecutive blocks per CTA.  Every thread encodes its block MSB-first into a shared-memory window that starts at the
// word holding the CTA's first bit; the first and last word of a block's code are shared with the neighbouring blocks, so
// they are combined with atomicOr, the words in between are the thread's own.
constexpr int EMIT_WORDS = (HUF_THREADS * JPEG_MAX_BLOCK_BITS + 31) / 32 + 2;

struct BitSink {
    uint32_t* win;          // shared-memory window (zeroed)
    int word;               // index of the word `acc` will be flushed to
    uint64_t acc;           // pending bits, left-aligned at bit 63
    int fill;               // number of valid bits in acc (counted from the top)
    bool first;
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        acc |= (uint64_t)code << (64 - fill - len);
        fill += len;
        if (fill >= 32) {
            const uint32_t w = (uint32_t)(acc >> 32);
            if (first) { atomicOr(win + word, w); first = false; } else win[word] = w;
            word++; acc <<= 32; fill -= 32;
        }
    }
    __device__ __forceinline__ void finish() { if (fill > 0) atomicOr(win + word, (uint32_t)(acc >> 32)); }
};

__global__ void __launch_bounds__(HUF_THREADS)
k_jpeg_emit(const int16_t* __restrict__ coef, const unsigned long long* __restrict__ blk_mask, JpegGeom g, const JpegTables* __restrict__ tab,
            const uint32_t* __restrict__ blk_bits, const uint32_t* __restrict__ cta_off, int ncta, uint32_t* __restrict__ bits32)
{
    __shared__ uint32_t sac[2][256];
    __shared__ uint32_t win[EMIT_WORDS];
    __shared__ uint32_t swsum[HUF_THREADS / 32];
    const int tid = threadIdx.x, z = blockIdx.y, lane = tid & 31;
    for (int i = tid; i < 512; i += HUF_THREADS) sac[i >> 8][i & 255] = tab->ac[i >> 8][i & 255];
    for (int i = tid; i < EMIT_WORDS; i += HUF_THREADS) win[i] = 0u;
    const int blk0 = blockIdx.x * HUF_THREADS;
    const int blk = blk0 + tid;
    // bit offset of every block of this CTA: the CTA's offset (k_jpeg_scan) + an exclusive scan of the 128 code lengths
    const uint32_t mylen = blk < g.nblk ? blk_bits[(size_t)z * g.nblk + blk] : 0u;
    uint32_t inc = mylen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) swsum[tid >> 5] = inc;
    const uint32_t base_bit = cta_off[(size_t)z * ncta + blockIdx.x];
    const uint32_t base_word = base_bit >> 5;
    __syncthreads();
    uint32_t myoff = base_bit + inc - mylen;
    for (int i = 0; i < (tid >> 5); i++) myoff += swsum[i];
    const uint32_t end_bit = base_bit + swsum[0] + swsum[1] + swsum[2] + swsum[3];
    if (blk < g.nblk) {
        const int16_t* pc = coef + (size_t)z * g.nblk * 64;
        const int16_t* mine = pc + (size_t)blk * 64;
        const int c = (blk % 6) >= 4 ? 1 : 0;
        unsigned long long m = blk_mask[(size_t)z * g.nblk + blk];
        const uint32_t p0 = myoff - (base_word << 5);                // bit position inside the window
        BitSink s{win, (int)(p0 >> 5), 0ull, (int)(p0 & 31), true};
        const int diff = (int)mine[0] - dc_predictor(pc, blk);
        int n = nbits_of(diff);
        const uint32_t dcode = tab->dc[c][n];
        s.put(dcode >> 8, (int)(dcode & 0xff));
        if (n) s.put((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << n) - 1), n);
        int prev = 0;
        while (m) {
            const int k = __ffsll((long long)m) - 1;
            m &= m - 1;
            int run = k - prev - 1;
            prev = k;
            const int v = (int)mine[k];
            while (run > 15) { const uint32_t zc = sac[c][0xf0]; s.put(zc >> 8, (int)(zc & 0xff)); run -= 16; }
            n = nbits_of(v);
            const uint32_t ac = sac[c][(run << 4) | n];
            // code and value bits in one go (<= 16 + 10 bits)
            s.put(((ac >> 8) << n) | ((uint32_t)(v < 0 ? v - 1 : v) & ((1u << n) - 1)), (int)(ac & 0xff) + n);
        }
        if (prev < 63) { const uint32_t eob = sac[c][0]; s.put(eob >> 8, (int)(eob & 0xff)); }
        s.finish();
    }
    __syncthreads();
    // window -> global, byte order of the stream (big-endian words).  First and last word may be shared with the neighbours.
    uint32_t* stream = bits32 + (size_t)z * (g.bits_cap / 4);
    const uint32_t last_word = end_bit == base_bit ? base_word : ((end_bit - 1) >> 5);
    const int nw = (int)(last_word - base_word) + 1;
    const bool shared_first = (base_bit & 31u) != 0u, shared_last = (end_bit & 31u) != 0u;     // a neighbour writes into the same word
    for (int i = tid; i < nw; i += HUF_THREADS) {
        const uint32_t v = __byte_perm(win[i], 0, 0x0123);
        if ((i == 0 && shared_first) || (i == nw - 1 && shared_last)) { if (v) atomicOr(stream + base_word + i, v); }
        else stream[base_word + i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// byte stuffing and compaction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stream_bytes(uint32_t total_bits) { return (total_bits + 7) >> 3; }

// byte i of picture z's unstuffed stream, the last byte padded with 1-bits (jchuff.c flush_bits)
__device__ __forceinline__ uint32_t load16_padded(const uint8_t* stream, uint32_t i0, uint32_t nbytes, uint32_t total_bits, uint8_t out[16])
{
    const uint4 q = *reinterpret_cast<const uint4*>(stream + i0);
    memcpy(out, &q, 16);
    const uint32_t cnt = i0 >= nbytes ? 0u : min(16u, nbytes - i0);
    if ((total_bits & 7) && i0 + cnt == nbytes && cnt > 0) out[cnt - 1] |= (uint8_t)((1u << (8 - (total_bits & 7))) - 1);
    return cnt;
}

constexpr int SEG_THREADS = JPEG_SEG / 16;               // 16 bytes per thread

__global__ void __launch_bounds__(SEG_THREADS)
k_jpeg_ffcount(const uint32_t* __restrict__ bits32, JpegGeom g, const uint32_t* __restrict__ total_bits, uint32_t* __restrict__ seg_ff)
{
    const int z = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x;
    const uint32_t tb = total_bits[z], nbytes = stream_bytes(tb);
    const uint32_t i0 = (uint32_t)seg * JPEG_SEG + tid * 16;
    uint32_t cnt = 0;
    if ((uint32_t)seg * JPEG_SEG < nbytes) {
        uint8_t b[16];
        const uint32_t n = load16_padded(reinterpret_cast<const uint8_t*>(bits32) + (size_t)z * g.bits_cap, i0, nbytes, tb, b);
        for (uint32_t i = 0; i < n; i++) cnt += b[i] == 0xff;
    }
    __shared__ uint32_t sw[SEG_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((tid & 31) == 0) sw[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int i = 0; i < SEG_THREADS / 32; i++) s += sw[i];
        seg_ff[(size_t)z * g.nseg_cap + seg] = s;
    }
}

// One CTA for the whole chunk: per picture the exclusive prefix of its segments' 0xFF counts and its final size; then the
// pictures are laid out back to back.  out_off[z] = byte offset of picture z in the chunk's output, sizes[z] its size,
// chunk_total[0] the sum, chunk_total[1] a flag (non-zero: a picture did not fit in out_cap).
__global__ void __launch_bounds__(1024)
k_jpeg_layout(uint32_t* __restrict__ seg_ff, JpegGeom g, const uint32_t* __restrict__ total_bits, int header_len, int batch, int nseg_launched,
              unsigned long long* __restrict__ out_off, uint32_t* __restrict__ sizes, unsigned long long* __restrict__ chunk_total)
{
    extern __shared__ unsigned long long ssize[];        // batch stream sizes
    __shared__ unsigned sflag;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) sflag = 0u;
    __syncthreads();
    for (int z = w; z < batch; z += 32) {                // one warp per picture: exclusive prefix of its segments' 0xFF counts
        const uint32_t nbytes = stream_bytes(total_bits[z]);
        const int nseg = (int)((nbytes + JPEG_SEG - 1) / JPEG_SEG);
        uint32_t* a = seg_ff + (size_t)z * g.nseg_cap;
        uint32_t carry = 0;
        const int nscan = min(nseg, nseg_launched);
        for (int s0 = 0; s0 < nscan; s0 += 32) {
            const int i = s0 + lane;
            const uint32_t v = i < nscan ? a[i] : 0u;
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            if (i < nscan) a[i] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            const unsigned long long sz = (unsigned long long)header_len + nbytes + carry + 2ull;
            ssize[z] = sz;
            sizes[z] = (uint32_t)sz;
            if (sz > g.out_cap || nseg > nseg_launched) atomicOr(&sflag, 1u);
        }
    }
    __syncthreads();
    if (tid == 0) {                                      // the pictures of the chunk back to back
        unsigned long long run = 0ull;
        for (int z = 0; z < batch; z++) { out_off[z] = run; run += ssize[z]; }
        chunk_total[0] = run; chunk_total[1] = sflag;
    }
}

__global__ void __launch_bounds__(SEG_THREADS)
k_jpeg_stuff(const uint32_t* __restrict__ bits32, JpegGeom g, const uint32_t* __restrict__ total_bits, const uint32_t* __restrict__ seg_ff,
             const unsigned long long* __restrict__ out_off, const JpegTables* __restrict__ tab, uint8_t* __restrict__ out)
{
    __shared__ uint8_t sbuf[2 * JPEG_SEG];
    __shared__ uint32_t swarp[SEG_THREADS / 32];
    const int z = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t tb = total_bits[z], nbytes = stream_bytes(tb);
    uint8_t* dst = out + out_off[z];
    const int hl = tab->header_len;
    if (seg == 0) for (int i = tid; i < hl; i += SEG_THREADS) dst[i] = tab->header[i];
    if ((uint32_t)seg * JPEG_SEG >= nbytes) {
        if (seg == 0 && tid == 0) { dst[hl] = 0xff; dst[hl + 1] = 0xd9; }       // empty stream cannot happen (every block codes >= 4 bits)
        return;
    }
    const uint32_t i0 = (uint32_t)seg * JPEG_SEG + tid * 16;
    uint8_t b[16];
    const uint32_t n = load16_padded(reinterpret_cast<const uint8_t*>(bits32) + (size_t)z * g.bits_cap, i0, nbytes, tb, b);
    uint32_t ff = 0;
    for (uint32_t i = 0; i < n; i++) ff += b[i] == 0xff;
    const uint32_t mine = n + ff;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) swarp[w] = inc;
    __syncthreads();
    uint32_t before = inc - mine;
    for (int i = 0; i < w; i++) before += swarp[i];
    uint32_t total = 0;
    for (int i = 0; i < SEG_THREADS / 32; i++) total += swarp[i];
    uint32_t o = before;
    for (uint32_t i = 0; i < n; i++) { sbuf[o++] = b[i]; if (b[i] == 0xff) sbuf[o++] = 0; }
    __syncthreads();
    uint8_t* d = dst + hl + (size_t)seg * JPEG_SEG + seg_ff[(size_t)z * g.nseg_cap + seg];
    for (uint32_t i = tid; i < total; i += SEG_THREADS) d[i] = sbuf[i];
    if ((uint32_t)(seg + 1) * JPEG_SEG >= nbytes && tid == 0) { d[total] = 0xff; d[total + 1] = 0xd9; }      // EOI after the last segment
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
void launch_jpeg_encode(Launch& L, const JpegWork& w, const uint8_t* bgr, size_t bgr_item, int batch, uint8_t* out, uint32_t* sizes,
                        unsigned long long* chunk_total)
{
    const JpegGeom& g = w.geom;
    L.run("jpeg_dct", [&](cudaStream_t s) {
        dim3 grid(divup(g.mcux, 4), g.mcuy, batch);
        k_jpeg_dct<<<grid, DCT_THREADS, 0, s>>>(bgr, bgr_item, g, w.tables, w.coef, w.blk_mask);
    });
    L.run("jpeg_count", [&](cudaStream_t s) {
        dim3 grid(divup(g.nblk, HUF_THREADS), batch);
        k_jpeg_count<<<grid, HUF_THREADS, 0, s>>>(w.coef, w.blk_mask, g, w.tables, w.blk_bits);
    });
    const int ncta = divup(g.nblk, HUF_THREADS);
    L.run("jpeg_scan", [&](cudaStream_t s) {
        k_jpeg_ctasum<<<dim3(ncta, batch), HUF_THREADS, 0, s>>>(w.blk_bits, g, ncta, w.cta_bits);
        k_jpeg_scan<<<batch, SCAN_THREADS, 0, s>>>(w.cta_bits, ncta, g, w.bits32, w.total_bits);
    });
    L.run("jpeg_emit", [&](cudaStream_t s) {
        k_jpeg_emit<<<dim3(ncta, batch), HUF_THREADS, 0, s>>>(w.coef, w.blk_mask, g, w.tables, w.blk_bits, w.cta_bits, ncta, w.bits32);
    });
    // the number of 4 KB segments a picture really has is only known on the device: launch for a bound derived from the
    // picture size (a q95 stream is far below 1 byte per pixel) and let empty segments return at once; the bound is checked
    // by k_jpeg_layout through out_cap
    const int nseg = std::min(g.nseg_cap, (int)(((size_t)g.W * g.H * 3 + JPEG_SEG - 1) / JPEG_SEG));
    L.run("jpeg_ffcount", [&](cudaStream_t s) {
        k_jpeg_ffcount<<<dim3(nseg, batch), SEG_THREADS, 0, s>>>(w.bits32, g, w.total_bits, w.seg_ff);
    });
    L.run("jpeg_layout", [&](cudaStream_t s) {
        k_jpeg_layout<<<1, 1024, sizeof(unsigned long long) * (size_t)batch, s>>>(w.seg_ff, g, w.total_bits, w.header_len, batch, nseg, w.out_off, sizes, chunk_total);
    });
    L.run("jpeg_stuff", [&](cudaStream_t s) {
        k_jpeg_stuff<<<dim3(nseg, batch), SEG_THREADS, 0, s>>>(w.bits32, g, w.total_bits, w.seg_ff, w.out_off, w.tables, out);
    });
}

}  // namespace ofb
