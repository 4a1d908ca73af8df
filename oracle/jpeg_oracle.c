/*
 * jpeg_oracle.c -- CPU restatement of the baseline JPEG encoder behind cv2.imwrite("flow_<ms>.jpeg", rgb)
 * (/root/reference/visualize_optical_flow.py:57-58).  TEST INFRASTRUCTURE ONLY: nothing under optical_flow_b200/
 * links or calls this file; it is the checker for the GPU encoder (optical_flow_b200/csrc/jpeg.cu).
 *
 * The arithmetic lives in an un-vendored dependency: OpenCV's imgcodecs hands the BGR picture to libjpeg(-turbo)
 * (this image: libjpeg-turbo 3.1.2 inside opencv-python-headless 4.13.0) with cv2's defaults: quality 95, 4:2:0 chroma
 * subsampling, baseline sequential DCT (SOF0), the standard Huffman tables of ITU-T T.81 Annex K (no optimisation),
 * no restart markers, a JFIF 1.01 APP0 segment.  The published algorithm restated here:
 *   1. BGR -> YCbCr, 16-bit fixed point (libjpeg jccolor.c):  Y = (19595 R + 38470 G + 7471 B + 32768) >> 16, ...
 *   2. chroma 2x2 box down-sample with the alternating bias 1, 2, 1, 2 (jcsample.c h2v2_downsample); image edges are
 *      padded by replication: columns BEFORE the down-sample, rows AFTER it (jcprepct.c)
 *   3. 8x8 forward DCT, the "islow" integer transform of jfdctint.c (CONST_BITS 13, PASS1_BITS 2) on samples - 128
 *   4. quantisation by 8*q with rounding half away from zero (jcdctmgr.c); tables = Annex K scaled for the quality
 *   5. blocks of an MCU beyond the image ("dummy blocks") are all-zero AC with the DC of the preceding block (jccoefct.c)
 *   6. Huffman coding of DC differences and AC run/size pairs, 0xFF byte stuffing, 1-bit padding at the end (jchuff.c)
 * Pinned byte-for-byte against cv2.imencode on this image's cv2 (tests/test_oracle_vs_golden.py, tests/golden/jpeg_*.npz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static const uint8_t STD_LUMA_Q[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t STD_CHROMA_Q[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t ZIGZAG[64] = {   /* zigzag position -> natural (row-major) index */
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ITU-T T.81 Annex K.3 */
static const uint8_t DC_LUMA_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t DC_CHROMA_BITS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t AC_LUMA_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t AC_LUMA_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
static const uint8_t AC_CHROMA_BITS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t AC_CHROMA_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};

typedef struct { uint16_t code[256]; uint8_t len[256]; } HuffTab;

/* T.81 Annex C: canonical codes from the BITS / HUFFVAL lists */
static void build_huff(const uint8_t* bits, const uint8_t* vals, int nvals, HuffTab* t)
{
    memset(t, 0, sizeof(*t));
    int k = 0; unsigned code = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1] && k < nvals; i++, k++) { t->code[vals[k]] = (uint16_t)code++; t->len[vals[k]] = (uint8_t)l; }
        code <<= 1;
    }
}

/* jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table (force_baseline) */
ORC_API void jpeg_oracle_quant_table(int quality, int chroma, uint8_t out[64])
{
    if (quality <= 0) quality = 1;
    if (quality > 100) quality = 100;
    int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    const uint8_t* base = chroma ? STD_CHROMA_Q : STD_LUMA_Q;
    for (int i = 0; i < 64; i++) {
        long t = ((long)base[i] * scale + 50L) / 100L;
        if (t <= 0) t = 1;
        if (t > 255) t = 255;
        out[i] = (uint8_t)t;
    }
}

/* jfdctint.c (islow): in-place on 64 ints, rows then columns; output scaled by 8 */
static void fdct_islow(int* d)
{
    enum { CB = 13, P1 = 2 };
    const int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137,
              F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
    for (int pass = 0; pass < 2; pass++) {
        const int st = pass == 0 ? 1 : 8, nx = pass == 0 ? 8 : 1;
        for (int i = 0; i < 8; i++) {
            int* p = d + i * nx;
            int t0 = p[0] + p[7 * st], t7 = p[0] - p[7 * st], t1 = p[st] + p[6 * st], t6 = p[st] - p[6 * st];
            int t2 = p[2 * st] + p[5 * st], t5 = p[2 * st] - p[5 * st], t3 = p[3 * st] + p[4 * st], t4 = p[3 * st] - p[4 * st];
            int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            if (pass == 0) { p[0] = (t10 + t11) << P1; p[4 * st] = (t10 - t11) << P1; }
            else { p[0] = DESCALE(t10 + t11, P1); p[4 * st] = DESCALE(t10 - t11, P1); }
            const int sh = pass == 0 ? CB - P1 : CB + P1;
            int z1 = (t12 + t13) * F0541;
            p[2 * st] = DESCALE(z1 + t13 * F0765, sh);
            p[6 * st] = DESCALE(z1 + t12 * (-F1847), sh);
            z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * F1175;
            t4 *= F0298; t5 *= F2053; t6 *= F3072; t7 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            p[7 * st] = DESCALE(t4 + z1 + z3, sh);
            p[5 * st] = DESCALE(t5 + z2 + z4, sh);
            p[3 * st] = DESCALE(t6 + z2 + z3, sh);
            p[st] = DESCALE(t7 + z1 + z4, sh);
        }
    }
#undef DESCALE
}

typedef struct { uint8_t* out; size_t cap, n; uint64_t acc; int nbits; int overflow; } BitW;

static void put_byte(BitW* w, unsigned b) { if (w->n < w->cap) w->out[w->n] = (uint8_t)b; else w->overflow = 1; w->n++; }
static void put_bits(BitW* w, unsigned code, int len)
{
    w->acc = (w->acc << len) | (code & ((1u << len) - 1));
    w->nbits += len;
    while (w->nbits >= 8) {
        unsigned b = (unsigned)(w->acc >> (w->nbits - 8)) & 0xff;
        put_byte(w, b);
        if (b == 0xff) put_byte(w, 0);
        w->nbits -= 8;
    }
}
static int nbits_of(int v) { int n = 0; if (v < 0) v = -v; while (v) { n++; v >>= 1; } return n; }

static void encode_block(BitW* w, const int16_t* zz /* zigzag order */, int* last_dc, const HuffTab* dc, const HuffTab* ac)
{
    int diff = zz[0] - *last_dc;
    *last_dc = zz[0];
    int n = nbits_of(diff);
    put_bits(w, dc->code[n], dc->len[n]);
    if (n) put_bits(w, (unsigned)(diff < 0 ? diff - 1 : diff), n);
    int run = 0;
    for (int k = 1; k < 64; k++) {
        int v = zz[k];
        if (v == 0) { run++; continue; }
        while (run > 15) { put_bits(w, ac->code[0xf0], ac->len[0xf0]); run -= 16; }
        n = nbits_of(v);
        int sym = (run << 4) | n;
        put_bits(w, ac->code[sym], ac->len[sym]);
        put_bits(w, (unsigned)(v < 0 ? v - 1 : v), n);
        run = 0;
    }
    if (run > 0) put_bits(w, ac->code[0], ac->len[0]);
}

static size_t put_marker_seg(uint8_t* o, size_t n, int marker, const uint8_t* payload, int len)
{
    o[n++] = 0xff; o[n++] = (uint8_t)marker; o[n++] = (uint8_t)((len + 2) >> 8); o[n++] = (uint8_t)((len + 2) & 0xff);
    memcpy(o + n, payload, (size_t)len);
    return n + (size_t)len;
}

/* SOI, APP0 (JFIF 1.01, density 1:1), DQT x2, SOF0 (4:2:0), DHT x4, SOS -- as jcmarker.c writes them.  Returns the size. */
ORC_API size_t jpeg_oracle_header(int W, int H, int quality, uint8_t* o)
{
    size_t n = 0;
    o[n++] = 0xff; o[n++] = 0xd8;
    const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    n = put_marker_seg(o, n, 0xe0, app0, 14);
    for (int c = 0; c < 2; c++) {
        uint8_t q[64], seg[65];
        jpeg_oracle_quant_table(quality, c, q);
        seg[0] = (uint8_t)c;
        for (int i = 0; i < 64; i++) seg[1 + i] = q[ZIGZAG[i]];
        n = put_marker_seg(o, n, 0xdb, seg, 65);
    }
    const uint8_t sof[15] = {8, (uint8_t)(H >> 8), (uint8_t)(H & 255), (uint8_t)(W >> 8), (uint8_t)(W & 255), 3,
                             1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    n = put_marker_seg(o, n, 0xc0, sof, 15);
    const uint8_t* bits[4] = {DC_LUMA_BITS, AC_LUMA_BITS, DC_CHROMA_BITS, AC_CHROMA_BITS};
    const uint8_t* vals[4] = {DC_VALS, AC_LUMA_VALS, DC_VALS, AC_CHROMA_VALS};
    const int nv[4] = {12, 162, 12, 162}, id[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; t++) {
        uint8_t seg[1 + 16 + 162];
        seg[0] = (uint8_t)id[t];
        memcpy(seg + 1, bits[t], 16);
        memcpy(seg + 17, vals[t], (size_t)nv[t]);
        n = put_marker_seg(o, n, 0xc4, seg, 17 + nv[t]);
    }
    const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    n = put_marker_seg(o, n, 0xda, sos, 10);
    return n;
}

/* Quantised coefficients of the whole picture in scan order: MCU-major, 6 blocks per MCU (Y00 Y01 Y10 Y11 Cb Cr), 64 int16
 * per block in ZIGZAG order.  coef must hold mcux*mcuy*6*64 values.  Also the stage the GPU test compares first. */
ORC_API void jpeg_oracle_coefficients(const uint8_t* bgr, int W, int H, int quality, int16_t* coef)
{
    const int mcux = (W + 15) / 16, mcuy = (H + 15) / 16;
    const int yw = ((W + 7) / 8) * 8, yh = ((H + 7) / 8);                 /* Y: padded width, block rows */
    const int cw_real = (W + 1) / 2, ch_real = (H + 1) / 2;
    const int cwb = (cw_real + 7) / 8, chb = (ch_real + 7) / 8;            /* chroma blocks */
    const int cw = cwb * 8;
    uint8_t ql[64], qc[64];
    jpeg_oracle_quant_table(quality, 0, ql);
    jpeg_oracle_quant_table(quality, 1, qc);
    /* full-resolution planes, columns replicated to 2*cw (>= yw), rows to an even count (jcprepct.c: the conversion buffer of
     * max_v_samp_factor rows is padded at the bottom of the image BEFORE down-sampling) */
    const int fw = 2 * cw > yw ? 2 * cw : yw, fh = (H + 1) & ~1;
    uint8_t* Y = (uint8_t*)malloc((size_t)fw * fh);
    uint8_t* Cb = (uint8_t*)malloc((size_t)fw * fh);
    uint8_t* Cr = (uint8_t*)malloc((size_t)fw * fh);
    for (int y = 0; y < fh; y++) {
        const int sy = y < H ? y : H - 1;
        for (int x = 0; x < fw; x++) {
            const int sx = x < W ? x : W - 1;
            const uint8_t* p = bgr + ((size_t)sy * W + sx) * 3;
            const int b = p[0], g = p[1], r = p[2];
            Y[(size_t)y * fw + x] = (uint8_t)((19595 * r + 38470 * g + 7471 * b + 32768) >> 16);
            Cb[(size_t)y * fw + x] = (uint8_t)((-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16);
            Cr[(size_t)y * fw + x] = (uint8_t)((32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16);
        }
    }
    /* h2v2 down-sample: bias alternates 1, 2 along a row; rows beyond the last real chroma row replicate it (AFTER sampling) */
    const int chp = chb * 8;
    uint8_t* cb2 = (uint8_t*)malloc((size_t)cw * chp);
    uint8_t* cr2 = (uint8_t*)malloc((size_t)cw * chp);
    for (int y = 0; y < chp; y++) {
        const int sy = y < ch_real ? y : ch_real - 1;
        for (int x = 0; x < cw; x++) {
            const int bias = 1 + (x & 1);
            const uint8_t* a = Cb + (size_t)(2 * sy) * fw + 2 * x;
            cb2[(size_t)y * cw + x] = (uint8_t)((a[0] + a[1] + a[fw] + a[fw + 1] + bias) >> 2);
            a = Cr + (size_t)(2 * sy) * fw + 2 * x;
            cr2[(size_t)y * cw + x] = (uint8_t)((a[0] + a[1] + a[fw] + a[fw + 1] + bias) >> 2);
        }
    }
    const int ywb = yw / 8;
    for (int my = 0; my < mcuy; my++)
        for (int mx = 0; mx < mcux; mx++) {
            int16_t* mc = coef + ((size_t)my * mcux + mx) * 6 * 64;
            for (int b = 0; b < 6; b++) {
                int16_t* out = mc + b * 64;
                int blk[64];
                int real;
                const uint8_t* q;
                if (b < 4) {
                    const int bx = 2 * mx + (b & 1), by = 2 * my + (b >> 1);
                    real = bx < ywb && by < yh;
                    q = ql;
                    if (real)
                        for (int i = 0; i < 8; i++) {
                            int sy = by * 8 + i;
                            if (sy >= H) sy = H - 1;                    /* rows padded after "down-sampling" (a copy for Y) */
                            for (int j = 0; j < 8; j++) blk[i * 8 + j] = (int)Y[(size_t)sy * fw + bx * 8 + j] - 128;
                        }
                } else {
                    real = mx < cwb && my < chb;
                    q = qc;
                    const uint8_t* src = b == 4 ? cb2 : cr2;
                    if (real)
                        for (int i = 0; i < 8; i++)
                            for (int j = 0; j < 8; j++) blk[i * 8 + j] = (int)src[(size_t)(my * 8 + i) * cw + mx * 8 + j] - 128;
                }
                if (!real) {                                             /* dummy block: zero AC, DC of the preceding block */
                    memset(out, 0, 64 * sizeof(int16_t));
                    out[0] = out[-64];
                    continue;
                }
                fdct_islow(blk);
                for (int k = 0; k < 64; k++) {
                    const int nat = ZIGZAG[k], d = 8 * q[nat];
                    int v = blk[nat];
                    v = v < 0 ? -((-v + d / 2) / d) : (v + d / 2) / d;
                    out[k] = (int16_t)v;
                }
            }
        }
    free(Y); free(Cb); free(Cr); free(cb2); free(cr2);
}

/* The whole file.  Returns the number of bytes the stream needs (> cap means it did not fit). */
ORC_API size_t jpeg_oracle_encode(const uint8_t* bgr, int W, int H, int quality, uint8_t* out, size_t cap)
{
    const int mcux = (W + 15) / 16, mcuy = (H + 15) / 16;
    int16_t* coef = (int16_t*)malloc((size_t)mcux * mcuy * 6 * 64 * sizeof(int16_t));
    jpeg_oracle_coefficients(bgr, W, H, quality, coef);
    uint8_t hdr[1024];
    size_t hn = jpeg_oracle_header(W, H, quality, hdr);
    if (cap >= hn) memcpy(out, hdr, hn);
    HuffTab dcl, acl, dcc, acc;
    build_huff(DC_LUMA_BITS, DC_VALS, 12, &dcl);
    build_huff(AC_LUMA_BITS, AC_LUMA_VALS, 162, &acl);
    build_huff(DC_CHROMA_BITS, DC_VALS, 12, &dcc);
    build_huff(AC_CHROMA_BITS, AC_CHROMA_VALS, 162, &acc);
    BitW w = {out, cap, hn, 0, 0, cap < hn};
    int last[3] = {0, 0, 0};
    for (size_t m = 0; m < (size_t)mcux * mcuy; m++)
        for (int b = 0; b < 6; b++) {
            const int c = b < 4 ? 0 : b - 3;
            encode_block(&w, coef + (m * 6 + b) * 64, &last[c], c ? &dcc : &dcl, c ? &acc : &acl);
        }
    if (w.nbits > 0) put_bits(&w, 0x7f, 8 - w.nbits);       /* pad the last byte with 1-bits */
    put_byte(&w, 0xff); put_byte(&w, 0xd9);
    free(coef);
    return w.n;
}
