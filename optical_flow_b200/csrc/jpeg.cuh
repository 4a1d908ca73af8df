// jpeg.cuh -- declarations of the GPU baseline-JPEG encoder (jpeg.cu): the picture of visualize_optical_flow.py:57-58 as the
// byte stream cv2.imwrite would write (quality 95, 4:2:0, Annex-K Huffman tables).
#pragma once
#include "common.cuh"

namespace ofb {

constexpr int JPEG_MAX_BLOCK_BITS = 1664;    // DC 9 + 11, 63 x (16 + 10) AC bits, rounded up
constexpr int JPEG_SEG = 4096;               // bytes of the unstuffed stream per stuffing CTA

struct JpegTables {                          // one per (W, H, quality), lives in device memory
    uint32_t recip[2][64];                   // ceil(2^32 / (8 q)) by natural index; [0] luma, [1] chroma
    uint16_t half[2][64];                    // 8 q / 2
    uint8_t zz_of_nat[64];                   // natural index -> zigzag position
    uint32_t dc[2][256];                     // (code << 8) | length by category
    uint32_t ac[2][256];                     // (code << 8) | length by (run << 4 | size)
    uint8_t header[640];                     // SOI .. SOS
    int header_len;
};

struct JpegGeom {
    int W, H, mcux, mcuy, nblk;              // 16 x 16 MCUs, 6 blocks each
    int ywb, yhb, ch_real;                   // luma blocks per row / column that are real; real chroma rows
    size_t bits_cap;                         // bytes reserved per picture for the unstuffed stream (worst case)
    int nseg_cap;
    size_t out_cap;                          // largest stuffed stream (with header) a picture may have
};

struct JpegWork {                            // device workspaces for `batch` pictures
    JpegGeom geom;
    const JpegTables* tables;
    int header_len;
    int16_t* coef;                           // batch x nblk x 64
    unsigned long long* blk_mask;            // batch x nblk: bit k = AC coefficient at zigzag position k is non-zero
    uint32_t* blk_bits;                      // batch x nblk: code length in bits
    uint32_t* cta_bits;                      // batch x ceil(nblk / 128): code length, then bit offset, of every emit CTA
    uint32_t* bits32;                        // batch x bits_cap bytes
    uint32_t* total_bits;                    // batch
    uint32_t* seg_ff;                        // batch x nseg_cap
    unsigned long long* out_off;             // batch
};

void jpeg_build_tables(int W, int H, int quality, JpegTables& t);
JpegGeom jpeg_geometry(int W, int H);
// `batch` pictures (H, W, 3) uint8 BGR, bgr_item bytes apart -> their JPEG streams packed back to back in `out`;
// sizes[z] = bytes of picture z, chunk_total[0] = sum, chunk_total[1] != 0 if a stream exceeded geom.out_cap.
void launch_jpeg_encode(Launch& L, const JpegWork& w, const uint8_t* bgr, size_t bgr_item, int batch, uint8_t* out, uint32_t* sizes,
                        unsigned long long* chunk_total);

}  // namespace ofb
