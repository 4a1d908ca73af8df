/*
 * optflow_b200.h -- C ABI of the B200-native Farneback + HSV-visualisation engine.
 *
 * This is the drop-in boundary for the ONE hot path of JacobLoe/optical_flow
 * (SURVEY.md section 8b).  Every entry point names the reference interface it replaces:
 *
 *   ofb_farneback_*       cv2.calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize,
 *                         iterations, poly_n, poly_sigma, flags)
 *                           /root/reference/optical_flow.py:51-59
 *                           /root/reference/visualize_optical_flow.py:38-46
 *   ofb_cart_to_polar_*   cv2.cartToPolar(flow[...,0], flow[...,1])
 *                           optical_flow.py:61, visualize_optical_flow.py:48
 *   ofb_sum_magnitude_*   np.sum(mag)                                  optical_flow.py:64
 *   ofb_flow_to_bgr_*     hsv[...,0]=ang*180/pi; hsv[...,1]=255; hsv[...,2]=normalize(mag,0,255,MINMAX);
 *                         cvtColor(hsv, COLOR_HSV2BGR)                 visualize_optical_flow.py:51-55
 *   ofb_pair_*            calculate_optical_flow(frame1, frame2)       optical_flow.py:49-66  (feature)
 *                         the loop body                                visualize_optical_flow.py:37-55 (picture)
 *   ofb_shot_*            the sequential per-pair loops                visualize_optical_flow.py:21-63,
 *                                                                      optical_flow.py:83-99
 *   ofb_*_jpeg            cv2.imwrite(os.path.join(images_path, 'flow_<ms>.jpeg'), rgb)   visualize_optical_flow.py:57-58
 *                         (the encoder behind it: libjpeg baseline, quality 95, 4:2:0 -- the same bytes, made on the GPU)
 *   ofb_bgr_to_gray_*     cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)      optical_flow.py:44,
 *                                                                      visualize_optical_flow.py:31,35
 *   ofb_resize_u8_*       cv2.resize(frame, (w, h))  [INTER_LINEAR]    optical_flow.py:25-31
 *   ofb_*_bgr_host        the same loops fed with the DECODED BGR frames (resize / gray conversion on the GPU)
 *
 * Plain C: pointers and sizes only, no C++ or torch types.  All functions return 0 on success or
 * a negative ofb_status; none throws.  ofb_last_error() gives the message for the last failure on
 * a context.  Calls on one context are stream-ordered; use one context per GPU per host thread.
 * There is NO CPU fallback: without a CUDA device ofb_create fails with OFB_ERR_NO_DEVICE.
 *
 * "_host" entry points take host pointers (pageable or pinned) and perform the H2D / D2H copies;
 * "_device" entry points take device pointers on the context's device and leave results there.
 */
#ifndef OPTFLOW_B200_H
#define OPTFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFB_ABI_VERSION 4

/* flags of cv2.calcOpticalFlowFarneback */
#define OFB_OPTFLOW_USE_INITIAL_FLOW 4
#define OFB_OPTFLOW_FARNEBACK_GAUSSIAN 256

typedef enum ofb_status {
    OFB_OK = 0,
    OFB_ERR_CUDA = -1,          /* a CUDA runtime call failed; see ofb_last_error */
    OFB_ERR_NO_DEVICE = -2,     /* no usable CUDA device: the engine never falls back to the CPU */
    OFB_ERR_BAD_ARG = -3,       /* null pointer, bad size, unsupported dtype */
    OFB_ERR_UNSUPPORTED = -4,   /* parameter outside what the kernels implement (message says which) */
    OFB_ERR_ASSERT = -215       /* cv2's CV_StsAssert: size mismatch, pyr_scale >= 1, bad initial flow */
} ofb_status;

/* pixel type of the input frames (cv2 converts any depth to f32 first) */
typedef enum ofb_dtype { OFB_U8 = 0, OFB_F32 = 1 } ofb_dtype;

/* The seven algorithm parameters, same meaning and order as the cv2 call. */
typedef struct ofb_params {
    double pyr_scale;
    int levels;
    int winsize;
    int iterations;
    int poly_n;
    double poly_sigma;
    int flags;
} ofb_params;

typedef struct ofb_context ofb_context;

/* ---- lifecycle ---------------------------------------------------------- */
int ofb_abi_version(void);
/* Error text for failures that happen before a context exists (ofb_create). Thread-local. */
const char* ofb_global_error(void);
int ofb_device_count(void);
int ofb_create(int device, ofb_context** out);
void ofb_destroy(ofb_context* ctx);
const char* ofb_last_error(const ofb_context* ctx);
int ofb_device_of(const ofb_context* ctx);
int ofb_sm_count(const ofb_context* ctx);
int ofb_synchronize(ofb_context* ctx);

/* Pinned host memory for overlapped uploads / downloads (cudaHostAlloc / cudaFreeHost). */
void* ofb_host_alloc(size_t bytes);
void ofb_host_free(void* p);
/* Plain device memory on the context's device (cudaMalloc / cudaFree), for the _device entry points. */
void* ofb_device_alloc(ofb_context* ctx, size_t bytes);
void ofb_device_free(ofb_context* ctx, void* p);
int ofb_memcpy_h2d(ofb_context* ctx, void* dst, const void* src, size_t bytes);
int ofb_memcpy_d2h(ofb_context* ctx, void* dst, const void* src, size_t bytes);

/* ---- the drop-in call ----------------------------------------------------
 * flow: (H, W, 2) float32, C-contiguous, channel 0 = dx, 1 = dy.  Written in place; also READ when
 * flags & OFB_OPTFLOW_USE_INITIAL_FLOW.  Pitches are in bytes (0 = tightly packed).
 * Errors mirror cv2: OFB_ERR_ASSERT for pyr_scale >= 1 or W/H <= 0. */
int ofb_farneback_host(ofb_context* ctx, const void* prev, const void* next, int dtype, int W, int H,
                       size_t prev_pitch, size_t next_pitch, float* flow, const ofb_params* p);
int ofb_farneback_device(ofb_context* ctx, const void* d_prev, const void* d_next, int dtype, int W, int H,
                         size_t prev_pitch, size_t next_pitch, float* d_flow, const ofb_params* p);

/* ---- companions of the call (Appendix B of SURVEY.md) ---------------------- */
int ofb_cart_to_polar_host(ofb_context* ctx, const float* flow, int W, int H, float* mag, float* ang);
/* the same with cv2's angleInDegrees flag */
int ofb_cart_to_polar_host2(ofb_context* ctx, const float* flow, int W, int H, float* mag, float* ang, int angle_in_degrees);
int ofb_sum_magnitude_host(ofb_context* ctx, const float* flow, int W, int H, float* out);
int ofb_flow_to_bgr_host(ofb_context* ctx, const float* flow, int W, int H, uint8_t* bgr);
int ofb_flow_to_bgr_device(ofb_context* ctx, const float* d_flow, int W, int H, uint8_t* d_bgr);
int ofb_sum_magnitude_device(ofb_context* ctx, const float* d_flow, int W, int H, float* d_out);

/* ---- fused per-pair and per-shot forms ------------------------------------
 * Any of bgr / magsum / flow may be NULL (not produced / not copied back).
 * ofb_shot_*: n_frames consecutive u8 frames (tightly packed, n_frames*H*W bytes) -> n_frames-1 results:
 *   bgr    (n_frames-1, H, W, 3) uint8      magsum (n_frames-1) float32     flow (n_frames-1, H, W, 2) float32
 * Per-frame work (pyramid, polynomial expansion) is computed once per frame and shared by the two
 * pairs the frame belongs to.  device_ms (may be NULL) receives the elapsed time between CUDA events
 * recorded on the engine's streams around the whole shot (copies included for the _host form). */
int ofb_pair_host(ofb_context* ctx, const void* prev, const void* next, int dtype, int W, int H,
                  const ofb_params* p, uint8_t* bgr, float* magsum, float* flow);
int ofb_shot_host(ofb_context* ctx, const uint8_t* frames, int n_frames, int W, int H, const ofb_params* p,
                  uint8_t* bgr, float* magsum, float* flow, float* device_ms);
/* ofb_shot_host with the frames at n_frames separate addresses (a decoder's own buffers, pinned or pageable): frame i is
 * frames[i], W*H bytes.  Nothing is assembled on the host; every frame is uploaded straight from where it lies. */
int ofb_shot_host_v(ofb_context* ctx, const uint8_t* const* frames, int n_frames, int W, int H, const ofb_params* p,
                    uint8_t* bgr, float* magsum, float* flow, float* device_ms);
/* ofb_shot_host whose pictures leave the GPU as JPEG files instead of raw BGR (visualize_optical_flow.py:57-58): the byte
 * streams cv2.imwrite(..., rgb) / cv2.imencode('.jpeg', rgb) produce for the same picture (baseline, `quality` as
 * IMWRITE_JPEG_QUALITY, cv2's default is 95; 4:2:0; standard Huffman tables), byte for byte.  The n_frames-1 streams are
 * written back to back into `jpeg` (capacity jpeg_cap bytes; OFB_ERR_BAD_ARG if it is too small), jpeg_sizes[i] = bytes of
 * stream i (stream i starts at the sum of the sizes before it).  magsum may be NULL. */
int ofb_shot_host_jpeg(ofb_context* ctx, const uint8_t* frames, int n_frames, int W, int H, const ofb_params* p, int quality,
                       uint8_t* jpeg, size_t jpeg_cap, uint32_t* jpeg_sizes, float* magsum, float* device_ms);
/* ofb_shot_host_jpeg with the frames at separate addresses (as ofb_shot_host_v). */
int ofb_shot_host_v_jpeg(ofb_context* ctx, const uint8_t* const* frames, int n_frames, int W, int H, const ofb_params* p, int quality,
                         uint8_t* jpeg, size_t jpeg_cap, uint32_t* jpeg_sizes, float* magsum, float* device_ms);
/* The same with decoded BGR frames as input (gray conversion / resize on the GPU, as ofb_shot_bgr_host). */
int ofb_shot_bgr_host_jpeg(ofb_context* ctx, const uint8_t* bgr_frames, int n_frames, int W, int H, int dW, int dH, const ofb_params* p,
                           int quality, uint8_t* jpeg, size_t jpeg_cap, uint32_t* jpeg_sizes, float* magsum, float* device_ms);
/* The encoder on its own: n pictures (n, H, W, 3) uint8 BGR -> n JPEG streams, packed as above. */
int ofb_jpeg_encode_host(ofb_context* ctx, const uint8_t* bgr, int n, int W, int H, int quality, uint8_t* jpeg, size_t jpeg_cap,
                         uint32_t* jpeg_sizes);
/* n_pairs INDEPENDENT pairs (prev[i], next[i]), each (H, W) uint8 tightly packed: the window loop of
 * optical_flow.py:83-99, whose pairs need not share frames.  Same outputs as ofb_shot_host. */
int ofb_pairs_host(ofb_context* ctx, const uint8_t* prev, const uint8_t* next, int n_pairs, int W, int H,
                   const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, float* device_ms);
int ofb_shot_device(ofb_context* ctx, const uint8_t* d_frames, int n_frames, int W, int H, const ofb_params* p,
                    uint8_t* d_bgr, float* d_magsum, float* d_flow, float* device_ms);
/* Pairs per launch ("chunk") the ofb_shot_* / ofb_pairs_* entry points use for W x H frames and a job of n_pairs pairs: option
 * "batch" when it is set, else enough pixels per launch to fill the GPU at the coarse scales (96e6 / (W*H), a multiple of 4; a
 * job shorter than four such chunks is cut into about four).  A long job runs chunks of B/4, B/2, B, ..., B, B/2, B/4 pairs; results never depend on the chunking (tests pin that), the
 * function exists so that tests and bench.py can aim their parity checks at the chunk seams. */
int ofb_shot_chunk(const ofb_context* ctx, int W, int H, int n_pairs);

/* ---- frame preprocessing on the GPU -------------------
 * Both are integer algorithms and bit-exact against cv2 (tests/golden/preprocess.npz).
 * ofb_bgr_to_gray_host: (H, W, 3) uint8 BGR -> (H, W) uint8, gray = (3735 B + 19235 G + 9798 R + 16384) >> 15.
 * ofb_resize_u8_host:   cv2.resize(src, (dW, dH)) with the default INTER_LINEAR for 1 or 3 interleaved channels;
 *                       to_gray != 0 (3 channels only) applies the gray conversion to the resized pixel and writes
 *                       (dH, dW) uint8 -- read_frame of optical_flow.py:34-46 in one pass.
 * ofb_shot_bgr_host / ofb_pairs_bgr_host: as ofb_shot_host / ofb_pairs_host, but the frames are the decoded BGR
 * frames (n, H, W, 3); dW = dH = 0 keeps the size, otherwise every frame is first resized to dW x dH.  All outputs
 * have the size of the gray frames (dW x dH).  `gray` (may be NULL) receives the n gray frames. */
int ofb_bgr_to_gray_host(ofb_context* ctx, const uint8_t* bgr, int W, int H, uint8_t* gray);
int ofb_resize_u8_host(ofb_context* ctx, const uint8_t* src, int W, int H, int channels, int dW, int dH, int to_gray,
                       uint8_t* dst);
int ofb_shot_bgr_host(ofb_context* ctx, const uint8_t* bgr_frames, int n_frames, int W, int H, int dW, int dH,
                      const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, uint8_t* gray, float* device_ms);
int ofb_pairs_bgr_host(ofb_context* ctx, const uint8_t* prev_bgr, const uint8_t* next_bgr, int n_pairs, int W, int H,
                       int dW, int dH, const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, float* device_ms);

/* ---- per-stage entry points (parity tests; host arrays in cv2's layouts) ---
 * R and M are (H, W, 5) float32 interleaved as in OpenCV; flow is (H, W, 2). */
int ofb_scale_count(int W, int H, double pyr_scale, int levels);           /* K: scales K..0 run */
int ofb_scale_geometry(int W, int H, double pyr_scale, int k, int* Wk, int* Hk, int* ksize, double* sigma);
int ofb_stage_level_image(ofb_context* ctx, const void* frame, int dtype, int W, int H, double pyr_scale, int k,
                          float* out /* Hk*Wk */);
int ofb_stage_polyexp(ofb_context* ctx, const float* img, int W, int H, int poly_n, double poly_sigma, float* R);
int ofb_stage_update_matrices(ofb_context* ctx, const float* R0, const float* R1, const float* flow,
                              int W, int H, float* M);
int ofb_stage_blur_solve(ofb_context* ctx, const float* M, int W, int H, int winsize, int gaussian, float* flow);
/* quantised DCT coefficients of one picture in scan order: ceil(W/16)*ceil(H/16) MCUs x 6 blocks x 64 int16 (zigzag order) */
int ofb_stage_jpeg_coefficients(ofb_context* ctx, const uint8_t* bgr, int W, int H, int quality, int16_t* coef);
int ofb_stage_upsample_flow(ofb_context* ctx, const float* prev_flow, int Wp, int Hp, int W, int H,
                            double pyr_scale, float* flow);

/* ---- options and measurement ----------------------------------------------
 * Options (all default 0):
 *   "generic_kernels"  1 = force the simple global-memory kernels (any winsize / poly_n)
 *   "profile"          1 = bracket every kernel launch with CUDA events on the launching stream
 *   "batch"            pairs per launch inside a shot (0 = chosen from the frame size)
 *   "batch_scale0"     pairs per launch at scale 0 (0 = same as batch)
 *   "iter_prefetch"    1 (default) = software L2 prefetch in the iteration kernel
 *   "hsv_table"        1 (default) = the picture kernel looks HSV->BGR up in a 65536-entry table built by the same
 *                      arithmetic at ofb_create; 0 = evaluates the conversion per pixel (identical output)
 *   "polyexp_tma"      1 = scale-0 polynomial expansion as a persistent grid whose halo tiles are staged by TMA
 *                      (cp.async.bulk.tensor + mbarrier, double-buffered); bit-identical results, measured slower
 *                      than the default one-tile-per-CTA kernel on B200, so off by default */
int ofb_set_option(ofb_context* ctx, const char* name, int value);
/* Every workspace of the current plan carries a 256-byte guard band on both sides: returns how many guard bytes have been
 * overwritten since the plan was built (0 = no out-of-bounds write next to a buffer), or a negative status. */
int ofb_debug_check_guards(ofb_context* ctx);

typedef struct ofb_kernel_stat {
    char name[48];
    uint64_t launches;
    double total_ms;          /* sum of CUDA-event durations (only with option "profile") */
} ofb_kernel_stat;
/* Copies up to max entries; returns the number of distinct kernels seen.  Launch counts are always
 * kept; durations only under "profile". */
int ofb_get_kernel_stats(ofb_context* ctx, ofb_kernel_stat* out, int max);
void ofb_reset_kernel_stats(ofb_context* ctx);

/* Algorithmic HBM bytes of one frame pair: B_pair, and B_viz = 19*W*H. */
double ofb_algorithmic_bytes_pair(int W, int H, const ofb_params* p);
double ofb_algorithmic_bytes_viz(int W, int H);

#ifdef __cplusplus
}
#endif
#endif /* OPTFLOW_B200_H */
