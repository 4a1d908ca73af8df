// engine.cu -- context, workspaces, the per-scale schedule of cv2.calcOpticalFlowFarneback
//  and the C ABI declared in include/optflow_b200.h.
//
// Host logic mirrored from the reference's dependency (FarnebackOpticalFlowImpl::calc):
//   K = number of extra scales (coarsest >= 32 px); for k = K..0:
//     flow_k   = zeros | area-resized initial flow * scale | bilinear up-sample of flow_{k+1} * 1/pyr_scale
//     R[i]     = polyexp(level_image(frame_i, k))           i = 0, 1
//     M        = UpdateMatrices(R0, R1, flow_k)
//     repeat iterations:  flow_k = solve(blur(M));  M = UpdateMatrices(...) unless last
//
// B200-first re-organisation:
//   * per-frame work (level images + polynomial expansion, every scale) is computed ONCE per frame into a
//     ring of frame slots and shared by the two pairs the frame belongs to;
//   * pairs of a shot are independent, so a CHUNK of `batch` pairs is processed per launch (blockIdx.z):
//     the coarse scales (a few CTAs per pair) fill the 148 SMs and launch latency is amortised;
//   * middle iterations run as ONE kernel (blur -> solve -> UpdateMatrices), M ping-pongs between two
//     buffers, the flow of a middle iteration never reaches memory, and the first UpdateMatrices of a
//     scale does the inter-scale up-sample on the fly (k_um0, k_iter in iter.cu);
//   * scale 0's pre-blur is fused into the polynomial expansion (k_polyexp2 SRC 1/2).
// Anything outside the fast paths (Gaussian window, winsize > 33, poly_n other than 3/5/7, iterations == 0,
// option "generic_kernels") runs the simple per-item global-memory kernels with identical results.
#include "common.cuh"
#include "launch.cuh"
#include "jpeg.cuh"
#include "../../include/optflow_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace ofb;

namespace {

thread_local std::string g_error;

struct Level {
    int k = 0, W = 0, H = 0, pitch = 0, ksize = 0;
    double sigma = 0, scale = 1;
    float* taps = nullptr;              // device, ksize floats
    std::vector<float> taps_h;          // host copy
    int* sx = nullptr; float* ax = nullptr;     // bilinear tables (k >= 1)
    int* sy = nullptr; float* ay = nullptr;
    int* ux = nullptr; float* uax = nullptr;    // bilinear tables from scale k+1 to this scale (flow up-sample)
    int* uy = nullptr; float* uay = nullptr;
    float* R = nullptr;                 // ring of nslots x 5 planes
    float2* flow = nullptr;             // batch x (H, W) float2 (k >= 1)
    size_t plane() const { return (size_t)H * pitch; }
    size_t slot_stride() const { return 5 * plane(); }
    size_t flow_item() const { return (size_t)W * H; }
};

struct Plan {
    bool valid = false;
    int W = 0, H = 0, dtype = 0, batch = 1, nslots = 2;
    ofb_params p{};
    std::vector<Level> lv;              // index = k (0 = full resolution)
    int K = 0;
    float* T = nullptr; size_t t_item = 0;      // batch x (H x pitch_1): row-pass intermediate of the pyramid
    float* I = nullptr; size_t i_item = 0;      // batch x level image
    float* If[3] = {nullptr, nullptr, nullptr}; // batch x level image of scales 1..3 (one-pass pyramid)
    bool pyr_fused = false;                     // geometry allows k_pyr_fused (pyr_scale 0.5, sizes multiple of 8)
    float* tmp3 = nullptr;              // 3 planes (generic polyexp, one item)
    float* M[2] = {nullptr, nullptr};   // batch x 5 planes, ping-pong
    size_t m_item = 0;
    double* btmp = nullptr;             // 5 planes f64 (generic blur, one item)
    float* poly = nullptr;              // g, xg, xxg : 3 * (2n+1)
    float* gtaps = nullptr;             // half taps of the Gaussian window, m+1
    PolyConst pc{};
    PolyArgs pa{};                      // constants part filled once
    uint8_t* f0 = nullptr;              // device copy of a single frame (first frame of a shot / `prev`)
    uint8_t* fstage[2] = {nullptr, nullptr};    // 2 x (2 * batch) frames: a chunk of a shot, or prev | next of a chunk of pairs
    float2* flow0[2] = {nullptr, nullptr};      // 2 x batch scale-0 flows
    uint8_t* bgr[2] = {nullptr, nullptr};       // 2 x batch pictures
    float* sums = nullptr;              // per-pair magnitude sums of a shot (host API), grown on demand
    size_t sums_cap = 0;
    std::vector<void*> allocs;
    std::vector<std::pair<unsigned char*, size_t>> guards;     // (allocation base, payload bytes) of every guarded workspace
    bool fast_poly = false, fast_iter = false;
    std::vector<float> gk;              // host copy of the Gaussian window half taps
};

}  // namespace

struct ofb_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr, s_expand = nullptr;
    cudaEvent_t ev_expanded[2]{}, ev_solved[2]{};
    bool overlap_expand = true;         // per-frame work of chunk c+1 on s_expand while the per-pair work of chunk c runs on s_compute
    cudaEvent_t ev_h2d[2]{}, ev_frame_free[2]{}, ev_out_ready[2]{}, ev_out_free[2]{}, ev_t0 = nullptr, ev_t1 = nullptr;
    std::string err;
    Profiler prof;
    bool generic = false;
    KernelOptions kopt;                 // per-context kernel options, copied into every Launch
    int batch = 0;                      // pairs per launch inside a shot (0 = choose from the frame size)
    int batch0 = 0;                     // pairs per launch at scale 0 (0 = same as batch)
    bool alt_order = false;             // alternate the grid direction of consecutive iteration launches (L2 reuse of M)
    Plan plan;
    unsigned* minmax = nullptr;         // 2 per batch item
    double* sumacc = nullptr;           // 1 per batch item
    float* sumout = nullptr;            // device staging of magnitude sums (batch items)
    unsigned* hsv_table = nullptr;      // (H byte << 8 | V byte) -> packed BGR, built once (viz.cu)
    bool use_hsv_table = true;          // option "hsv_table"
    float* stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes[6] = {0, 0, 0, 0, 0, 0};
    // GPU JPEG encoder (jpeg.cu): workspaces for `jpeg_batch` pictures of jpeg_W x jpeg_H at jpeg_quality
    JpegWork jw{};
    int jpeg_W = 0, jpeg_H = 0, jpeg_quality = 0, jpeg_batch = 0;
    std::vector<void*> jpeg_allocs;
    std::vector<std::pair<unsigned char*, size_t>> jpeg_guards;
    uint8_t* jout[2] = {nullptr, nullptr};              // packed streams of a chunk, by chunk parity
    unsigned long long* d_tot[2] = {nullptr, nullptr};  // {bytes of the chunk, overflow flag}
    unsigned long long* h_tot = nullptr;                // pinned mirror, 2 x 2
    cudaEvent_t ev_tot[2]{};
    // 8-bit bilinear resize tables of the last (source, destination) geometry (preprocess.cu)
    int rs_geom[4] = {0, 0, 0, 0};      // sW, sH, dW, dH
    int* rs_tab = nullptr;              // x0[dW] x1[dW] y0[dH] y1[dH] | short ax[2dW] ay[2dH]
    ResizeTab rs{};
};

namespace {

constexpr int MAX_BATCH = 1024;        // pairs per launch; small frames (the 129-px feature regime) need hundreds to fill 148 SMs

Launch make_launch(ofb_context* c, cudaStream_t s) { return Launch{s, &c->prof, c->device, c->sm_count, c->kopt}; }

int fail(ofb_context* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_error = msg;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, OFB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

// Error paths of the host entry points return through CU() while asynchronous copies to or from CALLER-owned buffers may
// still be queued: the guard drains the context's three streams before such a return hands the buffers back.
struct DrainOnError {
    ofb_context* c;
    bool armed = true;
    ~DrainOnError()
    {
        if (!armed) return;
        cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_expand); cudaStreamSynchronize(c->s_compute); cudaStreamSynchronize(c->s_d2h);
        cudaGetLastError();
    }
};

struct JpegOut { uint8_t* out; size_t cap; uint32_t* sizes; int quality; };      // host destination of the JPEG delivery

// ---- A.1 ------------------------------------------------------------------------------------
int num_scales(int W, int H, double pyr_scale, int levels)
{
    int k; double scale = 1.0;
    for (k = 0; k < levels; k++) {
        scale *= pyr_scale;
        if (W * scale < 32 || H * scale < 32) break;
    }
    return k;
}

void scale_geometry(int W, int H, double pyr_scale, int k, int* Wk, int* Hk, int* ksize, double* sigma, double* scale_out)
{
    double scale = 1.0;
    for (int i = 0; i < k; i++) scale *= pyr_scale;
    double s = (1.0 / scale - 1.0) * 0.5;
    int sz = (int)lrint(s * 5.0) | 1;
    if (sz < 3) sz = 3;
    *Wk = (int)lrint(W * scale);
    *Hk = (int)lrint(H * scale);
    *ksize = sz; *sigma = s;
    if (scale_out) *scale_out = scale;
}

// cv::getGaussianKernel(n, sigma, CV_32F)
void gaussian_taps(int n, double sigma, std::vector<float>& out)
{
    out.resize(n);
    if (sigma <= 0 && n <= 7 && (n & 1)) {
        static const float k1[] = {1.f}, k3[] = {0.25f, 0.5f, 0.25f}, k5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
                           k7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
        const float* f = n == 1 ? k1 : n == 3 ? k3 : n == 5 ? k5 : k7;
        for (int i = 0; i < n; i++) out[i] = f[i];
        return;
    }
    double sg = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2 = -0.5 / (sg * sg), sum = 0;
    std::vector<double> t(n);
    for (int i = 0; i < n; i++) { double x = i - (n - 1) * 0.5; t[i] = std::exp(scale2 * x * x); sum += t[i]; }
    sum = 1.0 / sum;
    for (int i = 0; i < n; i++) out[i] = (float)(t[i] * sum);
}

// A.5: FarnebackPrepareGaussian
void poly_constants(int n, double sigma, std::vector<float>& tab, double ig[4])
{
    int len = 2 * n + 1;
    tab.assign(3 * len, 0.f);
    float* g = tab.data() + n; float* xg = g + len; float* xxg = xg + len;
    if (sigma < FLT_EPSILON) sigma = n * 0.3;
    double s = 0.;
    for (int x = -n; x <= n; x++) { g[x] = (float)std::exp(-x * x / (2 * sigma * sigma)); s += g[x]; }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x] = (float)(g[x] * s);
        xg[x] = (float)(x * g[x]);
        xxg[x] = (float)(x * x * g[x]);
    }
    double G00 = 0, G11 = 0, G33 = 0, G55 = 0;
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            float gg = g[y] * g[x];
            G00 += gg; G11 += gg * x * x; G33 += gg * x * x * x * x; G55 += gg * x * x * y * y;
        }
    double a = G00, b = G11, c = G33, d = G55, den = a * (c + d) - 2 * b * b;
    ig[0] = 1.0 / G11; ig[1] = -b / den; ig[2] = (a * c - b * b) / ((c - d) * den); ig[3] = 1.0 / G55;
}

void fill_poly_args(PolyArgs& pa, int n, const std::vector<float>& tab, const double ig[4])
{
    memset(&pa, 0, sizeof(pa));
    pa.n = n;
    int len = 2 * n + 1;
    if (n <= 8)
        for (int k = 0; k <= n; k++) {
            pa.g[k] = tab[n + k]; pa.xg[k] = tab[len + n + k]; pa.xxg[k] = tab[2 * len + n + k];
            pa.gd[k] = pa.g[k]; pa.xgd[k] = pa.xg[k]; pa.xxgd[k] = pa.xxg[k];
        }
    pa.ig11 = ig[0]; pa.ig03 = ig[1]; pa.ig33 = ig[2]; pa.ig55 = ig[3];
}

void gauss_half_taps(int winsize, std::vector<float>& k)
{
    int m = winsize / 2;
    k.resize(m + 1);
    double sigma = m * 0.3, s = 1;
    k[0] = (float)s;
    for (int i = 1; i <= m; i++) { float t = (float)std::exp(-i * i / (2 * sigma * sigma)); k[i] = t; s += t * 2; }
    s = 1. / s;
    for (int i = 0; i <= m; i++) k[i] = (float)(k[i] * s);
}

// cv::resize(INTER_LINEAR) coordinate rule.  cv2 has two: the one-channel f32 resize of the level images goes
// through its IPP path, which keeps the coordinate in double; the two-channel resize of the flow field (A.2) runs OpenCV's
// generic path, which rounds the coordinate to f32 BEFORE the floor (`fx = (float)((dx+0.5)*scale_x - 0.5)`).  Both measured
// against the installed wheel (tests/test_oracle_vs_golden.py::test_resize_coordinate_rules_of_cv2); they only differ when
// the scale is not dyadic (pyr_scale != 0.5 or odd sizes).
void linear_table(int dst, int src, std::vector<int>& idx, std::vector<float>& w1, bool f32_coord = false)
{
    idx.resize(dst); w1.resize(dst);
    double scale = 1.0 / ((double)dst / src);
    for (int d = 0; d < dst; d++) {
        double f = (d + 0.5) * scale - 0.5;
        if (f32_coord) f = (double)(float)f;
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= src - 1) { s = src - 1; f = 0; }
        idx[d] = s; w1[d] = (float)f;
    }
}

// ---- plan -----------------------------------------------------------------------------------
void free_plan(Plan& pl)
{
    for (void* p : pl.allocs) cudaFree(p);
    pl = Plan();
}

// Every workspace is allocated with a 256-byte guard band in front and behind, filled with GUARD_BYTE; ofb_debug_check_guards
// counts the guard bytes that no longer hold it.  compute-sanitizer is not available on the B200 pool this was developed on,
// so this (plus the bitwise repeat / batch / shard invariance tests) is how out-of-bounds WRITES next to a buffer are caught.
constexpr size_t GUARD = 256;
constexpr int GUARD_BYTE = 0xA5;

int guarded_alloc(ofb_context* ctx, std::vector<void*>& owner, std::vector<std::pair<unsigned char*, size_t>>& guards, void** out, size_t bytes)
{
    unsigned char* p = nullptr;
    bytes = (bytes + 255) & ~(size_t)255;
    CU(cudaMalloc((void**)&p, bytes + 2 * GUARD));
    CU(cudaMemset(p, GUARD_BYTE, GUARD));
    CU(cudaMemset(p + GUARD + bytes, GUARD_BYTE, GUARD));
    owner.push_back(p);
    guards.push_back({p, bytes});
    *out = p + GUARD;
    return 0;
}

template <class T> int dalloc(ofb_context* ctx, Plan& pl, T** out, size_t count)
{
    void* p = nullptr;
    if (int rc = guarded_alloc(ctx, pl.allocs, pl.guards, &p, count * sizeof(T) + 256)) return rc;
    *out = (T*)p;
    return 0;
}

template <class T> int dupload(ofb_context* ctx, Plan& pl, T** out, const std::vector<T>& v)
{
    if (int rc = dalloc(ctx, pl, out, std::max<size_t>(v.size(), 1))) return rc;
    CU(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

bool same_params(const ofb_params& a, const ofb_params& b)
{
    return a.pyr_scale == b.pyr_scale && a.levels == b.levels && a.winsize == b.winsize && a.iterations == b.iterations &&
           a.poly_n == b.poly_n && a.poly_sigma == b.poly_sigma && a.flags == b.flags;
}

int validate(ofb_context* ctx, int W, int H, int dtype, const ofb_params* p)
{
    if (!ctx) return fail(nullptr, OFB_ERR_BAD_ARG, "null context");
    if (!p) return fail(ctx, OFB_ERR_BAD_ARG, "null params");
    if (dtype != OFB_U8 && dtype != OFB_F32) return fail(ctx, OFB_ERR_BAD_ARG, "dtype must be OFB_U8 or OFB_F32");
    if (W <= 0 || H <= 0 || !(p->pyr_scale < 1))
        return fail(ctx, OFB_ERR_ASSERT,
                    "prev0.size() == next0.size() && prev0.channels() == next0.channels() && prev0.channels() == 1 && pyrScale_ < 1");
    if (!(p->pyr_scale > 0)) return fail(ctx, OFB_ERR_UNSUPPORTED, "pyr_scale must be > 0");
    if (p->poly_n < 1 || p->poly_n > 64) return fail(ctx, OFB_ERR_UNSUPPORTED, "poly_n outside 1..64");
    if (p->winsize < 1 || p->winsize > 1024) return fail(ctx, OFB_ERR_UNSUPPORTED, "winsize outside 1..1024");
    if (p->iterations < 0) return fail(ctx, OFB_ERR_UNSUPPORTED, "iterations < 0");
    return 0;
}

int ensure_plan(ofb_context* ctx, int W, int H, int dtype, const ofb_params* p, int batch)
{
    Plan& pl = ctx->plan;
    batch = std::max(1, std::min(batch, MAX_BATCH));
    if (pl.valid && pl.W == W && pl.H == H && pl.dtype == dtype && same_params(pl.p, *p) && pl.batch >= batch) return 0;
    CU(cudaStreamSynchronize(ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_expand));
    CU(cudaStreamSynchronize(ctx->s_h2d));
    CU(cudaStreamSynchronize(ctx->s_d2h));
    free_plan(pl);
    pl.W = W; pl.H = H; pl.dtype = dtype; pl.p = *p; pl.batch = batch;
    pl.nslots = 2 * batch + 1;          // a chunk's b+1 frames stay readable while the next chunk's b frames are being expanded
    pl.K = num_scales(W, H, p->pyr_scale, p->levels);
    pl.lv.resize(pl.K + 1);
    const size_t B = (size_t)batch;
    for (int k = 0; k <= pl.K; k++) {
        Level& l = pl.lv[k];
        l.k = k;
        scale_geometry(W, H, p->pyr_scale, k, &l.W, &l.H, &l.ksize, &l.sigma, &l.scale);
        if (l.W < 1 || l.H < 1) return fail(ctx, OFB_ERR_UNSUPPORTED, "a pyramid level collapsed to zero size");
        l.pitch = round_up(l.W, 32);
        std::vector<float> taps;
        gaussian_taps(l.ksize, l.sigma, taps);
        if (int rc = dupload(ctx, pl, &l.taps, taps)) return rc;
        l.taps_h = taps;
        std::vector<int> ix, iy; std::vector<float> wx, wy;
        linear_table(l.W, W, ix, wx);
        linear_table(l.H, H, iy, wy);
        if (int rc = dupload(ctx, pl, &l.sx, ix)) return rc;
        if (int rc = dupload(ctx, pl, &l.ax, wx)) return rc;
        if (int rc = dupload(ctx, pl, &l.sy, iy)) return rc;
        if (int rc = dupload(ctx, pl, &l.ay, wy)) return rc;
        if (int rc = dalloc(ctx, pl, &l.R, (size_t)pl.nslots * l.slot_stride())) return rc;
        if (k < pl.K) {
            int Wc, Hc, ksc; double sgc;
            scale_geometry(W, H, p->pyr_scale, k + 1, &Wc, &Hc, &ksc, &sgc, nullptr);
            linear_table(l.W, Wc, ix, wx, true);             // flow up-sample: generic path, f32 coordinate
            linear_table(l.H, Hc, iy, wy, true);
            if (int rc = dupload(ctx, pl, &l.ux, ix)) return rc;
            if (int rc = dupload(ctx, pl, &l.uax, wx)) return rc;
            if (int rc = dupload(ctx, pl, &l.uy, iy)) return rc;
            if (int rc = dupload(ctx, pl, &l.uay, wy)) return rc;
        }
        if (k > 0)
            if (int rc = dalloc(ctx, pl, &l.flow, B * l.flow_item())) return rc;
    }
    const Level& l0 = pl.lv[0];
    const size_t plane0 = l0.plane();
    pl.t_item = (size_t)H * l0.pitch;
    pl.i_item = plane0;
    pl.m_item = 5 * plane0;
    if (int rc = dalloc(ctx, pl, &pl.T, B * pl.t_item)) return rc;
    if (int rc = dalloc(ctx, pl, &pl.I, B * pl.i_item)) return rc;
    if (int rc = dalloc(ctx, pl, &pl.tmp3, 3 * plane0)) return rc;
    if (int rc = dalloc(ctx, pl, &pl.M[0], B * pl.m_item)) return rc;
    if (int rc = dalloc(ctx, pl, &pl.M[1], B * pl.m_item)) return rc;
    if (int rc = dalloc(ctx, pl, &pl.btmp, 5 * plane0)) return rc;
    std::vector<float> tab; double ig[4];
    poly_constants(p->poly_n, p->poly_sigma, tab, ig);
    if (int rc = dupload(ctx, pl, &pl.poly, tab)) return rc;
    int len = 2 * p->poly_n + 1;
    pl.pc = PolyConst{pl.poly, pl.poly + len, pl.poly + 2 * len, p->poly_n, ig[0], ig[1], ig[2], ig[3]};
    fill_poly_args(pl.pa, p->poly_n, tab, ig);
    std::vector<float> gk;
    gauss_half_taps(p->winsize, gk);
    if (int rc = dupload(ctx, pl, &pl.gtaps, gk)) return rc;
    const size_t esz = dtype == OFB_U8 ? 1 : 4, n = (size_t)W * H;
    if (int rc = dalloc(ctx, pl, &pl.f0, n * esz)) return rc;
    for (int s = 0; s < 2; s++) {
        if (int rc = dalloc(ctx, pl, &pl.fstage[s], 2 * B * n * esz)) return rc;
        if (int rc = dalloc(ctx, pl, &pl.flow0[s], B * n)) return rc;
        if (int rc = dalloc(ctx, pl, &pl.bgr[s], B * n * 3)) return rc;
    }
    pl.pyr_fused = dtype == OFB_U8 && p->pyr_scale == 0.5 && W % 8 == 0 && H % 8 == 0 && W >= 64 && H >= 64 && pl.K >= 1;
    for (int k = 1; k <= std::min(pl.K, 3) && pl.pyr_fused; k++) {
        static const int want_ks[4] = {0, 3, 9, 19};
        if (pl.lv[k].ksize != want_ks[k]) { pl.pyr_fused = false; break; }
    }
    if (pl.pyr_fused)
        for (int k = 1; k <= std::min(pl.K, 3); k++)
            if (int rc = dalloc(ctx, pl, &pl.If[k - 1], B * pl.lv[k].plane())) return rc;
    pl.fast_poly = polyexp2_supported(p->poly_n);
    pl.fast_iter = iter_supported(p->winsize) && p->iterations >= 1;
    pl.gk = gk;
    pl.valid = true;
    return 0;
}

RView slot_planes(const Level& l, int slot)
{
    float* p = l.R + (size_t)slot * l.slot_stride();
    return RView{reinterpret_cast<float4*>(p), p + 4 * l.plane(), l.pitch};
}
SlotRing ring(const Plan& pl, const Level& l, int step = 1) { return SlotRing{l.R, l.slot_stride(), l.plane(), l.pitch, pl.nslots, step}; }
Planes5 m_planes(const Plan& pl, const Level& l, int which, int item)
{
    return Planes5{pl.M[which] + (size_t)item * pl.m_item, l.plane(), l.pitch};
}

// Per-frame part for `count` consecutive frames (first one = frame index f0 of the shot -> slot f0 % nslots).
// Frames are `item_bytes` apart starting at d_frames, rows `pitch_bytes` apart.
void expand_frames(ofb_context* ctx, Launch& L, const void* d_frames, size_t item_bytes, size_t pitch_bytes, int f0, int count,
                   int slot_step = 1)
{
    Plan& pl = ctx->plan;
    const bool fast = pl.fast_poly && !ctx->generic && !ctx->kopt.generic_polyexp;
    // scales 1..3 in one pass over the frames when the geometry allows it (pyramid.cu k_pyr_fused)
    int nfused = 0;
    if (fast && pl.pyr_fused && ctx->kopt.pyr_fused && pyr_fused_supported(pl.dtype, pl.W, pl.H, pl.p.pyr_scale, d_frames, pitch_bytes, item_bytes)) {
        nfused = std::min(pl.K, 3);
        PyrFusedLaunch f{};
        f.src = d_frames; f.src_item = item_bytes; f.src_pitch = pitch_bytes; f.W = pl.W; f.H = pl.H; f.nlev = nfused;
        for (int k = 1; k <= nfused; k++) {
            const Level& l = pl.lv[k];
            f.Wk[k - 1] = l.W; f.Hk[k - 1] = l.H; f.pitch[k - 1] = l.pitch;
            f.sx[k - 1] = l.sx; f.ax[k - 1] = l.ax; f.sy[k - 1] = l.sy; f.ay[k - 1] = l.ay;
            f.I[k - 1] = pl.If[k - 1]; f.i_item[k - 1] = l.plane(); f.taps[k - 1] = l.taps_h.data();
        }
        launch_pyr_fused(L, f, count);
    }
    for (int k = pl.K; k >= 0; k--) {
        Level& l = pl.lv[k];
        const int slot0 = f0 % pl.nslots;
        if (fast && k >= 1 && k <= nfused) {
            PolyArgs a = pl.pa;
            a.src = pl.If[k - 1]; a.src_item = l.plane() * sizeof(float); a.src_pitch = (size_t)l.pitch * sizeof(float);
            a.W = l.W; a.H = l.H; a.R = ring(pl, l, slot_step); a.slot0 = slot0;
            launch_polyexp2(L, 0, a, count);
            continue;
        }
        if (fast && k == 0) {
            PolyArgs a = pl.pa;
            a.src = d_frames; a.src_item = item_bytes; a.src_pitch = pitch_bytes;
            a.W = l.W; a.H = l.H; a.R = ring(pl, l, slot_step); a.slot0 = slot0;
            launch_polyexp2(L, pl.dtype == OFB_U8 ? 1 : 2, a, count);
            continue;
        }
        if (fast) {
            PyrArgs py{};
            py.src = d_frames; py.src_item = item_bytes; py.src_pitch = pitch_bytes;
            py.W = pl.W; py.H = pl.H; py.Wk = l.W; py.Hk = l.H; py.ksize = l.ksize; py.taps = l.taps;
            py.sx = l.sx; py.ax = l.ax; py.sy = l.sy; py.ay = l.ay;
            py.T = pl.T; py.t_item = pl.t_item; py.t_cap = pl.t_item; py.I = pl.I; py.i_item = pl.i_item; py.pitch = l.pitch;
            for (size_t q = 0; q < l.taps_h.size() && q < 80; q++) py.tapsv[q] = l.taps_h[q];
            launch_pyr2(L, pl.dtype, py, count);
            PolyArgs a = pl.pa;
            a.src = pl.I; a.src_item = pl.i_item * sizeof(float); a.src_pitch = (size_t)l.pitch * sizeof(float);
            a.W = l.W; a.H = l.H; a.R = ring(pl, l, slot_step); a.slot0 = slot0;
            launch_polyexp2(L, 0, a, count);
            continue;
        }
        for (int z = 0; z < count; z++) {
            const char* fr = (const char*)d_frames + (size_t)z * item_bytes;
            launch_pyr_h(L, fr, pl.dtype, pl.W, pl.H, pitch_bytes, l.taps, l.ksize, pl.T, l.W, l.pitch);
            launch_pyr_v(L, pl.T, pl.H, l.pitch, l.taps, l.ksize, pl.I, l.W, l.H, l.pitch);
            launch_polyexp(L, pl.I, l.W, l.H, l.pitch, pl.pc, pl.tmp3, slot_planes(l, (f0 + z * slot_step) % pl.nslots), ctx->generic);
        }
    }
}

// Per-pair part for `count` consecutive pairs; pair z uses the slots of frames t0+z and t0+z+1 and writes
// its scale-0 flow to d_flow + z * flow_item (float2 units).  `initial` = OPTFLOW_USE_INITIAL_FLOW (count == 1).
bool solve_pairs(ofb_context* ctx, Launch& L, int t0, int count, float2* d_flow, size_t flow_item, int step = 1,
                 bool want_minmax = false)
{
    Plan& pl = ctx->plan;
    const ofb_params& p = pl.p;
    const bool gaussian = (p.flags & OFB_OPTFLOW_FARNEBACK_GAUSSIAN) != 0;
    const bool initial = (p.flags & OFB_OPTFLOW_USE_INITIAL_FLOW) != 0;
    const bool fast = pl.fast_iter && !ctx->generic;
    const float c4 = (float)(1e-3 * (double)p.winsize * p.winsize * p.winsize * p.winsize);
    const float up_mul = (float)(1. / p.pyr_scale);
    const bool fold_minmax = want_minmax && fast;          // min / max of |flow| folded into the last launch of scale 0
    if (fold_minmax) launch_minmax_reset_batch(L, ctx->minmax, count);
    for (int k = pl.K; k >= 0; k--) {
        Level& l = pl.lv[k];
        float2* flow = k == 0 ? d_flow : l.flow;
        const size_t fitem = k == 0 ? flow_item : l.flow_item();
        const int slot0 = t0 % pl.nslots;
        if (k == pl.K && initial && k > 0)
            launch_area_flow(L, d_flow, pl.W, pl.H, flow, l.W, l.H, (float)l.scale);   // count == 1
        if (fast) {
            const int sub = (k == 0 && ctx->batch0 > 0) ? std::min(ctx->batch0, count) : count;
            for (int z0 = 0; z0 < count; z0 += sub) {
                const int nb = std::min(sub, count - z0);
                Um0Args u{};
                u.R = ring(pl, l, step); u.slot0 = (slot0 + z0 * step) % pl.nslots;
                u.M = pl.M[0] + (size_t)z0 * pl.m_item; u.m_item = pl.m_item; u.plane = l.plane(); u.pitch = l.pitch;
                u.W = l.W; u.H = l.H;
                int src = 0;
                if (k == pl.K) {
                    if (initial) { src = 1; u.flow = flow + (size_t)z0 * fitem; u.flow_item = fitem; }
                } else {
                    const Level& cl = pl.lv[k + 1];
                    src = 2;
                    u.flow = cl.flow + (size_t)z0 * cl.flow_item(); u.flow_item = cl.flow_item();
                    u.Wp = cl.W; u.Hp = cl.H;
                    u.ux = l.ux; u.uax = l.uax; u.uy = l.uy; u.uay = l.uay;
                    u.mul = up_mul;
                }
                launch_um0(L, src, u, nb);
                int cur = 0;
                for (int i = 0; i < p.iterations; i++) {
                    const bool last = i == p.iterations - 1;
                    IterArgs a{};
                    a.Min = pl.M[cur] + (size_t)z0 * pl.m_item; a.Mout = pl.M[cur ^ 1] + (size_t)z0 * pl.m_item;
                    a.m_item = pl.m_item; a.plane = l.plane(); a.pitch = l.pitch;
                    a.R = ring(pl, l, step); a.slot0 = (slot0 + z0 * step) % pl.nslots;
                    a.flow = flow + (size_t)z0 * fitem; a.flow_item = fitem;
                    a.W = l.W; a.H = l.H; a.c = gaussian ? 1e-3f : c4;
                    a.c64 = 1e-3 * (double)p.winsize * p.winsize * p.winsize * p.winsize;
                    a.gauss = gaussian ? 1 : 0;
                    if (gaussian) for (size_t q = 0; q < pl.gk.size() && q < 17; q++) a.gk[q] = pl.gk[q];
                    a.minmax = (fold_minmax && last && k == 0) ? ctx->minmax + 2 * z0 : nullptr;
                    a.reverse = (ctx->alt_order && (i % 2 == 0)) ? 1 : 0;     // um0 wrote M forwards; alternate from there
                    launch_iter(L, a, p.winsize, !last, nb);
                    cur ^= 1;
                }
            }
            continue;
        }
        // generic path: explicit flow buffers, one item at a time
        for (int z = 0; z < count; z++) {
            float2* fl = flow + (size_t)z * fitem;
            if (k == pl.K) {
                if (!initial) cudaMemsetAsync(fl, 0, sizeof(float2) * l.flow_item(), L.stream);
            } else {
                const Level& cl = pl.lv[k + 1];
                launch_upsample_flow(L, cl.flow + (size_t)z * cl.flow_item(), cl.W, cl.H, fl, l.W, l.H, up_mul);
            }
            RView R0 = slot_planes(l, (slot0 + z * step) % pl.nslots), R1 = slot_planes(l, (slot0 + z * step + 1) % pl.nslots);
            Planes5 M = m_planes(pl, l, 0, 0);
            launch_update_matrices(L, R0, R1, fl, l.W, l.H, M);
            for (int i = 0; i < p.iterations; i++) {
                if (gaussian) launch_blur_solve_gauss(L, M, l.W, l.H, p.winsize, pl.gtaps, (float*)pl.btmp, fl, true);
                else launch_blur_solve_box(L, M, l.W, l.H, p.winsize, pl.btmp, fl, true);
                if (i < p.iterations - 1) launch_update_matrices(L, R0, R1, fl, l.W, l.H, M);
            }
        }
    }
    return fold_minmax;
}

int stage_buf(ofb_context* ctx, int i, size_t bytes, float** out)
{
    if (ctx->stage_bytes[i] < bytes) {
        if (ctx->stage[i]) cudaFree(ctx->stage[i]);
        ctx->stage[i] = nullptr; ctx->stage_bytes[i] = 0;
        CU(cudaMalloc((void**)&ctx->stage[i], bytes + 256));
        ctx->stage_bytes[i] = bytes;
    }
    *out = ctx->stage[i];
    return 0;
}

int upload_frame(ofb_context* ctx, const void* src, size_t pitch, int W, int H, int dtype, void* dst, cudaStream_t s)
{
    size_t row = (size_t)W * (dtype == OFB_U8 ? 1 : 4);
    if (pitch == 0) pitch = row;
    if (pitch < row) return fail(ctx, OFB_ERR_BAD_ARG, "row pitch smaller than a row");
    CU(cudaMemcpy2DAsync(dst, row, src, pitch, row, H, cudaMemcpyHostToDevice, s));
    return 0;
}

void picture(ofb_context* ctx, Launch& L, const float2* d_flow, size_t flow_item, size_t n, uint8_t* d_bgr, size_t bgr_item, int count,
             bool minmax_done = false)
{
    launch_picture_batch(L, d_flow, flow_item, n, ctx->minmax, d_bgr, bgr_item, count, minmax_done,
                         ctx->use_hsv_table ? ctx->hsv_table : nullptr);
}

// Pairs per launch inside a shot: enough pixels per launch to fill 148 SMs at the coarse scales and to
// amortise launch latency, bounded by workspace size (~90 MB of M / R / flow per 1080p pair).
int shot_batch(const ofb_context* ctx, int W, int H, int n_pairs)
{
    int b = ctx->batch;
    if (b <= 0) {
        // 1080p: 48.  Measured on B200 (300-pair shot, default arithmetic, profiles/r2x_ab_batch.log): 24 / 37 / 48 / 74 pairs per
        // launch give 4078 / 4069 / 4171 / 4061 pairs/s.  k_iter64 walks a whole column per CTA, so a launch has batch x 24 long
        // CTAs for 296 resident ones: 48 is 3.9 waves at scale 0 and fills the GPU at the two coarser scales (24 left half of
        // it idle at 480x270); beyond that the per-launch working set leaves L2.
        double px = (double)W * H;
        b = (int)std::ceil(96.0e6 / px);
        b = (b + 3) / 4 * 4;
        b = std::max(4, std::min(b, 512));
        // A job shorter than four such chunks (a short shot, one rank's share of a sharded shot) is cut into about four, but not
        // below b / 4: upload, kernels and download of consecutive chunks overlap, one big chunk would serialise them.
        if (n_pairs < 4 * b) {
            const int q = ((n_pairs + 3) / 4 + 3) / 4 * 4;
            b = std::max(std::max(4, b / 4), std::min(q, b));
        }
    }
    return std::max(1, std::min(b, std::min(n_pairs, MAX_BATCH)));
}

int no_initial_flow(ofb_context* ctx, const ofb_params* p)
{
    if (p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW)
        return fail(ctx, OFB_ERR_UNSUPPORTED, "OPTFLOW_USE_INITIAL_FLOW is only meaningful through ofb_farneback_*");
    return 0;
}


// ---- frame preprocessing ---------------------------------------------------------------------
// Source indices and 11-bit weights of cv::resize(INTER_LINEAR) for 8-bit images, as cv2 builds them: the coordinate
// (d + 0.5) * scale - 0.5 is formed in double and rounded to f32 before the floor; columns move an out-of-range index
// inside and zero its fraction, rows only clip the two indices (oracle/preprocess_oracle.c coord_u8).
void resize_coord(int d, double scale, int slen, bool clamp_weights, int* s0, int* s1, short* w)
{
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= (float)s;
    if (clamp_weights) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= slen - 1) { s = slen - 1; f = 0.f; }
    }
    *s0 = std::min(std::max(s, 0), slen - 1);
    *s1 = std::min(std::max(s + 1, 0), slen - 1);
    w[0] = (short)std::lrint((1.f - f) * 2048.f);
    w[1] = (short)std::lrint(f * 2048.f);
}

int ensure_resize(ofb_context* ctx, int sW, int sH, int dW, int dH)
{
    if (ctx->rs_tab && ctx->rs_geom[0] == sW && ctx->rs_geom[1] == sH && ctx->rs_geom[2] == dW && ctx->rs_geom[3] == dH) return 0;
    CU(cudaStreamSynchronize(ctx->s_compute));
    if (ctx->rs_tab) { cudaFree(ctx->rs_tab); ctx->rs_tab = nullptr; }
    const size_t nint = 2 * (size_t)dW + 2 * (size_t)dH;
    std::vector<int> buf(nint + (nint + 1) / 2);            // ints, then the shorts (2 per destination coordinate)
    int *x0 = buf.data(), *x1 = x0 + dW, *y0 = x1 + dW, *y1 = y0 + dH;
    short* ax = reinterpret_cast<short*>(buf.data() + nint);
    short* ay = ax + 2 * (size_t)dW;
    const double sx = 1.0 / ((double)dW / sW), sy = 1.0 / ((double)dH / sH);
    for (int x = 0; x < dW; x++) resize_coord(x, sx, sW, true, &x0[x], &x1[x], ax + 2 * x);
    for (int y = 0; y < dH; y++) resize_coord(y, sy, sH, false, &y0[y], &y1[y], ay + 2 * y);
    CU(cudaMalloc((void**)&ctx->rs_tab, buf.size() * sizeof(int)));
    CU(cudaMemcpy(ctx->rs_tab, buf.data(), buf.size() * sizeof(int), cudaMemcpyHostToDevice));
    int* d = ctx->rs_tab;
    const short* dax = reinterpret_cast<const short*>(d + nint);
    ctx->rs = ResizeTab{d, d + dW, dax, d + 2 * dW, d + 2 * dW + dH, dax + 2 * (size_t)dW};
    ctx->rs_geom[0] = sW; ctx->rs_geom[1] = sH; ctx->rs_geom[2] = dW; ctx->rs_geom[3] = dH;
    return 0;
}

// `count` decoded BGR frames (sW x sH, tightly packed, src_item bytes apart) -> gray frames of dW x dH (gray_item apart):
// cv2.resize (only when the size changes) followed by cvtColor(BGR2GRAY), optical_flow.py:42-44.
void preprocess_frames(ofb_context* ctx, Launch& L, const uint8_t* d_src, size_t src_item, int sW, int sH, uint8_t* d_gray,
                       size_t gray_item, int dW, int dH, int count)
{
    if (sW == dW && sH == dH) launch_bgr2gray(L, d_src, src_item, (size_t)sW * 3, d_gray, gray_item, (size_t)dW, dW, dH, count);
    else launch_resize_u8(L, d_src, src_item, (size_t)sW * 3, 3, true, d_gray, gray_item, (size_t)dW, dW, dH, ctx->rs, count);
}

// Geometry check of the BGR entry points; dW / dH = 0 means "no resize".
int bgr_geometry(ofb_context* ctx, int sW, int sH, int* dW, int* dH)
{
    if (sW <= 0 || sH <= 0 || *dW < 0 || *dH < 0 || ((*dW == 0) != (*dH == 0))) return fail(ctx, OFB_ERR_BAD_ARG, "bad frame or target size");
    if (*dW == 0) { *dW = sW; *dH = sH; }
    if (*dW != sW || *dH != sH) return ensure_resize(ctx, sW, sH, *dW, *dH);
    return 0;
}

// ---- GPU JPEG encoder workspaces --------------------------------------------------------------------------------
void free_jpeg(ofb_context* ctx)
{
    for (void* p : ctx->jpeg_allocs) cudaFree(p);
    ctx->jpeg_allocs.clear();
    ctx->jpeg_guards.clear();
    ctx->jpeg_W = ctx->jpeg_H = ctx->jpeg_quality = ctx->jpeg_batch = 0;
}

int ensure_jpeg(ofb_context* ctx, int W, int H, int quality, int batch)
{
    if (W > 65535 || H > 65535) return fail(ctx, OFB_ERR_UNSUPPORTED, "JPEG pictures are limited to 65535 x 65535");
    quality = std::max(1, std::min(quality, 100));
    if (ctx->jpeg_W == W && ctx->jpeg_H == H && ctx->jpeg_quality == quality && ctx->jpeg_batch >= batch) return 0;
    CU(cudaStreamSynchronize(ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_d2h));
    free_jpeg(ctx);
    JpegWork& w = ctx->jw;
    w.geom = jpeg_geometry(W, H);
    const JpegGeom& g = w.geom;
    const size_t B = (size_t)batch;
    auto alloc = [&](void** out, size_t bytes) -> int { return guarded_alloc(ctx, ctx->jpeg_allocs, ctx->jpeg_guards, out, bytes + 256); };
    JpegTables host_tab;
    jpeg_build_tables(W, H, quality, host_tab);
    void* q;
    if (int rc = alloc(&q, sizeof(JpegTables))) return rc;
    CU(cudaMemcpy(q, &host_tab, sizeof(JpegTables), cudaMemcpyHostToDevice));
    w.tables = (const JpegTables*)q; w.header_len = host_tab.header_len;
    if (int rc = alloc((void**)&w.coef, B * g.nblk * 64 * sizeof(int16_t))) return rc;
    if (int rc = alloc((void**)&w.blk_mask, B * g.nblk * sizeof(unsigned long long))) return rc;
    if (int rc = alloc((void**)&w.blk_bits, B * g.nblk * sizeof(uint32_t))) return rc;
    if (int rc = alloc((void**)&w.cta_bits, B * ((size_t)g.nblk / 128 + 2) * sizeof(uint32_t))) return rc;
    if (int rc = alloc((void**)&w.bits32, B * g.bits_cap + JPEG_SEG)) return rc;
    if (int rc = alloc((void**)&w.total_bits, B * sizeof(uint32_t))) return rc;
    if (int rc = alloc((void**)&w.seg_ff, B * g.nseg_cap * sizeof(uint32_t))) return rc;
    if (int rc = alloc((void**)&w.out_off, B * sizeof(unsigned long long))) return rc;
    for (int s = 0; s < 2; s++) {
        if (int rc = alloc((void**)&ctx->jout[s], B * g.out_cap)) return rc;
        if (int rc = alloc((void**)&ctx->d_tot[s], 2 * sizeof(unsigned long long))) return rc;
    }
    ctx->jpeg_W = W; ctx->jpeg_H = H; ctx->jpeg_quality = quality; ctx->jpeg_batch = batch;
    return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {
#pragma GCC visibility push(default)

int ofb_abi_version(void) { return OFB_ABI_VERSION; }
const char* ofb_global_error(void) { return g_error.c_str(); }

int ofb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ofb_create(int device, ofb_context** out)
{
    ofb_context* ctx = nullptr;
    if (!out) return fail(nullptr, OFB_ERR_BAD_ARG, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, OFB_ERR_NO_DEVICE,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this engine has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, OFB_ERR_BAD_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, OFB_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100-class; kernels are built for sm_100a only");
    ofb_context* c = new ofb_context();
    ctx = c;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t rc = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (rc == cudaSuccess) rc = r; };
    ok(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&c->s_expand, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        ok(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&c->ev_frame_free[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&c->ev_out_ready[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&c->ev_out_free[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&c->ev_expanded[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&c->ev_solved[i], cudaEventDisableTiming));
    }
    ok(cudaEventCreate(&c->ev_t0));
    ok(cudaEventCreate(&c->ev_t1));
    ok(cudaMalloc((void**)&c->minmax, sizeof(unsigned) * 2 * MAX_BATCH + 256));
    ok(cudaMalloc((void**)&c->sumacc, sizeof(double) * MAX_BATCH + 256));
    ok(cudaMalloc((void**)&c->sumout, sizeof(float) * MAX_BATCH + 256));
    ok(cudaMalloc((void**)&c->hsv_table, sizeof(unsigned) * 65536));
    ok(cudaHostAlloc((void**)&c->h_tot, 4 * sizeof(unsigned long long), cudaHostAllocPortable));
    for (int i = 0; i < 2; i++) ok(cudaEventCreateWithFlags(&c->ev_tot[i], cudaEventDisableTiming));
    if (rc != cudaSuccess) {
        std::string m = std::string("context setup: ") + cudaGetErrorString(rc);
        delete c;
        return fail(nullptr, OFB_ERR_CUDA, m);
    }
    {
        Launch L = make_launch(c, c->s_compute);
        launch_build_hsv_table(L, c->hsv_table);
        if (cudaStreamSynchronize(c->s_compute) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            delete c;
            return fail(nullptr, OFB_ERR_CUDA, "context setup: building the HSV->BGR table failed");
        }
        c->prof.reset();
    }
    *out = c;
    return OFB_OK;
}

void ofb_destroy(ofb_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->prof.collect();
    free_plan(ctx->plan);
    free_jpeg(ctx);
    if (ctx->h_tot) cudaFreeHost(ctx->h_tot);
    for (int i = 0; i < 2; i++) cudaEventDestroy(ctx->ev_tot[i]);
    for (int i = 0; i < 6; i++) if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    if (ctx->rs_tab) cudaFree(ctx->rs_tab);
    cudaFree(ctx->minmax); cudaFree(ctx->sumacc); cudaFree(ctx->sumout); cudaFree(ctx->hsv_table);
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(ctx->ev_h2d[i]); cudaEventDestroy(ctx->ev_frame_free[i]);
        cudaEventDestroy(ctx->ev_out_ready[i]); cudaEventDestroy(ctx->ev_out_free[i]);
        cudaEventDestroy(ctx->ev_expanded[i]); cudaEventDestroy(ctx->ev_solved[i]);
    }
    cudaEventDestroy(ctx->ev_t0); cudaEventDestroy(ctx->ev_t1);
    cudaStreamDestroy(ctx->s_compute); cudaStreamDestroy(ctx->s_h2d); cudaStreamDestroy(ctx->s_d2h); cudaStreamDestroy(ctx->s_expand);
    delete ctx;
}

const char* ofb_last_error(const ofb_context* ctx) { return ctx ? ctx->err.c_str() : g_error.c_str(); }
int ofb_device_of(const ofb_context* ctx) { return ctx ? ctx->device : -1; }
int ofb_sm_count(const ofb_context* ctx) { return ctx ? ctx->sm_count : 0; }

int ofb_synchronize(ofb_context* ctx)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->s_h2d));
    CU(cudaStreamSynchronize(ctx->s_expand));
    CU(cudaStreamSynchronize(ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_d2h));
    return OFB_OK;
}

void* ofb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void ofb_host_free(void* p) { if (p) cudaFreeHost(p); }

void* ofb_device_alloc(ofb_context* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { ctx->err = "cudaMalloc failed"; cudaGetLastError(); return nullptr; }
    return p;
}
void ofb_device_free(ofb_context* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }

int ofb_memcpy_h2d(ofb_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_compute));
    return OFB_OK;
}
int ofb_memcpy_d2h(ofb_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_compute));
    return OFB_OK;
}

// ---- the drop-in call ---------------------------------------------------------------------------
int ofb_farneback_device(ofb_context* ctx, const void* d_prev, const void* d_next, int dtype, int W, int H,
                         size_t prev_pitch, size_t next_pitch, float* d_flow, const ofb_params* p)
{
    if (int rc = validate(ctx, W, H, dtype, p)) return rc;
    if (!d_prev || !d_next || !d_flow) return fail(ctx, OFB_ERR_BAD_ARG, "null frame or flow pointer");
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_plan(ctx, W, H, dtype, p, 1)) return rc;
    size_t row = (size_t)W * (dtype == OFB_U8 ? 1 : 4);
    Launch L = make_launch(ctx, ctx->s_compute);
    expand_frames(ctx, L, d_prev, 0, prev_pitch ? prev_pitch : row, 0, 1);
    expand_frames(ctx, L, d_next, 0, next_pitch ? next_pitch : row, 1, 1);
    solve_pairs(ctx, L, 0, 1, (float2*)d_flow, (size_t)W * H);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->s_compute));
    return OFB_OK;
}

int ofb_farneback_host(ofb_context* ctx, const void* prev, const void* next, int dtype, int W, int H,
                       size_t prev_pitch, size_t next_pitch, float* flow, const ofb_params* p)
{
    if (int rc = validate(ctx, W, H, dtype, p)) return rc;
    if (!prev || !next || !flow) return fail(ctx, OFB_ERR_BAD_ARG, "null frame or flow pointer");
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_plan(ctx, W, H, dtype, p, 1)) return rc;
    Plan& pl = ctx->plan;
    cudaStream_t s = ctx->s_compute;
    if (int rc = upload_frame(ctx, prev, prev_pitch, W, H, dtype, pl.f0, s)) return rc;
    if (int rc = upload_frame(ctx, next, next_pitch, W, H, dtype, pl.fstage[0], s)) return rc;
    size_t fbytes = sizeof(float2) * (size_t)W * H;
    if (p->flags & OFB_OPTFLOW_USE_INITIAL_FLOW) CU(cudaMemcpyAsync(pl.flow0[0], flow, fbytes, cudaMemcpyHostToDevice, s));
    size_t row = (size_t)W * (dtype == OFB_U8 ? 1 : 4);
    Launch L = make_launch(ctx, s);
    expand_frames(ctx, L, pl.f0, 0, row, 0, 1);
    expand_frames(ctx, L, pl.fstage[0], 0, row, 1, 1);
    solve_pairs(ctx, L, 0, 1, pl.flow0[0], (size_t)W * H);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(flow, pl.flow0[0], fbytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

// ---- companions -------------------------------------------------------------------------------------
int ofb_flow_to_bgr_device(ofb_context* ctx, const float* d_flow, int W, int H, uint8_t* d_bgr)
{
    if (!ctx || !d_flow || !d_bgr || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    Launch L = make_launch(ctx, ctx->s_compute);
    picture(ctx, L, (const float2*)d_flow, 0, (size_t)W * H, d_bgr, 0, 1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->s_compute));
    return OFB_OK;
}

int ofb_flow_to_bgr_host(ofb_context* ctx, const float* flow, int W, int H, uint8_t* bgr)
{
    if (!ctx || !flow || !bgr || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    size_t n = (size_t)W * H;
    float *df, *db;
    if (int rc = stage_buf(ctx, 0, n * 8, &df)) return rc;
    if (int rc = stage_buf(ctx, 1, n * 3 + 16, &db)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(df, flow, n * 8, cudaMemcpyHostToDevice, s));
    Launch L = make_launch(ctx, s);
    picture(ctx, L, (const float2*)df, 0, n, (uint8_t*)db, 0, 1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(bgr, db, n * 3, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_sum_magnitude_device(ofb_context* ctx, const float* d_flow, int W, int H, float* d_out)
{
    if (!ctx || !d_flow || !d_out || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    Launch L = make_launch(ctx, ctx->s_compute);
    launch_sum_magnitude_batch(L, (const float2*)d_flow, 0, (size_t)W * H, ctx->sumacc, d_out, 1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->s_compute));
    return OFB_OK;
}

int ofb_sum_magnitude_host(ofb_context* ctx, const float* flow, int W, int H, float* out)
{
    if (!ctx || !flow || !out || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    size_t n = (size_t)W * H;
    float* df;
    if (int rc = stage_buf(ctx, 0, n * 8, &df)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(df, flow, n * 8, cudaMemcpyHostToDevice, s));
    Launch L = make_launch(ctx, s);
    launch_sum_magnitude_batch(L, (const float2*)df, 0, n, ctx->sumacc, ctx->sumout, 1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->sumout, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_cart_to_polar_host(ofb_context* ctx, const float* flow, int W, int H, float* mag, float* ang)
{
    return ofb_cart_to_polar_host2(ctx, flow, W, H, mag, ang, 0);
}

int ofb_cart_to_polar_host2(ofb_context* ctx, const float* flow, int W, int H, float* mag, float* ang, int angle_in_degrees)
{
    if (!ctx || !flow || !mag || !ang || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    size_t n = (size_t)W * H;
    float *df, *dm, *da;
    if (int rc = stage_buf(ctx, 0, n * 8, &df)) return rc;
    if (int rc = stage_buf(ctx, 1, n * 4, &dm)) return rc;
    if (int rc = stage_buf(ctx, 2, n * 4, &da)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(df, flow, n * 8, cudaMemcpyHostToDevice, s));
    Launch L = make_launch(ctx, s);
    launch_cart_to_polar(L, (const float2*)df, n, dm, da, angle_in_degrees != 0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(mag, dm, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ang, da, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

// ---- fused pair / shot ------------------------------------------------------------------------------
int ofb_pair_host(ofb_context* ctx, const void* prev, const void* next, int dtype, int W, int H,
                  const ofb_params* p, uint8_t* bgr, float* magsum, float* flow)
{
    if (int rc = validate(ctx, W, H, dtype, p)) return rc;
    if (!prev || !next) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer");
    if (int rc = no_initial_flow(ctx, p)) return rc;
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_plan(ctx, W, H, dtype, p, 1)) return rc;
    Plan& pl = ctx->plan;
    cudaStream_t s = ctx->s_compute;
    if (int rc = upload_frame(ctx, prev, 0, W, H, dtype, pl.f0, s)) return rc;
    if (int rc = upload_frame(ctx, next, 0, W, H, dtype, pl.fstage[0], s)) return rc;
    size_t row = (size_t)W * (dtype == OFB_U8 ? 1 : 4), n = (size_t)W * H;
    Launch L = make_launch(ctx, s);
    expand_frames(ctx, L, pl.f0, 0, row, 0, 1);
    expand_frames(ctx, L, pl.fstage[0], 0, row, 1, 1);
    const bool mm = solve_pairs(ctx, L, 0, 1, pl.flow0[0], n, 1, bgr != nullptr);
    if (bgr) picture(ctx, L, pl.flow0[0], 0, n, pl.bgr[0], 0, 1, mm);
    if (magsum) launch_sum_magnitude_batch(L, pl.flow0[0], 0, n, ctx->sumacc, ctx->sumout, 1);
    CU(cudaGetLastError());
    if (bgr) CU(cudaMemcpyAsync(bgr, pl.bgr[0], n * 3, cudaMemcpyDeviceToHost, s));
    if (magsum) CU(cudaMemcpyAsync(magsum, ctx->sumout, 4, cudaMemcpyDeviceToHost, s));
    if (flow) CU(cudaMemcpyAsync(flow, pl.flow0[0], n * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_shot_device(ofb_context* ctx, const uint8_t* d_frames, int n_frames, int W, int H, const ofb_params* p,
                    uint8_t* d_bgr, float* d_magsum, float* d_flow, float* device_ms)
{
    if (int rc = validate(ctx, W, H, OFB_U8, p)) return rc;
    if (!d_frames || n_frames < 2) return fail(ctx, OFB_ERR_BAD_ARG, "need at least two frames");
    if (int rc = no_initial_flow(ctx, p)) return rc;
    CU(cudaSetDevice(ctx->device));
    const int B = shot_batch(ctx, W, H, n_frames - 1);
    if (int rc = ensure_plan(ctx, W, H, OFB_U8, p, B)) return rc;
    Plan& pl = ctx->plan;
    cudaStream_t s = ctx->s_compute;
    const size_t n = (size_t)W * H;
    Launch L = make_launch(ctx, s);
    cudaStream_t se = ctx->overlap_expand ? ctx->s_expand : s;
    Launch LE = make_launch(ctx, se);
    CU(cudaEventRecord(ctx->ev_t0, s));
    if (se != s) CU(cudaStreamWaitEvent(se, ctx->ev_t0, 0));
    // chunk c+1 is expanded on `se` while the pairs of chunk c are solved on `s` (ring of 2 * batch + 1 frame slots)
    auto expand_chunk = [&](int c) -> int {
        const int t0 = c * B, b = std::min(B, n_frames - 1 - t0), par = c & 1;
        if (se != s && c >= 2) CU(cudaStreamWaitEvent(se, ctx->ev_solved[par], 0));
        if (c == 0) expand_frames(ctx, LE, d_frames, n, (size_t)W, 0, 1);
        expand_frames(ctx, LE, d_frames + (size_t)(t0 + 1) * n, n, (size_t)W, t0 + 1, b);
        if (se != s) CU(cudaEventRecord(ctx->ev_expanded[par], se));
        return 0;
    };
    const int n_chunks = (n_frames - 1 + B - 1) / B;
    if (int rc = expand_chunk(0)) return rc;
    for (int c = 0; c < n_chunks; c++) {
        const int t0 = c * B, b = std::min(B, n_frames - 1 - t0), par = c & 1;
        if (se != s) {
            CU(cudaStreamWaitEvent(s, ctx->ev_expanded[par], 0));
            if (c + 1 < n_chunks) if (int rc = expand_chunk(c + 1)) return rc;
        }
        float2* fl = d_flow ? (float2*)d_flow + (size_t)t0 * n : pl.flow0[0];
        const bool mm = solve_pairs(ctx, L, t0, b, fl, n, 1, d_bgr != nullptr);
        if (d_bgr) picture(ctx, L, fl, n, n, d_bgr + (size_t)t0 * n * 3, n * 3, b, mm);
        if (d_magsum) launch_sum_magnitude_batch(L, fl, n, n, ctx->sumacc, d_magsum + t0, b);
        CU(cudaEventRecord(ctx->ev_solved[par], s));
        if (se == s && c + 1 < n_chunks) if (int rc = expand_chunk(c + 1)) return rc;
    }
    CU(cudaEventRecord(ctx->ev_t1, s));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    CU(cudaStreamSynchronize(se));
    if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev_t0, ctx->ev_t1));
    return OFB_OK;
}

// Shared body of the four host entry points.
//   shot  (next == nullptr): `first` holds n_pairs+1 consecutive frames; pair t = frames t, t+1; a frame is expanded once.
//   pairs (next != nullptr): pair t = (first[t], next[t]), independent frames (the window loop of optical_flow.py:83-99).
//   sW > 0: the frames are decoded BGR frames of sW x sH; gray frames of W x H are produced on the GPU (gray_out: shot only).
// Three streams: frames of chunk c+1 upload (s_h2d) while chunk c computes (s_compute) and the results of chunk c-1 download
// (s_d2h); staging buffers are double-buffered by chunk parity and guarded by events, no host synchronisation inside.
//   fptr != nullptr (shot only): frame i lives at fptr[i] (a decoder's own buffers; nothing is assembled on the host).
static int host_impl(ofb_context* ctx, const uint8_t* first, const uint8_t* next, int n_pairs, int W, int H, int sW, int sH,
                     const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, uint8_t* gray_out, float* device_ms,
                     const uint8_t* const* fptr = nullptr, const JpegOut* jpg = nullptr)
{
    if (fptr && !first) first = fptr[0];
    const bool want_jpeg = jpg != nullptr;
    if (int rc = validate(ctx, W, H, OFB_U8, p)) return rc;
    const bool pairs = next != nullptr;
    if (!first || n_pairs < 1) return fail(ctx, OFB_ERR_BAD_ARG, pairs ? "need at least one pair" : "need at least two frames");
    if (int rc = no_initial_flow(ctx, p)) return rc;
    CU(cudaSetDevice(ctx->device));
    const int B = shot_batch(ctx, W, H, n_pairs);
    if (int rc = ensure_plan(ctx, W, H, OFB_U8, p, B)) return rc;
    if (want_jpeg) if (int rc = ensure_jpeg(ctx, W, H, jpg->quality, B)) return rc;
    Plan& pl = ctx->plan;
    cudaStream_t sc = ctx->s_compute, su = ctx->s_h2d, sd = ctx->s_d2h;
    const size_t n = (size_t)W * H;
    const bool from_bgr = sW > 0;
    uint32_t* d_jsizes = nullptr;
    if (want_jpeg) { float* q; if (int rc = stage_buf(ctx, 5, sizeof(uint32_t) * (size_t)n_pairs, &q)) return rc; d_jsizes = (uint32_t*)q; }
    size_t jpeg_done = 0;                                              // bytes of finished streams already queued for download
    // JPEG streams of chunk c are packed in jout[c & 1]; their total size is only known on the device, so the download of
    // chunk c is issued one iteration later, after a 16-byte read-back of the total (the GPU already runs chunk c+1 by then).
    auto flush_jpeg = [&](int c) -> int {
        const int par = c & 1;
        CU(cudaEventSynchronize(ctx->ev_tot[par]));
        const unsigned long long total = ctx->h_tot[2 * par], flag = ctx->h_tot[2 * par + 1];
        if (flag) return fail(ctx, OFB_ERR_UNSUPPORTED, "a JPEG stream is larger than the raw picture (not a picture this encoder expects)");
        if (jpeg_done + total > jpg->cap) return fail(ctx, OFB_ERR_BAD_ARG, "jpeg output buffer too small");
        CU(cudaMemcpyAsync(jpg->out + jpeg_done, ctx->jout[par], (size_t)total, cudaMemcpyDeviceToHost, sd));
        CU(cudaEventRecord(ctx->ev_out_free[par], sd));
        jpeg_done += (size_t)total;
        return 0;
    };
    const size_t sn = from_bgr ? (size_t)sW * sH * 3 : n;                // bytes of one source frame
    float* d_sums = nullptr;
    if (magsum) if (int rc = stage_buf(ctx, 3, sizeof(float) * (size_t)n_pairs, &d_sums)) return rc;
    // BGR staging: [first frame of a shot] [parity 0: B + B frames] [parity 1: B + B frames]
    uint8_t *bs0 = nullptr, *bs[2] = {nullptr, nullptr};
    const size_t sna = (sn + 15) & ~(size_t)15;
    if (from_bgr) {
        float* q;
        if (int rc = stage_buf(ctx, 4, sna * (4 * (size_t)B + 1), &q)) return rc;
        bs0 = (uint8_t*)q; bs[0] = bs0 + sna; bs[1] = bs[0] + sna * 2 * B;
    }
    Launch L = make_launch(ctx, sc);
    DrainOnError drain{ctx};
    // Chunk schedule: full chunks of B pairs, but a long job starts and ends with short ones (B/4, B/2, ..., B/2, B/4):
    // the first upload and the last download are the only copies that nothing overlaps, so they are kept small.
    std::vector<int> cstart;                                              // first pair of every chunk, then n_pairs
    {
        std::vector<int> head, tail;
        if (n_pairs >= 4 * B && B >= 8) { head = {B / 4, B / 2}; tail = {B / 2, B / 4}; }
        int t = 0, tail_sum = 0;
        for (int v : tail) tail_sum += v;
        for (int v : head) { cstart.push_back(t); t += v; }
        while (n_pairs - tail_sum - t > 0) { cstart.push_back(t); t += std::min(B, n_pairs - tail_sum - t); }
        for (int v : tail) { cstart.push_back(t); t += v; }
        cstart.push_back(n_pairs);
    }
    const int n_chunks = (int)cstart.size() - 1;
    // where the source frames of a chunk land on the device: `lo` = frames t0+1.. (shot) or prev (pairs), `hi` = next (pairs)
    auto src_lo = [&](int par) { return from_bgr ? bs[par] : pl.fstage[par]; };
    auto src_hi = [&](int par) { return from_bgr ? bs[par] + sna * B : pl.fstage[par] + n * B; };

    CU(cudaEventRecord(ctx->ev_t0, su));
    if (!pairs) CU(cudaMemcpyAsync(from_bgr ? bs0 : pl.f0, first, sn, cudaMemcpyHostToDevice, su));
    auto upload = [&](int c) -> int {
        const int t0 = cstart[c], b = cstart[c + 1] - t0, par = c & 1;
        if (c >= 2) CU(cudaStreamWaitEvent(su, ctx->ev_frame_free[par], 0));
        if (pairs) {
            CU(cudaMemcpyAsync(src_lo(par), first + (size_t)t0 * sn, (size_t)b * sn, cudaMemcpyHostToDevice, su));
            CU(cudaMemcpyAsync(src_hi(par), next + (size_t)t0 * sn, (size_t)b * sn, cudaMemcpyHostToDevice, su));
        } else if (fptr) {
            for (int i = 0; i < b; i++)
                CU(cudaMemcpyAsync(src_lo(par) + (size_t)i * sn, fptr[t0 + 1 + i], sn, cudaMemcpyHostToDevice, su));
        } else {
            CU(cudaMemcpyAsync(src_lo(par), first + (size_t)(t0 + 1) * sn, (size_t)b * sn, cudaMemcpyHostToDevice, su));
        }
        CU(cudaEventRecord(ctx->ev_h2d[par], su));
        return 0;
    };
    if (int rc = upload(0)) return rc;
    for (int c = 0; c < n_chunks; c++) {
        const int t0 = cstart[c], b = cstart[c + 1] - t0, par = c & 1;
        if (c + 1 < n_chunks) if (int rc = upload(c + 1)) return rc;
        // Per-frame work of this chunk on `se`: with overlap (shot mode) that is s_expand, so it runs beside the per-pair kernels of
        // the PREVIOUS chunk on s_compute; its ring slots were last read by the chunk two back (nslots = 2 * batch + 1).
        cudaStream_t se = (ctx->overlap_expand && !pairs) ? ctx->s_expand : sc;
        Launch LE = make_launch(ctx, se);
        CU(cudaStreamWaitEvent(se, ctx->ev_h2d[par], 0));
        if (se != sc && c >= 2) CU(cudaStreamWaitEvent(se, ctx->ev_solved[par], 0));
        uint8_t* g_lo = pl.fstage[par];
        uint8_t* g_hi = pl.fstage[par] + n * B;
        if (from_bgr) {
            if (!pairs && c == 0) preprocess_frames(ctx, LE, bs0, sn, sW, sH, pl.f0, n, W, H, 1);
            preprocess_frames(ctx, LE, bs[par], sn, sW, sH, g_lo, n, W, H, b);
            if (pairs) preprocess_frames(ctx, LE, bs[par] + sna * B, sn, sW, sH, g_hi, n, W, H, b);
            if (gray_out && !pairs) {   // the gray frames the reference would have computed on the host (stream-ordered copy)
                if (c == 0) CU(cudaMemcpyAsync(gray_out, pl.f0, n, cudaMemcpyDeviceToHost, se));
                CU(cudaMemcpyAsync(gray_out + (size_t)(t0 + 1) * n, g_lo, (size_t)b * n, cudaMemcpyDeviceToHost, se));
            }
        }
        if (pairs) {
            expand_frames(ctx, LE, g_lo, n, (size_t)W, 0, b, 2);          // prev[z] -> slot 2z
            expand_frames(ctx, LE, g_hi, n, (size_t)W, 1, b, 2);          // next[z] -> slot 2z+1
        } else {
            if (c == 0) expand_frames(ctx, LE, pl.f0, n, (size_t)W, 0, 1);
            expand_frames(ctx, LE, g_lo, n, (size_t)W, t0 + 1, b);
        }
        CU(cudaEventRecord(ctx->ev_frame_free[par], se));
        if (se != sc) {
            CU(cudaEventRecord(ctx->ev_expanded[par], se));
            CU(cudaStreamWaitEvent(sc, ctx->ev_expanded[par], 0));
        }
        if (c >= 2) CU(cudaStreamWaitEvent(sc, ctx->ev_out_free[par], 0));
        const bool want_pic = bgr != nullptr || want_jpeg;
        const bool mm = pairs ? solve_pairs(ctx, L, 0, b, pl.flow0[par], n, 2, want_pic)
                              : solve_pairs(ctx, L, t0, b, pl.flow0[par], n, 1, want_pic);
        if (want_pic) picture(ctx, L, pl.flow0[par], n, n, pl.bgr[par], n * 3, b, mm);
        if (want_jpeg) launch_jpeg_encode(L, ctx->jw, pl.bgr[par], n * 3, b, ctx->jout[par], d_jsizes + t0, ctx->d_tot[par]);
        if (magsum) launch_sum_magnitude_batch(L, pl.flow0[par], n, n, ctx->sumacc, d_sums + t0, b);
        CU(cudaEventRecord(ctx->ev_out_ready[par], sc));
        CU(cudaEventRecord(ctx->ev_solved[par], sc));
        if (bgr || flow || want_jpeg) {
            if (!want_jpeg || bgr || flow) CU(cudaStreamWaitEvent(sd, ctx->ev_out_ready[par], 0));
            if (bgr) CU(cudaMemcpyAsync(bgr + (size_t)t0 * n * 3, pl.bgr[par], (size_t)b * n * 3, cudaMemcpyDeviceToHost, sd));
            if (flow) CU(cudaMemcpyAsync(flow + (size_t)t0 * n * 2, pl.flow0[par], (size_t)b * n * 8, cudaMemcpyDeviceToHost, sd));
            if (want_jpeg) {
                // first the streams of the PREVIOUS chunk (their size is known by now), so that their download does not queue
                // behind this chunk's kernels on s_d2h; then the 16-byte read-back of this chunk's size
                if (c >= 1) if (int rc = flush_jpeg(c - 1)) return rc;      // records ev_out_free of the previous chunk
                CU(cudaStreamWaitEvent(sd, ctx->ev_out_ready[par], 0));
                CU(cudaMemcpyAsync(ctx->h_tot + 2 * par, ctx->d_tot[par], 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sd));
                CU(cudaEventRecord(ctx->ev_tot[par], sd));
            } else {
                CU(cudaEventRecord(ctx->ev_out_free[par], sd));
            }
        } else {
            CU(cudaEventRecord(ctx->ev_out_free[par], sc));
        }
    }
    if (want_jpeg) {
        if (int rc = flush_jpeg(n_chunks - 1)) return rc;
        CU(cudaMemcpyAsync(jpg->sizes, d_jsizes, sizeof(uint32_t) * (size_t)n_pairs, cudaMemcpyDeviceToHost, sd));
    }
    CU(cudaStreamWaitEvent(sd, ctx->ev_out_ready[(n_chunks - 1) & 1], 0));
    if (magsum) CU(cudaMemcpyAsync(magsum, d_sums, sizeof(float) * (size_t)n_pairs, cudaMemcpyDeviceToHost, sd));
    CU(cudaEventRecord(ctx->ev_t1, sd));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sd));
    CU(cudaStreamSynchronize(sc));
    CU(cudaStreamSynchronize(ctx->s_expand));
    CU(cudaStreamSynchronize(su));
    drain.armed = false;
    if (device_ms) CU(cudaEventElapsedTime(device_ms, ctx->ev_t0, ctx->ev_t1));
    return OFB_OK;
}

int ofb_shot_host(ofb_context* ctx, const uint8_t* frames, int n_frames, int W, int H, const ofb_params* p,
                  uint8_t* bgr, float* magsum, float* flow, float* device_ms)
{
    return host_impl(ctx, frames, nullptr, n_frames - 1, W, H, 0, 0, p, bgr, magsum, flow, nullptr, device_ms);
}

int ofb_shot_host_v(ofb_context* ctx, const uint8_t* const* frames, int n_frames, int W, int H, const ofb_params* p,
                    uint8_t* bgr, float* magsum, float* flow, float* device_ms)
{
    if (!frames) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer table");
    for (int i = 0; i < n_frames; i++) if (!frames[i]) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer");
    return host_impl(ctx, nullptr, nullptr, n_frames - 1, W, H, 0, 0, p, bgr, magsum, flow, nullptr, device_ms, frames);
}

int ofb_shot_host_jpeg(ofb_context* ctx, const uint8_t* frames, int n_frames, int W, int H, const ofb_params* p, int quality,
                       uint8_t* jpeg, size_t jpeg_cap, uint32_t* jpeg_sizes, float* magsum, float* device_ms)
{
    if (!jpeg || !jpeg_sizes) return fail(ctx, OFB_ERR_BAD_ARG, "null jpeg output");
    const JpegOut jo{jpeg, jpeg_cap, jpeg_sizes, quality};
    return host_impl(ctx, frames, nullptr, n_frames - 1, W, H, 0, 0, p, nullptr, magsum, nullptr, nullptr, device_ms, nullptr, &jo);
}

int ofb_shot_host_v_jpeg(ofb_context* ctx, const uint8_t* const* frames, int n_frames, int W, int H, const ofb_params* p, int quality,
                         uint8_t* jpeg, size_t jpeg_cap, uint32_t* jpeg_sizes, float* magsum, float* device_ms)
{
    if (!frames) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer table");
    for (int i = 0; i < n_frames; i++) if (!frames[i]) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer");
    if (!jpeg || !jpeg_sizes) return fail(ctx, OFB_ERR_BAD_ARG, "null jpeg output");
    const JpegOut jo{jpeg, jpeg_cap, jpeg_sizes, quality};
    return host_impl(ctx, nullptr, nullptr, n_frames - 1, W, H, 0, 0, p, nullptr, magsum, nullptr, nullptr, device_ms, frames, &jo);
}

int ofb_shot_bgr_host_jpeg(ofb_context* ctx, const uint8_t* bgr_frames, int n_frames, int W, int H, int dW, int dH, const ofb_params* p,
                           int quality, uint8_t* jpeg, size_t jpeg_cap, uint32_t* jpeg_sizes, float* magsum, float* device_ms)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    if (!jpeg || !jpeg_sizes) return fail(ctx, OFB_ERR_BAD_ARG, "null jpeg output");
    CU(cudaSetDevice(ctx->device));
    if (int rc = bgr_geometry(ctx, W, H, &dW, &dH)) return rc;
    const JpegOut jo{jpeg, jpeg_cap, jpeg_sizes, quality};
    return host_impl(ctx, bgr_frames, nullptr, n_frames - 1, dW, dH, W, H, p, nullptr, magsum, nullptr, nullptr, device_ms, nullptr, &jo);
}

int ofb_jpeg_encode_host(ofb_context* ctx, const uint8_t* bgr, int n, int W, int H, int quality, uint8_t* jpeg, size_t jpeg_cap,
                         uint32_t* jpeg_sizes)
{
    if (!ctx || !bgr || !jpeg || !jpeg_sizes || n < 1 || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    const size_t px = (size_t)W * H * 3;
    const int B = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, std::min<size_t>(64, (size_t)(256.0e6 / (double)px) + 1)));
    if (int rc = ensure_jpeg(ctx, W, H, quality, B)) return rc;
    float *dsrc, *dsz;
    if (int rc = stage_buf(ctx, 0, px * B + 16, &dsrc)) return rc;
    if (int rc = stage_buf(ctx, 1, sizeof(uint32_t) * (size_t)B, &dsz)) return rc;
    cudaStream_t s = ctx->s_compute;
    Launch L = make_launch(ctx, s);
    size_t done = 0;
    for (int t0 = 0; t0 < n; t0 += B) {
        const int b = std::min(B, n - t0);
        CU(cudaMemcpyAsync(dsrc, bgr + (size_t)t0 * px, px * b, cudaMemcpyHostToDevice, s));
        launch_jpeg_encode(L, ctx->jw, (const uint8_t*)dsrc, px, b, ctx->jout[0], (uint32_t*)dsz, ctx->d_tot[0]);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ctx->h_tot, ctx->d_tot[0], 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(jpeg_sizes + t0, dsz, sizeof(uint32_t) * (size_t)b, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (ctx->h_tot[1]) return fail(ctx, OFB_ERR_UNSUPPORTED, "a JPEG stream is larger than the raw picture");
        if (done + ctx->h_tot[0] > jpeg_cap) return fail(ctx, OFB_ERR_BAD_ARG, "jpeg output buffer too small");
        CU(cudaMemcpy(jpeg + done, ctx->jout[0], (size_t)ctx->h_tot[0], cudaMemcpyDeviceToHost));
        done += (size_t)ctx->h_tot[0];
    }
    return OFB_OK;
}

int ofb_stage_jpeg_coefficients(ofb_context* ctx, const uint8_t* bgr, int W, int H, int quality, int16_t* coef)
{
    if (!ctx || !bgr || !coef || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_jpeg(ctx, W, H, quality, 1)) return rc;
    const size_t px = (size_t)W * H * 3;
    float *dsrc, *dsz;
    if (int rc = stage_buf(ctx, 0, px + 16, &dsrc)) return rc;
    if (int rc = stage_buf(ctx, 1, 64, &dsz)) return rc;
    cudaStream_t s = ctx->s_compute;
    Launch L = make_launch(ctx, s);
    CU(cudaMemcpyAsync(dsrc, bgr, px, cudaMemcpyHostToDevice, s));
    launch_jpeg_encode(L, ctx->jw, (const uint8_t*)dsrc, px, 1, ctx->jout[0], (uint32_t*)dsz, ctx->d_tot[0]);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(coef, ctx->jw.coef, sizeof(int16_t) * 64 * (size_t)ctx->jw.geom.nblk, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_pairs_host(ofb_context* ctx, const uint8_t* prev, const uint8_t* next, int n_pairs, int W, int H,
                   const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, float* device_ms)
{
    if (!next) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer");
    return host_impl(ctx, prev, next, n_pairs, W, H, 0, 0, p, bgr, magsum, flow, nullptr, device_ms);
}

int ofb_shot_bgr_host(ofb_context* ctx, const uint8_t* bgr_frames, int n_frames, int W, int H, int dW, int dH,
                      const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, uint8_t* gray, float* device_ms)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    if (int rc = bgr_geometry(ctx, W, H, &dW, &dH)) return rc;
    return host_impl(ctx, bgr_frames, nullptr, n_frames - 1, dW, dH, W, H, p, bgr, magsum, flow, gray, device_ms);
}

int ofb_pairs_bgr_host(ofb_context* ctx, const uint8_t* prev_bgr, const uint8_t* next_bgr, int n_pairs, int W, int H, int dW, int dH,
                       const ofb_params* p, uint8_t* bgr, float* magsum, float* flow, float* device_ms)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    if (!next_bgr) return fail(ctx, OFB_ERR_BAD_ARG, "null frame pointer");
    CU(cudaSetDevice(ctx->device));
    if (int rc = bgr_geometry(ctx, W, H, &dW, &dH)) return rc;
    return host_impl(ctx, prev_bgr, next_bgr, n_pairs, dW, dH, W, H, p, bgr, magsum, flow, nullptr, device_ms);
}

// ---- frame preprocessing on its own (parity tests) ---------------------------------------------
int ofb_bgr_to_gray_host(ofb_context* ctx, const uint8_t* bgr, int W, int H, uint8_t* gray)
{
    if (!ctx || !bgr || !gray || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)W * H;
    float *ds, *dd;
    if (int rc = stage_buf(ctx, 0, n * 3 + 16, &ds)) return rc;
    if (int rc = stage_buf(ctx, 1, n + 16, &dd)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(ds, bgr, n * 3, cudaMemcpyHostToDevice, s));
    Launch L = make_launch(ctx, s);
    launch_bgr2gray(L, (const uint8_t*)ds, 0, (size_t)W * 3, (uint8_t*)dd, 0, (size_t)W, W, H, 1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(gray, dd, n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_resize_u8_host(ofb_context* ctx, const uint8_t* src, int W, int H, int channels, int dW, int dH, int to_gray, uint8_t* dst)
{
    if (!ctx || !src || !dst || W <= 0 || H <= 0 || dW <= 0 || dH <= 0 || (channels != 1 && channels != 3) || (to_gray && channels != 3))
        return fail(ctx, OFB_ERR_BAD_ARG, "bad argument (channels must be 1 or 3; to_gray needs 3)");
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure_resize(ctx, W, H, dW, dH)) return rc;
    const size_t sn = (size_t)W * H * channels, dn = (size_t)dW * dH * (to_gray ? 1 : channels);
    float *ds, *dd;
    if (int rc = stage_buf(ctx, 0, sn + 16, &ds)) return rc;
    if (int rc = stage_buf(ctx, 1, dn + 16, &dd)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(ds, src, sn, cudaMemcpyHostToDevice, s));
    Launch L = make_launch(ctx, s);
    launch_resize_u8(L, (const uint8_t*)ds, 0, (size_t)W * channels, channels, to_gray != 0, (uint8_t*)dd, 0,
                     (size_t)dW * (to_gray ? 1 : channels), dW, dH, ctx->rs, 1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, dd, dn, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

// ---- per-stage entry points ------------------------------------------------------------------------
int ofb_scale_count(int W, int H, double pyr_scale, int levels) { return num_scales(W, H, pyr_scale, levels); }

int ofb_scale_geometry(int W, int H, double pyr_scale, int k, int* Wk, int* Hk, int* ksize, double* sigma)
{
    if (!Wk || !Hk || !ksize || !sigma || k < 0) return OFB_ERR_BAD_ARG;
    scale_geometry(W, H, pyr_scale, k, Wk, Hk, ksize, sigma, nullptr);
    return OFB_OK;
}

int ofb_stage_level_image(ofb_context* ctx, const void* frame, int dtype, int W, int H, double pyr_scale, int k, float* out)
{
    if (!ctx || !frame || !out || W <= 0 || H <= 0 || k < 0 || (dtype != OFB_U8 && dtype != OFB_F32))
        return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    int Wk, Hk, ksize; double sigma;
    scale_geometry(W, H, pyr_scale, k, &Wk, &Hk, &ksize, &sigma, nullptr);
    int pitch = round_up(Wk, 32);
    size_t esz = dtype == OFB_U8 ? 1 : 4;
    float *dfr, *dT, *dI, *dtab;
    if (int rc = stage_buf(ctx, 0, (size_t)W * H * esz, &dfr)) return rc;
    const size_t t_cap = std::max((size_t)H * pitch, (size_t)Hk * W);
    if (int rc = stage_buf(ctx, 1, sizeof(float) * t_cap, &dT)) return rc;
    if (int rc = stage_buf(ctx, 2, sizeof(float) * (size_t)Hk * pitch, &dI)) return rc;
    std::vector<float> taps;
    gaussian_taps(ksize, sigma, taps);
    std::vector<int> ix, iy; std::vector<float> wx, wy;
    linear_table(Wk, W, ix, wx);
    linear_table(Hk, H, iy, wy);
    // one staging block: taps | sx | ax | sy | ay
    size_t o_taps = 0, o_sx = o_taps + taps.size(), o_ax = o_sx + ix.size(), o_sy = o_ax + wx.size(), o_ay = o_sy + iy.size();
    if (int rc = stage_buf(ctx, 3, sizeof(float) * (o_ay + wy.size()), &dtab)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(dtab + o_taps, taps.data(), 4 * taps.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dtab + o_sx, ix.data(), 4 * ix.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dtab + o_ax, wx.data(), 4 * wx.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dtab + o_sy, iy.data(), 4 * iy.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dtab + o_ay, wy.data(), 4 * wy.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dfr, frame, (size_t)W * H * esz, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));      // host vectors go out of scope below
    Launch L = make_launch(ctx, s);
    if (ctx->generic) {
        launch_pyr_h(L, dfr, dtype, W, H, (size_t)W * esz, dtab + o_taps, ksize, dT, Wk, pitch);
        launch_pyr_v(L, dT, H, pitch, dtab + o_taps, ksize, dI, Wk, Hk, pitch);
    } else {
        PyrArgs py{};
        py.src = dfr; py.src_item = 0; py.src_pitch = (size_t)W * esz;
        py.W = W; py.H = H; py.Wk = Wk; py.Hk = Hk; py.ksize = ksize; py.taps = dtab + o_taps;
        py.sx = (const int*)(dtab + o_sx); py.ax = dtab + o_ax; py.sy = (const int*)(dtab + o_sy); py.ay = dtab + o_ay;
        py.T = dT; py.t_item = 0; py.t_cap = t_cap; py.I = dI; py.i_item = 0; py.pitch = pitch;
        for (size_t q = 0; q < taps.size() && q < 80; q++) py.tapsv[q] = taps[q];
        launch_pyr2(L, dtype, py, 1);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpy2DAsync(out, sizeof(float) * Wk, dI, sizeof(float) * pitch, sizeof(float) * Wk, Hk, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_stage_polyexp(ofb_context* ctx, const float* img, int W, int H, int poly_n, double poly_sigma, float* R)
{
    if (!ctx || !img || !R || W <= 0 || H <= 0 || poly_n < 1 || poly_n > 64) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    int pitch = round_up(W, 32);
    size_t plane = (size_t)H * pitch;
    float *dI, *dtmp, *dR, *dout, *dtab;
    if (int rc = stage_buf(ctx, 0, sizeof(float) * plane, &dI)) return rc;
    if (int rc = stage_buf(ctx, 1, sizeof(float) * 3 * plane, &dtmp)) return rc;
    if (int rc = stage_buf(ctx, 2, sizeof(float) * 5 * plane, &dR)) return rc;
    if (int rc = stage_buf(ctx, 3, sizeof(float) * 5 * (size_t)W * H, &dout)) return rc;
    std::vector<float> tab; double ig[4];
    poly_constants(poly_n, poly_sigma, tab, ig);
    if (int rc = stage_buf(ctx, 4, sizeof(float) * tab.size(), &dtab)) return rc;
    cudaStream_t s = ctx->s_compute;
    CU(cudaMemcpyAsync(dtab, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpy2DAsync(dI, sizeof(float) * pitch, img, sizeof(float) * W, sizeof(float) * W, H, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    int len = 2 * poly_n + 1;
    PolyConst pc{dtab, dtab + len, dtab + 2 * len, poly_n, ig[0], ig[1], ig[2], ig[3]};
    Launch L = make_launch(ctx, s);
    RView Rp{reinterpret_cast<float4*>(dR), dR + 4 * plane, pitch};
    if (!ctx->generic && polyexp2_supported(poly_n)) {
        PolyArgs a;
        fill_poly_args(a, poly_n, tab, ig);
        a.src = dI; a.src_item = 0; a.src_pitch = sizeof(float) * (size_t)pitch; a.W = W; a.H = H;
        a.R = SlotRing{dR, 5 * plane, plane, pitch, 1, 1}; a.slot0 = 0;
        launch_polyexp2(L, 0, a, 1);
    } else {
        launch_polyexp(L, dI, W, H, pitch, pc, dtmp, Rp, ctx->generic);
    }
    launch_r_interleave(L, Rp, W, H, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(R, dout, sizeof(float) * 5 * (size_t)W * H, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_stage_update_matrices(ofb_context* ctx, const float* R0, const float* R1, const float* flow, int W, int H, float* M)
{
    if (!ctx || !R0 || !R1 || !flow || !M || W <= 0 || H <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    int pitch = round_up(W, 32);
    size_t plane = (size_t)H * pitch, n = (size_t)W * H;
    float *din, *dR, *dM, *dfl;
    if (int rc = stage_buf(ctx, 0, sizeof(float) * 5 * n, &din)) return rc;
    if (int rc = stage_buf(ctx, 1, sizeof(float) * 10 * plane, &dR)) return rc;     // two slots
    if (int rc = stage_buf(ctx, 3, sizeof(float) * 5 * plane, &dM)) return rc;
    if (int rc = stage_buf(ctx, 4, sizeof(float) * 2 * n, &dfl)) return rc;
    cudaStream_t s = ctx->s_compute;
    Launch L = make_launch(ctx, s);
    RView p0{reinterpret_cast<float4*>(dR), dR + 4 * plane, pitch};
    RView p1{reinterpret_cast<float4*>(dR + 5 * plane), dR + 9 * plane, pitch};
    Planes5 pm{dM, plane, pitch};
    CU(cudaMemcpyAsync(din, R0, sizeof(float) * 5 * n, cudaMemcpyHostToDevice, s));
    launch_r_deinterleave(L, din, W, H, p0);
    CU(cudaMemcpyAsync(din, R1, sizeof(float) * 5 * n, cudaMemcpyHostToDevice, s));
    launch_r_deinterleave(L, din, W, H, p1);
    CU(cudaMemcpyAsync(dfl, flow, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, s));
    if (ctx->generic) {
        launch_update_matrices(L, p0, p1, (const float2*)dfl, W, H, pm);
    } else {
        Um0Args u{};
        u.flow = (const float2*)dfl; u.flow_item = 0;
        u.R = SlotRing{dR, 5 * plane, plane, pitch, 2, 1}; u.slot0 = 0;
        u.M = dM; u.m_item = 0; u.plane = plane; u.pitch = pitch; u.W = W; u.H = H;
        launch_um0(L, 1, u, 1);
    }
    launch_interleave5(L, pm, W, H, din);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(M, din, sizeof(float) * 5 * n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_stage_blur_solve(ofb_context* ctx, const float* M, int W, int H, int winsize, int gaussian, float* flow)
{
    if (!ctx || !M || !flow || W <= 0 || H <= 0 || winsize < 1 || winsize > 1024) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    int pitch = round_up(W, 32);
    size_t plane = (size_t)H * pitch, n = (size_t)W * H;
    float *din, *dM, *dtmp, *dfl, *dk;
    if (int rc = stage_buf(ctx, 0, sizeof(float) * 5 * n, &din)) return rc;
    if (int rc = stage_buf(ctx, 1, sizeof(float) * 5 * plane, &dM)) return rc;
    if (int rc = stage_buf(ctx, 2, sizeof(double) * 5 * plane, &dtmp)) return rc;
    if (int rc = stage_buf(ctx, 3, sizeof(float) * 2 * n, &dfl)) return rc;
    std::vector<float> gk;
    gauss_half_taps(winsize, gk);
    if (int rc = stage_buf(ctx, 4, sizeof(float) * gk.size(), &dk)) return rc;
    cudaStream_t s = ctx->s_compute;
    Launch L = make_launch(ctx, s);
    Planes5 pm{dM, plane, pitch};
    CU(cudaMemcpyAsync(dk, gk.data(), sizeof(float) * gk.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(din, M, sizeof(float) * 5 * n, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    launch_deinterleave5(L, din, W, H, pm);
    if (!ctx->generic && iter_supported(winsize)) {
        IterArgs a{};
        a.Min = dM; a.Mout = nullptr; a.m_item = 0; a.plane = plane; a.pitch = pitch;
        a.flow = (float2*)dfl; a.flow_item = 0; a.W = W; a.H = H;
        a.c = gaussian ? 1e-3f : (float)(1e-3 * (double)winsize * winsize * winsize * winsize);
        a.c64 = 1e-3 * (double)winsize * winsize * winsize * winsize;
        a.gauss = gaussian ? 1 : 0;
        if (gaussian) for (size_t q = 0; q < gk.size() && q < 17; q++) a.gk[q] = gk[q];
        launch_iter(L, a, winsize, false, 1);
    } else if (gaussian) {
        launch_blur_solve_gauss(L, pm, W, H, winsize, dk, dtmp, (float2*)dfl, true);
    } else {
        launch_blur_solve_box(L, pm, W, H, winsize, (double*)dtmp, (float2*)dfl, true);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(flow, dfl, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

int ofb_stage_upsample_flow(ofb_context* ctx, const float* prev_flow, int Wp, int Hp, int W, int H, double pyr_scale, float* flow)
{
    if (!ctx || !prev_flow || !flow || W <= 0 || H <= 0 || Wp <= 0 || Hp <= 0) return fail(ctx, OFB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    float *dp, *df;
    if (int rc = stage_buf(ctx, 0, sizeof(float) * 2 * (size_t)Wp * Hp, &dp)) return rc;
    if (int rc = stage_buf(ctx, 1, sizeof(float) * 2 * (size_t)W * H, &df)) return rc;
    cudaStream_t s = ctx->s_compute;
    Launch L = make_launch(ctx, s);
    CU(cudaMemcpyAsync(dp, prev_flow, sizeof(float) * 2 * (size_t)Wp * Hp, cudaMemcpyHostToDevice, s));
    launch_upsample_flow(L, (const float2*)dp, Wp, Hp, (float2*)df, W, H, (float)(1. / pyr_scale));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(flow, df, sizeof(float) * 2 * (size_t)W * H, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return OFB_OK;
}

// ---- options and measurement -----------------------------------------------------------------------
int ofb_debug_check_guards(ofb_context* ctx)
{
    if (!ctx) return OFB_ERR_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    std::vector<unsigned char> h(2 * GUARD);
    int bad = 0;
    auto scan = [&](const std::vector<std::pair<unsigned char*, size_t>>& gs) -> int {
        for (const auto& g : gs) {
            CU(cudaMemcpy(h.data(), g.first, GUARD, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(h.data() + GUARD, g.first + GUARD + g.second, GUARD, cudaMemcpyDeviceToHost));
            for (unsigned char b : h) bad += b != GUARD_BYTE;
        }
        return 0;
    };
    if (int rc = scan(ctx->plan.guards)) return rc;
    if (int rc = scan(ctx->jpeg_guards)) return rc;
    return bad;
}

int ofb_shot_chunk(const ofb_context* ctx, int W, int H, int n_pairs)
{
    if (!ctx || W <= 0 || H <= 0 || n_pairs <= 0) return 0;
    return shot_batch(ctx, W, H, n_pairs);
}

int ofb_set_option(ofb_context* ctx, const char* name, int value)
{
    if (!ctx || !name) return OFB_ERR_BAD_ARG;
    if (!strcmp(name, "generic_kernels")) { ctx->generic = value != 0; return OFB_OK; }
    if (!strcmp(name, "alt_order")) { ctx->alt_order = value != 0; return OFB_OK; }
    if (!strcmp(name, "iter_ilp")) { ctx->kopt.iter_ilp = value; return OFB_OK; }
    if (!strcmp(name, "iter_prefetch")) { ctx->kopt.iter_prefetch = value; return OFB_OK; }
    if (!strcmp(name, "polyexp_tma")) { ctx->kopt.polyexp_tma = value; return OFB_OK; }
    if (!strcmp(name, "overlap_expand")) { ctx->overlap_expand = value != 0; return OFB_OK; }
    if (!strcmp(name, "generic_polyexp")) { ctx->kopt.generic_polyexp = value; return OFB_OK; }
    if (!strcmp(name, "pyr_fused")) { ctx->kopt.pyr_fused = value; return OFB_OK; }
    if (!strcmp(name, "polyexp_fast")) { ctx->kopt.polyexp_fast = value; return OFB_OK; }
    if (!strcmp(name, "exact_window_sums")) { ctx->kopt.exact_window_sums = value; return OFB_OK; }
    if (!strcmp(name, "polyexp_exact")) { ctx->kopt.polyexp_exact = value; return OFB_OK; }
    if (!strcmp(name, "exact_arithmetic")) { ctx->kopt.exact_window_sums = ctx->kopt.polyexp_exact = (value != 0); return OFB_OK; }
    if (!strcmp(name, "fast_arithmetic")) { ctx->kopt.exact_window_sums = ctx->kopt.polyexp_exact = (value == 0); return OFB_OK; }
    if (!strcmp(name, "hsv_table")) { ctx->use_hsv_table = value != 0; return OFB_OK; }
    if (!strcmp(name, "batch")) { ctx->batch = std::max(0, std::min(value, MAX_BATCH)); return OFB_OK; }
    if (!strcmp(name, "batch_scale0")) { ctx->batch0 = std::max(0, std::min(value, MAX_BATCH)); return OFB_OK; }
    if (!strcmp(name, "profile")) {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        ctx->prof.collect();
        ctx->prof.timing = value != 0;
        return OFB_OK;
    }
    return fail(ctx, OFB_ERR_BAD_ARG, std::string("unknown option: ") + name);
}

int ofb_get_kernel_stats(ofb_context* ctx, ofb_kernel_stat* out, int max)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    ctx->prof.collect();
    int n = (int)ctx->prof.stats.size();
    for (int i = 0; i < n && i < max && out; i++) {
        memset(&out[i], 0, sizeof(out[i]));
        strncpy(out[i].name, ctx->prof.stats[i].name.c_str(), sizeof(out[i].name) - 1);
        out[i].launches = ctx->prof.stats[i].launches;
        out[i].total_ms = ctx->prof.stats[i].total_ms;
    }
    return n;
}

void ofb_reset_kernel_stats(ofb_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->prof.reset();
}

double ofb_algorithmic_bytes_pair(int W, int H, const ofb_params* p)
{
    if (!p || W <= 0 || H <= 0 || !(p->pyr_scale < 1) || !(p->pyr_scale > 0)) return 0;
    int K = num_scales(W, H, p->pyr_scale, p->levels);
    double N = (double)W * H, total = 0, prevNk = 0;
    for (int k = K; k >= 0; k--) {
        int Wk, Hk, ks; double sg;
        scale_geometry(W, H, p->pyr_scale, k, &Wk, &Hk, &ks, &sg, nullptr);
        double Nk = (double)Wk * Hk;
        total += 2 * N + (64.0 + 96.0 * p->iterations) * Nk + 8.0 * prevNk;
        prevNk = Nk;
    }
    return total;
}

double ofb_algorithmic_bytes_viz(int W, int H) { return 19.0 * (double)W * H; }

#pragma GCC visibility pop
}  // extern "C"
