"""ctypes wrapper over oracle/farneback_oracle.c.  TEST INFRASTRUCTURE ONLY.

Each function names the reference behaviour it restates (SURVEY.md Appendix A/B;
call sites /root/reference/optical_flow.py:51-64, visualize_optical_flow.py:38-55).
Arrays use cv2's layouts: R / M are (H, W, 5) interleaved f32, flow is (H, W, 2) f32.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_FARNEBACK_GAUSSIAN = 256


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("farneback_oracle.c", "preprocess_oracle.c", "jpeg_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        u8p = C.POINTER(C.c_uint8)
        L.orc_num_scales.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_num_scales.restype = C.c_int
        L.orc_scale_geometry.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_gaussian_taps.argtypes = [C.c_int, C.c_double, fp]
        L.orc_level_image.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, fp]
        L.orc_gaussian_blur.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_double, fp]
        L.orc_resize_linear.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_resize_area.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int]
        L.orc_polyexp_constants.argtypes = [C.c_int, C.c_double, fp, fp, fp, C.POINTER(C.c_double)]
        L.orc_polyexp.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_double, fp]
        L.orc_update_matrices.argtypes = [fp, fp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int]
        L.orc_blur_solve_box.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.orc_blur_solve_gauss.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.orc_gauss_taps_half.argtypes = [C.c_int, fp]
        L.orc_upsample_flow.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_farneback.argtypes = [fp, fp, C.c_int, C.c_int, fp, C.c_double, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_farneback.restype = C.c_int
        L.orc_cart_to_polar.argtypes = [fp, fp, C.c_size_t, C.c_size_t, fp, fp]
        L.orc_sum_f32.argtypes = [fp, C.c_size_t]
        L.orc_sum_f32.restype = C.c_float
        L.orc_hsv2bgr_pixel.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_bgr2gray.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.orc_resize_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_viz.argtypes = [fp, C.c_int, C.c_int, C.c_int, u8p, u8p, u8p]
        L.orc_sum_magnitude.argtypes = [fp, C.c_int, C.c_int]
        L.orc_sum_magnitude.restype = C.c_float
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --- A.1 ---------------------------------------------------------------------
def scale_schedule(W, H, pyr_scale, levels):
    """[(k, W_k, H_k, ksize_k, sigma_k, scale_k)] for k = K..0."""
    L = lib()
    K = L.orc_num_scales(W, H, float(pyr_scale), int(levels))
    out = []
    for k in range(K, -1, -1):
        wk, hk, ks = C.c_int(), C.c_int(), C.c_int()
        sg, sc = C.c_double(), C.c_double()
        L.orc_scale_geometry(W, H, float(pyr_scale), k, C.byref(wk), C.byref(hk), C.byref(ks), C.byref(sg), C.byref(sc))
        out.append((k, wk.value, hk.value, ks.value, sg.value, sc.value))
    return out


# --- A.3 / A.4 ---------------------------------------------------------------
def gaussian_taps(ksize, sigma):
    out = np.empty(ksize, np.float32)
    lib().orc_gaussian_taps(ksize, float(sigma), _f(out))
    return out


def gaussian_blur(img, ksize, sigma):
    img = _f32(img)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_gaussian_blur(_f(img), W, H, ksize, float(sigma), _f(out))
    return out


def resize_linear(img, Wd, Hd, float_coords=0):
    img = _f32(img)
    cn = 1 if img.ndim == 2 else img.shape[2]
    Hs, Ws = img.shape[:2]
    out = np.empty((Hd, Wd) if img.ndim == 2 else (Hd, Wd, cn), np.float32)
    lib().orc_resize_linear(_f(img), Ws, Hs, _f(out), Wd, Hd, cn, float_coords)
    return out


def resize_area(img, Wd, Hd):
    img = _f32(img)
    cn = 1 if img.ndim == 2 else img.shape[2]
    Hs, Ws = img.shape[:2]
    out = np.empty((Hd, Wd) if img.ndim == 2 else (Hd, Wd, cn), np.float32)
    lib().orc_resize_area(_f(img), Ws, Hs, _f(out), Wd, Hd, cn)
    return out


def level_image(frame, Wk, Hk, ksize, sigma, float_coords=0):
    """I_k of SURVEY.md A.3: convertTo(f32) -> GaussianBlur -> resize(INTER_LINEAR)."""
    f = _f32(frame)
    H, W = f.shape
    out = np.empty((Hk, Wk), np.float32)
    lib().orc_level_image(_f(f), W, H, Wk, Hk, ksize, float(sigma), float_coords, _f(out))
    return out


# --- A.5 / A.6 ---------------------------------------------------------------
def polyexp_constants(n, sigma):
    g = np.empty(2 * n + 1, np.float32)
    xg = np.empty_like(g)
    xxg = np.empty_like(g)
    ig = (C.c_double * 4)()
    lib().orc_polyexp_constants(n, float(sigma), _f(g), _f(xg), _f(xxg), ig)
    return g, xg, xxg, tuple(ig)


def polyexp(img, n, sigma):
    img = _f32(img)
    H, W = img.shape
    out = np.empty((H, W, 5), np.float32)
    lib().orc_polyexp(_f(img), W, H, n, float(sigma), _f(out))
    return out


# --- A.8 - A.10 ---------------------------------------------------------------
def update_matrices(R0, R1, flow):
    R0, R1, flow = _f32(R0), _f32(R1), _f32(flow)
    H, W = flow.shape[:2]
    M = np.empty((H, W, 5), np.float32)
    lib().orc_update_matrices(_f(R0), _f(R1), _f(flow), W, H, _f(M), 0, H)
    return M


def blur_solve(M, winsize, gaussian=False):
    M = _f32(M)
    H, W = M.shape[:2]
    flow = np.empty((H, W, 2), np.float32)
    (lib().orc_blur_solve_gauss if gaussian else lib().orc_blur_solve_box)(_f(M), W, H, int(winsize), _f(flow))
    return flow


def gauss_taps_half(winsize):
    k = np.empty(winsize // 2 + 1, np.float32)
    lib().orc_gauss_taps_half(int(winsize), _f(k))
    return k


def upsample_flow(prev, W, H, pyr_scale, float_coords=1):
    prev = _f32(prev)
    Hp, Wp = prev.shape[:2]
    out = np.empty((H, W, 2), np.float32)
    lib().orc_upsample_flow(_f(prev), Wp, Hp, _f(out), W, H, float(pyr_scale), float_coords)
    return out


# --- the whole call -----------------------------------------------------------
def farneback(prev, next_, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags, float_coords=0):
    """Restatement of cv2.calcOpticalFlowFarneback (optical_flow.py:51-59, visualize_optical_flow.py:38-46)."""
    p, n = _f32(prev), _f32(next_)
    H, W = p.shape
    if flags & OPTFLOW_USE_INITIAL_FLOW:
        out = np.array(flow, dtype=np.float32, order="C", copy=True)
    else:
        out = np.zeros((H, W, 2), np.float32)
    rc = lib().orc_farneback(_f(p), _f(n), W, H, _f(out), float(pyr_scale), int(levels), int(winsize),
                             int(iterations), int(poly_n), float(poly_sigma), int(flags), float_coords)
    if rc != 0:
        raise ValueError("oracle farneback error %d" % rc)
    return out


# --- Appendix B ---------------------------------------------------------------
def cart_to_polar(flow):
    """cv2.cartToPolar(flow[...,0], flow[...,1]) (optical_flow.py:61)."""
    flow = _f32(flow)
    H, W = flow.shape[:2]
    mag = np.empty((H, W), np.float32)
    ang = np.empty((H, W), np.float32)
    base = flow.ctypes.data
    x = C.cast(base, C.POINTER(C.c_float))
    y = C.cast(base + 4, C.POINTER(C.c_float))
    lib().orc_cart_to_polar(x, y, W * H, 2, _f(mag), _f(ang))
    return mag, ang


def viz(flow, hsv_round=0, return_hv=False):
    """visualize_optical_flow.py:48-55 -> BGR u8 picture."""
    flow = _f32(flow)
    H, W = flow.shape[:2]
    bgr = np.empty((H, W, 3), np.uint8)
    hue = np.empty((H, W), np.uint8)
    val = np.empty((H, W), np.uint8)
    lib().orc_viz(_f(flow), W, H, hsv_round, _u8(bgr), _u8(hue), _u8(val))
    return (bgr, hue, val) if return_hv else bgr


def hsv_table(hsv_round=0):
    """(256,256,3) table BGR[H,V] at S=255."""
    t = np.empty((256, 256, 3), np.uint8)
    px = (C.c_uint8 * 3)()
    L = lib()
    for h in range(256):
        for v in range(256):
            L.orc_hsv2bgr_pixel(h, 255, v, hsv_round, px)
            t[h, v] = (px[0], px[1], px[2])
    return t


def sum_magnitude(flow):
    """optical_flow.py:61-64: np.sum(cartToPolar(...)[0])."""
    flow = _f32(flow)
    H, W = flow.shape[:2]
    return float(lib().orc_sum_magnitude(_f(flow), W, H))


# ---- frame preprocessing (oracle/preprocess_oracle.c) -------------------------------------
def bgr2gray(bgr):
    """cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) for uint8 (optical_flow.py:44, visualize_optical_flow.py:31,35)."""
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    H, W = bgr.shape[:2]
    out = np.empty((H, W), np.uint8)
    lib().orc_bgr2gray(_u8(bgr), W, H, _u8(out))
    return out


def resize_u8(src, dsize):
    """cv2.resize(src, (w, h)) with the default INTER_LINEAR for uint8, 1 or 3 channels (optical_flow.py:25-31)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    H, W = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    dW, dH = int(dsize[0]), int(dsize[1])
    out = np.empty((dH, dW) if src.ndim == 2 else (dH, dW, cn), np.uint8)
    lib().orc_resize_u8(_u8(src), W, H, cn, _u8(out), dW, dH)
    return out


# ---- the JPEG artefact (visualize_optical_flow.py:57-58; oracle/jpeg_oracle.c) -----------------------------------
def jpeg_encode(bgr, quality=95):
    """cv2.imencode('.jpeg', bgr)[1] for an (H, W, 3) uint8 picture with cv2's defaults (quality 95, 4:2:0, standard tables)."""
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    H, W = bgr.shape[:2]
    L = lib()
    L.jpeg_oracle_encode.restype = C.c_size_t
    L.jpeg_oracle_encode.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_size_t]
    cap = W * H * 4 + 4096
    out = np.empty(cap, np.uint8)
    n = L.jpeg_oracle_encode(_u8(bgr), W, H, int(quality), _u8(out), cap)
    assert n <= cap
    return out[:n].copy()


def jpeg_coefficients(bgr, quality=95):
    """Quantised DCT coefficients in scan order: (mcu_rows * mcu_cols, 6, 64) int16, zigzag order inside a block."""
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    H, W = bgr.shape[:2]
    n_mcu = ((W + 15) // 16) * ((H + 15) // 16)
    out = np.empty((n_mcu, 6, 64), np.int16)
    L = lib()
    L.jpeg_oracle_coefficients.restype = None
    L.jpeg_oracle_coefficients.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int16)]
    L.jpeg_oracle_coefficients(_u8(bgr), W, H, int(quality), out.ctypes.data_as(C.POINTER(C.c_int16)))
    return out


def jpeg_header(W, H, quality=95):
    L = lib()
    L.jpeg_oracle_header.restype = C.c_size_t
    L.jpeg_oracle_header.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    out = np.empty(1024, np.uint8)
    n = L.jpeg_oracle_header(W, H, int(quality), _u8(out))
    return out[:n].copy()
