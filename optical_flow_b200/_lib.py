"""ctypes binding of include/optflow_b200.h (the C-ABI of the sm_100a engine).

There is no CPU fallback anywhere in this package: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C optical_flow_b200/csrc`) loading
fails with ImportError, and creating an engine without a CUDA device fails with RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OFB_LIB_PATH selects another build of the SAME library (A/B runs of a kernel change on one GPU box).
LIB_PATH = os.environ.get("OFB_LIB_PATH") or os.path.join(_HERE, "lib", "libofb200.so")

OFB_OK = 0
OFB_ERR_CUDA = -1
OFB_ERR_NO_DEVICE = -2
OFB_ERR_BAD_ARG = -3
OFB_ERR_UNSUPPORTED = -4
OFB_ERR_ASSERT = -215

OFB_U8 = 0
OFB_F32 = 1


class Params(C.Structure):
    _fields_ = [("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iterations", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("flags", C.c_int)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint64), ("total_ms", C.c_double)]


_lib = None


def load():
    """Loads libofb200.so once and declares every prototype of include/optflow_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "optical_flow_b200: %s is missing. Build it with `make -C optical_flow_b200/csrc` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, fp, u8p, ip, dp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_double)
    pp = C.POINTER(Params)
    sz = C.c_size_t
    i = C.c_int

    def proto(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    proto("ofb_abi_version", i)
    proto("ofb_global_error", C.c_char_p)
    proto("ofb_device_count", i)
    proto("ofb_create", i, i, C.POINTER(vp))
    proto("ofb_destroy", None, vp)
    proto("ofb_last_error", C.c_char_p, vp)
    proto("ofb_device_of", i, vp)
    proto("ofb_sm_count", i, vp)
    proto("ofb_synchronize", i, vp)
    proto("ofb_host_alloc", vp, sz)
    proto("ofb_host_free", None, vp)
    proto("ofb_device_alloc", vp, vp, sz)
    proto("ofb_device_free", None, vp, vp)
    proto("ofb_memcpy_h2d", i, vp, vp, vp, sz)
    proto("ofb_memcpy_d2h", i, vp, vp, vp, sz)
    proto("ofb_farneback_host", i, vp, vp, vp, i, i, i, sz, sz, vp, pp)
    proto("ofb_farneback_device", i, vp, vp, vp, i, i, i, sz, sz, vp, pp)
    proto("ofb_cart_to_polar_host", i, vp, vp, i, i, vp, vp)
    proto("ofb_cart_to_polar_host2", i, vp, vp, i, i, vp, vp, i)
    proto("ofb_sum_magnitude_host", i, vp, vp, i, i, vp)
    proto("ofb_flow_to_bgr_host", i, vp, vp, i, i, vp)
    proto("ofb_flow_to_bgr_device", i, vp, vp, i, i, vp)
    proto("ofb_sum_magnitude_device", i, vp, vp, i, i, vp)
    proto("ofb_pair_host", i, vp, vp, vp, i, i, i, pp, vp, vp, vp)
    proto("ofb_shot_host", i, vp, vp, i, i, i, pp, vp, vp, vp, fp)
    proto("ofb_shot_host_v", i, vp, C.POINTER(vp), i, i, i, pp, vp, vp, vp, fp)
    proto("ofb_shot_host_jpeg", i, vp, vp, i, i, i, pp, i, vp, sz, vp, vp, fp)
    proto("ofb_shot_host_v_jpeg", i, vp, C.POINTER(vp), i, i, i, pp, i, vp, sz, vp, vp, fp)
    proto("ofb_shot_bgr_host_jpeg", i, vp, vp, i, i, i, i, i, pp, i, vp, sz, vp, vp, fp)
    proto("ofb_jpeg_encode_host", i, vp, vp, i, i, i, i, vp, sz, vp)
    proto("ofb_stage_jpeg_coefficients", i, vp, vp, i, i, i, vp)
    proto("ofb_shot_device", i, vp, vp, i, i, i, pp, vp, vp, vp, fp)
    proto("ofb_pairs_host", i, vp, vp, vp, i, i, i, pp, vp, vp, vp, fp)
    proto("ofb_bgr_to_gray_host", i, vp, vp, i, i, vp)
    proto("ofb_resize_u8_host", i, vp, vp, i, i, i, i, i, i, vp)
    proto("ofb_shot_bgr_host", i, vp, vp, i, i, i, i, i, pp, vp, vp, vp, vp, fp)
    proto("ofb_pairs_bgr_host", i, vp, vp, vp, i, i, i, i, i, pp, vp, vp, vp, fp)
    proto("ofb_scale_count", i, i, i, C.c_double, i)
    proto("ofb_scale_geometry", i, i, i, C.c_double, i, ip, ip, ip, dp)
    proto("ofb_stage_level_image", i, vp, vp, i, i, i, C.c_double, i, vp)
    proto("ofb_stage_polyexp", i, vp, vp, i, i, i, C.c_double, vp)
    proto("ofb_stage_update_matrices", i, vp, vp, vp, vp, i, i, vp)
    proto("ofb_stage_blur_solve", i, vp, vp, i, i, i, i, vp)
    proto("ofb_stage_upsample_flow", i, vp, vp, i, i, i, i, C.c_double, vp)
    proto("ofb_debug_check_guards", i, vp)
    proto("ofb_set_option", i, vp, C.c_char_p, i)
    proto("ofb_shot_chunk", i, vp, i, i, i)
    proto("ofb_get_kernel_stats", i, vp, C.POINTER(KernelStat), i)
    proto("ofb_reset_kernel_stats", None, vp)
    proto("ofb_algorithmic_bytes_pair", C.c_double, i, i, pp)
    proto("ofb_algorithmic_bytes_viz", C.c_double, i, i)
    _lib = L
    return L
