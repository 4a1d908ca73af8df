"""Host-side mirror of the reference's operator interface for the Farneback + HSV hot path.

`Farneback` owns one C-ABI context (one GPU, its streams and workspaces).  The module-level functions in
`optical_flow_b200/__init__.py` (`calcOpticalFlowFarneback`, `cartToPolar`) keep cv2's names, argument
meaning and error behaviour so that the two call sites of the reference,
/root/reference/optical_flow.py:51-64 and /root/reference/visualize_optical_flow.py:38-55, can switch
by changing one import.

No CPU fallback: every method runs CUDA kernels through include/optflow_b200.h or raises.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib

OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_FARNEBACK_GAUSSIAN = 256

#: the seven literals hard-coded at optical_flow.py:53-59 and visualize_optical_flow.py:40-46
REFERENCE_PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)

_ASSERT_FRAMES = ("prev0.size() == next0.size() && prev0.channels() == next0.channels() && "
                  "prev0.channels() == 1 && pyrScale_ < 1")
_ASSERT_FLOW = "_flow0.size() == prev0.size() && _flow0.channels() == 2 && _flow0.depth() == CV_32F"


try:                                    # cv2 is only used for its exception TYPE: `except cv2.error:` at a call site that switched to
    import cv2 as _cv2                  # this package keeps working; nothing of cv2's arithmetic is reachable from here
    _ErrorBase = _cv2.error
except Exception:                       # pragma: no cover - cv2 missing
    _ErrorBase = Exception


class error(_ErrorBase):
    """cv2.error for the argument errors of calcOpticalFlowFarneback (code -215): a subclass of the installed cv2's own
    exception type when cv2 is importable, so existing `except cv2.error` handlers still catch it."""

    def __init__(self, msg, code=-215, func="calc"):
        super().__init__("OpenCV-compatible(%d) error: (%d:Assertion failed) %s in function '%s'" % (code, code, msg, func))
        self.code = code
        self.err = msg
        self.func = func
        self.msg = str(self)


def make_params(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0):
    return _lib.Params(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n), float(poly_sigma), int(flags))


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


def pinned_empty(shape, dtype):
    """NumPy array backed by page-locked host memory (cudaHostAlloc), for overlapped H2D / D2H."""
    L = _lib.load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = L.ofb_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("cudaHostAlloc(%d bytes) failed" % n)
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, L.ofb_host_free, p)
    return arr


def _single_channel(a, name):
    a = np.asarray(a)
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[..., 0]
    if a.ndim != 2:
        raise error(_ASSERT_FRAMES)
    return a


def _as_frame(a):
    """cv2 converts any depth to f32 first: u8 goes to the u8 kernels, the rest via f32."""
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a), _lib.OFB_U8
    if a.dtype.kind in "uifb":
        return np.ascontiguousarray(a, dtype=np.float32), _lib.OFB_F32
    raise TypeError("unsupported frame dtype %s" % a.dtype)


def validate_call(prev, next, flow, pyr_scale, flags):
    """Argument contract of cv2.calcOpticalFlowFarneback, GPU-free:
    returns (prev, next, dtype_code, out_flow) or raises `error` with cv2's -215 texts."""
    prev = _single_channel(prev, "prev")
    next = _single_channel(next, "next")
    if prev.shape != next.shape or not (float(pyr_scale) < 1):
        raise error(_ASSERT_FRAMES)
    H, W = prev.shape
    flags = int(flags)
    good = isinstance(flow, np.ndarray) and flow.dtype == np.float32 and flow.shape == (H, W, 2)
    inplace = good and flow.flags.c_contiguous and flow.flags.writeable
    if flags & OPTFLOW_USE_INITIAL_FLOW:
        if not good:
            raise error(_ASSERT_FLOW)
        out = flow if inplace else np.ascontiguousarray(flow).copy()
    else:
        out = flow if inplace else np.empty((H, W, 2), np.float32)   # a wrong `flow` is silently ignored (8b)
    p, dt = _as_frame(prev)
    n, dt2 = _as_frame(next)
    if dt != dt2:                                                  # cv2 accepts mixed depths
        p, n, dt = p.astype(np.float32), n.astype(np.float32), _lib.OFB_F32
    return p, n, dt, out


class Farneback:
    """One engine context on one GPU (C-ABI: ofb_create / ofb_destroy)."""

    def __init__(self, device=0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.ofb_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("optical_flow_b200: cannot create engine on device %d: %s (status %d)"
                               % (device, self._L.ofb_global_error().decode(), rc))
        self._h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, self._L.ofb_destroy, h)

    def close(self):
        self._fin()

    # -- helpers -----------------------------------------------------------------------------------
    def _check(self, rc, func="calc"):
        if rc == 0:
            return
        msg = self._L.ofb_last_error(self._h).decode()
        if rc == _lib.OFB_ERR_ASSERT:
            raise error(msg, func=func)
        if rc == _lib.OFB_ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        if rc == _lib.OFB_ERR_BAD_ARG:
            raise ValueError(msg)
        raise RuntimeError("optical_flow_b200 CUDA failure (%d): %s" % (rc, msg))

    @property
    def sm_count(self):
        return self._L.ofb_sm_count(self._h)

    def set_option(self, name, value):
        self._check(self._L.ofb_set_option(self._h, name.encode(), int(value)))

    def shot_chunk(self, W, H, n_pairs):
        """Pairs per launch the shot / pairs entry points use for this geometry (option "batch" or the engine's default)."""
        return self._L.ofb_shot_chunk(self._h, int(W), int(H), int(n_pairs))

    def chunk_starts(self, W, H, n_pairs):
        """First pair of every chunk of a shot of n_pairs pairs: B/4, B/2, B, ..., B, B/2, B/4 (engine.cu host_impl)."""
        B = self.shot_chunk(W, H, n_pairs)
        head, tail = ([B // 4, B // 2], [B // 2, B // 4]) if (n_pairs >= 4 * B and B >= 8) else ([], [])
        starts, t = [], 0
        for v in head:
            starts.append(t); t += v
        while n_pairs - sum(tail) - t > 0:
            starts.append(t); t += min(B, n_pairs - sum(tail) - t)
        for v in tail:
            starts.append(t); t += v
        return starts

    def check_guards(self):
        """Number of overwritten guard bytes around the current workspaces (0 = no out-of-bounds write next to a buffer)."""
        n = self._L.ofb_debug_check_guards(self._h)
        if n < 0:
            self._check(n)
        return n

    def synchronize(self):
        self._check(self._L.ofb_synchronize(self._h))

    def kernel_stats(self):
        arr = (_lib.KernelStat * 128)()
        n = self._L.ofb_get_kernel_stats(self._h, arr, 128)
        return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].total_ms)) for i in range(min(n, 128))}

    def reset_kernel_stats(self):
        self._L.ofb_reset_kernel_stats(self._h)

    # -- the drop-in call ----------------------------------------------------------------------------
    def calc(self, prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
        """cv2.calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n,
        poly_sigma, flags) -> flow   (optical_flow.py:51-59, visualize_optical_flow.py:38-46)."""
        p, n, dt, out = validate_call(prev, next, flow, pyr_scale, flags)
        H, W = p.shape
        prm = make_params(pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
        self._check(self._L.ofb_farneback_host(self._h, _ptr(p), _ptr(n), dt, W, H, 0, 0, _ptr(out), C.byref(prm)))
        return out

    # -- companions ----------------------------------------------------------------------------------
    def cart_to_polar(self, flow, angle_in_degrees=False):
        """cv2.cartToPolar(flow[...,0], flow[...,1], angleInDegrees=...) -> (magnitude, angle)."""
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        H, W = flow.shape[:2]
        mag = np.empty((H, W), np.float32)
        ang = np.empty((H, W), np.float32)
        self._check(self._L.ofb_cart_to_polar_host2(self._h, _ptr(flow), W, H, _ptr(mag), _ptr(ang), int(bool(angle_in_degrees))), "cartToPolar")
        return mag, ang

    def sum_magnitude(self, flow):
        """np.sum(cv2.cartToPolar(...)[0])  (optical_flow.py:61-64)."""
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        H, W = flow.shape[:2]
        out = np.zeros(1, np.float32)
        self._check(self._L.ofb_sum_magnitude_host(self._h, _ptr(flow), W, H, _ptr(out)))
        return np.float32(out[0])

    def flow_to_bgr(self, flow):
        """The HSV picture of visualize_optical_flow.py:48-55 as one fused GPU pass."""
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        H, W = flow.shape[:2]
        bgr = np.empty((H, W, 3), np.uint8)
        self._check(self._L.ofb_flow_to_bgr_host(self._h, _ptr(flow), W, H, _ptr(bgr)))
        return bgr

    # -- fused pair / shot ---------------------------------------------------------------------------
    def pair(self, prev, next, want_bgr=True, want_magsum=False, want_flow=False, **params):
        """One loop body of the reference: frames in, picture and/or summed magnitude out; the f32 flow
        stays on the GPU unless want_flow."""
        prev = _single_channel(prev, "prev")
        next = _single_channel(next, "next")
        if prev.shape != next.shape:
            raise error(_ASSERT_FRAMES)
        H, W = prev.shape
        p, dt = _as_frame(prev)
        n, dt2 = _as_frame(next)
        if dt != dt2:
            p, n, dt = p.astype(np.float32), n.astype(np.float32), _lib.OFB_F32
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        bgr = np.empty((H, W, 3), np.uint8) if want_bgr else None
        ms = np.zeros(1, np.float32) if want_magsum else None
        fl = np.empty((H, W, 2), np.float32) if want_flow else None
        self._check(self._L.ofb_pair_host(self._h, _ptr(p), _ptr(n), dt, W, H, C.byref(prm), _ptr(bgr), _ptr(ms), _ptr(fl)))
        return {"bgr": bgr, "magsum": None if ms is None else np.float32(ms[0]), "flow": fl}

    def shot(self, frames, want_bgr=True, want_magsum=False, want_flow=False, out_bgr=None, **params):
        """All consecutive pairs of a shot: frames (n, H, W) uint8 -> n-1 results.  Per-frame work is
        shared between the two pairs of each frame; uploads / downloads overlap compute."""
        frames = np.ascontiguousarray(frames)
        if frames.ndim != 3 or frames.dtype != np.uint8 or frames.shape[0] < 2:
            raise ValueError("frames must be (n>=2, H, W) uint8")
        n, H, W = frames.shape
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        bgr = None
        if want_bgr:
            bgr = out_bgr if out_bgr is not None else np.empty((n - 1, H, W, 3), np.uint8)
            assert bgr.shape == (n - 1, H, W, 3) and bgr.dtype == np.uint8 and bgr.flags.c_contiguous
        ms = np.zeros(n - 1, np.float32) if want_magsum else None
        fl = np.empty((n - 1, H, W, 2), np.float32) if want_flow else None
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_host(self._h, _ptr(frames), n, W, H, C.byref(prm), _ptr(bgr), _ptr(ms), _ptr(fl),
                                          C.byref(dev_ms)))
        return {"bgr": bgr, "magsum": ms, "flow": fl, "device_ms": float(dev_ms.value)}

    def shot_frames(self, frame_list, want_bgr=True, want_magsum=False, want_flow=False, out_bgr=None, **params):
        """`shot` for frames that live in separate buffers (what a decoder hands out): a sequence of (H, W) uint8
        C-contiguous arrays.  No host-side assembly -- each frame is uploaded from where it lies (ofb_shot_host_v)."""
        n = len(frame_list)
        if n < 2:
            raise ValueError("need at least two frames")
        H, W = frame_list[0].shape
        for f in frame_list:
            if f.shape != (H, W) or f.dtype != np.uint8 or not f.flags.c_contiguous:
                raise ValueError("every frame must be a C-contiguous (H, W) uint8 array of the same size")
        table = (C.c_void_p * n)(*[f.ctypes.data for f in frame_list])
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        bgr = None
        if want_bgr:
            bgr = out_bgr if out_bgr is not None else np.empty((n - 1, H, W, 3), np.uint8)
            assert bgr.shape[0] >= n - 1 and bgr.shape[1:] == (H, W, 3) and bgr.dtype == np.uint8 and bgr.flags.c_contiguous
        ms = np.zeros(n - 1, np.float32) if want_magsum else None
        fl = np.empty((n - 1, H, W, 2), np.float32) if want_flow else None
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_host_v(self._h, table, n, W, H, C.byref(prm), _ptr(bgr), _ptr(ms), _ptr(fl), C.byref(dev_ms)))
        return {"bgr": bgr, "magsum": ms, "flow": fl, "device_ms": float(dev_ms.value)}

    def shot_jpeg(self, frames, quality=95, out=None, want_magsum=False, **params):
        """`shot` whose pictures come back as the JPEG files the reference writes (visualize_optical_flow.py:57-58): the bytes of
        cv2.imencode('.jpeg', picture) for every pair, encoded on the GPU.  Returns {"jpeg": uint8 buffer, "sizes", "offsets"};
        stream i is jpeg[offsets[i]:offsets[i] + sizes[i]].  `out` may be a pre-allocated (pinned) uint8 buffer."""
        frames = np.ascontiguousarray(frames)
        if frames.ndim != 3 or frames.dtype != np.uint8 or frames.shape[0] < 2:
            raise ValueError("frames must be (n>=2, H, W) uint8")
        n, H, W = frames.shape
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        if out is None:
            out = np.empty((n - 1) * (W * H + 4096), np.uint8)         # 1 byte per pixel: ~10x a flow picture at quality 95
        assert out.dtype == np.uint8 and out.flags.c_contiguous
        sizes = np.zeros(n - 1, np.uint32)
        ms = np.zeros(n - 1, np.float32) if want_magsum else None
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_host_jpeg(self._h, _ptr(frames), n, W, H, C.byref(prm), int(quality), _ptr(out), out.size,
                                               _ptr(sizes), _ptr(ms), C.byref(dev_ms)))
        offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.int64)])
        return {"jpeg": out, "sizes": sizes, "offsets": offsets, "magsum": ms, "device_ms": float(dev_ms.value)}

    def shot_bgr_jpeg(self, frames_bgr, dsize=None, quality=95, **params):
        """`shot_jpeg` fed with the decoded BGR frames (n, H, W, 3), as shot_bgr: what visualize_optical_flow.py's loop does
        from `vid.read()` to `cv2.imwrite(flow_<ms>.jpeg)` (:23-58), with only the frames going up and the files coming back."""
        f = self._bgr_stack(frames_bgr, "frames_bgr")
        if f.shape[0] < 2:
            raise ValueError("need at least two frames")
        n, H, W = f.shape[:3]
        dW, dH = (W, H) if dsize is None else (int(dsize[0]), int(dsize[1]))
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        out = np.empty((n - 1) * (dW * dH + 4096), np.uint8)
        sizes = np.zeros(n - 1, np.uint32)
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_bgr_host_jpeg(self._h, _ptr(f), n, W, H, 0 if dsize is None else dW, 0 if dsize is None else dH,
                                                   C.byref(prm), int(quality), _ptr(out), out.size, _ptr(sizes), None, C.byref(dev_ms)))
        offs = np.concatenate([[0], np.cumsum(sizes, dtype=np.int64)])
        return {"files": [out[offs[i]:offs[i + 1]] for i in range(n - 1)], "sizes": sizes, "device_ms": float(dev_ms.value)}

    def jpeg_encode(self, pictures, quality=95):
        """cv2.imencode('.jpeg', picture)[1] for every (H, W, 3) uint8 BGR picture of `pictures` (n, H, W, 3), on the GPU.
        Returns a list of uint8 arrays."""
        pics = np.ascontiguousarray(pictures)
        if pics.ndim == 3:
            pics = pics[None]
        if pics.ndim != 4 or pics.shape[3] != 3 or pics.dtype != np.uint8:
            raise ValueError("pictures must be (n, H, W, 3) uint8")
        n, H, W = pics.shape[:3]
        out = np.empty(n * (W * H * 3 + 4096), np.uint8)
        sizes = np.zeros(n, np.uint32)
        self._check(self._L.ofb_jpeg_encode_host(self._h, _ptr(pics), n, W, H, int(quality), _ptr(out), out.size, _ptr(sizes)), "imencode")
        offs = np.concatenate([[0], np.cumsum(sizes, dtype=np.int64)])
        return [out[offs[i]:offs[i + 1]].copy() for i in range(n)]

    def stage_jpeg_coefficients(self, picture, quality=95):
        pic = np.ascontiguousarray(picture)
        H, W = pic.shape[:2]
        out = np.empty((((W + 15) // 16) * ((H + 15) // 16), 6, 64), np.int16)
        self._check(self._L.ofb_stage_jpeg_coefficients(self._h, _ptr(pic), W, H, int(quality), _ptr(out)))
        return out

    def shot_frames_jpeg(self, frame_list, quality=95, out=None, **params):
        """`shot_jpeg` for frames in separate buffers (ofb_shot_host_v_jpeg)."""
        n = len(frame_list)
        if n < 2:
            raise ValueError("need at least two frames")
        H, W = frame_list[0].shape
        for f in frame_list:
            if f.shape != (H, W) or f.dtype != np.uint8 or not f.flags.c_contiguous:
                raise ValueError("every frame must be a C-contiguous (H, W) uint8 array of the same size")
        table = (C.c_void_p * n)(*[f.ctypes.data for f in frame_list])
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        if out is None:
            out = np.empty((n - 1) * (W * H + 4096), np.uint8)
        sizes = np.zeros(n - 1, np.uint32)
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_host_v_jpeg(self._h, table, n, W, H, C.byref(prm), int(quality), _ptr(out), out.size, _ptr(sizes),
                                                 None, C.byref(dev_ms)))
        offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.int64)])
        return {"jpeg": out, "sizes": sizes, "offsets": offsets, "device_ms": float(dev_ms.value)}

    def pairs(self, prev_frames, next_frames, want_bgr=False, want_magsum=True, want_flow=False, **params):
        """n independent pairs (prev_frames[i], next_frames[i]) in one batched submission: the window loop of
        optical_flow.py:83-99 (its pairs need not share frames)."""
        a = np.ascontiguousarray(prev_frames)
        b = np.ascontiguousarray(next_frames)
        if a.shape != b.shape or a.ndim != 3 or a.dtype != np.uint8 or b.dtype != np.uint8 or a.shape[0] < 1:
            raise ValueError("prev_frames / next_frames must both be (n>=1, H, W) uint8")
        n, H, W = a.shape
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        bgr = np.empty((n, H, W, 3), np.uint8) if want_bgr else None
        ms = np.zeros(n, np.float32) if want_magsum else None
        fl = np.empty((n, H, W, 2), np.float32) if want_flow else None
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_pairs_host(self._h, _ptr(a), _ptr(b), n, W, H, C.byref(prm), _ptr(bgr), _ptr(ms), _ptr(fl),
                                           C.byref(dev_ms)))
        return {"bgr": bgr, "magsum": ms, "flow": fl, "device_ms": float(dev_ms.value)}

    # -- frame preprocessing on the GPU -------------------------------------------
    def bgr_to_gray(self, bgr):
        """cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) for (H, W, 3) uint8 (optical_flow.py:44, visualize_optical_flow.py:31,35)."""
        bgr = np.ascontiguousarray(bgr)
        if bgr.ndim != 3 or bgr.shape[2] != 3 or bgr.dtype != np.uint8:
            raise ValueError("bgr must be (H, W, 3) uint8")
        H, W = bgr.shape[:2]
        gray = np.empty((H, W), np.uint8)
        self._check(self._L.ofb_bgr_to_gray_host(self._h, _ptr(bgr), W, H, _ptr(gray)), "cvtColor")
        return gray

    def resize(self, src, dsize, to_gray=False):
        """cv2.resize(src, (w, h)) with the default INTER_LINEAR, uint8, 1 or 3 channels (optical_flow.py:25-31);
        to_gray=True additionally applies cvtColor(BGR2GRAY) to the resized frame (read_frame, optical_flow.py:34-46)."""
        src = np.ascontiguousarray(src)
        if src.dtype != np.uint8 or src.ndim not in (2, 3) or (src.ndim == 3 and src.shape[2] not in (1, 3)):
            raise ValueError("src must be uint8 (H, W) or (H, W, 3)")
        H, W = src.shape[:2]
        cn = 1 if src.ndim == 2 else src.shape[2]
        dW, dH = int(dsize[0]), int(dsize[1])
        out = np.empty((dH, dW) if (src.ndim == 2 or to_gray) else (dH, dW, cn), np.uint8)
        self._check(self._L.ofb_resize_u8_host(self._h, _ptr(src), W, H, cn, dW, dH, int(bool(to_gray)), _ptr(out)), "resize")
        return out

    @staticmethod
    def _bgr_stack(frames, name):
        frames = np.ascontiguousarray(frames)
        if frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8:
            raise ValueError("%s must be (n, H, W, 3) uint8" % name)
        return frames

    def shot_bgr(self, frames_bgr, dsize=None, want_bgr=True, want_magsum=False, want_flow=False, want_gray=False, **params):
        """`shot` fed with the decoded BGR frames (n, H, W, 3): the optional cv2.resize to dsize = (w, h) and the gray
        conversion run on the GPU, so each frame crosses PCIe once and is never touched by the host again."""
        f = self._bgr_stack(frames_bgr, "frames_bgr")
        if f.shape[0] < 2:
            raise ValueError("need at least two frames")
        n, H, W = f.shape[:3]
        dW, dH = (W, H) if dsize is None else (int(dsize[0]), int(dsize[1]))
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        bgr = np.empty((n - 1, dH, dW, 3), np.uint8) if want_bgr else None
        ms = np.zeros(n - 1, np.float32) if want_magsum else None
        fl = np.empty((n - 1, dH, dW, 2), np.float32) if want_flow else None
        gr = np.empty((n, dH, dW), np.uint8) if want_gray else None
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_bgr_host(self._h, _ptr(f), n, W, H, 0 if dsize is None else dW, 0 if dsize is None else dH,
                                              C.byref(prm), _ptr(bgr), _ptr(ms), _ptr(fl), _ptr(gr), C.byref(dev_ms)))
        return {"bgr": bgr, "magsum": ms, "flow": fl, "gray": gr, "device_ms": float(dev_ms.value)}

    def pairs_bgr(self, prev_bgr, next_bgr, dsize=None, want_bgr=False, want_magsum=True, want_flow=False, **params):
        """`pairs` fed with decoded BGR frames; resize / gray conversion on the GPU (optical_flow.py:83-99 with read_frame)."""
        a = self._bgr_stack(prev_bgr, "prev_bgr")
        b = self._bgr_stack(next_bgr, "next_bgr")
        if a.shape != b.shape or a.shape[0] < 1:
            raise ValueError("prev_bgr / next_bgr must have the same shape (n>=1, H, W, 3)")
        n, H, W = a.shape[:3]
        dW, dH = (W, H) if dsize is None else (int(dsize[0]), int(dsize[1]))
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        bgr = np.empty((n, dH, dW, 3), np.uint8) if want_bgr else None
        ms = np.zeros(n, np.float32) if want_magsum else None
        fl = np.empty((n, dH, dW, 2), np.float32) if want_flow else None
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_pairs_bgr_host(self._h, _ptr(a), _ptr(b), n, W, H, 0 if dsize is None else dW,
                                               0 if dsize is None else dH, C.byref(prm), _ptr(bgr), _ptr(ms), _ptr(fl), C.byref(dev_ms)))
        return {"bgr": bgr, "magsum": ms, "flow": fl, "device_ms": float(dev_ms.value)}

    # -- device-resident shot (bench: inputs already in HBM) -------------------------------------------
    def device_alloc(self, nbytes):
        p = self._L.ofb_device_alloc(self._h, int(nbytes))
        if not p:
            raise MemoryError("cudaMalloc(%d) failed" % nbytes)
        return p

    def device_free(self, p):
        self._L.ofb_device_free(self._h, C.c_void_p(p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self._L.ofb_memcpy_h2d(self._h, C.c_void_p(dptr), _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        assert arr.flags.c_contiguous
        self._check(self._L.ofb_memcpy_d2h(self._h, _ptr(arr), C.c_void_p(dptr), arr.nbytes))

    def shot_device(self, d_frames, n_frames, W, H, d_bgr=None, d_magsum=None, d_flow=None, **params):
        prm = make_params(**{**REFERENCE_PARAMS, **params})
        dev_ms = C.c_float(0)
        self._check(self._L.ofb_shot_device(self._h, C.c_void_p(d_frames), int(n_frames), W, H, C.byref(prm),
                                            C.c_void_p(d_bgr), C.c_void_p(d_magsum), C.c_void_p(d_flow), C.byref(dev_ms)))
        return float(dev_ms.value)

    # -- per-stage entry points (parity tests) ---------------------------------------------------------
    def stage_level_image(self, frame, pyr_scale, k):
        f, dt = _as_frame(_single_channel(frame, "frame"))
        H, W = f.shape
        wk, hk, ks = C.c_int(), C.c_int(), C.c_int()
        sg = C.c_double()
        self._L.ofb_scale_geometry(W, H, float(pyr_scale), int(k), C.byref(wk), C.byref(hk), C.byref(ks), C.byref(sg))
        out = np.empty((hk.value, wk.value), np.float32)
        self._check(self._L.ofb_stage_level_image(self._h, _ptr(f), dt, W, H, float(pyr_scale), int(k), _ptr(out)))
        return out

    def stage_polyexp(self, img, poly_n, poly_sigma):
        img = np.ascontiguousarray(img, dtype=np.float32)
        H, W = img.shape
        R = np.empty((H, W, 5), np.float32)
        self._check(self._L.ofb_stage_polyexp(self._h, _ptr(img), W, H, int(poly_n), float(poly_sigma), _ptr(R)))
        return R

    def stage_update_matrices(self, R0, R1, flow):
        R0 = np.ascontiguousarray(R0, dtype=np.float32)
        R1 = np.ascontiguousarray(R1, dtype=np.float32)
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        H, W = flow.shape[:2]
        M = np.empty((H, W, 5), np.float32)
        self._check(self._L.ofb_stage_update_matrices(self._h, _ptr(R0), _ptr(R1), _ptr(flow), W, H, _ptr(M)))
        return M

    def stage_blur_solve(self, M, winsize, gaussian=False):
        M = np.ascontiguousarray(M, dtype=np.float32)
        H, W = M.shape[:2]
        flow = np.empty((H, W, 2), np.float32)
        self._check(self._L.ofb_stage_blur_solve(self._h, _ptr(M), W, H, int(winsize), int(bool(gaussian)), _ptr(flow)))
        return flow

    def stage_upsample_flow(self, prev_flow, W, H, pyr_scale):
        prev_flow = np.ascontiguousarray(prev_flow, dtype=np.float32)
        Hp, Wp = prev_flow.shape[:2]
        flow = np.empty((H, W, 2), np.float32)
        self._check(self._L.ofb_stage_upsample_flow(self._h, _ptr(prev_flow), Wp, Hp, W, H, float(pyr_scale), _ptr(flow)))
        return flow


def scale_schedule(W, H, pyr_scale, levels):
    """[(k, W_k, H_k, ksize_k, sigma_k)] for k = K..0, computed by the C library."""
    L = _lib.load()
    K = L.ofb_scale_count(W, H, float(pyr_scale), int(levels))
    out = []
    for k in range(K, -1, -1):
        wk, hk, ks = C.c_int(), C.c_int(), C.c_int()
        sg = C.c_double()
        L.ofb_scale_geometry(W, H, float(pyr_scale), k, C.byref(wk), C.byref(hk), C.byref(ks), C.byref(sg))
        out.append((k, wk.value, hk.value, ks.value, sg.value))
    return out


def algorithmic_bytes(W, H, with_viz=True, **params):
    """compulsory HBM bytes of one frame pair (B_pair, plus B_viz = 19*W*H)."""
    L = _lib.load()
    prm = make_params(**{**REFERENCE_PARAMS, **params})
    b = L.ofb_algorithmic_bytes_pair(W, H, C.byref(prm))
    return b + (L.ofb_algorithmic_bytes_viz(W, H) if with_viz else 0.0)
