"""Deterministic synthetic frames with known affine motion.  Test and bench input data
(not product code, not the oracle).

No network, no video files: texture = Gaussian-blurred uniform noise on a padded canvas;
frame t = canvas warped by the t-th power of a small affine map, cropped to (H, W).
Every pair therefore has non-zero motion at the borders (avoids the A.8 branch chaos).
"""
import numpy as np


def _canvas(W, H, seed, sigma=2.0, pad=64):
    import cv2
    rng = np.random.default_rng(seed)
    c = rng.random((H + 2 * pad, W + 2 * pad), dtype=np.float32)
    c = cv2.GaussianBlur(c, (0, 0), sigma)
    c -= c.min()
    c /= max(float(c.max()), 1e-12)
    return c


def affine_step(angle_deg=0.05, scale=1.0005, tx=1.5, ty=-2.0, cx=0.0, cy=0.0):
    a = np.deg2rad(angle_deg)
    ca, sa = np.cos(a) * scale, np.sin(a) * scale
    A = np.array([[ca, -sa, 0.0], [sa, ca, 0.0], [0, 0, 1.0]])
    T = np.array([[1, 0, cx], [0, 1, cy], [0, 0, 1.0]])
    Ti = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1.0]])
    S = np.array([[1, 0, tx], [0, 1, ty], [0, 0, 1.0]])
    return S @ T @ A @ Ti


def shot(W, H, n_frames, seed=0, sigma=2.0, pad=64, step=None, out=None):
    """(n_frames, H, W) uint8 frames; frame t = warp(canvas, A^t)."""
    import cv2
    c = _canvas(W, H, seed, sigma, pad)
    Hc, Wc = c.shape
    if step is None:
        step = affine_step(cx=Wc / 2, cy=Hc / 2)
    frames = out if out is not None else np.empty((n_frames, H, W), np.uint8)
    A = np.eye(3)
    for t in range(n_frames):
        w = cv2.warpAffine(c, A[:2], (Wc, Hc), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        frames[t] = np.clip(w[pad:pad + H, pad:pad + W] * 255.0, 0, 255).astype(np.uint8)
        A = step @ A
    return frames


def rough_field(W, H, seed=0, lo=10.0, hi=20.0, cells=(6, 4)):
    """A piecewise-constant displacement field (H, W, 2): the frame is cut into cells[0] x cells[1] blocks and every block
    moves by its own vector of 10-20 px in a random direction (SURVEY.md 8d asks for flows of ~1-20 px; this is the far,
    discontinuous end, where the UpdateMatrices gathers of neighbouring pixels diverge at every block edge)."""
    rng = np.random.default_rng(seed)
    nx, ny = cells
    mag = rng.uniform(lo, hi, (ny, nx))
    ang = rng.uniform(0, 2 * np.pi, (ny, nx))
    d = np.stack([mag * np.cos(ang), mag * np.sin(ang)], -1).astype(np.float32)
    ys = np.minimum(np.arange(H) * ny // H, ny - 1)
    xs = np.minimum(np.arange(W) * nx // W, nx - 1)
    return d[ys][:, xs]


def shot_rough(W, H, n_frames, seed=0, sigma=2.0, out=None, lo=10.0, hi=20.0):
    """(n_frames, H, W) uint8 frames with ROUGH motion: frame t = canvas sampled at p + t * D(p) on a periodic canvas,
    D = rough_field (10-20 px per pair, piecewise constant, random directions)."""
    import cv2
    c = _canvas(W, H, seed, sigma, pad=0)
    D = rough_field(W, H, seed, lo, hi)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    frames = out if out is not None else np.empty((n_frames, H, W), np.uint8)
    for t in range(n_frames):
        mx = np.mod(xs + np.float32(t) * D[..., 0], np.float32(W - 1))
        my = np.mod(ys + np.float32(t) * D[..., 1], np.float32(H - 1))
        w = cv2.remap(c, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        frames[t] = np.clip(w * 255.0, 0, 255).astype(np.uint8)
    return frames


def pair(W, H, seed=0, **kw):
    f = shot(W, H, 2, seed, **kw)
    return f[0], f[1]


def analytic_flow(W, H, pad=64, step=None):
    """Ground-truth forward flow of one step at every pixel: frame1(p + flow) = frame0(p)."""
    Hc, Wc = H + 2 * pad, W + 2 * pad
    if step is None:
        step = affine_step(cx=Wc / 2, cy=Hc / 2)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    px, py = xs + pad, ys + pad
    qx = step[0, 0] * px + step[0, 1] * py + step[0, 2]
    qy = step[1, 0] * px + step[1, 1] * py + step[1, 2]
    return np.stack([qx - px, qy - py], -1).astype(np.float32)


def stress_pair(kind, W, H, seed=0):
    """Extra parity inputs: white-noise shift, coarse texture, flat field + square, identical."""
    import cv2
    rng = np.random.default_rng(seed)
    if kind == "noise_shift":
        big = rng.integers(0, 256, (H + 16, W + 16), dtype=np.uint8)
        return big[8:8 + H, 8:8 + W].copy(), big[6:6 + H, 9:9 + W].copy()
    if kind == "coarse":
        c = cv2.GaussianBlur(rng.random((H + 64, W + 64), dtype=np.float32), (0, 0), 8.0)
        c = ((c - c.min()) / (c.max() - c.min()) * 255).astype(np.uint8)
        return c[32:32 + H, 32:32 + W].copy(), c[20:20 + H, 12:12 + W].copy()
    if kind == "flat_square":
        a = np.full((H, W), 200, np.uint8)
        b = a.copy()
        a[H // 3:H // 3 + 40, W // 3:W // 3 + 40] = 30
        b[H // 3 + 3:H // 3 + 43, W // 3 + 5:W // 3 + 45] = 30
        return a, b
    if kind == "identical":
        a = (cv2.GaussianBlur(rng.random((H, W), dtype=np.float32), (0, 0), 2.0))
        a = ((a - a.min()) / (a.max() - a.min()) * 255).astype(np.uint8)
        return a, a.copy()
    raise ValueError(kind)
