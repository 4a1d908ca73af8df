"""GPU parity tests (run on the B200 box with `-m gpu`).  Everything goes through the C-ABI
(include/optflow_b200.h) via optical_flow_b200; the checker is the CPU oracle (oracle/), the committed
cv2 golden vectors (tests/golden/) and -- when the box has it -- cv2 itself.

Tolerance (BASELINE.json north_star): endpoint difference vs cv2  mean <= 1e-3 px, max <= 1e-2 px;
HSV picture within +-1 on >= 99.9 % of pixels.  Integer / byte results that have an exact definition
(magnitude, angle, hue, value, the picture for a given flow) are compared bit-exactly.
"""
import numpy as np
import pytest

import os

from conftest import golden_cases, load_golden, epe, EPE_MEAN_TOL, EPE_MAX_TOL, GOLDEN_DIR

pytestmark = pytest.mark.gpu

# what the kernels are expected to reach (reported; the hard gate is the north_star tolerance above)
TIGHT_MEAN, TIGHT_MAX = 2e-5, 1e-3


@pytest.fixture(scope="module")
def eng():
    import optical_flow_b200 as ofb
    return ofb.Farneback(0)


def _textured(W, H, seed):
    """cv2-free synthetic texture (smooth noise) so the stage tests do not depend on cv2."""
    rng = np.random.default_rng(seed)
    a = rng.random((H + 16, W + 16)).astype(np.float32)
    for _ in range(3):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5
    a = (a - a.min()) / (a.max() - a.min())
    f0 = (a[8:8 + H, 8:8 + W] * 255).astype(np.uint8)
    f1 = (a[7:7 + H, 10:10 + W] * 255).astype(np.uint8)
    return f0, f1


# ------------------------------------------------------------------------------------------------
# per-stage parity against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,sigma", [(5, 1.2), (7, 1.5), (3, 0.0)])
def test_stage_polyexp_exact_mode_is_bit_identical_to_the_oracle(eng, oracle, n, sigma):
    """The default batched kernel (option polyexp_exact = 1) with cv2's own float / double mix in the horizontal pass (float sums and
    differences, four float products, separate double multiply and add) gives the oracle's R bit for bit."""
    f0, _ = _textured(203, 97, 2)
    img = oracle.gaussian_blur(f0.astype(np.float32), 3, 0.0)
    got = eng.stage_polyexp(img, n, sigma)               # polyexp_exact is the default
    assert np.array_equal(got, oracle.polyexp(img, n, sigma)), float(np.abs(got - oracle.polyexp(img, n, sigma)).max())


@pytest.mark.parametrize("winsize", [3, 9, 15, 16, 33])
def test_stage_blur_solve_exact_mode_follows_cv2_running_sums(eng, oracle, winsize):
    """The default box-window kernel (k_iter64, option exact_window_sums = 1): the oracle's flow to f32 rounding even on a matrix field whose windows are
    exactly rank-deficient (one edge direction), where the default f32 sums differ visibly."""
    rng = np.random.default_rng(winsize)
    H, W = 150, 203
    ys, xs = np.mgrid[0:H, 0:W]
    gx = np.where((xs // 12) % 2 == 0, 80.0, -80.0).astype(np.float32)          # a pure-x pattern: G is rank 1 everywhere
    M = np.zeros((H, W, 5), np.float32)
    M[..., 0] = gx * gx; M[..., 3] = gx * 0.37
    M += rng.random((H, W, 5)).astype(np.float32) * 1e-3
    ref = oracle.blur_solve(M, winsize)
    got = eng.stage_blur_solve(M, winsize)               # exact_window_sums is the default
    err = np.abs(got - ref).max()
    assert err <= 1e-5 * max(1.0, float(np.abs(ref).max())), (winsize, float(err))


@pytest.mark.parametrize("W,H,pyr,levels", [(320, 180, 0.5, 3), (129, 77, 0.5, 3), (200, 150, 0.7, 4), (640, 360, 0.5, 3)])
def test_stage_level_image(eng, oracle, W, H, pyr, levels):
    f0, _ = _textured(W, H, 1)
    for (k, wk, hk, ks, sg, sc) in oracle.scale_schedule(W, H, pyr, levels):
        ref = oracle.level_image(f0, wk, hk, ks, sg)
        got = eng.stage_level_image(f0, pyr, k)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 2e-4, (k, np.abs(got - ref).max())      # 0..255 scale
    # f32 input path gives the same image as the u8 path
    a = eng.stage_level_image(f0, pyr, 1)
    b = eng.stage_level_image(f0.astype(np.float32), pyr, 1)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("generic", [0, 1, 2])
@pytest.mark.parametrize("n,sigma", [(5, 1.2), (7, 1.5), (3, 0.0), (1, 1.2), (9, 2.0)])
def test_stage_polyexp(eng, oracle, n, sigma, generic):
    """generic 0 = default (cv2's mix, bit-exact), 1 = the simple kernels, 2 = option fast_arithmetic (f64-FMA horizontal pass,
    FFMA2 vertical pass on interior tiles)."""
    f0, _ = _textured(203, 97, 2)
    img = oracle.gaussian_blur(f0.astype(np.float32), 3, 0.0)
    ref = oracle.polyexp(img, n, sigma)
    eng.set_option("generic_kernels", 1 if generic == 1 else 0)
    eng.set_option("fast_arithmetic", 1 if generic == 2 else 0)
    try:
        got = eng.stage_polyexp(img, n, sigma)
    finally:
        eng.set_option("generic_kernels", 0)
        eng.set_option("fast_arithmetic", 0)
    if generic != 2 or n not in (3, 5, 7):
        # same expressions, same order, no contraction: bit-exact
        assert np.array_equal(got, ref), float(np.abs(got - ref).max())
    else:
        # unrolled kernel: horizontal pass entirely in f64 (cv2 rounds a few terms to f32 first)
        assert np.abs(got - ref).max() <= 1e-4, float(np.abs(got - ref).max())


@pytest.mark.parametrize("generic", [0, 1])
def test_stage_update_matrices(eng, oracle, generic):
    rng = np.random.default_rng(3)
    H, W = 75, 131
    R0 = rng.normal(0, 5, (H, W, 5)).astype(np.float32)
    R1 = rng.normal(0, 5, (H, W, 5)).astype(np.float32)
    flow = rng.normal(0, 3, (H, W, 2)).astype(np.float32)
    flow[:10] *= 20                                      # push some samples out of bounds
    ref = oracle.update_matrices(R0, R1, flow)
    eng.set_option("generic_kernels", generic)
    try:
        got = eng.stage_update_matrices(R0, R1, flow)
    finally:
        eng.set_option("generic_kernels", 0)
    assert np.array_equal(got, ref), float(np.abs(got - ref).max())


@pytest.mark.parametrize("generic", [0, 1, 2])
@pytest.mark.parametrize("winsize", [15, 9, 31, 16, 3, 2, 1, 33, 41])
def test_stage_blur_solve_box(eng, oracle, winsize, generic):
    """generic 0 = the default kernel (k_iter64: cv2's double running sums), 1 = the simple global-memory kernels,
    2 = option fast_arithmetic (k_iter: f32 van Herk sums, compensated f32 solve)."""
    rng = np.random.default_rng(4)
    H, W = 143, 211
    r = rng.normal(0, 3, (H, W, 5)).astype(np.float32)
    M = np.empty_like(r)                                  # a plausible M: G SPD-ish, h arbitrary
    M[..., 0] = r[..., 0] ** 2 + r[..., 2] ** 2
    M[..., 1] = (r[..., 0] + r[..., 1]) * r[..., 2]
    M[..., 2] = r[..., 1] ** 2 + r[..., 2] ** 2
    M[..., 3] = r[..., 3]
    M[..., 4] = r[..., 4]
    ref = oracle.blur_solve(M, winsize, gaussian=False)
    eng.set_option("generic_kernels", 1 if generic == 1 else 0)
    eng.set_option("fast_arithmetic", 1 if generic == 2 else 0)
    try:
        got = eng.stage_blur_solve(M, winsize, gaussian=False)
    finally:
        eng.set_option("generic_kernels", 0)
        eng.set_option("fast_arithmetic", 0)
    mean, mx = epe(got, ref)
    scale = max(1.0, float(np.abs(ref).max()))
    # default / generic: f64 sums.  fast_arithmetic: f32 van Herk window sums + Kahan-compensated f32 solve.
    tol = 2e-4 if generic == 2 else 2e-5
    print("blur_solve box win %d generic %d: max diff %.2e (|flow| max %.2f)" % (winsize, generic, mx, scale))
    assert mx <= tol * scale, (winsize, mean, mx, scale)


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("winsize", [15, 16, 9, 31, 1, 41])
def test_stage_blur_solve_gauss(eng, oracle, winsize, generic):
    rng = np.random.default_rng(5)
    H, W = 90, 160
    r = rng.normal(0, 3, (H, W, 5)).astype(np.float32)
    M = np.empty_like(r)
    M[..., 0] = r[..., 0] ** 2 + r[..., 2] ** 2
    M[..., 1] = (r[..., 0] + r[..., 1]) * r[..., 2]
    M[..., 2] = r[..., 1] ** 2 + r[..., 2] ** 2
    M[..., 3] = r[..., 3]
    M[..., 4] = r[..., 4]
    ref = oracle.blur_solve(M, winsize, gaussian=True)
    eng.set_option("generic_kernels", generic)
    try:
        got = eng.stage_blur_solve(M, winsize, gaussian=True)
    finally:
        eng.set_option("generic_kernels", 0)
    if generic or winsize // 2 < 1 or winsize // 2 > 16:
        assert np.array_equal(got, ref), epe(got, ref)       # same taps, same order, f64 solve: bit-exact
    else:
        # fast path: the blurred field is bit-identical to cv2's, the solve is compensated f32 instead of f64
        mean, mx = epe(got, ref)
        assert mx <= 2e-5 * max(1.0, float(np.abs(ref).max())), (winsize, mean, mx)


def test_stage_upsample_flow(eng, oracle):
    rng = np.random.default_rng(6)
    for (wp, hp, w, h, ps) in [(80, 45, 160, 90, 0.5), (64, 38, 129, 77, 0.5), (115, 65, 165, 93, 0.7)]:
        prev = rng.normal(0, 2, (hp, wp, 2)).astype(np.float32)
        ref = oracle.upsample_flow(prev, w, h, ps)
        got = eng.stage_upsample_flow(prev, w, h, ps)
        assert np.array_equal(got, ref), float(np.abs(got - ref).max())


# ------------------------------------------------------------------------------------------------
# the whole call against the committed cv2 golden vectors and the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("name", golden_cases())
def test_farneback_matches_cv2_golden(eng, oracle, name, generic):
    g = load_golden(name)
    init = g.get("init_flow")
    eng.set_option("generic_kernels", generic)
    try:
        flow = eng.calc(g["prev"], g["next"], None if init is None else init.copy(), **g["kw"])
    finally:
        eng.set_option("generic_kernels", 0)
    assert flow.shape == g["flow"].shape and flow.dtype == np.float32
    mean, mx = epe(flow, g["flow"])
    print("golden %-32s generic=%d  EPE vs cv2: mean %.2e max %.2e" % (name, generic, mean, mx))
    assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL, (name, mean, mx)
    assert mean <= TIGHT_MEAN and mx <= TIGHT_MAX, (name, mean, mx)
    ref = oracle.farneback(g["prev"], g["next"], None if init is None else init.copy(), **g["kw"])
    mean, mx = epe(flow, ref)
    assert mean <= TIGHT_MEAN and mx <= TIGHT_MAX, ("vs oracle", name, mean, mx)


@pytest.mark.parametrize("name", golden_cases())
def test_picture_on_cv2_flow(eng, oracle, name):
    g = load_golden(name)
    bgr = eng.flow_to_bgr(g["flow"])
    # bit-exact against the oracle's truncating form (== cv2's vector body over the whole (H,V) table)
    assert np.array_equal(bgr, oracle.viz(g["flow"], 0))
    d = np.abs(bgr.astype(np.int16) - g["bgr"].astype(np.int16))
    assert d.max() <= 1 and (d == 0).all(-1).mean() > 0.85


@pytest.mark.parametrize("name", golden_cases())
def test_polar_and_feature_on_cv2_flow(eng, oracle, name):
    g = load_golden(name)
    mag, ang = eng.cart_to_polar(g["flow"])
    omag, oang = oracle.cart_to_polar(g["flow"])
    assert np.array_equal(mag, omag) and np.array_equal(ang, oang)
    s = float(eng.sum_magnitude(g["flow"]))
    assert abs(s - float(g["magsum"])) <= 1e-5 * abs(float(g["magsum"]))


def test_picture_hsv_table_exhaustive(eng):
    """Every (hue, value) byte pair: build a flow field whose quantised H,V sweep the table."""
    z = np.load(__import__("conftest").GOLDEN_DIR + "/hsv2bgr_table.npz")
    from oracle import c_oracle
    # polar grid: angle spans 0..2pi (hue bytes 0..255 via the mod-256 wrap), magnitude 0..1
    ang = np.linspace(0, 2 * np.pi, 1440, endpoint=False, dtype=np.float64)[:, None]
    mag = np.linspace(0, 1, 1024, dtype=np.float64)[None, :]
    flow = np.stack([mag * np.cos(ang), mag * np.sin(ang)], -1).astype(np.float32)
    bgr = eng.flow_to_bgr(flow)
    ob, hue, val = c_oracle.viz(flow, 0, return_hv=True)
    assert np.array_equal(bgr, ob)
    assert np.array_equal(bgr, z["body"][hue, val])          # cv2's own table, vector body
    assert len(np.unique(hue)) == 256 and len(np.unique(val)) == 256


# ------------------------------------------------------------------------------------------------
# the drop-in contract
# ------------------------------------------------------------------------------------------------
def test_dropin_contract(eng):
    import optical_flow_b200 as ofb
    g = load_golden("featurepath_129x77")
    kw = g["kw"]
    a = ofb.calcOpticalFlowFarneback(g["prev"], g["next"], None, kw["pyr_scale"], kw["levels"], kw["winsize"],
                                     kw["iterations"], kw["poly_n"], kw["poly_sigma"], kw["flags"])
    b = ofb.calcOpticalFlowFarneback(prev=g["prev"], next=g["next"], flow=None, **kw)
    assert np.array_equal(a, b)                               # deterministic, positional == keyword
    assert a.flags.c_contiguous and a.strides == (8 * 129, 8, 4)
    buf = np.full((77, 129, 2), 7, np.float32)
    r = ofb.calcOpticalFlowFarneback(g["prev"], g["next"], buf, **kw)
    assert r is buf and np.array_equal(buf, a)               # written in place, same object returned
    bad = np.zeros((10, 10, 2), np.float32)
    r = ofb.calcOpticalFlowFarneback(g["prev"], g["next"], bad, **kw)
    assert r is not bad and np.array_equal(r, a) and not bad.any()
    for dt in (np.uint16, np.int16, np.float32, np.float64):  # any depth, identical result
        r = ofb.calcOpticalFlowFarneback(g["prev"].astype(dt), g["next"].astype(dt), None, **kw)
        assert np.array_equal(r, a), dt
    r = ofb.calcOpticalFlowFarneback(g["prev"], g["next"].astype(np.float64), None, **kw)
    assert np.array_equal(r, a)
    big = np.zeros((154, 258), np.uint8)
    big[::2, ::2] = g["prev"]
    r = ofb.calcOpticalFlowFarneback(big[::2, ::2], g["next"], None, **kw)
    assert np.array_equal(r, a)                               # non-contiguous view
    with pytest.raises(ofb.error):
        ofb.calcOpticalFlowFarneback(g["prev"], g["next"][:-1], None, **kw)
    with pytest.raises(ofb.error):
        ofb.calcOpticalFlowFarneback(g["prev"], g["next"], None, **dict(kw, pyr_scale=1.0))
    with pytest.raises(ofb.error):
        ofb.calcOpticalFlowFarneback(g["prev"], g["next"], None, **dict(kw, flags=4))
    z = ofb.calcOpticalFlowFarneback(g["prev"], g["next"], None, **dict(kw, iterations=0))
    assert z.shape == a.shape and not z.any()                 # iterations=0 returns zeros
    mag, ang = ofb.cartToPolar(a[..., 0], a[..., 1])
    assert mag.shape == (77, 129) and ang.shape == (77, 129)


def test_degenerate_parameters_are_finite(eng, oracle):
    g = load_golden("levels0_64x48")
    for kw in (dict(winsize=1), dict(poly_n=1), dict(levels=-2), dict(levels=0, winsize=2)):
        p = dict(g["kw"]); p.update(kw)
        f = eng.calc(g["prev"], g["next"], None, **p)
        assert np.isfinite(f).all()
    # single scale, winsize=1: no chaotic amplification across scales -> tight parity is still expected
    p = dict(g["kw"], winsize=1, levels=0, iterations=1)
    f = eng.calc(g["prev"], g["next"], None, **p)
    ref = oracle.farneback(g["prev"], g["next"], None, **p)
    mean, mx = epe(f, ref)
    assert mx <= 1e-3 * max(1.0, float(np.abs(ref).max())), (mean, mx)


@pytest.mark.parametrize("W,H", [(8, 8), (16, 9), (5, 40), (1, 50), (50, 1), (2, 2), (31, 33), (100, 7), (97, 131)])
def test_tiny_and_odd_frames(eng, oracle, W, H):
    """cv2 accepts any frame size (scales that would be < 32 px are dropped); so must the kernels."""
    rng = np.random.default_rng(W * 100 + H)
    a = rng.integers(0, 256, (H, W), dtype=np.uint8)
    b = np.roll(a, 1, 1)
    kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    for generic in (0, 1):
        eng.set_option("generic_kernels", generic)
        try:
            f = eng.calc(a, b, None, **kw)
        finally:
            eng.set_option("generic_kernels", 0)
        ref = oracle.farneback(a, b, None, **kw)
        assert f.shape == (H, W, 2) and np.isfinite(f).all()
        mean, mx = epe(f, ref)
        assert mx <= 1e-3 * max(1.0, float(np.abs(ref).max())), (W, H, generic, mean, mx)
    pic = eng.pair(a, b, want_bgr=True, want_magsum=True, want_flow=True, **kw)
    assert np.array_equal(pic["bgr"], oracle.viz(pic["flow"], 0))


@pytest.mark.parametrize("kw", [dict(winsize=41), dict(winsize=64, iterations=2), dict(poly_n=9, poly_sigma=2.0), dict(poly_n=2),
                                dict(winsize=41, flags=256), dict(pyr_scale=0.8, levels=6), dict(pyr_scale=0.3, levels=2),
                                dict(winsize=33, poly_n=7, iterations=1)])
def test_parameters_outside_the_fast_paths(eng, oracle, kw):
    """winsize > 33, poly_n not in {3,5,7}, unusual pyr_scale: generic kernels, same results."""
    f0, f1 = _smooth_pair(352, 288, 41)
    p = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    p.update(kw)
    flow = eng.calc(f0, f1, None, **p)
    ref = oracle.farneback(f0, f1, None, **p)
    mean, mx = epe(flow, ref)
    print("params %s: mean %.2e max %.2e" % (kw, mean, mx))
    assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL, (kw, mean, mx)


def test_device_pointer_entry_points_and_pitch(eng):
    """ofb_farneback_device / ofb_flow_to_bgr_device / ofb_sum_magnitude_device with a row pitch larger than the row."""
    import ctypes as C
    from optical_flow_b200 import _lib, make_params
    L = _lib.load()
    f0, f1 = _smooth_pair(200, 120, 43)
    H, W = f0.shape
    pitch = 256
    pad0 = np.zeros((H, pitch), np.uint8); pad0[:, :W] = f0
    pad1 = np.zeros((H, pitch), np.uint8); pad1[:, :W] = f1
    d0, d1 = eng.device_alloc(pad0.nbytes), eng.device_alloc(pad1.nbytes)
    dflow, dbgr, dsum = eng.device_alloc(W * H * 8), eng.device_alloc(W * H * 3), eng.device_alloc(4)
    eng.h2d(d0, pad0); eng.h2d(d1, pad1)
    prm = make_params()
    vp = C.c_void_p
    assert L.ofb_farneback_device(eng._h, vp(d0), vp(d1), 0, W, H, pitch, pitch, vp(dflow), C.byref(prm)) == 0
    assert L.ofb_flow_to_bgr_device(eng._h, vp(dflow), W, H, vp(dbgr)) == 0
    assert L.ofb_sum_magnitude_device(eng._h, vp(dflow), W, H, vp(dsum)) == 0
    flow = np.empty((H, W, 2), np.float32); bgr = np.empty((H, W, 3), np.uint8); sm = np.empty(1, np.float32)
    eng.d2h(flow, dflow); eng.d2h(bgr, dbgr); eng.d2h(sm, dsum)
    ref = eng.calc(f0, f1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    assert np.array_equal(flow, ref)
    assert np.array_equal(bgr, eng.flow_to_bgr(ref))
    assert abs(float(sm[0]) - float(eng.sum_magnitude(ref))) <= 1e-6 * float(sm[0])
    for p in (d0, d1, dflow, dbgr, dsum):
        eng.device_free(p)
    # error paths of the C ABI: status codes, messages, no exceptions
    bad = make_params(pyr_scale=1.0)
    assert L.ofb_farneback_device(eng._h, vp(1), vp(1), 0, W, H, 0, 0, vp(1), C.byref(bad)) == _lib.OFB_ERR_ASSERT
    assert b"pyrScale_ < 1" in L.ofb_last_error(eng._h)
    assert L.ofb_farneback_device(eng._h, vp(None), vp(1), 0, W, H, 0, 0, vp(1), C.byref(prm)) == _lib.OFB_ERR_BAD_ARG
    assert L.ofb_set_option(eng._h, b"no_such_option", 1) == _lib.OFB_ERR_BAD_ARG


# ------------------------------------------------------------------------------------------------
# fused pair / shot forms
# ------------------------------------------------------------------------------------------------
def test_pair_and_shot_equal_the_per_call_results(eng, oracle):
    W, H, n = 192, 108, 5
    rng = np.random.default_rng(8)
    base = rng.random((H + 40, W + 40)).astype(np.float32)
    for _ in range(3):
        base = (base + np.roll(base, 1, 0) + np.roll(base, -1, 0) + np.roll(base, 1, 1) + np.roll(base, -1, 1)) / 5
    base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
    frames = np.stack([base[10 + t:10 + t + H, 12 + 2 * t:12 + 2 * t + W] for t in range(n)])
    kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    flows = [eng.calc(frames[t], frames[t + 1], None, **kw) for t in range(n - 1)]
    res = eng.shot(frames, want_bgr=True, want_magsum=True, want_flow=True, **kw)
    assert res["device_ms"] > 0
    for t in range(n - 1):
        assert np.array_equal(res["flow"][t], flows[t]), t           # same kernels, same inputs: bitwise
        assert np.array_equal(res["bgr"][t], eng.flow_to_bgr(flows[t]))
        assert np.array_equal(res["bgr"][t], oracle.viz(flows[t], 0))
        s = float(eng.sum_magnitude(flows[t]))
        assert abs(float(res["magsum"][t]) - s) <= 1e-6 * s
        assert abs(s - oracle.sum_magnitude(flows[t])) <= 1e-5 * s
    one = eng.pair(frames[0], frames[1], want_bgr=True, want_magsum=True, want_flow=True, **kw)
    assert np.array_equal(one["flow"], flows[0]) and np.array_equal(one["bgr"], res["bgr"][0])
    # two frames = one pair; pinned buffers work as inputs and outputs
    import optical_flow_b200 as ofb
    pin = ofb.pinned_empty(frames.shape, np.uint8)
    pin[:] = frames
    out = ofb.pinned_empty((n - 1, H, W, 3), np.uint8)
    res2 = eng.shot(pin, want_bgr=True, out_bgr=out, **kw)
    assert res2["bgr"] is out and np.array_equal(out, res["bgr"])


def test_batch_size_does_not_change_results(eng):
    """Pairs are processed in chunks of `batch` per launch; a pair's result must not depend on the chunking."""
    W, H, n = 224, 136, 8
    f0, _ = _textured(W + 40, H + 40, 10)
    frames = np.stack([f0[4 + t:4 + t + H, 2 * t:2 * t + W] for t in range(n)])
    ref = None
    try:
        for b, b0 in ((1, 0), (3, 0), (4, 1), (7, 2), (16, 0)):
            eng.set_option("batch", b)
            eng.set_option("batch_scale0", b0)
            res = eng.shot(frames, want_bgr=True, want_flow=True, want_magsum=True)
            if ref is None:
                ref = res
            else:
                assert np.array_equal(res["flow"], ref["flow"]), (b, b0)
                assert np.array_equal(res["bgr"], ref["bgr"]), (b, b0)
    finally:
        eng.set_option("batch", 0)
        eng.set_option("batch_scale0", 0)


def test_shot_sharded_ranges_reassemble(eng):
    """Multi-GPU partitioning is by contiguous pair ranges with one overlap frame:
    processing the ranges separately gives exactly the unsharded shot."""
    import optical_flow_b200 as ofb
    W, H, n = 160, 96, 9
    f0, f1 = _textured(W + 40, H + 40, 9)
    frames = np.stack([f0[5 + t:5 + t + H, 3 * t:3 * t + W] for t in range(n)])
    whole = eng.shot(frames, want_bgr=True, want_magsum=True)
    for ws in (2, 3):
        parts_bgr, parts_sum = [], []
        for r in range(ws):
            s, e = ofb.shard_pairs(n - 1, ws, r)
            res = eng.shot(frames[s:e + 1], want_bgr=True, want_magsum=True)
            parts_bgr.append(res["bgr"]); parts_sum.append(res["magsum"])
        assert np.array_equal(np.concatenate(parts_bgr), whole["bgr"])
        assert np.array_equal(np.concatenate(parts_sum), whole["magsum"])


# ------------------------------------------------------------------------------------------------
# BASELINE.json sizes: 1080p reference parameters, the 4K Gaussian config, the parameter sweep
# ------------------------------------------------------------------------------------------------
def _cv2_or_none():
    try:
        import cv2
        return cv2
    except Exception:
        return None


def _smooth_pair(W, H, seed, shift=(2.3, -1.7)):
    """cv2-free pair with smooth, NON-INTEGER, spatially varying motion (a small affine map, bilinear sampling).
    Integer translations put border pixels exactly on the in/out-of-bounds switch of A.8, where a 1e-7 change of
    the flow flips the branch and moves a 15x15 neighbourhood by ~1e-2 px on cv2 itself."""
    rng = np.random.default_rng(seed)
    pad = 24
    a = rng.random(((H + 2 * pad) // 4 + 2, (W + 2 * pad) // 4 + 2)).astype(np.float32)
    a = np.kron(a, np.ones((4, 4), np.float32))[:H + 2 * pad, :W + 2 * pad]
    for _ in range(4):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5
    a = (a - a.min()) / (a.max() - a.min()) * 255
    f0 = a[pad:pad + H, pad:pad + W]
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    sx = xs + pad + np.float32(shift[0]) + np.float32(0.0021) * ys - np.float32(0.0013) * xs
    sy = ys + pad + np.float32(shift[1]) + np.float32(0.0017) * xs + np.float32(0.0011) * ys
    x0 = np.floor(sx).astype(np.int64); y0 = np.floor(sy).astype(np.int64)
    fx = sx - x0; fy = sy - y0
    x0 = np.clip(x0, 0, a.shape[1] - 2); y0 = np.clip(y0, 0, a.shape[0] - 2)
    f1 = (a[y0, x0] * (1 - fx) * (1 - fy) + a[y0, x0 + 1] * fx * (1 - fy) +
          a[y0 + 1, x0] * (1 - fx) * fy + a[y0 + 1, x0 + 1] * fx * fy)
    return np.ascontiguousarray(f0.astype(np.uint8)), np.ascontiguousarray(f1.astype(np.uint8))


def test_1080p_reference_parameters(eng, oracle):
    f0, f1 = _smooth_pair(1920, 1080, 21)
    kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    flow = eng.calc(f0, f1, None, **kw)
    ref = oracle.farneback(f0, f1, None, **kw)
    mean, mx = epe(flow, ref)
    print("1080p vs oracle: mean %.2e max %.2e  |flow|max %.2f" % (mean, mx, np.abs(ref).max()))
    assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL
    cv2 = _cv2_or_none()
    if cv2 is not None:
        cf = cv2.calcOpticalFlowFarneback(f0, f1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        mean, mx = epe(flow, cf)
        print("1080p vs cv2 %s: mean %.2e max %.2e" % (cv2.__version__, mean, mx))
        assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL
        from oracle import cv2_reference
        bgr = eng.flow_to_bgr(flow)
        cb = cv2_reference.viz(cf)
        within1 = (np.abs(bgr.astype(np.int16) - cb.astype(np.int16)) <= 1).all(-1).mean()
        print("1080p picture end-to-end within +-1: %.5f" % within1)
        assert within1 >= 0.999
    # properties that do not need a reference: determinism, and the picture / feature of this very flow
    assert np.array_equal(flow, eng.calc(f0, f1, None, **kw))
    res = eng.pair(f0, f1, want_bgr=True, want_magsum=True, want_flow=True, **kw)
    assert np.array_equal(res["flow"], flow)
    assert np.array_equal(res["bgr"], oracle.viz(flow, 0))
    assert abs(float(res["magsum"]) - oracle.sum_magnitude(flow)) <= 1e-5 * oracle.sum_magnitude(flow)


def test_4k_gaussian_config(eng, oracle):
    """BASELINE configs[2]: 3840x2160, levels 5, poly_n 7, poly_sigma 1.5, OPTFLOW_FARNEBACK_GAUSSIAN."""
    kw = dict(pyr_scale=0.5, levels=5, winsize=15, iterations=3, poly_n=7, poly_sigma=1.5, flags=256)
    f0, f1 = _smooth_pair(3840, 2160, 22, shift=(3.4, 5.2))
    flow = eng.calc(f0, f1, None, **kw)
    assert np.isfinite(flow).all()
    cv2 = _cv2_or_none()
    if cv2 is not None:
        cf = cv2.calcOpticalFlowFarneback(f0, f1, None, 0.5, 5, 15, 3, 7, 1.5, 256)
        mean, mx = epe(flow, cf)
        print("4K gaussian vs cv2: mean %.2e max %.2e" % (mean, mx))
    else:
        ref = oracle.farneback(f0, f1, None, **kw)
        mean, mx = epe(flow, ref)
        print("4K gaussian vs oracle: mean %.2e max %.2e" % (mean, mx))
    assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL


@pytest.mark.parametrize("winsize", [9, 15, 31])
@pytest.mark.parametrize("iterations", [3, 10])
def test_parameter_sweep_720p(eng, oracle, winsize, iterations):
    """BASELINE configs[4] at 1280x720 (the larger sizes run in bench.py --sweep)."""
    kw = dict(pyr_scale=0.5, levels=3, winsize=winsize, iterations=iterations, poly_n=5, poly_sigma=1.2, flags=0)
    f0, f1 = _smooth_pair(1280, 720, 23 + winsize)
    flow = eng.calc(f0, f1, None, **kw)
    ref = oracle.farneback(f0, f1, None, **kw)
    mean, mx = epe(flow, ref)
    print("720p win %d it %d vs oracle: mean %.2e max %.2e" % (winsize, iterations, mean, mx))
    assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL


def test_stress_inputs_reported(eng, oracle):
    """SURVEY.md 8d stress inputs: white-noise shift, flat field with a moving square, identical frames.
    Gated on the mean; the max is reported (branch flips of A.8 at ~0 motion are precision-independent)."""
    rng = np.random.default_rng(31)
    H, W = 270, 480
    big = rng.integers(0, 256, (H + 16, W + 16), dtype=np.uint8)
    cases = {"noise_shift": (big[8:8 + H, 8:8 + W].copy(), big[6:6 + H, 9:9 + W].copy())}
    a = np.full((H, W), 200, np.uint8); b = a.copy()
    a[90:130, 160:200] = 30; b[93:133, 165:205] = 30
    cases["flat_square"] = (a, b)
    t, _ = _textured(W, H, 32)
    cases["identical"] = (t, t.copy())
    kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    for name, (f0, f1) in cases.items():
        flow = eng.calc(f0, f1, None, **kw)
        ref = oracle.farneback(f0, f1, None, **kw)
        mean, mx = epe(flow, ref)
        n_bad = int((np.sqrt(((flow - ref) ** 2).sum(-1)) > EPE_MAX_TOL).sum())
        print("stress %-12s mean %.2e max %.2e  pixels > 1e-2: %d" % (name, mean, mx, n_bad))
        assert mean <= EPE_MEAN_TOL, name
        if name != "identical":
            assert mx <= EPE_MAX_TOL, name


def test_native_library_is_what_ran(eng):
    """The loaded code is the in-tree CUDA library, and kernels were actually launched."""
    from optical_flow_b200 import _lib
    maps = open("/proc/self/maps").read()
    assert "libofb200.so" in maps
    stats = eng.kernel_stats()
    assert stats.get("iter_fused", (0, 0))[0] > 0 and stats.get("polyexp_scale0", (0, 0))[0] > 0


# ------------------------------------------------------------------------------------------------
# frame preprocessing on the GPU: integer algorithms, bit-exact
# ------------------------------------------------------------------------------------------------
def _preprocess_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "preprocess.npz"))
    return z, sorted(k[:-len("_src")] for k in z.files if k.endswith("_src"))


@pytest.mark.parametrize("name", _preprocess_cases()[1])
def test_preprocess_matches_cv2_golden_bit_exactly(eng, name):
    z, _ = _preprocess_cases()
    src, resized, gray = z[name + "_src"], z[name + "_resized"], z[name + "_gray"]
    dh, dw = resized.shape[:2]
    assert np.array_equal(eng.resize(src, (dw, dh)), resized)
    assert np.array_equal(eng.resize(src[..., 1].copy(), (dw, dh)), z[name + "_resized_c1"])
    assert np.array_equal(eng.resize(src, (dw, dh), to_gray=True), gray)          # fused resize + gray
    assert np.array_equal(eng.bgr_to_gray(resized), gray)
    assert np.array_equal(eng.bgr_to_gray(src), z[name + "_gray_fullres"])


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 129, 72), (1280, 720, 129, 72), (640, 360, 320, 180), (333, 222, 129, 86),
                                         (96, 64, 129, 86), (1920, 1080, 1920, 1080)])
def test_preprocess_matches_oracle_at_video_sizes(eng, oracle, sw, sh, dw, dh):
    rng = np.random.default_rng(sw * 7 + dw)
    src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    ref = oracle.resize_u8(src, (dw, dh))
    assert np.array_equal(eng.resize(src, (dw, dh)), ref)
    assert np.array_equal(eng.resize(src, (dw, dh), to_gray=True), oracle.bgr2gray(ref))
    assert np.array_equal(eng.bgr_to_gray(src), oracle.bgr2gray(src))


def test_shot_bgr_equals_host_preprocessing_then_shot(eng, oracle):
    """Feeding decoded BGR frames (resize + gray on the GPU) gives bit-identical pictures, sums and gray frames to
    preprocessing with the oracle and feeding gray frames; both with and without the resize."""
    rng = np.random.default_rng(77)
    base = (rng.random((5, 120 + 16, 160 + 16, 3)) * 255).astype(np.uint8)
    import scipy.ndimage as ndi
    base = ndi.gaussian_filter(base.astype(np.float32), (0, 2, 2, 0))
    base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
    frames = np.stack([base[0, i:i + 120, 2 * i:2 * i + 160] for i in range(5)])          # moving crop of one texture
    for dsize in (None, (129, 96)):
        gray = np.stack([oracle.bgr2gray(f if dsize is None else oracle.resize_u8(f, dsize)) for f in frames])
        a = eng.shot_bgr(frames, dsize=dsize, want_bgr=True, want_magsum=True, want_gray=True)
        b = eng.shot(gray, want_bgr=True, want_magsum=True)
        assert np.array_equal(a["gray"], gray)
        assert np.array_equal(a["bgr"], b["bgr"])
        assert np.array_equal(a["magsum"], b["magsum"])
        c = eng.pairs_bgr(frames[:-1], frames[1:], dsize=dsize, want_magsum=True)
        assert np.array_equal(c["magsum"], b["magsum"])


def test_polyexp_tma_path_is_bit_identical(eng):
    """Engine option "polyexp_tma": the persistent scale-0 kernel with TMA-staged halo tiles gives the same bits as the
    default kernel (same arithmetic; only where the raw patch comes from differs), on a frame with interior and border tiles."""
    f = _textured(448, 200, 5)[0]
    frames = np.stack([np.roll(f, (i, 2 * i), (0, 1)) for i in range(4)])
    eng.set_option("fast_arithmetic", 1)              # the TMA kernel is a variant of the fast-arithmetic path
    eng.set_option("polyexp_fast", 0)                 # ... with pe_tile's arithmetic (pe_tile_fast uses FFMA2 in the vertical pass)
    try:
        ref = eng.shot(frames, want_bgr=True, want_flow=True)
        eng.set_option("polyexp_tma", 1)
        got = eng.shot(frames, want_bgr=True, want_flow=True)
    finally:
        eng.set_option("polyexp_tma", 0)
        eng.set_option("polyexp_fast", 1)
        eng.set_option("fast_arithmetic", 0)
    assert np.array_equal(got["flow"], ref["flow"])
    assert np.array_equal(got["bgr"], ref["bgr"])


def test_hsv_table_and_arithmetic_pictures_are_identical(eng, oracle):
    """Option "hsv_table": the looked-up colour conversion equals the per-pixel arithmetic and the oracle, bit for bit,
    on a flow field that reaches every hue and a wide range of magnitudes."""
    rng = np.random.default_rng(9)
    H, W = 120, 256
    ang = rng.random((H, W)) * 2 * np.pi
    mag = rng.random((H, W)) ** 3 * 40
    flow = np.stack([mag * np.cos(ang), mag * np.sin(ang)], -1).astype(np.float32)
    frames = np.zeros((2, H, W), np.uint8)                       # the shot path needs frames; the picture is what is tested
    a = eng.flow_to_bgr(flow)
    eng.set_option("hsv_table", 0)
    try:
        b = eng.flow_to_bgr(flow)
    finally:
        eng.set_option("hsv_table", 1)
    assert np.array_equal(a, b)
    assert np.array_equal(a, oracle.viz(flow, 0))


@pytest.mark.parametrize("W,H", [(64, 48), (68, 52), (72, 33), (132, 76), (260, 140), (324, 200), (448, 17)])
def test_aligned_width_geometries_shot_equals_pairs_and_oracle(eng, oracle, W, H):
    """Widths that are multiples of 4 take the vectorised staging paths (polyexp, column-first pyramid) with every
    mix of interior / border tiles; the batched shot must equal per-pair calls bit for bit and the oracle within tolerance."""
    f0, f1 = _smooth_pair(W, H, 7 * W + H)
    f2 = np.ascontiguousarray(np.flipud(f0))
    frames = np.stack([f0, f1, f2, f1])
    res = eng.shot(frames, want_bgr=True, want_flow=True)
    kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    for t in range(3):
        one = eng.calc(frames[t], frames[t + 1], None, **kw)
        assert np.array_equal(res["flow"][t], one), (W, H, t)
    ref = oracle.farneback(f0, f1, None, **kw)
    mean, mx = epe(res["flow"][0], ref)
    assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL, (W, H, mean, mx)


@pytest.mark.parametrize("sw,sh,dsize", [(131, 77, None), (131, 77, (64, 38)), (200, 113, (129, 72)), (66, 50, None)])
def test_shot_bgr_with_unaligned_frames(eng, oracle, sw, sh, dsize):
    """Decoded frames whose byte size is not a multiple of 4 (scalar gray kernel, unaligned staging offsets inside a chunk)."""
    rng = np.random.default_rng(sw + sh)
    frames = rng.integers(0, 256, (7, sh, sw, 3), dtype=np.uint8)
    gray = np.stack([oracle.bgr2gray(f if dsize is None else oracle.resize_u8(f, dsize)) for f in frames])
    eng.set_option("batch", 3)                          # several chunks, frames at odd offsets inside the staging buffer
    try:
        a = eng.shot_bgr(frames, dsize=dsize, want_bgr=True, want_magsum=True, want_gray=True)
        b = eng.shot(gray, want_bgr=True, want_magsum=True)
    finally:
        eng.set_option("batch", 0)
    assert np.array_equal(a["gray"], gray)
    assert np.array_equal(a["bgr"], b["bgr"]) and np.array_equal(a["magsum"], b["magsum"])


@pytest.mark.parametrize("kind", ["zeros", "white", "stripes_x", "stripes_y", "checker", "step_at_tile_edge"])
def test_degenerate_frames_through_the_shot_path(eng, oracle, kind):
    """Flat, saturated and hard-edged frames (edges placed on the 64-column / 16-row tile boundaries of the staging
    kernels): finite output, flat frames give exactly zero flow, everything else stays within the parity tolerance of
    the oracle away from cv2's own chaotic in/out-of-bounds switch (A.8) -- compared on the 3-scale median."""
    W, H = 256, 96
    ys, xs = np.mgrid[0:H, 0:W]
    base = {"zeros": np.zeros((H, W)), "white": np.full((H, W), 255), "stripes_x": ((xs // 8) % 2) * 255,
            "stripes_y": ((ys // 8) % 2) * 255, "checker": (((xs // 16) + (ys // 16)) % 2) * 200 + 20,
            "step_at_tile_edge": np.where(xs < 128, 30, 220) + np.where(ys < 48, 0, 25)}[kind].astype(np.uint8)
    nxt = np.roll(base, (1, 2), (0, 1))
    frames = np.stack([base, nxt, base])
    res = eng.shot(frames, want_bgr=True, want_flow=True, want_magsum=True)
    assert np.isfinite(res["flow"]).all() and np.isfinite(res["magsum"]).all()
    if kind in ("zeros", "white"):
        assert not res["flow"].any() and not res["bgr"].any()
        return
    kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    ref = oracle.farneback(base, nxt, None, **kw)
    d = np.sqrt(((res["flow"][0].astype(np.float64) - ref) ** 2).sum(-1))
    assert np.median(d) <= 1e-4 and d.mean() <= 2e-2, (kind, float(np.median(d)), float(d.mean()), float(d.max()))
    assert np.array_equal(res["bgr"][0], oracle.viz(res["flow"][0], 0))


# ------------------------------------------------------------------------------------------------
# memory safety and races without compute-sanitizer (closed on the B200 pool; profiles/r2b_memcheck.out)
# ------------------------------------------------------------------------------------------------
def test_no_workspace_guard_band_is_overwritten():
    """Every engine workspace carries a 256-byte guard band on both sides (engine.cu guarded_alloc).  After a workload that
    exercises every fast-path kernel family on frames with interior and border tiles (tools/sanitize_case.py's list), no guard
    byte may have changed: an out-of-bounds write next to a buffer -- the kind a halo / tile-edge bug makes -- shows up here."""
    import optical_flow_b200 as ofb
    eng = ofb.Farneback(0)
    rng = np.random.default_rng(0)
    ref = dict(ofb.REFERENCE_PARAMS)
    for (W, H) in ((448, 200), (203, 97), (640, 360), (129, 72), (64, 64), (72, 136)):
        fr = np.stack([_textured(W, H, 40 + t)[0] for t in range(5)])
        for kw in (ref, dict(ref, winsize=9, iterations=2), dict(ref, flags=256, poly_n=7, poly_sigma=1.5), dict(ref, winsize=33)):
            eng.shot(fr, want_bgr=True, want_magsum=True, want_flow=True, **kw)
            assert eng.check_guards() == 0, (W, H, kw)
        eng.pairs(fr[:-1], fr[1:], want_magsum=True, **ref)
        eng.shot_jpeg(fr, **ref)
        f = eng.calc(fr[0], fr[1], None, **ref)
        eng.calc(fr[0], fr[1], f.copy(), **dict(ref, flags=4))
        eng.calc(fr[0].astype(np.float32), fr[1].astype(np.float32), None, **ref)
        assert eng.check_guards() == 0, (W, H)
    bgr = rng.integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
    eng.shot_bgr(bgr, dsize=(129, 96), want_bgr=True, want_magsum=True, want_gray=True)
    assert eng.check_guards() == 0


def test_repeated_runs_are_bitwise_identical(eng):
    """k_iter / k_iter64 reuse shared memory across steps with one barrier fewer than phases, k_polyexp2 aliases its staging
    buffers, the JPEG emit kernel combines words with atomicOr: a race would show as run-to-run differences.  Eight
    repetitions of a shot with many strips and tiles must agree bit for bit, with the software prefetch on and off."""
    f = _textured(640, 360, 77)[0]
    frames = np.stack([np.roll(f, (t, 2 * t), (0, 1)) for t in range(6)])
    ref = None
    for rep in range(8):
        eng.set_option("iter_prefetch", rep % 2)
        try:
            res = eng.shot(frames, want_bgr=True, want_flow=True, want_magsum=True)
            jp = eng.shot_jpeg(frames)
        finally:
            eng.set_option("iter_prefetch", 1)
        cur = (res["flow"].tobytes(), res["bgr"].tobytes(), res["magsum"].tobytes(), jp["jpeg"][:int(jp["sizes"].sum())].tobytes())
        if ref is None:
            ref = cur
        assert cur == ref, rep


def test_cart_to_polar_in_degrees_and_in_place_outputs(eng):
    """cv2.cartToPolar(x, y, angleInDegrees=True) and the optional magnitude / angle destinations (boundary nits of round 1)."""
    cv2 = _cv2_or_none()
    import optical_flow_b200 as ofb
    rng = np.random.default_rng(2)
    x = rng.normal(0, 3, (37, 53)).astype(np.float32)
    y = rng.normal(0, 3, (37, 53)).astype(np.float32)
    mag, ang = ofb.cartToPolar(x, y, angleInDegrees=True)
    m2, a2 = ofb.cartToPolar(x, y)
    assert np.array_equal(mag, m2)
    assert np.allclose(ang, np.degrees(a2.astype(np.float64)), rtol=1e-5, atol=1e-4)          # radians = degrees * (pi/180) in f32
    dst_m, dst_a = np.empty_like(x), np.empty_like(x)
    rm, ra = ofb.cartToPolar(x, y, dst_m, dst_a)
    assert rm is dst_m and ra is dst_a and np.array_equal(dst_m, m2) and np.array_equal(dst_a, a2)
    if cv2 is not None:
        cm, ca = cv2.cartToPolar(x, y, angleInDegrees=True)
        assert np.array_equal(mag, cm) and np.array_equal(ang, ca)
        try:
            ofb.calcOpticalFlowFarneback(np.zeros((8, 8), np.uint8), np.zeros((8, 9), np.uint8), None, 0.5, 3, 15, 3, 5, 1.2, 0)
            assert False, "size mismatch must raise"
        except cv2.error as e:                     # ofb.error IS a cv2.error
            assert "prev0.size() == next0.size()" in str(e)


def test_two_devices_in_one_process():
    """Function attributes (dynamic shared memory above 48 KB) and kernel options are per device / per context: a second engine
    on another GPU of the same process must work and give the same bits (round-1 advisor finding).  Needs two GPUs."""
    import optical_flow_b200 as ofb
    from optical_flow_b200 import _lib
    if _lib.load().ofb_device_count() < 2:
        pytest.skip("one GPU visible")
    f0, f1 = _textured(448, 200, 3)
    kw = dict(ofb.REFERENCE_PARAMS)
    e0, e1 = ofb.Farneback(0), ofb.Farneback(1)
    e1.set_option("fast_arithmetic", 1)                      # options do not leak between contexts
    a = e0.pair(f0, f1, want_bgr=True, want_flow=True, **kw)
    e1.set_option("fast_arithmetic", 0)
    b = e1.pair(f0, f1, want_bgr=True, want_flow=True, **dict(kw, poly_n=7, poly_sigma=1.5))      # another kernel instance first
    b = e1.pair(f0, f1, want_bgr=True, want_flow=True, **kw)
    assert np.array_equal(a["flow"], b["flow"]) and np.array_equal(a["bgr"], b["bgr"])
    assert e0.check_guards() == 0 and e1.check_guards() == 0
