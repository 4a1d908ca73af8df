"""Parity of the path bench.py times (run on the B200 box with `-m gpu`): the BATCHED shot entry points at
BASELINE.json's full sizes, compared with cv2 itself (the reference's dependency, importable on the GPU box) at the
pairs where the engine's schedule has a seam -- chunk boundaries of the B/4, B/2, B, ..., B/2, B/4 chunking, the wrap of
the 2*B-slot frame ring, and the shard boundary of a 2-rank split.

Reference loop being matched: /root/reference/visualize_optical_flow.py:21-63 (call at :38-46, picture at :48-55).
Tolerance (BASELINE.json north_star): endpoint difference mean <= 1e-3 px, max <= 1e-2 px; picture +-1 on >= 99.9 %.
"""
import numpy as np
import pytest

from conftest import epe, EPE_MEAN_TOL, EPE_MAX_TOL

pytestmark = pytest.mark.gpu

REF = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


@pytest.fixture(scope="module")
def eng():
    import optical_flow_b200 as ofb
    return ofb.Farneback(0)


@pytest.fixture(scope="module")
def cv2():
    return pytest.importorskip("cv2")


def _cv2_pair(cv2, f0, f1, kw):
    """The reference's own lines: visualize_optical_flow.py:38-46 and :48-55."""
    from oracle import cv2_reference
    flow = cv2.calcOpticalFlowFarneback(f0, f1, None, kw["pyr_scale"], kw["levels"], kw["winsize"], kw["iterations"],
                                        kw["poly_n"], kw["poly_sigma"], kw["flags"])
    return flow, cv2_reference.viz(flow)


def _within1(a, b):
    return float((np.abs(a.astype(np.int16) - b.astype(np.int16)) <= 1).all(-1).mean())


def _check_pairs(cv2, frames, flows, pictures, pairs, kw, tag):
    """Every listed pair against cv2: north_star tolerance on the flow and the picture.  A pair that misses the MAX gate is
    examined with the oracle (the same algorithm as cv2, f64 sums, another implementation): the gate then applies wherever
    cv2 agrees with the oracle, the GPU must agree with the oracle everywhere, and the pixels where cv2 departs from its own
    algorithm's restatement (its A.8 in/out-of-bounds knife edge, profiles/r2b_diag4k.log) must be a handful."""
    worst = (0.0, 0.0, 1.0)
    for t in pairs:
        cf, cb = _cv2_pair(cv2, frames[t], frames[t + 1], kw)
        d = np.sqrt(((flows[t].astype(np.float64) - cf) ** 2).sum(-1))
        mean, mx = float(d.mean()), float(d.max())
        w1 = _within1(pictures[t], cb)
        assert mean <= EPE_MEAN_TOL, (tag, t, mean)
        if mx > EPE_MAX_TOL:
            from oracle import c_oracle
            c_oracle.build()
            ref = c_oracle.farneback(frames[t], frames[t + 1], None, **kw)
            d_ref = np.sqrt(((ref.astype(np.float64) - cf) ** 2).sum(-1))
            d_gpu = np.sqrt(((ref.astype(np.float64) - flows[t]) ** 2).sum(-1))
            n_bad, n_cv2 = int((d > EPE_MAX_TOL).sum()), int((d_ref > EPE_MAX_TOL).sum())
            print("%s pair %d: %d px beyond 1e-2 vs cv2 (max %.2e); cv2 vs oracle has %d such px (max %.2e); GPU vs oracle max %.2e"
                  % (tag, t, n_bad, mx, n_cv2, d_ref.max(), d_gpu.max()))
            assert d_gpu.max() <= EPE_MAX_TOL, (tag, t, float(d_gpu.max()))
            assert d[d_ref <= 1e-3].max() <= EPE_MAX_TOL, (tag, t)
            assert n_bad <= 1e-5 * d.size and n_cv2 >= n_bad // 2, (tag, t, n_bad, n_cv2)
            mx = float(d[d_ref <= 1e-3].max())
        worst = (max(worst[0], mean), max(worst[1], mx), min(worst[2], w1))
        assert w1 >= 0.999, (tag, t, w1)
    print("%s: %d pairs vs cv2 %s: worst mean %.2e max %.2e px, picture within +-1 >= %.5f"
          % (tag, len(pairs), cv2.__version__, worst[0], worst[1], worst[2]))


@pytest.fixture(scope="module")
def shot_1080p():
    import synth_frames
    return synth_frames.shot(1920, 1080, 301, seed=100)         # bench.py's rank-0 shot


def test_1080p_300_pair_shot_at_default_batch_matches_cv2_at_every_seam(eng, cv2, shot_1080p):
    """configs[1] exactly as bench.py runs it: 301 frames, default chunk (48 pairs per launch), host API and device-resident
    API.  Chunks start at pairs 0 | 12 | 36 | 84 | ... | 228 | 264 | 288; the 2B+1-slot frame ring wraps near frame 96.
    Compared with cv2: both sides of every kind of seam, the ring wrap, the middle and the last pair."""
    frames = shot_1080p
    P, H, W = frames.shape[0] - 1, frames.shape[1], frames.shape[2]
    n = W * H
    B = eng.shot_chunk(W, H, P)
    starts = eng.chunk_starts(W, H, P)
    assert starts[0] == 0 and len(starts) >= 5 and starts[1] == B // 4 and P - starts[-1] == B // 4, starts
    seams = starts[1:4] + starts[-2:]
    pairs = sorted(set([0, 2 * B - 1, 2 * B, 2 * B + 1, P // 2, P - 1] + [s - 1 for s in seams] + seams))
    # device-resident entry point (bench.py's device leg): all pictures and all flows stay in HBM
    d_frames = eng.device_alloc(frames.nbytes)
    d_bgr = eng.device_alloc(P * n * 3)
    d_flow = eng.device_alloc(P * n * 8)
    try:
        eng.h2d(d_frames, frames)
        eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, d_flow=d_flow, **REF)
        flows, pics = {}, {}
        for t in pairs:
            flows[t] = np.empty((H, W, 2), np.float32); eng.d2h(flows[t], d_flow + t * n * 8)
            pics[t] = np.empty((H, W, 3), np.uint8); eng.d2h(pics[t], d_bgr + t * n * 3)
        _check_pairs(cv2, frames, flows, pics, pairs, REF, "1080p shot_device chunk %d" % B)
        # host entry point (bench.py's e2e leg): every picture equals the device-resident one, bit for bit
        host = eng.shot(frames, want_bgr=True, **REF)["bgr"]
        all_dev = np.empty((P, H, W, 3), np.uint8)
        eng.d2h(all_dev, d_bgr)
        assert np.array_equal(host, all_dev), "ofb_shot_host and ofb_shot_device disagree"
        # and the single-pair drop-in call gives the same flow as the batched ring-slot path
        for t in (0, pairs[3], pairs[4], P - 1):
            one = eng.calc(frames[t], frames[t + 1], None, **REF)
            assert np.array_equal(one, flows[t]), t
    finally:
        for p in (d_frames, d_bgr, d_flow):
            eng.device_free(p)


def test_1080p_two_rank_shard_seam_matches_cv2_and_the_unsharded_shot(eng, cv2, shot_1080p):
    """a shot cut into contiguous pair ranges with one overlap frame.  The two shards of a 60-pair shot,
    run separately, reproduce the unsharded pictures bit for bit; the pairs either side of the seam match cv2."""
    import optical_flow_b200 as ofb
    frames = shot_1080p[:61]
    whole = eng.shot(frames, want_bgr=True, want_flow=True, **REF)
    parts = []
    for r in range(2):
        s, e = ofb.shard_pairs(60, 2, r)
        parts.append(eng.shot(frames[s:e + 1], want_bgr=True, **REF)["bgr"])
    assert np.array_equal(np.concatenate(parts), whole["bgr"])
    seam = ofb.shard_pairs(60, 2, 0)[1]
    _check_pairs(cv2, frames, whole["flow"], whole["bgr"], [seam - 1, seam], REF, "2-rank shard seam")


def test_4k_gaussian_shot_matches_cv2(eng, cv2):
    """configs[2] through the batched shot path: 3840x2160, levels 5, poly_n 7, sigma 1.5, OPTFLOW_FARNEBACK_GAUSSIAN, chunks of
    6 pairs (the 4K default is 12; 6 puts a chunk seam at pair 6 and a last chunk of one pair inside a 13-pair shot)."""
    import synth_frames
    kw = dict(pyr_scale=0.5, levels=5, winsize=15, iterations=3, poly_n=7, poly_sigma=1.5, flags=256)
    frames = synth_frames.shot(3840, 2160, 14, seed=7)
    assert eng.shot_chunk(3840, 2160, 100) == 12
    eng.set_option("batch", 6)
    try:
        res = eng.shot(frames, want_bgr=True, want_flow=True, **kw)
    finally:
        eng.set_option("batch", 0)
    _check_pairs(cv2, frames, res["flow"], res["bgr"], [0, 5, 6, 12], kw, "4K gaussian shot")


# ------------------------------------------------------------------------------------------------
# adversarial inputs for the f32 van Herk window sums + compensated f32 solve (cv2: f64 running sums, f64 solve)
# ------------------------------------------------------------------------------------------------
def _subpixel_shift(a, dx, dy):
    """Bilinear resampling of a float image by a non-integer shift (replicate border)."""
    H, W = a.shape
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    sx = np.clip(xs + np.float32(dx), 0, W - 1.001); sy = np.clip(ys + np.float32(dy), 0, H - 1.001)
    x0 = np.floor(sx).astype(np.int64); y0 = np.floor(sy).astype(np.int64)
    fx = sx - x0; fy = sy - y0
    return (a[y0, x0] * (1 - fx) * (1 - fy) + a[y0, x0 + 1] * fx * (1 - fy) +
            a[y0 + 1, x0] * (1 - fx) * fy + a[y0 + 1, x0 + 1] * fx * fy)


def _stress_frames(kind, W=1920, H=1080):
    rng = np.random.default_rng(5)
    ys, xs = np.mgrid[0:H, 0:W]
    if kind == "flat_field_moving_square":               # a flat 200 field with one dark 120x120 square, moved by (5, 3)
        a = np.full((H, W), 200.0, np.float32)
        a[400:520, 800:920] = 30
        b = np.full((H, W), 200.0, np.float32)
        b[403:523, 805:925] = 30
        return a.astype(np.uint8), b.astype(np.uint8)
    if kind == "high_contrast_checker":                  # 0 / 255 squares of 24 px, sub-pixel motion
        a = ((((xs // 24) + (ys // 24)) % 2) * 255).astype(np.float32)
        return a.astype(np.uint8), np.clip(_subpixel_shift(a, 1.37, -0.81), 0, 255).astype(np.uint8)
    if kind == "low_texture":                            # grey 128 +- 3 levels of smooth noise: det ~ 1e-3 * w^4 dominates
        t = rng.random((H // 8 + 2, W // 8 + 2)).astype(np.float32)
        t = np.kron(t, np.ones((8, 8), np.float32))[:H, :W]
        for _ in range(6):
            t = (t + np.roll(t, 1, 0) + np.roll(t, -1, 0) + np.roll(t, 1, 1) + np.roll(t, -1, 1)) / 5
        a = 128 + 6 * (t - t.mean()) / (t.max() - t.min())
        return np.round(a).astype(np.uint8), np.round(_subpixel_shift(a, 1.3, -0.7)).astype(np.uint8)
    if kind == "step_edges":                             # vertical and horizontal hard steps on a flat field, sub-pixel motion
        a = (np.where(xs < 700, 20.0, 235.0) + np.where(ys < 500, 0.0, 20.0)).astype(np.float32)
        a[:, 1200:1210] = 255
        return a.astype(np.uint8), np.clip(_subpixel_shift(a, 2.4, 0.6), 0, 255).astype(np.uint8)
    raise ValueError(kind)


def _flow_in_mode(eng, frames, kw, exact):
    eng.set_option("fast_arithmetic", 0 if exact else 1)          # the default is the exact arithmetic
    try:
        return eng.shot(frames, want_flow=True, want_bgr=False, **kw)["flow"][0]
    finally:
        eng.set_option("fast_arithmetic", 0)


def _gate_both_modes(eng, cv2, f0, f1, kw, tag):
    """Both arithmetic modes of the engine against cv2 on one pair.

    default = exact arithmetic (cv2's own running sums and float / double mix): the north_star tolerance on EVERY pixel where
    cv2 agrees with the oracle, its own algorithm in a second implementation (on 256 x 96 frames of pure stripes cv2 and the
    oracle themselves are 6e-2 apart in places).
    fast_arithmetic = 1 (f32 van Herk sums, f64-FMA polynomial expansion): the same tolerance on the mean, and on the max wherever the two
    modes agree to 1e-3 -- they differ only in summation order and rounding, so where they disagree the window is
    rank-deficient and the flow is decided by rounding history (cv2's included); that set must stay a small fraction."""
    from oracle import c_oracle
    c_oracle.build()
    frames = np.stack([f0, f1, f0])
    cf, _ = _cv2_pair(cv2, f0, f1, kw)
    ref = c_oracle.farneback(f0, f1, None, **kw)
    d_ref = np.sqrt(((ref.astype(np.float64) - cf) ** 2).sum(-1))
    fast = _flow_in_mode(eng, frames, kw, False)
    exact = _flow_in_mode(eng, frames, kw, True)
    assert np.isfinite(fast).all() and np.isfinite(exact).all()
    d_fast = np.sqrt(((fast.astype(np.float64) - cf) ** 2).sum(-1))
    d_exact = np.sqrt(((exact.astype(np.float64) - cf) ** 2).sum(-1))
    d_modes = np.sqrt(((fast.astype(np.float64) - exact) ** 2).sum(-1))
    sensitive = d_modes > 1e-3
    print("%-44s default (exact) vs cv2 mean %.2e max %.2e (>1e-2: %d) | fast_arithmetic vs cv2 mean %.2e max %.2e (>1e-2: %d) | oracle vs cv2 max %.2e | "
          "rank-deficient-sensitive px %.4f" % (tag, d_exact.mean(), d_exact.max(), int((d_exact > EPE_MAX_TOL).sum()), d_fast.mean(),
                                                d_fast.max(), int((d_fast > EPE_MAX_TOL).sum()), d_ref.max(), float(sensitive.mean())))
    agree = d_ref <= 1e-3
    assert agree.mean() >= 0.95, tag
    # where cv2 and the oracle themselves are > 1e-2 apart somewhere (integer motion on a periodic pattern: A.8 branch flips, which
    # no arithmetic reproduces), the max gate is the distance cv2 keeps from its own restatement
    max_gate = EPE_MAX_TOL if d_ref.max() <= EPE_MAX_TOL else 5.0 * float(d_ref.max())
    assert d_exact.mean() <= EPE_MEAN_TOL and d_exact[agree].max() <= max_gate, (tag, "exact", float(d_exact.mean()), float(d_exact[agree].max()))
    assert d_fast.mean() <= 2 * EPE_MEAN_TOL, (tag, "fast mean", float(d_fast.mean()))
    ok = agree & ~sensitive
    assert d_fast[ok].max() <= max_gate, (tag, "fast", float(d_fast[ok].max()))
    assert sensitive.mean() <= (0.05 if d_ref.max() <= EPE_MAX_TOL else 0.15), (tag, float(sensitive.mean()))
    return d_fast, d_exact


@pytest.mark.parametrize("winsize", [15, 31, 33])
@pytest.mark.parametrize("kind", ["flat_field_moving_square", "high_contrast_checker", "low_texture", "step_edges"])
def test_1080p_adversarial_inputs_in_both_arithmetic_modes(eng, cv2, kind, winsize):
    """Flat fields, 0/255 edges, a barely textured field, a perfectly periodic checkerboard; winsize 31 / 33 are the longest
    window sums.  What the round-2 diagnosis found (profiles/r2e_diag_*.log): the f32 kernels of round 1 (now option
    fast_arithmetic) are inside the tolerance on all of these EXCEPT where a window is exactly rank-deficient (the replicated
    bottom rows of the checkerboard: 1 % of the pixels, up to 0.2 px), and there the difference is not f32 against f64 -- exact
    f64 sums land equally far -- but cv2's running-sum drift (it adds FLOAT differences to its double column sums).  The
    default arithmetic reproduces cv2's."""
    f0, f1 = _stress_frames(kind)
    d_fast, d_exact = _gate_both_modes(eng, cv2, f0, f1, dict(REF, winsize=winsize), "stress %s winsize %d" % (kind, winsize))
    assert d_exact.max() <= EPE_MAX_TOL, (kind, winsize, float(d_exact.max()))          # the default: every pixel of these inputs
    if kind != "high_contrast_checker":                       # nothing rank-deficient here: the fast path holds the gate too
        assert d_fast.mean() <= EPE_MEAN_TOL and d_fast.max() <= EPE_MAX_TOL, (kind, winsize, float(d_fast.max()))


@pytest.mark.parametrize("kind", ["stripes_x", "stripes_y", "checker", "step_at_tile_edge", "identical"])
def test_degenerate_frames_against_cv2(eng, cv2, kind):
    """Hard-edged 256 x 96 frames with INTEGER motion: windows that hold a single edge direction, pixels exactly on
    UpdateMatrices' in/out-of-bounds switch (A.8).  Same gates as above, in both modes."""
    W, H = 256, 96
    ys, xs = np.mgrid[0:H, 0:W]
    base = {"stripes_x": ((xs // 8) % 2) * 255, "stripes_y": ((ys // 8) % 2) * 255,
            "checker": (((xs // 16) + (ys // 16)) % 2) * 200 + 20,
            "step_at_tile_edge": np.where(xs < 128, 30, 220) + np.where(ys < 48, 0, 25),
            "identical": ((xs * 7 + ys * 13) % 256)}[kind].astype(np.uint8)
    nxt = base.copy() if kind == "identical" else np.roll(base, (1, 2), (0, 1))
    _gate_both_modes(eng, cv2, base, nxt, REF, "degenerate %s" % kind)
    res = eng.shot(np.stack([base, nxt, base]), want_bgr=True, want_flow=True, **REF)
    from oracle import c_oracle
    assert np.array_equal(res["bgr"][0], c_oracle.viz(res["flow"][0], 0))


def test_flat_regions_with_sensor_noise_where_cv2_itself_is_not_reproducible(eng, cv2):
    """Large flat regions and long clean edges plus +-1..2 grey levels of camera noise.  In the flat parts the flow is driven by
    the noise and is chaotic IN cv2: its optimised and its plain build (cv2.setUseOptimized) differ by up to several pixels on
    0.3 % of this frame, and so does the oracle.  Parity can only be stated where cv2 reproduces itself: on those pixels both
    modes of the engine hold the north_star tolerance up to the 99.9th percentile and on the mean; the rest is counted.
    This input is why the exact arithmetic is the default: over the WHOLE frame the default is as far from cv2 (mean 4.7e-4 px,
    summed magnitude 8e-4 relative) as cv2's two builds are from each other (6.4e-4), the f32 window sums five times farther
    (3.1e-3, 6.7e-3): noise-driven flow in flat regions is where summation rounding shows (profiles/r2o_diag_modes_noisy.log)."""
    from oracle import c_oracle
    c_oracle.build()
    rng = np.random.default_rng(12)
    H, W = 1080, 1920
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)

    def scene(dx, dy):
        a = np.full((H, W), 60.0, np.float32)
        a[(xs - dx) > 700] = 190.0                                            # a long vertical edge
        a[((ys - dy) > 300) & ((ys - dy) < 420)] += 40.0                      # a horizontal band
        a[(xs - dx - 300) ** 2 + (ys - dy - 700) ** 2 < 150 ** 2] = 230.0     # a disc
        a[((xs - dx) * 0.6 + (ys - dy)) > 1500] = 20.0                        # a slanted edge
        return a
    f0 = cv2.GaussianBlur(scene(0.0, 0.0), (0, 0), 0.8) + rng.normal(0, 1.2, (H, W)).astype(np.float32)
    f1 = cv2.GaussianBlur(scene(2.3, -1.4), (0, 0), 0.8) + rng.normal(0, 1.2, (H, W)).astype(np.float32)
    f0, f1 = np.clip(np.round(f0), 0, 255).astype(np.uint8), np.clip(np.round(f1), 0, 255).astype(np.uint8)
    cf, _ = _cv2_pair(cv2, f0, f1, REF)
    cv2.setUseOptimized(False)
    try:
        cf_plain, _ = _cv2_pair(cv2, f0, f1, REF)
    finally:
        cv2.setUseOptimized(True)
    ref = c_oracle.farneback(f0, f1, None, **REF)
    dist = lambda a, b: np.sqrt(((a.astype(np.float64) - b) ** 2).sum(-1))
    stable = (dist(cf_plain, cf) <= 1e-3) & (dist(ref, cf) <= 1e-3)
    frames = np.stack([f0, f1, f0])
    for name, exact in (("fast_arithmetic", False), ("default (exact)", True)):
        d = dist(_flow_in_mode(eng, frames, REF, exact), cf)
        print("noisy flat regions, %-16s vs cv2: mean %.2e, on cv2-stable px (%.4f of the frame): mean %.2e p99.9 %.2e max %.2e, beyond 1e-2: %d px | "
              "cv2 plain vs optimised max %.2f px, oracle vs cv2 max %.2f px"
              % (name, d.mean(), stable.mean(), d[stable].mean(), np.quantile(d[stable], 0.999), d[stable].max(), int((d[stable] > EPE_MAX_TOL).sum()),
                 dist(cf_plain, cf).max(), dist(ref, cf).max()))
        assert stable.mean() >= 0.98
        assert d[stable].mean() <= EPE_MEAN_TOL and np.quantile(d[stable], 0.999) <= EPE_MAX_TOL, name
