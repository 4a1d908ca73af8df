"""Pins the CPU oracle (oracle/farneback_oracle.c) against golden vectors produced by the
reference's own dependency, cv2 (tests/golden/make_golden.py).  CPU only.

The oracle is far inside the north_star tolerance (mean 1e-3 / max 1e-2 px); the gate here is
two orders tighter so that the oracle can serve as the per-stage checker for the CUDA kernels.
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, epe, GOLDEN_DIR

ORACLE_MEAN_TOL = 1e-5
ORACLE_MAX_TOL = 2e-4


@pytest.mark.parametrize("name", golden_cases())
def test_farneback_flow_matches_cv2_golden(oracle, name):
    g = load_golden(name)
    init = g.get("init_flow")
    flow = oracle.farneback(g["prev"], g["next"], None if init is None else init.copy(), **g["kw"])
    assert flow.shape == g["flow"].shape and flow.dtype == np.float32
    mean, mx = epe(flow, g["flow"])
    assert mean <= ORACLE_MEAN_TOL and mx <= ORACLE_MAX_TOL, (name, mean, mx)


@pytest.mark.parametrize("name", golden_cases())
def test_viz_on_cv2_flow_is_bit_exact_in_hue_and_value(oracle, name):
    g = load_golden(name)
    bgr, hue, val = oracle.viz(g["flow"], 0, return_hv=True)
    assert np.array_equal(hue, g["hue"])
    assert np.array_equal(val, g["val"])
    d = np.abs(bgr.astype(np.int16) - g["bgr"].astype(np.int16))
    assert d.max() <= 1                     # cv2's body truncates, its tail rounds
    assert (d == 0).all(-1).mean() > 0.85


@pytest.mark.parametrize("name", golden_cases())
def test_summed_magnitude(oracle, name):
    g = load_golden(name)
    s = oracle.sum_magnitude(g["flow"])
    assert abs(s - float(g["magsum"])) <= 1e-6 * abs(float(g["magsum"]))


def test_cart_to_polar_matches_golden_hue_path(oracle):
    g = load_golden("ref_320x180")
    mag, ang = oracle.cart_to_polar(g["flow"])
    hue = ((ang * np.float32(180)) / np.float32(np.pi)).astype(np.int64) & 255
    assert np.array_equal(hue.astype(np.uint8), g["hue"])
    assert mag.min() >= 0 and ang.min() >= 0 and ang.max() < 2 * np.pi + 1e-5


def test_hsv2bgr_table_body_and_tail(oracle):
    z = np.load(GOLDEN_DIR + "/hsv2bgr_table.npz")
    assert np.array_equal(oracle.hsv_table(0), z["body"])    # truncating form == cv2's vector body, 65536/65536
    assert np.array_equal(oracle.hsv_table(1), z["tail"])    # rounding form   == cv2's scalar tail, 65536/65536


def test_scale_schedule_matches_survey_probes(oracle):
    s = oracle.scale_schedule(1920, 1080, 0.5, 3)
    assert [(w, h, k) for _, w, h, k, _, _ in s] == [(240, 135, 19), (480, 270, 9), (960, 540, 3), (1920, 1080, 3)]
    s = oracle.scale_schedule(3840, 2160, 0.5, 5)
    assert [(w, h, k) for _, w, h, k, _, _ in s] == [(120, 68, 79), (240, 135, 39), (480, 270, 19), (960, 540, 9),
                                                    (1920, 1080, 3), (3840, 2160, 3)]
    s = oracle.scale_schedule(129, 77, 0.5, 3)
    assert [(w, h) for _, w, h, _, _, _ in s] == [(64, 38), (129, 77)]
    assert len(oracle.scale_schedule(40, 40, 0.5, 3)) == 1


def test_degenerate_parameters_are_finite(oracle):
    g = load_golden("levels0_64x48")
    for kw in (dict(iterations=0), dict(winsize=1), dict(poly_n=1), dict(levels=-2)):
        p = dict(g["kw"]); p.update(kw)
        f = oracle.farneback(g["prev"], g["next"], None, **p)
        assert np.isfinite(f).all()
    p = dict(g["kw"]); p["iterations"] = 0
    assert not oracle.farneback(g["prev"], g["next"], None, **p).any()


def test_oracle_rejects_pyr_scale_ge_1(oracle):
    g = load_golden("levels0_64x48")
    p = dict(g["kw"]); p["pyr_scale"] = 1.0
    with pytest.raises(ValueError):
        oracle.farneback(g["prev"], g["next"], None, **p)


# ---- frame preprocessing: integer algorithms, bit-exact -------------------------------------
def _preprocess_golden():
    z = np.load(GOLDEN_DIR + "/preprocess.npz")
    names = sorted(k[:-len("_src")] for k in z.files if k.endswith("_src"))
    return z, names


@pytest.mark.parametrize("name", _preprocess_golden()[1])
def test_preprocess_oracle_is_bit_exact_against_cv2(oracle, name):
    z, _ = _preprocess_golden()
    src, resized, gray = z[name + "_src"], z[name + "_resized"], z[name + "_gray"]
    dh, dw = resized.shape[:2]
    assert np.array_equal(oracle.resize_u8(src, (dw, dh)), resized)                       # cv2.resize, 3 channels
    assert np.array_equal(oracle.resize_u8(src[..., 1].copy(), (dw, dh)), z[name + "_resized_c1"])   # 1 channel
    assert np.array_equal(oracle.bgr2gray(resized), gray)                                 # cvtColor(BGR2GRAY)
    assert np.array_equal(oracle.bgr2gray(src), z[name + "_gray_fullres"])


def test_bgr2gray_oracle_exhaustive_channels(oracle):
    """Every value of each channel against the closed form, and the weights sum to one (gray of a gray pixel is itself)."""
    v = np.arange(256, dtype=np.uint8)
    for c in range(3):
        px = np.zeros((1, 256, 3), np.uint8)
        px[0, :, c] = v
        w = (3735, 19235, 9798)[c]
        assert np.array_equal(oracle.bgr2gray(px)[0], ((v.astype(np.int64) * w + 16384) >> 15).astype(np.uint8))
    g = np.repeat(v[None, :, None], 3, axis=2)
    assert np.array_equal(oracle.bgr2gray(g)[0], v)


def test_preprocess_oracle_fuzz_against_installed_cv2(oracle):
    """When cv2 is importable (this container, the GPU box): random geometries, up- and down-scaling, 1 and 3 channels."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2024)
    for _ in range(40):
        sw, sh = int(rng.integers(2, 300)), int(rng.integers(2, 200))
        dw, dh = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        cn = int(rng.choice([1, 3]))
        src = rng.integers(0, 256, (sh, sw, cn), dtype=np.uint8)
        if cn == 1:
            src = src[..., 0].copy()
        assert np.array_equal(oracle.resize_u8(src, (dw, dh)), cv2.resize(src, (dw, dh))), (sw, sh, dw, dh, cn)
        if cn == 3:
            assert np.array_equal(oracle.bgr2gray(src), cv2.cvtColor(src, cv2.COLOR_BGR2GRAY))


# ---- the JPEG artefact (visualize_optical_flow.py:57-58): oracle/jpeg_oracle.c against cv2.imencode's bytes ----------------
def _jpeg_cases():
    import os
    z = np.load(os.path.join(GOLDEN_DIR, "jpeg_cases.npz"))
    return z, sorted(k[:-4] for k in z.files if k.endswith("_img"))


@pytest.mark.parametrize("name", _jpeg_cases()[1])
def test_jpeg_oracle_reproduces_cv2_imencode_byte_for_byte(oracle, name):
    z, _ = _jpeg_cases()
    img, ref, q = z[name + "_img"], z[name + "_jpg"], int(z[name + "_q"])
    got = oracle.jpeg_encode(img, q)
    assert got.size == ref.size and np.array_equal(got, ref), (name, got.size, ref.size)
    hdr = oracle.jpeg_header(img.shape[1], img.shape[0], q)
    assert np.array_equal(hdr, ref[:hdr.size])


def test_jpeg_oracle_against_the_installed_cv2_on_fresh_pictures(oracle):
    """When cv2 is importable (it is in this image), fuzz the oracle against it beyond the committed fixtures."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for i in range(12):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        q = int(rng.choice([95, 95, 30, 75, 100]))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if i % 3 == 0:                                        # smooth content: long zero runs, ZRL codes
            img = (np.add.outer(np.arange(h) * 3, np.arange(w) * 2)[..., None] % 256 + np.array([0, 40, 90])).astype(np.uint8)
        ref = cv2.imencode(".jpeg", img, [cv2.IMWRITE_JPEG_QUALITY, q])[1].ravel()
        got = oracle.jpeg_encode(img, q)
        assert np.array_equal(got, ref), (h, w, q)


def test_resize_coordinate_rules_of_cv2(oracle):
    """cv::resize(INTER_LINEAR) has two coordinate rules in the wheel (round-1 advisor finding): the one-channel f32 resize of
    the level images keeps the source coordinate in double, the two-channel resize of the flow field rounds it
    to f32 before the floor.  The oracle (and the engine's tables) use each where cv2 does.  Runs against the installed cv2;
    the wrong rule is two orders of magnitude further away on a non-dyadic scale."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for ws, hs, wd, hd in [(140, 105, 200, 150), (864, 486, 1080, 608), (98, 74, 140, 105)]:
        flow = rng.normal(0, 100, (hs, ws, 2)).astype(np.float32)
        ref = cv2.resize(flow, (wd, hd), interpolation=cv2.INTER_LINEAR)
        right = np.abs(oracle.resize_linear(flow, wd, hd, float_coords=1) - ref).max()
        wrong = np.abs(oracle.resize_linear(flow, wd, hd, float_coords=0) - ref).max()
        assert right <= 1e-4 and wrong > 10 * right, ("flow", ws, wd, right, wrong)        # 1e-4 on values of ~100: 1-2 ulp
        assert np.array_equal(oracle.upsample_flow(flow, wd, hd, 0.7),
                              oracle.resize_linear(flow, wd, hd, float_coords=1) * np.float32(1 / 0.7))
        img = rng.normal(0, 100, (hs, ws)).astype(np.float32)
        ref = cv2.resize(img, (wd, hd), interpolation=cv2.INTER_LINEAR)
        right = np.abs(oracle.resize_linear(img, wd, hd, float_coords=0) - ref).max()
        wrong = np.abs(oracle.resize_linear(img, wd, hd, float_coords=1) - ref).max()
        assert right <= 1e-4 and wrong > 10 * right, ("image", ws, wd, right, wrong)
