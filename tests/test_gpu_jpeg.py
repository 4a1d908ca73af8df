"""GPU parity of the JPEG delivery (run on the B200 box with `-m gpu`): the flow picture leaves the GPU as the file the
reference writes with cv2.imwrite(.., 'flow_<ms>.jpeg') (/root/reference/visualize_optical_flow.py:57-58).

Bar: BYTE-IDENTICAL to cv2.imencode('.jpeg', picture) (same libjpeg algorithm: integer colour conversion, islow DCT,
quantisation, Annex-K Huffman tables) -- against the committed cv2 fixtures, the CPU oracle, and cv2 itself on the box."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import optical_flow_b200 as ofb
    return ofb.Farneback(0)


def _cases():
    z = np.load(os.path.join(GOLDEN_DIR, "jpeg_cases.npz"))
    return z, sorted(k[:-4] for k in z.files if k.endswith("_img"))


@pytest.mark.parametrize("name", _cases()[1])
def test_jpeg_stream_equals_cv2_golden_bytes(eng, oracle, name):
    z, _ = _cases()
    img, ref, q = z[name + "_img"], z[name + "_jpg"], int(z[name + "_q"])
    coef = eng.stage_jpeg_coefficients(img, q)
    assert np.array_equal(coef, oracle.jpeg_coefficients(img, q)), name          # the integer DCT stage on its own
    got = eng.jpeg_encode(img, q)[0]
    assert got.size == ref.size and np.array_equal(got, ref), (name, got.size, ref.size)


@pytest.mark.parametrize("w,h", [(16, 16), (17, 9), (64, 48), (129, 72), (250, 131), (640, 360), (1920, 1080)])
def test_jpeg_geometries_against_the_oracle_and_cv2(eng, oracle, w, h):
    """Every mix of full and partial MCUs, pictures that span several 4 KB stuffing segments and several emit CTAs; random
    content (long codes, many 0xFF bytes) and smooth content (zero runs, ZRL)."""
    rng = np.random.default_rng(w * 31 + h)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ramp = (np.add.outer(np.arange(h) * 2, np.arange(w))[..., None] // 3 % 256 + np.array([0, 70, 140])).astype(np.uint8)
    sat = np.where(rng.random((h, w, 3)) < 0.5, 0, 255).astype(np.uint8)
    pics = np.stack([noise, ramp, sat])
    for q in (95, 100, 40):
        got = eng.jpeg_encode(pics, q)
        for i in range(3):
            ref = oracle.jpeg_encode(pics[i], q)
            assert got[i].size == ref.size and np.array_equal(got[i], ref), (w, h, q, i, got[i].size, ref.size)
    try:
        import cv2
    except Exception:
        return
    got = eng.jpeg_encode(pics)
    for i in range(3):
        assert np.array_equal(got[i], cv2.imencode(".jpeg", pics[i])[1].ravel()), (w, h, i)


def test_shot_jpeg_delivers_the_files_cv2_would_write(eng):
    """The whole path: frames in, JPEG files out, several chunks (deferred download of the packed streams); every stream is
    the cv2.imencode of the raw picture the same shot delivers, and decodes to a picture within JPEG's own error of it."""
    cv2 = pytest.importorskip("cv2")
    import synth_frames
    frames = synth_frames.shot(448, 200, 12, seed=4)
    raw = eng.shot(frames, want_bgr=True)["bgr"]
    for batch in (0, 3):
        eng.set_option("batch", batch)
        try:
            res = eng.shot_jpeg(frames, want_magsum=True)
        finally:
            eng.set_option("batch", 0)
        assert res["sizes"].shape == (11,) and (res["sizes"] > 600).all()
        for t in range(11):
            stream = res["jpeg"][res["offsets"][t]:res["offsets"][t] + res["sizes"][t]]
            ref = cv2.imencode(".jpeg", raw[t])[1].ravel()
            assert stream.size == ref.size and np.array_equal(stream, ref), (batch, t)
        dec = cv2.imdecode(np.asarray(res["jpeg"][:res["sizes"][0]]), cv2.IMREAD_COLOR)
        assert dec.shape == raw[0].shape and np.abs(dec.astype(int) - raw[0]).mean() < 3.0
    with pytest.raises(ValueError):                      # a too small output buffer is an argument error, not a crash
        eng.shot_jpeg(frames, out=np.empty(1000, np.uint8))


def test_shot_jpeg_1080p_default_batch(eng):
    cv2 = pytest.importorskip("cv2")
    import synth_frames
    frames = synth_frames.shot(1920, 1080, 61, seed=100)
    raw = eng.shot(frames, want_bgr=True)["bgr"]
    res = eng.shot_jpeg(frames)
    for t in (0, 5, 6, 17, 18, 41, 42, 59):
        stream = res["jpeg"][res["offsets"][t]:res["offsets"][t] + res["sizes"][t]]
        assert np.array_equal(stream, cv2.imencode(".jpeg", raw[t])[1].ravel()), t
    print("1080p JPEG streams: mean %.0f bytes (raw picture %d bytes)" % (res["sizes"].mean(), raw[0].nbytes))


def test_frame_list_and_bgr_entry_points_deliver_the_same_files(eng):
    """ofb_shot_host_v / ofb_shot_host_v_jpeg (frames in a decoder's own buffers) and ofb_shot_bgr_host_jpeg (decoded BGR frames in)
    against the contiguous gray entry points: same pictures, same JPEG bytes."""
    cv2 = pytest.importorskip("cv2")
    import synth_frames
    frames = synth_frames.shot(320, 184, 7, seed=9)
    flist = [np.ascontiguousarray(f) for f in frames]
    a = eng.shot(frames, want_bgr=True, want_magsum=True)
    b = eng.shot_frames(flist, want_bgr=True, want_magsum=True)
    assert np.array_equal(a["bgr"], b["bgr"]) and np.array_equal(a["magsum"], b["magsum"])
    ja, jb = eng.shot_jpeg(frames), eng.shot_frames_jpeg(flist)
    assert np.array_equal(ja["sizes"], jb["sizes"])
    n = int(ja["sizes"].sum())
    assert np.array_equal(ja["jpeg"][:n], jb["jpeg"][:n])
    bgr = np.stack([cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in frames])      # B = G = R: the gray conversion gives the frame back
    jc = eng.shot_bgr_jpeg(bgr)
    for t in range(6):
        assert np.array_equal(jc["files"][t], cv2.imencode(".jpeg", a["bgr"][t])[1].ravel()), t
