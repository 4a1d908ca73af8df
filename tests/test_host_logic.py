"""CPU-only tests: the C-ABI library loads and exports every symbol of include/optflow_b200.h, the host
side mirrors cv2's argument contract, scale schedule / byte model / sharding logic."""
import os
import re

import numpy as np
import pytest

import optical_flow_b200 as ofb
from optical_flow_b200 import _lib
from optical_flow_b200.engine import validate_call
from conftest import ROOT, GOLDEN_DIR


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "optflow_b200.h")).read()
    names = sorted(set(re.findall(r"\b(ofb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = _lib.load()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.ofb_abi_version() == 4


def test_no_cpu_fallback_without_a_device():
    L = _lib.load()
    if L.ofb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ofb.Farneback(0)


def test_product_never_imports_the_oracle_or_cv2_compute():
    pkg = os.path.join(ROOT, "optical_flow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src, f
                assert "cv2.calcOpticalFlowFarneback(" not in src.replace("cv2.calcOpticalFlowFarneback(prev", ""), f
    # the entry-point scripts and the tools are not allowed to reach the oracle either: only tests/ (diagnostics included),
    # __graft_entry__.smoke() and bench.py (cpu_baseline / reference arm / the post-timing parity check) do
    for f in ["optical_flow.py", "visualize_optical_flow.py", "synth_frames.py"] + \
             [os.path.join("tools", t) for t in os.listdir(os.path.join(ROOT, "tools")) if t.endswith((".py", ".sh"))]:
        src = open(os.path.join(ROOT, f)).read()
        assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


def test_scale_schedule_matches_oracle(oracle):
    for (W, H, ps, lv) in [(1920, 1080, 0.5, 3), (3840, 2160, 0.5, 5), (129, 77, 0.5, 3), (480, 270, 0.7, 4),
                           (40, 40, 0.5, 3), (640, 360, 0.5, 0), (640, 360, 0.5, -1), (33, 65, 0.5, 3)]:
        a = ofb.scale_schedule(W, H, ps, lv)
        b = [(k, w, h, ks, sg) for (k, w, h, ks, sg, _) in oracle.scale_schedule(W, H, ps, lv)]
        assert a == b


def test_algorithmic_bytes_match_survey_8d():
    mb = lambda **kw: ofb.algorithmic_bytes(with_viz=False, **kw) / 1e6
    assert abs(mb(W=1920, H=1080) - 991.4) < 0.1
    assert abs(mb(W=640, H=360) - 110.2) < 0.1
    assert abs(mb(W=1280, H=720) - 440.6) < 0.1
    assert abs(mb(W=1280, H=720, iterations=10) - 1263.2) < 0.2
    assert abs(mb(W=1920, H=1080, iterations=10) - 2842.1) < 0.2
    assert abs(mb(W=3840, H=2160, levels=5) - 4013.5) < 0.2
    assert abs(ofb.algorithmic_bytes(1920, 1080) / 1e6 - (991.4 + 39.4)) < 0.1


def test_validate_call_errors_mirror_cv2():
    a = np.zeros((40, 50), np.uint8)
    with pytest.raises(ofb.error) as e:
        validate_call(a, np.zeros((40, 51), np.uint8), None, 0.5, 0)
    assert e.value.code == -215 and "prev0.size() == next0.size()" in str(e.value)
    with pytest.raises(ofb.error):
        validate_call(np.zeros((40, 50, 3), np.uint8), np.zeros((40, 50, 3), np.uint8), None, 0.5, 0)
    with pytest.raises(ofb.error):
        validate_call(a, a, None, 1.0, 0)
    with pytest.raises(ofb.error) as e:
        validate_call(a, a, None, 0.5, ofb.OPTFLOW_USE_INITIAL_FLOW)
    assert "_flow0.size() == prev0.size()" in str(e.value)
    with pytest.raises(ofb.error):
        validate_call(a, a, np.zeros((40, 50, 2), np.float64), 0.5, ofb.OPTFLOW_USE_INITIAL_FLOW)


def test_validate_call_flow_ownership_and_depths():
    a = np.zeros((40, 50), np.uint8)
    good = np.zeros((40, 50, 2), np.float32)
    p, n, dt, out = validate_call(a, a, good, 0.5, 0)
    assert out is good and dt == _lib.OFB_U8
    # wrong flow silently ignored without the flag
    for bad in (np.zeros((40, 50, 2), np.float64), np.zeros((10, 10, 2), np.float32), None):
        _, _, _, out = validate_call(a, a, bad, 0.5, 0)
        assert out is not bad and out.shape == (40, 50, 2) and out.dtype == np.float32
    # any depth is accepted and routed through f32; mixed depths too; views are copied
    for dt_in in (np.uint16, np.int16, np.float32, np.float64):
        p, n, dt, _ = validate_call(a.astype(dt_in), a.astype(dt_in), None, 0.5, 0)
        assert dt == _lib.OFB_F32 and p.dtype == np.float32
    p, n, dt, _ = validate_call(a, a.astype(np.float64), None, 0.5, 0)
    assert dt == _lib.OFB_F32 and p.dtype == np.float32 and n.dtype == np.float32
    big = np.zeros((80, 100), np.uint8)
    p, n, dt, _ = validate_call(big[::2, ::2], big[::2, ::2], None, 0.5, 0)
    assert p.flags.c_contiguous and p.shape == (40, 50)
    p, _, _, _ = validate_call(a[..., None], a[..., None], None, 0.5, 0)
    assert p.shape == (40, 50)


def test_shard_pairs_cover_exactly_once():
    for n in (0, 1, 7, 300, 20000):
        for ws in (1, 2, 3, 4, 8):
            got = []
            for r in range(ws):
                s, e = ofb.shard_pairs(n, ws, r)
                assert 0 <= s <= e <= n
                got += list(range(s, e))
            assert got == list(range(n))
            sizes = [ofb.shard_pairs(n, ws, r)[1] - ofb.shard_pairs(n, ws, r)[0] for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1


def test_shard_shots_balanced_and_complete():
    rng = np.random.default_rng(0)
    lengths = rng.integers(50, 401, 80).tolist()
    for ws in (1, 2, 4, 8):
        parts = ofb.shard_shots(lengths, ws)
        seen = {}
        loads = []
        for r in parts:
            loads.append(sum(p[2] for p in r))
            for (i, off, n) in r:
                for t in range(off, off + n):
                    assert (i, t) not in seen
                    seen[(i, t)] = 1
        assert len(seen) == sum(lengths)
        assert max(loads) - min(loads) <= max(lengths)
    parts = ofb.shard_shots([1000], 4)
    assert sorted(p for r in parts for p in r) == [(0, 0, 250), (0, 250, 250), (0, 500, 250), (0, 750, 250)]


# ---- host-side video helpers of the entry-point scripts ------------------------------
def test_frame_reader_returns_the_frames_a_seek_per_frame_returns():
    """Decoding forward to a frame that lies ahead gives the frame `set(CAP_PROP_POS_FRAMES); read()` gives, for the
    access patterns of both scripts (float positions, window pairs, a backward jump) on the golden video."""
    cv2 = pytest.importorskip("cv2")
    from optical_flow_b200.video import FrameReader
    path = os.path.join(GOLDEN_DIR, "scripts", "vidA.mp4")
    patterns = {
        "visualize": [0.0, 7.0, 14.0, 21.0, 28.0, 35.0, 42.0],                 # fps*start/1000 + k*int(fps*0.3)
        "float": [2.5, 9.5, 16.5, 23.5],
        "windows": [0, 3, 4, 10, 11, 17, 18, 24, 25, 31, 32, 38, 39, 45, 46, 47],
        "backward": [10, 20, 5, 6, 30, 30, 31],
        "past_end": [40, 47, 48, 3],
    }
    for name, positions in patterns.items():
        ref = cv2.VideoCapture(path)
        want = []
        for p in positions:
            ref.set(cv2.CAP_PROP_POS_FRAMES, p)
            want.append(ref.read())
        ref.release()
        vid = cv2.VideoCapture(path)
        rd = FrameReader(vid)
        got = [rd.read_at(p) for p in positions]
        vid.release()
        for p, (ok_w, f_w), (ok_g, f_g) in zip(positions, want, got):
            assert ok_w == ok_g, (name, p)
            if ok_w:
                assert np.array_equal(f_w, f_g), (name, p)
        if name in ("visualize", "windows"):
            assert rd.seeks == 1 and rd.grabs > 0, (name, rd.seeks, rd.grabs)      # one seek, then forward decoding


def test_jpeg_writer_writes_the_same_bytes_as_imwrite(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from optical_flow_b200.video import JpegWriter
    rng = np.random.default_rng(0)
    imgs = [cv2.GaussianBlur(rng.integers(0, 256, (90, 120, 3), dtype=np.uint8), (0, 0), 1.5) for _ in range(6)]
    with JpegWriter(workers=3) as w:
        for i, im in enumerate(imgs):
            w.imwrite(str(tmp_path / ("a%d.jpeg" % i)), im)
    for i, im in enumerate(imgs):
        cv2.imwrite(str(tmp_path / ("b%d.jpeg" % i)), im)
        assert (tmp_path / ("a%d.jpeg" % i)).read_bytes() == (tmp_path / ("b%d.jpeg" % i)).read_bytes()


def test_error_is_a_cv2_error_and_argument_errors_need_no_gpu():
    """ofb.error subclasses the installed cv2.error (so `except cv2.error` keeps working); the argument contract of
    calcOpticalFlowFarneback is checked on the host before any GPU call."""
    cv2 = pytest.importorskip("cv2")
    from optical_flow_b200.engine import validate_call, error
    assert issubclass(error, cv2.error)
    with pytest.raises(cv2.error):
        validate_call(np.zeros((8, 8), np.uint8), np.zeros((8, 9), np.uint8), None, 0.5, 0)
    with pytest.raises(cv2.error):
        validate_call(np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint8), None, 1.0, 0)


def test_update_matrices_kernels_hold_no_contracted_packed_fma():
    """The UpdateMatrices kernels use packed f32x2 multiplies; ptxas contracts a packed multiply feeding a packed add into FFMA2
    even under .rn / -fmad=false, which would round once where cv2 rounds twice (csrc/um_device.cuh).  The built library's SASS
    must hold no FFMA2 in k_um0 / k_iter / k_iter64 / k_update_matrices (tools/check_sass.sh; needs cuobjdump, no GPU)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    r = subprocess.run([os.path.join(ROOT, "tools", "check_sass.sh")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_chunk_starts_mirror_the_engine_schedule():
    """Farneback.chunk_starts restates engine.cu's chunk schedule (B/4, B/2, B, ..., B, B/2, B/4 for a long job, plain chunks for a
    short one) from the chunk size the library reports; here with the library call replaced by the documented rule."""
    eng = ofb.Farneback.__new__(ofb.Farneback)          # no context: only the pure-Python schedule is exercised

    def rule(W, H, n_pairs):                             # ofb_shot_chunk's documented default
        b = -(-96_000_000 // (W * H))
        b = max(4, min((b + 3) // 4 * 4, 512))
        if n_pairs < 4 * b:
            q = ((n_pairs + 3) // 4 + 3) // 4 * 4
            b = max(max(4, b // 4), min(q, b))
        return max(1, min(b, n_pairs))

    eng.shot_chunk = rule
    assert rule(1920, 1080, 300) == 48 and rule(3840, 2160, 100) == 12 and rule(129, 72, 4096) == 512
    assert eng.chunk_starts(1920, 1080, 300) == [0, 12, 36, 84, 132, 180, 228, 264, 288]
    assert rule(1920, 1080, 38) == 12 and eng.chunk_starts(1920, 1080, 38) == [0, 12, 24, 36]      # one rank of an 8-way sharded shot
    assert rule(1920, 1080, 1) == 1 and eng.chunk_starts(1920, 1080, 1) == [0]
    for n in (1, 5, 47, 48, 191, 192, 193, 300, 1000):
        st = eng.chunk_starts(1920, 1080, n)
        assert st[0] == 0 and st == sorted(set(st)) and st[-1] < n
        assert max(b - a for a, b in zip(st, st[1:] + [n])) <= rule(1920, 1080, n)
