"""Entry-point parity on the GPU box: our optical_flow.py / visualize_optical_flow.py against what the
UNMODIFIED reference scripts wrote for the same synthetic video (tests/golden/make_script_golden.py)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, GOLDEN_DIR

pytestmark = pytest.mark.gpu
G = os.path.join(GOLDEN_DIR, "scripts")


def _run(script, args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    p = subprocess.run([sys.executable, os.path.join(ROOT, script)] + args, cwd=cwd, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()
    return p.stdout.decode()


def test_optical_flow_script_writes_the_reference_csv(tmp_path):
    cv2 = pytest.importorskip("cv2")
    media = tmp_path / "vidA" / "media"
    media.mkdir(parents=True)
    shutil.copy(os.path.join(G, "vidA.mp4"), media / "vidA.mp4")
    _run("optical_flow.py", [str(tmp_path), "vidA"], str(tmp_path))
    out_dir = tmp_path / "vidA" / "opticalflow"
    got = (out_dir / "vidA.csv").read_text()
    want = open(os.path.join(G, "expected_vidA.csv")).read()
    assert (out_dir / ".done").read_text() == open(os.path.join(G, "expected_done.txt")).read()
    g, w = got.split("\t"), want.split("\t")
    assert g[:2] == w[:2]                                    # start / end ms
    gv, wv = np.array(g[2].split(), float), np.array(w[2].split(), float)
    assert gv.shape == wv.shape and np.abs(gv - wv).max() <= 0.011, (got, want)   # 2-decimal rounding boundary
    if got != want:
        print("CSV differs only at a rounding boundary:", got, want)
    # second run: the .done file makes it a no-op unless forced (optical_flow.py:149-168)
    os.remove(out_dir / "vidA.csv")
    log = _run("optical_flow.py", [str(tmp_path), "vidA"], str(tmp_path))
    assert "already done" in log and not (out_dir / "vidA.csv").exists()
    _run("optical_flow.py", [str(tmp_path), "vidA", "--force_run", "True"], str(tmp_path))
    assert (out_dir / "vidA.csv").exists()


def test_visualize_script_writes_the_reference_pictures(tmp_path):
    cv2 = pytest.importorskip("cv2")
    out = tmp_path / "viz"
    _run("visualize_optical_flow.py", [os.path.join(G, "vidA.mp4"), str(out), "0", "1800"], str(tmp_path))
    want_dir = os.path.join(G, "expected_viz")
    assert sorted(os.listdir(out)) == sorted(os.listdir(want_dir))
    for name in sorted(os.listdir(want_dir)):
        a = cv2.imread(str(out / name)).astype(np.int16)
        b = cv2.imread(os.path.join(want_dir, name)).astype(np.int16)
        assert a.shape == b.shape
        d = np.abs(a - b)
        if name.startswith("source_"):
            assert d.max() == 0, name                        # same decoder, same encoder: identical files
        else:
            # pictures agree to +-1 before JPEG; after the lossy encode allow a small spread
            assert d.mean() < 0.5 and (d <= 3).mean() > 0.995, (name, d.mean(), d.max())


def test_pairs_api_equals_single_pair_calls():
    import optical_flow_b200 as ofb
    eng = ofb.Farneback(0)
    rng = np.random.default_rng(3)
    base = rng.random((120, 200)).astype(np.float32)
    for _ in range(3):
        base = (base + np.roll(base, 1, 0) + np.roll(base, -1, 0) + np.roll(base, 1, 1) + np.roll(base, -1, 1)) / 5
    base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
    prev = np.stack([base[8 + i:8 + i + 77, 10:139] for i in range(7)])
    nxt = np.stack([base[9 + i:9 + i + 77, 12 + i:141 + i] for i in range(7)])
    res = eng.pairs(prev, nxt, want_bgr=True, want_magsum=True, want_flow=True)
    for i in range(7):
        one = eng.pair(prev[i], nxt[i], want_bgr=True, want_magsum=True, want_flow=True)
        assert np.array_equal(res["flow"][i], one["flow"]), i
        assert np.array_equal(res["bgr"][i], one["bgr"]), i
        assert abs(float(res["magsum"][i]) - float(one["magsum"])) <= 1e-6 * float(one["magsum"])
