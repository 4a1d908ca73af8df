"""Generates tests/golden/scripts/* by running the UNMODIFIED reference scripts (from /root/reference, through
runpy) on a small synthetic video, in the build container:

    python tests/golden/make_script_golden.py

  vidA.mp4                    the synthetic video (mp4v, 320x240, 25 fps, 48 frames, known affine motion)
  expected_vidA.csv / expected_done.txt      what /root/reference/optical_flow.py writes for it (defaults)
  expected_viz/flow_<ms>.jpeg, source_<ms>.jpeg   what /root/reference/visualize_optical_flow.py writes for 0..1800 ms

One stub is needed to run optical_flow.py with the headless cv2 wheel: cv2.destroyAllWindows (optical_flow.py:104)
raises there, so it is replaced by a no-op for the run.
"""
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "scripts")
REF = "/root/reference"

import cv2  # noqa: E402
import synth_frames  # noqa: E402


def make_video(path, W=320, H=240, n=48, fps=25.0):
    gray = synth_frames.shot(W, H, n, seed=77, step=synth_frames.affine_step(0.4, 1.002, 1.8, -1.2, (W + 128) / 2, (H + 128) / 2))
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (W, H))
    assert vw.isOpened()
    for t in range(n):
        g = gray[t].astype(np.float32)
        bgr = np.stack([g * 0.9, g, np.clip(g * 1.05, 0, 255)], -1).astype(np.uint8)
        vw.write(bgr)
    vw.release()


def run_script(script, argv):
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [script] + argv
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)


def main():
    os.makedirs(OUT, exist_ok=True)
    video = os.path.join(OUT, "vidA.mp4")
    make_video(video)
    with tempfile.TemporaryDirectory() as tmp:
        media = os.path.join(tmp, "vidA", "media")
        os.makedirs(media)
        shutil.copy(video, os.path.join(media, "vidA.mp4"))
        cv2.destroyAllWindows = lambda: None                      # headless wheel: the one stub
        run_script(os.path.join(REF, "optical_flow.py"), [tmp, "vidA"])
        shutil.copy(os.path.join(tmp, "vidA", "opticalflow", "vidA.csv"), os.path.join(OUT, "expected_vidA.csv"))
        shutil.copy(os.path.join(tmp, "vidA", "opticalflow", ".done"), os.path.join(OUT, "expected_done.txt"))
        viz = os.path.join(tmp, "viz")
        run_script(os.path.join(REF, "visualize_optical_flow.py"), [video, viz, "0", "1800"])
        dst = os.path.join(OUT, "expected_viz")
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(viz, dst)
    print(open(os.path.join(OUT, "expected_vidA.csv")).read())
    print(sorted(os.listdir(os.path.join(OUT, "expected_viz"))))


if __name__ == "__main__":
    main()
