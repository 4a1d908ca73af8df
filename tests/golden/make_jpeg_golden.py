"""Generates tests/golden/jpeg_cases.npz: small BGR pictures and the bytes cv2.imencode('.jpeg', picture) -- the encoder
behind cv2.imwrite in /root/reference/visualize_optical_flow.py:57-58 -- produces for them on THIS image's cv2
(opencv-python-headless 4.13.0, libjpeg-turbo 3.1.2).  Run from the repo root:  python tests/golden/make_jpeg_golden.py
The fixtures pin oracle/jpeg_oracle.c (CPU tests) and, through it and directly, the GPU encoder (GPU tests)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synth_frames  # noqa: E402
from oracle import cv2_reference  # noqa: E402

rng = np.random.default_rng(2026)
cases = {}


def add(name, img, quality=None):
    args = [] if quality is None else [cv2.IMWRITE_JPEG_QUALITY, quality]
    ok, buf = cv2.imencode(".jpeg", img, args)
    assert ok
    cases[name + "_img"] = img
    cases[name + "_jpg"] = np.frombuffer(buf.tobytes(), np.uint8)
    cases[name + "_q"] = np.int32(95 if quality is None else quality)


f = synth_frames.shot(320, 180, 3, seed=5)
flow = cv2_reference.farneback(f[0], f[1])
add("flowpic_320x180", cv2_reference.viz(flow))                       # the reference's own artefact, cv2 defaults
add("flowpic_320x180_q50", cv2_reference.viz(flow), 50)
add("flowpic_320x180_q100", cv2_reference.viz(flow), 100)
add("noise_40x56", rng.integers(0, 256, (40, 56, 3), dtype=np.uint8))
add("noise_odd_33x17", rng.integers(0, 256, (17, 33, 3), dtype=np.uint8))     # dummy blocks right and bottom, odd chroma
add("noise_129x77", rng.integers(0, 256, (77, 129, 3), dtype=np.uint8))
add("noise_1x1", rng.integers(0, 256, (1, 1, 3), dtype=np.uint8))
add("noise_8x8_q10", rng.integers(0, 256, (8, 8, 3), dtype=np.uint8), 10)
add("saturated_24x24", np.where(rng.random((24, 24, 3)) < 0.5, 0, 255).astype(np.uint8), 100)   # largest coefficients, many 0xFF bytes
add("gray_64x48", np.dstack([f[0][:48, :64]] * 3))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "jpeg_cases.npz"), **cases)
print("wrote", len(cases) // 3, "cases,", sum(v.nbytes for v in cases.values()), "bytes")
