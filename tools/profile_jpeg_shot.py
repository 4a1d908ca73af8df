"""Small host-API shot with JPEG delivery for ncu: `python tools/profile_jpeg_shot.py [pairs] [W] [H]`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optical_flow_b200 as ofb  # noqa: E402
import synth_frames  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 24
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
eng = ofb.Farneback(0)
frames = synth_frames.shot(W, H, P + 1, seed=5)
r = eng.shot_jpeg(frames, **ofb.REFERENCE_PARAMS)
print("pairs %d  %dx%d  %.3f ms/pair  mean JPEG %.0f bytes" % (P, W, H, r["device_ms"] / P, r["sizes"].mean()))
