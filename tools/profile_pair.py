"""Small device-resident shot for ncu: `python tools/profile_pair.py [pairs] [W] [H]` (no host copies in the loop)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optical_flow_b200 as ofb  # noqa: E402
import synth_frames  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
eng = ofb.Farneback(0)
if os.environ.get("OFB_BATCH"):
    eng.set_option("batch", int(os.environ["OFB_BATCH"]))
frames = synth_frames.shot(W, H, P + 1, seed=5)
d_frames = eng.device_alloc(frames.nbytes)
d_bgr = eng.device_alloc(P * W * H * 3)
eng.h2d(d_frames, frames)
ms = eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, **ofb.REFERENCE_PARAMS)
out = np.empty((H, W, 3), np.uint8)
eng.d2h(out, d_bgr + (P - 1) * W * H * 3)
print("pairs %d  %dx%d  %.3f ms/pair  picture mean %.2f" % (P, W, H, ms / P, out.mean()))
