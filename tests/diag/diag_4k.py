"""Diagnostic for a parity outlier of the 4K Gaussian shot (tests/test_gpu_benchpath.py): where are the pixels whose
endpoint difference to cv2 exceeds 1e-2 px, what do cv2 / the oracle / the single-pair call give there."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cv2
import optical_flow_b200 as ofb
import synth_frames
from oracle import c_oracle
c_oracle.build()
kw = dict(pyr_scale=0.5, levels=5, winsize=15, iterations=3, poly_n=7, poly_sigma=1.5, flags=256)
frames = synth_frames.shot(3840, 2160, 14, seed=7)
eng = ofb.Farneback(0)
res = eng.shot(frames, want_flow=True, want_bgr=False, **kw)
for t in (6, 12, 3):
    cf = cv2.calcOpticalFlowFarneback(frames[t], frames[t + 1], None, 0.5, 5, 15, 3, 7, 1.5, 256)
    one = eng.calc(frames[t], frames[t + 1], None, **kw)
    d = np.sqrt(((res["flow"][t].astype(np.float64) - cf) ** 2).sum(-1))
    ys, xs = np.nonzero(d > 1e-2)
    print("pair", t, "shot==calc bitwise:", np.array_equal(one, res["flow"][t]), "mean %.2e max %.2e n_bad %d" % (d.mean(), d.max(), len(ys)))
    if len(ys):
        print("  bad pixel bbox x %d..%d y %d..%d" % (xs.min(), xs.max(), ys.min(), ys.max()))
        o = c_oracle.farneback(frames[t], frames[t + 1], None, **kw)
        do = np.sqrt(((o.astype(np.float64) - cf) ** 2).sum(-1))
        dg = np.sqrt(((o.astype(np.float64) - res["flow"][t]) ** 2).sum(-1))
        print("  oracle vs cv2: mean %.2e max %.2e n_bad %d ; oracle vs GPU: max %.2e n_bad %d" % (do.mean(), do.max(), int((do > 1e-2).sum()), dg.max(), int((dg > 1e-2).sum())))
        y, x = np.unravel_index(np.argmax(d), d.shape)
        print("  worst at (x=%d, y=%d): cv2 %s gpu %s oracle %s" % (x, y, cf[y, x], res["flow"][t][y, x], o[y, x]))
        print("  cv2 flow dx along x=0..2 at that row:", cf[y, 0:3, 0], " dy along bottom rows at that col:", cf[-3:, x, 1])
