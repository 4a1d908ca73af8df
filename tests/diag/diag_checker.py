"""Which stage makes the high-contrast checkerboard (rank-deficient windows) leave the parity tolerance?
GPU vs cv2 on the 1080p stress input for several engine-option sets."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import cv2
import optical_flow_b200 as ofb
import importlib.util
spec = importlib.util.spec_from_file_location("tb", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "test_gpu_benchpath.py"))
tb = importlib.util.module_from_spec(spec); spec.loader.exec_module(tb)
from oracle import c_oracle
c_oracle.build()
eng = ofb.Farneback(0)
eng.set_option("fast_arithmetic", 1)          # the options below are switched on one at a time from the fast baseline
for kind, ws in (("high_contrast_checker", 15), ("high_contrast_checker", 33), ("flat_field_moving_square", 15), ("step_edges", 15)):
    f0, f1 = tb._stress_frames(kind)
    kw = dict(tb.REF, winsize=ws)
    cf, _ = tb._cv2_pair(cv2, f0, f1, kw)
    ref = c_oracle.farneback(f0, f1, None, **kw)
    for name, opts in (("default (f32 van Herk)", {}), ("exact_window_sums", {"exact_window_sums": 1}),
                       ("exact + generic_polyexp", {"generic_polyexp": 1, "exact_window_sums": 1}),
                       ("exact + slow polyexp path", {"polyexp_fast": 0, "exact_window_sums": 1}), ("all_generic", {"generic_kernels": 1})):
        for k, v in opts.items():
            eng.set_option(k, v)
        fl = eng.shot(np.stack([f0, f1, f0]), want_flow=True, want_bgr=False, **kw)["flow"][0]
        for k in opts:
            eng.set_option(k, 0 if k != "polyexp_fast" else 1)
        d = np.sqrt(((fl.astype(np.float64) - cf) ** 2).sum(-1))
        do = np.sqrt(((fl.astype(np.float64) - ref) ** 2).sum(-1))
        print("%-26s %-24s vs cv2 mean %.2e max %.2e n>1e-2 %6d | vs oracle mean %.2e max %.2e" % (kind, name, d.mean(), d.max(), (d > 1e-2).sum(), do.mean(), do.max()))
