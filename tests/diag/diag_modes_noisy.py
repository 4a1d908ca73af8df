"""Which part of exact_arithmetic matters on flat regions with sensor noise (tests/test_gpu_benchpath.py's scene)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2
import optical_flow_b200 as ofb
rng = np.random.default_rng(12)
H, W = 1080, 1920
ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
def scene(dx, dy):
    a = np.full((H, W), 60.0, np.float32)
    a[(xs - dx) > 700] = 190.0
    a[((ys - dy) > 300) & ((ys - dy) < 420)] += 40.0
    a[(xs - dx - 300) ** 2 + (ys - dy - 700) ** 2 < 150 ** 2] = 230.0
    a[((xs - dx) * 0.6 + (ys - dy)) > 1500] = 20.0
    return a
f0 = cv2.GaussianBlur(scene(0.0, 0.0), (0, 0), 0.8) + rng.normal(0, 1.2, (H, W)).astype(np.float32)
f1 = cv2.GaussianBlur(scene(2.3, -1.4), (0, 0), 0.8) + rng.normal(0, 1.2, (H, W)).astype(np.float32)
f0, f1 = np.clip(np.round(f0), 0, 255).astype(np.uint8), np.clip(np.round(f1), 0, 255).astype(np.uint8)
cf = cv2.calcOpticalFlowFarneback(f0, f1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
cv2.setUseOptimized(False); cp = cv2.calcOpticalFlowFarneback(f0, f1, None, 0.5, 3, 15, 3, 5, 1.2, 0); cv2.setUseOptimized(True)
dist = lambda a, b: np.sqrt(((a.astype(np.float64) - b) ** 2).sum(-1))
print("cv2 plain vs optimised: mean %.2e  sum|flow| rel diff %.2e" % (dist(cp, cf).mean(), abs(np.sqrt((cp**2).sum(-1)).sum() / np.sqrt((cf**2).sum(-1)).sum() - 1)))
eng = ofb.Farneback(0)
eng.set_option("fast_arithmetic", 1)          # the options below are switched on one at a time from the fast baseline
for name, opts in (("default", {}), ("exact_window_sums", {"exact_window_sums": 1}), ("polyexp_exact", {"polyexp_exact": 1}), ("exact_arithmetic", {"exact_arithmetic": 1})):
    for k, v in opts.items(): eng.set_option(k, v)
    fl = eng.calc(f0, f1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    for k in opts: eng.set_option(k, 0)
    d = dist(fl, cf)
    print("%-20s vs cv2: mean %.2e median %.2e p99 %.2e p99.9 %.2e max %.2e  n>1e-2 %6d  sum|flow| rel diff %.2e"
          % (name, d.mean(), np.median(d), np.quantile(d, 0.99), np.quantile(d, 0.999), d.max(), (d > 1e-2).sum(),
             abs(np.sqrt((fl**2).sum(-1)).sum() / np.sqrt((cf**2).sum(-1)).sum() - 1)))
