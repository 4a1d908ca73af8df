"""Stage-swap diagnosis on the GPU box: the CPU oracle's per-scale loop with ONE stage's output taken from the GPU engine.
Tells which stage's (tolerated) per-stage difference is amplified into the end-to-end deviation on rank-deficient input."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2
import optical_flow_b200 as ofb
from oracle import c_oracle as orc
import importlib.util
spec = importlib.util.spec_from_file_location("tb", os.path.join(ROOT, "tests", "test_gpu_benchpath.py"))
tb = importlib.util.module_from_spec(spec); spec.loader.exec_module(tb)
orc.build()
eng = ofb.Farneback(0)
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 544)
kind = sys.argv[3] if len(sys.argv) > 3 else "high_contrast_checker"
f0, f1 = tb._stress_frames(kind, W, H)
kw = dict(tb.REF)


def pipeline(level_fn, poly_fn, um_fn, solve_fn, up_fn):
    flow = None
    for (k, wk, hk, ks, sg, sc) in orc.scale_schedule(W, H, 0.5, 3):
        I0, I1 = level_fn(f0, k, wk, hk, ks, sg), level_fn(f1, k, wk, hk, ks, sg)
        R0, R1 = poly_fn(I0), poly_fn(I1)
        flow = np.zeros((hk, wk, 2), np.float32) if flow is None else up_fn(flow, wk, hk)
        M = um_fn(R0, R1, flow)
        for it in range(3):
            flow = solve_fn(M)
            if it < 2:
                M = um_fn(R0, R1, flow)
    return flow


o_level = lambda f, k, wk, hk, ks, sg: orc.level_image(f, wk, hk, ks, sg)
g_level = lambda f, k, wk, hk, ks, sg: eng.stage_level_image(f, 0.5, k)
o_poly = lambda I: orc.polyexp(I, 5, 1.2)
g_poly = lambda I: eng.stage_polyexp(I, 5, 1.2)
o_um, g_um = orc.update_matrices, eng.stage_update_matrices
o_solve = lambda M: orc.blur_solve(M, 15)
g_solve = lambda M: eng.stage_blur_solve(M, 15)
o_up = lambda fl, wk, hk: orc.upsample_flow(fl, wk, hk, 0.5)
g_up = lambda fl, wk, hk: eng.stage_upsample_flow(fl, wk, hk, 0.5)

cf = cv2.calcOpticalFlowFarneback(f0, f1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
full = orc.farneback(f0, f1, None, **kw)
gpu = eng.calc(f0, f1, None, **kw)
def rep(name, fl):
    d = np.sqrt(((fl.astype(np.float64) - cf) ** 2).sum(-1))
    print("%-34s vs cv2 mean %.2e max %.2e n>1e-2 %6d" % (name, d.mean(), d.max(), (d > 1e-2).sum()))
print(kind, W, H)
rep("oracle.farneback", full)
rep("GPU calc", gpu)
rep("oracle loop, all oracle stages", pipeline(o_level, o_poly, o_um, o_solve, o_up))
rep("oracle loop + GPU level images", pipeline(g_level, o_poly, o_um, o_solve, o_up))
rep("oracle loop + GPU polyexp", pipeline(o_level, g_poly, o_um, o_solve, o_up))
rep("oracle loop + GPU UpdateMatrices", pipeline(o_level, o_poly, g_um, o_solve, o_up))
rep("oracle loop + GPU blur+solve", pipeline(o_level, o_poly, o_um, g_solve, o_up))
rep("oracle loop + GPU upsample", pipeline(o_level, o_poly, o_um, o_solve, g_up))
rep("oracle loop, all GPU stages", pipeline(g_level, g_poly, g_um, g_solve, g_up))
for (k, wk, hk, ks, sg, sc) in orc.scale_schedule(W, H, 0.5, 3):
    a, b = o_level(f0, k, wk, hk, ks, sg), g_level(f0, k, wk, hk, ks, sg)
    print("level %d image GPU vs oracle: max abs %.3e (0..255 scale)" % (k, np.abs(a - b).max()))
d = np.sqrt(((gpu.astype(np.float64) - cf) ** 2).sum(-1))
ys, xs = np.nonzero(d > 1e-2)
if len(ys):
    print("bad pixels: x %d..%d y %d..%d ; within 40 px of the border: %.3f" % (xs.min(), xs.max(), ys.min(), ys.max(),
          float(((xs < 40) | (xs >= W - 40) | (ys < 40) | (ys >= H - 40)).mean())))
