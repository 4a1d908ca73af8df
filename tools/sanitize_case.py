"""Small workload for compute-sanitizer (memcheck / racecheck): every fast-path kernel family once, on frames that have
interior AND border tiles.  Run as:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py

  k_polyexp2<5|7, 0|1|2>  interior + border tiles, aligned and unaligned widths, f32 frames; option polyexp_tma
  k_pyr_*                 column-first and row-first pyramid passes
  k_um0<0,1,2>            zero flow, OPTFLOW_USE_INITIAL_FLOW, up-sampled flow
  k_iter<M, fused|last>   box window (winsize 15, 9), Gaussian window (flags 256), min/max folded into the last launch
  viz / preprocess        picture, magnitude sum, BGR->gray, u8 resize
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optical_flow_b200 as ofb  # noqa: E402

rng = np.random.default_rng(0)


def textured(n, H, W):
    a = rng.random((H + 2 * n + 8, W + 2 * n + 8)).astype(np.float32)
    for _ in range(3):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5
    a = ((a - a.min()) / (a.max() - a.min()) * 255).astype(np.uint8)
    return np.stack([a[4 + t:4 + t + H, 2 * t:2 * t + W] for t in range(n)])


eng = ofb.Farneback(0)
ref = dict(ofb.REFERENCE_PARAMS)
for (W, H) in ((448, 200), (203, 97)):                       # aligned (vector paths) and unaligned (scalar paths)
    fr = textured(4, H, W)
    r = eng.shot(fr, want_bgr=True, want_magsum=True, want_flow=True, **ref)                       # box, fused + last, min/max folded
    eng.shot(fr, want_bgr=True, **dict(ref, winsize=9, iterations=2))
    eng.shot(fr, want_bgr=True, **dict(ref, flags=256, poly_n=7, poly_sigma=1.5))                  # Gaussian window, poly_n 7
    eng.pairs(fr[:-1], fr[1:], want_magsum=True, **ref)                                             # independent pairs (slot step 2)
    f = eng.calc(fr[0], fr[1], None, **ref)
    eng.calc(fr[0], fr[1], f.copy(), **dict(ref, flags=4))                                          # k_um0<1> + area-resized initial flow
    eng.calc(fr[0].astype(np.float32), fr[1].astype(np.float32), None, **ref)                       # SRC 2 (f32 frames)
    eng.flow_to_bgr(f); eng.sum_magnitude(f); eng.cart_to_polar(f)
eng.set_option("polyexp_tma", 1)
eng.shot(textured(4, 200, 448), want_bgr=True, **ref)
eng.set_option("polyexp_tma", 0)
bgr = rng.integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
eng.shot_bgr(bgr, dsize=(129, 96), want_bgr=True, want_magsum=True, want_gray=True)
eng.shot_bgr(bgr, want_bgr=True)
if hasattr(eng, "shot_jpeg"):
    eng.shot_jpeg(textured(4, 200, 448), **ref)
    eng.shot_jpeg(textured(3, 97, 203), **ref)
eng.synchronize()
print("sanitize_case: done, kernels:", ", ".join(sorted(eng.kernel_stats())))
