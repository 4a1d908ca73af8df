"""Parameter sweep of BASELINE.json configs[2] and configs[4] on one B200, with cv2 on the host cores beside it.

    python tools/sweep.py [--quick] > profiles/r1_sweep.json

For every (size, winsize, iterations, flags, levels, poly_n, poly_sigma): device-resident pairs/s of a short shot
(CUDA events), endpoint difference vs cv2 on one pair, the algorithmic-byte roofline fraction, and cv2's pairs/s
with one single-threaded process per host core.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optical_flow_b200 as ofb  # noqa: E402
import synth_frames  # noqa: E402


def _init():
    sys.path.insert(0, ROOT)
    import cv2
    cv2.setNumThreads(1)


def _cv2_pair(args):
    prev, nxt, prm = args
    import cv2
    f = cv2.calcOpticalFlowFarneback(prev, nxt, None, prm["pyr_scale"], prm["levels"], prm["winsize"], prm["iterations"],
                                     prm["poly_n"], prm["poly_sigma"], prm["flags"])
    # the four picture lines of visualize_optical_flow.py:48-55, as bench.py's cpu_baseline leg runs them
    import numpy as np
    mag, ang = cv2.cartToPolar(f[..., 0], f[..., 1])
    hsv = np.zeros(prev.shape + (3,), np.uint8)
    hsv[..., 1] = 255
    hsv[..., 0] = ang * 180 / np.pi
    hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    return int(cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)[0, 0, 0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import cv2
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    base = dict(ofb.REFERENCE_PARAMS)
    sizes = [(1280, 720), (1920, 1080), (3840, 2160)]
    cfgs = []
    for (W, H) in sizes:
        for ws in (9, 15, 31):
            for it in (3, 10):
                cfgs.append((W, H, dict(base, winsize=ws, iterations=it)))
    cfgs.append((3840, 2160, dict(base, levels=5, poly_n=7, poly_sigma=1.5, flags=256)))     # configs[2]
    cfgs.append((1920, 1080, dict(base, flags=256)))
    cfgs.append((640, 360, dict(base)))
    cfgs.append((129, 72, dict(base)))                                                        # --frame_width 129 regime
    if args.quick:
        cfgs = cfgs[:2] + cfgs[-4:]
    eng = ofb.Farneback(0)
    cores = os.cpu_count() or 1
    pool = mp.get_context("spawn").Pool(cores, initializer=_init)
    out = []
    for (W, H, prm) in cfgs:
        n = W * H
        P = max(8, min(64, int(64e6 // n)))
        frames = synth_frames.shot(W, H, P + 1, seed=W + prm["winsize"])
        d_frames = eng.device_alloc(frames.nbytes)
        d_bgr = eng.device_alloc(P * n * 3)
        eng.h2d(d_frames, frames)
        eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, **prm)
        ms = min(eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, **prm) for _ in range(3))
        rate = P / (ms / 1e3)
        flow = eng.calc(frames[0], frames[1], None, **prm)
        cf = cv2.calcOpticalFlowFarneback(frames[0], frames[1], None, prm["pyr_scale"], prm["levels"], prm["winsize"],
                                          prm["iterations"], prm["poly_n"], prm["poly_sigma"], prm["flags"])
        d = np.sqrt(((flow.astype(np.float64) - cf) ** 2).sum(-1))
        ntask = cores if n >= 1920 * 1080 else 2 * cores
        tasks = [(frames[i % P], frames[i % P + 1], prm) for i in range(ntask)]
        pool.map(_cv2_pair, tasks[:cores])
        t0 = time.perf_counter()
        pool.map(_cv2_pair, tasks, chunksize=1)
        cpu_rate = ntask / (time.perf_counter() - t0)
        alg = ofb.algorithmic_bytes(W, H, with_viz=True, **prm)
        rec = {"size": "%dx%d" % (W, H), "params": prm, "pairs": P, "gpu_pairs_per_s": round(rate, 1),
               "alg_MB_per_pair": round(alg / 1e6, 1), "roofline_frac": round(rate * alg / 1e9 / peak, 3),
               "epe_vs_cv2_mean": float(d.mean()), "epe_vs_cv2_max": float(d.max()),
               "cv2_pairs_per_s": round(cpu_rate, 2), "cv2_cores": cores, "speedup": round(rate / cpu_rate, 1)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
        eng.device_free(d_frames); eng.device_free(d_bgr)
    pool.close()


if __name__ == "__main__":
    main()
