"""Generates tests/golden/*.npz from the reference's own dependency (cv2), run in the build container.

    python tests/golden/make_golden.py

Each farneback_*.npz holds the two uint8 frames, the parameters, and what the reference's lines
produce for them through cv2 (version recorded in the file):
    flow  = cv2.calcOpticalFlowFarneback(...)            optical_flow.py:51-59 / visualize_optical_flow.py:38-46
    bgr,hue,val = the HSV picture                           visualize_optical_flow.py:48-55
    magsum = np.sum(cartToPolar(...)[0])                    optical_flow.py:61-64
hsv2bgr_table.npz pins cvtColor(COLOR_HSV2BGR) at S=255 over the whole (H,V) domain, separately for
the vectorised body (wide image) and the scalar tail (3-pixel-wide image) -- cv2 rounds them differently
(SURVEY.md B.6).
The GPU box has no /root/reference and is not guaranteed to have cv2, so these files are committed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402
from oracle import cv2_reference as R, synth  # noqa: E402

CASES = {
    # name: (W, H, seed, overrides of the reference's parameters)
    "ref_320x180": (320, 180, 11, {}),
    "gauss_poly7_192x128": (192, 128, 12, dict(levels=5, poly_n=7, poly_sigma=1.5, flags=256)),
    "featurepath_129x77": (129, 77, 13, {}),
    "pyr07_win9_it10_200x150": (200, 150, 14, dict(pyr_scale=0.7, levels=4, winsize=9, iterations=10)),
    "win31_160x120": (160, 120, 15, dict(winsize=31)),
    "evenwin16_gauss_160x120": (160, 120, 16, dict(winsize=16, flags=256)),
    "initflow_160x120": (160, 120, 17, dict(flags=4)),
    "levels0_64x48": (64, 48, 18, dict(levels=0)),
    "polysigma0_n3_it1_160x120": (160, 120, 19, dict(poly_sigma=0.0, poly_n=3, iterations=1)),
}


# name: (source W, H, destination W, H, seed).  Destination heights follow optical_flow.py:27-29
# (int(frame_width / (w / h))) where the case mimics --frame_width 129.
PREPROCESS_CASES = {
    "down_320x180_to_129x72": (320, 180, 129, 72, 31),        # the feature path's default regime
    "down_odd_203x151_to_129x95": (203, 151, 129, 95, 32),
    "half_128x96_to_64x48": (128, 96, 64, 48, 33),            # exact ratio 2
    "up_100x60_to_129x77": (100, 60, 129, 77, 34),            # frame narrower than --frame_width
    "same_129x72": (129, 72, 129, 72, 35),
    "tiny_5x4_to_17x9": (5, 4, 17, 9, 36),
}


def main():
    for name, (W, H, seed, kw) in CASES.items():
        prev, nxt = synth.pair(W, H, seed)
        prm = dict(R.REFERENCE_PARAMS)
        prm.update(kw)
        init = None
        if prm["flags"] & 4:
            base = R.farneback(prev, nxt)
            init = (base + np.random.default_rng(seed).normal(0, 0.3, base.shape)).astype(np.float32)
        flow = R.farneback(prev, nxt, None if init is None else init.copy(), **kw)
        bgr, hue, val = R.viz(flow, return_hv=True)
        magsum = np.float32(R.summed_magnitude(flow))
        out = dict(prev=prev, next=nxt, flow=flow, bgr=bgr, hue=hue, val=val, magsum=magsum,
                   params=np.array([prm["pyr_scale"], prm["levels"], prm["winsize"], prm["iterations"],
                                    prm["poly_n"], prm["poly_sigma"], prm["flags"]], np.float64),
                   cv2_version=np.array(cv2.__version__))
        if init is not None:
            out["init_flow"] = init
        np.savez_compressed(os.path.join(HERE, "farneback_%s.npz" % name), **out)
        print(name, flow.shape, "max|flow| %.3f" % np.abs(flow).max())

    hv = np.zeros((256, 256, 3), np.uint8)
    hv[..., 0] = np.arange(256)[:, None]
    hv[..., 1] = 255
    hv[..., 2] = np.arange(256)[None, :]
    body = cv2.cvtColor(np.tile(hv, (1, 16, 1)), cv2.COLOR_HSV2BGR)[:, :256].copy()
    tail = np.empty_like(hv)
    for v0 in range(0, 256, 2):  # 2-pixel-wide strips never reach the vector body
        tail[:, v0:v0 + 2] = cv2.cvtColor(hv[:, v0:v0 + 2].copy(), cv2.COLOR_HSV2BGR)
    np.savez_compressed(os.path.join(HERE, "hsv2bgr_table.npz"), body=body, tail=tail,
                        cv2_version=np.array(cv2.__version__))
    print("hsv table: body/tail differ on %.1f%% of (H,V)" % (100 * (body != tail).any(-1).mean()))

    # frame preprocessing either side of the hot path: cv2.resize (default INTER_LINEAR) of
    # the decoded BGR frame, then cvtColor(BGR2GRAY) -- optical_flow.py:25-31, :42-44; visualize_optical_flow.py:31,35
    pre = {}
    for name, (sw, sh, dw, dh, seed) in PREPROCESS_CASES.items():
        rng = np.random.default_rng(seed)
        noise = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        smooth = cv2.GaussianBlur(noise, (0, 0), 2.5)
        src = np.where(rng.random((sh, sw, 1)) < 0.5, noise, smooth).astype(np.uint8)     # hard and soft texture mixed
        resized = cv2.resize(src, (dw, dh))
        pre[name + "_src"] = src
        pre[name + "_resized"] = resized
        pre[name + "_gray"] = cv2.cvtColor(resized, cv2.COLOR_BGR2GRAY)
        pre[name + "_gray_fullres"] = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
        pre[name + "_resized_c1"] = cv2.resize(src[..., 1].copy(), (dw, dh))
        print("preprocess", name, src.shape, "->", resized.shape)
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), cv2_version=np.array(cv2.__version__), **pre)


if __name__ == "__main__":
    main()
