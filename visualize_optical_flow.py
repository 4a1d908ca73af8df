#!/usr/bin/env python
"""visualize_optical_flow.py -- B200 version of the reference's shot visualiser, same command line:

    python visualize_optical_flow.py <video_dir> <images_path> <shot_begin ms> <shot_end ms>

It writes exactly the files the reference writes (/root/reference/visualize_optical_flow.py:9-63): for every
sampled frame after the first, `flow_<ms>.jpeg` (the HSV-coded Farneback flow against the previous sampled frame)
and `source_<ms>.jpeg` (the frame itself), sampling every 300 ms between shot_begin and shot_end.

What differs is only WHERE the hot path runs: the reference calls cv2.calcOpticalFlowFarneback + cartToPolar +
normalize + cvtColor once per pair in a Python loop (:37-55); here all sampled frames of the shot are collected
first and handed to the GPU engine as ONE shot (optical_flow_b200.Farneback.shot -> C-ABI ofb_shot_host), which
expands each frame once, batches the pairs, and returns the BGR pictures.  Decoding (cv2.VideoCapture), the
BGR->gray conversion and JPEG encoding stay on the host, as in the reference.
"""
import argparse
import os

import cv2
import numpy as np

import optical_flow_b200 as ofb

STEP_SIZE = 300     # ms between sampled frames (reference: module constant of the same name)


def sample_shot(v_path, start_ms, end_ms):
    """Frames the reference's loop would visit (visualize_optical_flow.py:14-27, :63): positions start at the
    *float* fps*start_ms/1000 and advance by int(fps*STEP_SIZE/1000); stops at the first unreadable frame.
    Returns (positions, BGR frames, fps)."""
    vid = cv2.VideoCapture(v_path)
    fps = vid.get(cv2.CAP_PROP_FPS)
    pos = fps * start_ms / 1000
    last = int(fps * end_ms / 1000)
    stride = int(fps * STEP_SIZE / 1000)
    positions, frames = [], []
    while pos < last:
        vid.set(cv2.CAP_PROP_POS_FRAMES, pos)
        ok, bgr = vid.read()
        if not ok:
            break
        positions.append(pos)
        frames.append(bgr)
        if stride <= 0:          # the reference would spin forever on this input; one frame is all it can mean
            break
        pos += stride
    vid.release()
    return positions, frames, fps


def get_optical_flow(v_path, images_path, start_ms, end_ms, engine=None):
    """Same signature and artefacts as the reference's get_optical_flow (visualize_optical_flow.py:9)."""
    os.makedirs(images_path, exist_ok=True)
    positions, frames, fps = sample_shot(v_path, start_ms, end_ms)
    if len(frames) < 2:
        return []
    gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames])
    eng = engine or ofb.default_engine()
    pictures = eng.shot(gray, want_bgr=True, **ofb.REFERENCE_PARAMS)["bgr"]     # hot path: one GPU submission
    written = []
    for k in range(1, len(frames)):
        stamp = str(int(positions[k] / fps * 1000))
        path_flow = os.path.join(images_path, "flow_" + stamp + ".jpeg")
        path_source = os.path.join(images_path, "source_" + stamp + ".jpeg")
        cv2.imwrite(path_flow, pictures[k - 1])
        cv2.imwrite(path_source, frames[k])
        written += [path_flow, path_source]
    return written


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("video_dir", help="the directory where the video-files are stored")
    parser.add_argument("images_path", help="the directory where the images are saved")
    parser.add_argument("shot_begin", type=int, help="the begin of a shot in milliseconds")
    parser.add_argument("shot_end", type=int, help="the end of a shot in milliseconds")
    args = parser.parse_args()
    get_optical_flow(args.video_dir, args.images_path, args.shot_begin, args.shot_end)
