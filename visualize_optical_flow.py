#!/usr/bin/env python
"""visualize_optical_flow.py -- B200 version of the reference's shot visualiser, same command line:

    python visualize_optical_flow.py <video_dir> <images_path> <shot_begin ms> <shot_end ms>

It writes exactly the files the reference writes (/root/reference/visualize_optical_flow.py:9-63): for every
sampled frame after the first, `flow_<ms>.jpeg` (the HSV-coded Farneback flow against the previous sampled frame)
and `source_<ms>.jpeg` (the frame itself), sampling every 300 ms between shot_begin and shot_end.

What differs is only WHERE the hot path runs: the reference calls cvtColor(BGR2GRAY) + cv2.calcOpticalFlowFarneback +
cartToPolar + normalize + cvtColor once per pair in a Python loop (:31-55); here the sampled BGR frames of the shot are
handed to the GPU engine as ONE shot (optical_flow_b200.Farneback.shot_bgr -> C-ABI ofb_shot_bgr_host), which converts
to gray on the device, expands each frame once, batches the pairs, and returns the BGR pictures.  Decoding
(cv2.VideoCapture) and JPEG encoding stay with cv2 on the host, as in the reference; frames ahead of the decoder are
reached by decoding forward instead of a seek per frame, and the JPEGs are encoded by a small thread pool
(optical_flow_b200/video.py).
"""
import argparse
import os

import cv2
import numpy as np

import optical_flow_b200 as ofb
from optical_flow_b200.video import FrameReader, JpegWriter

STEP_SIZE = 300     # ms between sampled frames (reference: module constant of the same name)
MAX_FRAMES = 96     # decoded frames per GPU submission


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
    reader = FrameReader(vid)
    while pos < last:
        ok, bgr = reader.read_at(pos)           # == vid.set(CAP_PROP_POS_FRAMES, pos); vid.read()
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
    eng = engine or ofb.default_engine()
    written = []
    with JpegWriter() as out:
        # hot path: one GPU submission per chunk of consecutive frames (chunks share their boundary frame, so a long
        # shot never holds more than MAX_FRAMES decoded frames in one array; results do not depend on the chunking)
        for c0 in range(0, len(frames) - 1, MAX_FRAMES - 1):
            chunk = frames[c0:c0 + MAX_FRAMES]
            pictures = eng.shot_bgr(np.stack(chunk), want_bgr=True, **ofb.REFERENCE_PARAMS)["bgr"]
            for k in range(1, len(chunk)):
                stamp = str(int(positions[c0 + k] / fps * 1000))
                path_flow = os.path.join(images_path, "flow_" + stamp + ".jpeg")
                path_source = os.path.join(images_path, "source_" + stamp + ".jpeg")
                out.imwrite(path_flow, pictures[k - 1])
                out.imwrite(path_source, chunk[k])
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
