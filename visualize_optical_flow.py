#!/usr/bin/env python
"""visualize_optical_flow.py -- B200 version of the reference's shot visualiser, same command line:

    python visualize_optical_flow.py <video_dir> <images_path> <shot_begin ms> <shot_end ms>

It writes exactly the files the reference writes (/root/reference/visualize_optical_flow.py:9-63): for every
sampled frame after the first, `flow_<ms>.jpeg` (the HSV-coded Farneback flow against the previous sampled frame)
and `source_<ms>.jpeg` (the frame itself), sampling every 300 ms between shot_begin and shot_end.

What differs is only WHERE the hot path runs: the reference calls cvtColor(BGR2GRAY) + cv2.calcOpticalFlowFarneback +
cartToPolar + normalize + cvtColor once per pair in a Python loop (:31-55); here the sampled BGR frames of the shot are
handed to the GPU engine as ONE shot (optical_flow_b200.Farneback.shot_bgr -> C-ABI ofb_shot_bgr_host), which converts
to gray on the device, expands each frame once, batches the pairs, encodes every flow picture as the baseline JPEG
cv2.imwrite would produce (same bytes: optical_flow_b200/csrc/jpeg.cu) and returns the FILES -- ~0.2 MB per 1080p pair
cross PCIe instead of a 6.2 MB raw picture.  Decoding (cv2.VideoCapture) and the JPEG of the source frame stay with cv2 on
the host, as in the reference; frames ahead of the decoder are reached by decoding forward instead of a seek per frame,
and the source JPEGs are encoded by a small thread pool (optical_flow_b200/video.py).
"""
import argparse
import os
import queue
import threading

import cv2
import numpy as np

import optical_flow_b200 as ofb
from optical_flow_b200.video import FrameReader, JpegWriter

STEP_SIZE = 300     # ms between sampled frames (reference: module constant of the same name)
MAX_FRAMES = 17     # decoded frames per GPU submission (16 pairs)


def sample_shot(v_path, start_ms, end_ms):
    """Generator over the frames the reference's loop would visit (visualize_optical_flow.py:14-27, :63): positions
    start at the *float* fps*start_ms/1000 and advance by int(fps*STEP_SIZE/1000); it stops at the first unreadable
    frame.  Yields (position, BGR frame, fps)."""
    vid = cv2.VideoCapture(v_path)
    fps = vid.get(cv2.CAP_PROP_FPS)
    pos = fps * start_ms / 1000
    last = int(fps * end_ms / 1000)
    stride = int(fps * STEP_SIZE / 1000)
    reader = FrameReader(vid)
    try:
        while pos < last:
            ok, bgr = reader.read_at(pos)           # == vid.set(CAP_PROP_POS_FRAMES, pos); vid.read()
            if not ok:
                break
            yield pos, bgr, fps
            if stride <= 0:          # the reference would spin forever on this input; one frame is all it can mean
                break
            pos += stride
    finally:
        vid.release()


def get_optical_flow(v_path, images_path, start_ms, end_ms, engine=None):
    """Same signature and artefacts as the reference's get_optical_flow (visualize_optical_flow.py:9).

    Streaming: a decoder thread runs ahead of the GPU (cv2 releases the GIL while decoding); every MAX_FRAMES
    consecutive sampled frames are one GPU submission (chunks share their boundary frame; results do not depend on
    the chunking), and the JPEGs of a finished chunk are encoded by the writer pool while the next chunk is decoded."""
    os.makedirs(images_path, exist_ok=True)
    eng = engine or ofb.default_engine()
    written = []
    frames_q = queue.Queue(maxsize=2 * MAX_FRAMES)

    failure = []

    def decode():
        try:
            for item in sample_shot(v_path, start_ms, end_ms):
                frames_q.put(item)
        except BaseException as e:          # re-raised in the caller's thread: the reference would have crashed here too
            failure.append(e)
        finally:
            frames_q.put(None)

    threading.Thread(target=decode, daemon=True).start()
    with JpegWriter() as out:
        chunk, positions, fps, done = [], [], 0.0, False
        while not done:
            item = frames_q.get()
            if item is None:
                done = True
            else:
                positions.append(item[0]); chunk.append(item[1]); fps = item[2]
            if len(chunk) >= 2 and (done or len(chunk) == MAX_FRAMES):
                files = eng.shot_bgr_jpeg(np.stack(chunk), **ofb.REFERENCE_PARAMS)["files"]    # hot path, :31-58
                for k in range(1, len(chunk)):
                    stamp = str(int(positions[k] / fps * 1000))
                    path_flow = os.path.join(images_path, "flow_" + stamp + ".jpeg")
                    path_source = os.path.join(images_path, "source_" + stamp + ".jpeg")
                    with open(path_flow, "wb") as fh:          # the bytes cv2.imwrite(path_flow, picture) would write
                        fh.write(files[k - 1].tobytes())
                    out.imwrite(path_source, chunk[k])
                    written += [path_flow, path_source]
                chunk, positions = chunk[-1:], positions[-1:]
    if failure:
        raise failure[0]
    return written


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("video_dir", help="the directory where the video-files are stored")
    parser.add_argument("images_path", help="the directory where the images are saved")
    parser.add_argument("shot_begin", type=int, help="the begin of a shot in milliseconds")
    parser.add_argument("shot_end", type=int, help="the end of a shot in milliseconds")
    args = parser.parse_args()
    get_optical_flow(args.video_dir, args.images_path, args.shot_begin, args.shot_end)
