#!/usr/bin/env python
"""optical_flow.py -- B200 version of the reference's batch feature extractor, same command line and artefacts:

    python optical_flow.py <features_root> [videoids ...] [--frame_width 129] [--step_size 300]
                           [--window_size 300] [--top_percentile 5] [--force_run False]

For every video id it reads <features_root>/<id>/media/<id>.mp4, samples one frame pair per window, reduces each
pair's Farneback flow to the summed magnitude, averages the windows covering each step position, scales by the
`top_percentile`-th percentile and writes <features_root>/<id>/opticalflow/<id>.csv plus a `.done` version file
(/root/reference/optical_flow.py:69-168).

The per-pair hot path (calculate_optical_flow, optical_flow.py:49-66: cv2.calcOpticalFlowFarneback ->
cartToPolar -> np.sum) does not run on the CPU here, and neither does read_frame's resize + BGR->gray (:25-46): the
decoded BGR frames of PAIRS_PER_SUBMISSION windows at a time go to the GPU engine as one batch of independent pairs
(optical_flow_b200.Farneback.pairs_bgr -> C-ABI ofb_pairs_bgr_host; the resize and the gray conversion are bit-exact
re-implementations of cv2's).  Decoding stays with cv2; a frame that lies just ahead of the decoder is reached by
decoding forward instead of a seek (optical_flow_b200/video.py).
"""
import argparse
import logging
import os

import cv2
import numpy as np
from tqdm import tqdm

import optical_flow_b200 as ofb
from optical_flow_b200.video import FrameReader

EXTRACTOR = "opticalflow"
VERSION = '20201209'      # kept equal to the reference's, so existing .done files stay valid
STANDALONE = True         # True: write .done files; False: always recompute and never write them
PAIRS_PER_SUBMISSION = 32 # decoded frame pairs handed to the GPU at a time (bounds host memory on long videos)

logger = logging.getLogger(__name__)
logging.basicConfig(level=logging.INFO)
_handler = logging.StreamHandler()
_handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
logger.addHandler(_handler)
logger.propagate = False


def resized_size(frame, frame_width):
    """Target (width, height) of resize_frame (optical_flow.py:25-31): keep the aspect ratio, width `frame_width`."""
    h, w = frame.shape[0], frame.shape[1]
    return frame_width, int(frame_width / (w / h))


def calculate_optical_flow_batch(starts, ends, frame_width, engine=None):
    """The batched counterpart of read_frame's preprocessing + calculate_optical_flow (optical_flow.py:25-66) for
    decoded BGR frames: resize to `frame_width` (if set), BGR->gray, Farneback, one summed magnitude per pair."""
    eng = engine or ofb.default_engine()
    dsize = resized_size(starts[0], frame_width) if frame_width else None
    res = eng.pairs_bgr(np.stack(starts), np.stack(ends), dsize=dsize, want_magsum=True, **ofb.REFERENCE_PARAMS)
    return res["magsum"]


def get_optical_flow(v_path, frame_width, step_size, window_size, engine=None):
    """Same signature and return value as the reference (optical_flow.py:69-117):
    ([mean summed magnitude per step position], [start_ms, end_ms])."""
    vid = cv2.VideoCapture(v_path)
    if not vid.isOpened():
        raise IOError("Unable to read from video: '{v_path}'".format(v_path=v_path))
    tot_frames = int(vid.get(cv2.CAP_PROP_FRAME_COUNT))
    fps = vid.get(cv2.CAP_PROP_FPS)
    step = int(fps * step_size / 1000)
    half = int(int(fps * window_size / 1000) / 2.)
    windows = [(max(0, c - half), min(tot_frames - 1, c + half)) for c in range(0, tot_frames, step)]

    # host: decode every window's two frames; the first unreadable frame ends the video (optical_flow.py:87-96).
    # GPU: PAIRS_PER_SUBMISSION windows per batch.
    reader = FrameReader(vid)
    mags, spans, starts, ends = [], [], [], []

    def flush():
        if spans:
            sums = calculate_optical_flow_batch(starts, ends, frame_width, engine)
            mags.extend((s, e, m) for (s, e), m in zip(spans, sums))
            del spans[:], starts[:], ends[:]

    for first, last in windows:
        ok, a = reader.read_at(first)
        if not ok or a is None:
            break
        ok, b = reader.read_at(last)
        if not ok or b is None:
            break
        spans.append((first, last)); starts.append(a); ends.append(b)
        if len(spans) >= PAIRS_PER_SUBMISSION:
            flush()
    flush()
    if not mags:
        raise Exception('Unable to extract the optical flow, no frames where found.')
    vid.release()

    agg = []
    for pos in range(0, tot_frames, step):
        covering = [m for (s, e, m) in mags if s <= pos < e]
        if covering:
            agg.append((pos, np.mean(covering)))
        else:
            logger.info("WARN: no entry for pos={pos}".format(pos=pos))
    start_ms = int(agg[0][0] / fps * 1000)
    end_ms = int(agg[-1][0] / fps * 1000)
    return [m for _, m in agg], [start_ms, end_ms]


def scale_magnitudes(mag, top_percentile):
    scaled = np.clip(mag / np.percentile(mag, top_percentile), a_min=0, a_max=1) * 100.
    return list(np.round(scaled, decimals=2))


def write_mag_to_csv(f_path, mag, segment_timestamps):
    with open(f_path, 'w', newline='') as f:
        f.write(str(segment_timestamps[0]) + '\t' + str(segment_timestamps[1]) + '\t' + " ".join(str(m) for m in mag))


def main(features_root, frame_width, step_size, window_size, top_percentile, videoids, force_run):
    logger.info("Computing optical flow for {0} videos".format(len(videoids)))
    engine = None                     # created on the first video that needs work: an all-up-to-date run touches no GPU
    for videoid in tqdm(videoids):
        features_dir = os.path.join(features_root, videoid, EXTRACTOR)
        v_path = os.path.join(features_root, videoid, 'media', videoid + '.mp4')
        os.makedirs(features_dir, exist_ok=True)
        f_path_csv = os.path.join(features_dir, "{videoid}.csv".format(videoid=videoid))
        done_file_path = os.path.join(features_dir, '.done')
        done_version = '\n'.join([VERSION, str(frame_width), str(step_size), str(window_size), str(top_percentile)])
        up_to_date = os.path.isfile(done_file_path) and open(done_file_path, 'r').read() == done_version
        if up_to_date and force_run != 'True':
            logger.info('optical flow was already done')
            continue
        engine = engine or ofb.default_engine()
        segments, timestamps = get_optical_flow(v_path, frame_width, step_size, window_size, engine)
        write_mag_to_csv(f_path_csv, scale_magnitudes(segments, top_percentile), timestamps)
        if STANDALONE:
            with open(done_file_path, 'w') as d:
                d.write(done_version)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("features_root", help="the directory where the images are to be stored")
    parser.add_argument("videoids", help="List of video ids. If empty, entire corpus is iterated.", nargs='*')
    parser.add_argument("--frame_width", type=int, default=129, help="set the width at which to which the frames are rescaled, default is 129")
    parser.add_argument("--step_size", type=int, default=300, help="defines at which distances the optical flow is calculated, in milliseconds, default is 300")
    parser.add_argument("--window_size", type=int, default=300,
                        help="defines the range in which images for optical flow calculation are extracted,"
                             " if window_size is equal to step_size two frames are extracted, default is 300")
    parser.add_argument("--top_percentile", type=int, default=5, help="set the percentage of magnitudes that are used to determine the max magnitude,")
    parser.add_argument("--force_run", default='False', help='sets whether the script runs regardless of the version of .done-files')
    args = parser.parse_args()
    main(args.features_root, args.frame_width, args.step_size, args.window_size, args.top_percentile, args.videoids, args.force_run)
