"""optical_flow_b200 -- B200-native (sm_100a) drop-in for the one hot path of JacobLoe/optical_flow:

    cv2.calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
    + cartToPolar / min-max normalise / HSV->BGR (visualize_optical_flow.py:48-55) or np.sum(mag) (optical_flow.py:61-64)

Usage at the reference's call sites (optical_flow.py:51, visualize_optical_flow.py:38):

    import optical_flow_b200 as ofb
    flow = ofb.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    mag, ang = ofb.cartToPolar(flow[..., 0], flow[..., 1])
    bgr = ofb.flow_to_bgr(flow)                 # the four visualisation lines, fused on the GPU

Python here is host glue over a C-ABI (include/optflow_b200.h); all arithmetic runs in hand-written
CUDA kernels.  There is no CPU fallback.
"""
import threading

import numpy as np

from .engine import (Farneback, error, make_params, pinned_empty, scale_schedule, algorithmic_bytes,
                     REFERENCE_PARAMS, OPTFLOW_USE_INITIAL_FLOW, OPTFLOW_FARNEBACK_GAUSSIAN)
from .sharding import shard_pairs, shard_shots

__all__ = ["Farneback", "error", "calcOpticalFlowFarneback", "cartToPolar", "flow_to_bgr", "sum_magnitude",
           "default_engine", "pinned_empty", "scale_schedule", "algorithmic_bytes", "shard_pairs", "shard_shots",
           "REFERENCE_PARAMS", "OPTFLOW_USE_INITIAL_FLOW", "OPTFLOW_FARNEBACK_GAUSSIAN"]

_tls = threading.local()


def default_engine(device=None):
    """One engine per Python thread (cv2 releases the GIL and is re-entrant; so is this, per thread)."""
    eng = getattr(_tls, "engine", None)
    if eng is None or (device is not None and eng.device != device):
        eng = Farneback(0 if device is None else device)
        _tls.engine = eng
    return eng


def calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
    """Drop-in for cv2.calcOpticalFlowFarneback (same positional / keyword names, same errors, in-place `flow`)."""
    return default_engine().calc(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)


def cartToPolar(x, y, magnitude=None, angle=None, angleInDegrees=False):
    """Drop-in for cv2.cartToPolar(x, y) -> (magnitude, angle) as the reference calls it
    (optical_flow.py:61, visualize_optical_flow.py:48: x, y are the two planes of a flow field)."""
    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    if x.shape != y.shape:
        raise error("x.size() == y.size() && x.type() == y.type()", func="cartToPolar")
    shp = x.shape
    fl = np.stack([x.reshape(-1), y.reshape(-1)], -1).reshape(1, -1, 2)
    mag, ang = default_engine().cart_to_polar(fl, angle_in_degrees=bool(angleInDegrees))
    mag, ang = mag.reshape(shp), ang.reshape(shp)
    if magnitude is not None and isinstance(magnitude, np.ndarray) and magnitude.shape == shp and magnitude.dtype == np.float32:
        magnitude[...] = mag; mag = magnitude           # cv2 fills a correctly typed destination in place
    if angle is not None and isinstance(angle, np.ndarray) and angle.shape == shp and angle.dtype == np.float32:
        angle[...] = ang; ang = angle
    return mag, ang


def flow_to_bgr(flow):
    return default_engine().flow_to_bgr(flow)


def sum_magnitude(flow):
    return default_engine().sum_magnitude(flow)
