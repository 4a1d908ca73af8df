"""The reference's own hot-path lines executed through the installed cv2.  TEST INFRASTRUCTURE ONLY.

These functions are the *reference arm*: they call exactly what
/root/reference/optical_flow.py:51-64 and /root/reference/visualize_optical_flow.py:38-55 call,
with the same arguments.  cv2 is the un-vendored dependency that holds the arithmetic
(requirements_optical_flow.txt:3 pins opencv-python 4.2.0.32; this image has 4.13.0 headless).
"""
import numpy as np

REFERENCE_PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def farneback(prev, next_, flow=None, **params):
    """optical_flow.py:51-59 / visualize_optical_flow.py:38-46."""
    import cv2
    p = dict(REFERENCE_PARAMS)
    p.update(params)
    return cv2.calcOpticalFlowFarneback(prev, next_, flow, p["pyr_scale"], p["levels"], p["winsize"],
                                        p["iterations"], p["poly_n"], p["poly_sigma"], p["flags"])


def summed_magnitude(flow):
    """optical_flow.py:61-64."""
    import cv2
    mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    return np.sum(mag)


def viz(flow, return_hv=False):
    """visualize_optical_flow.py:48-55 (hsv has the shape of the BGR frame, :51)."""
    import cv2
    mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
    hsv = np.zeros(flow.shape[:2] + (3,), np.uint8)
    hsv[..., 1] = 255
    with np.errstate(all="ignore"):
        hsv[..., 0] = ang * 180 / np.pi
        hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX)
    bgr = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)
    if return_hv:
        return bgr, hsv[..., 0].copy(), hsv[..., 2].copy()
    return bgr


def pair_viz(prev, next_, **params):
    """One iteration of the shot loop, visualize_optical_flow.py:38-55."""
    return viz(farneback(prev, next_, None, **params))


def pair_feature(prev, next_, **params):
    """calculate_optical_flow, optical_flow.py:49-66."""
    return summed_magnitude(farneback(prev, next_, None, **params))
