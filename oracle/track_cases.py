"""Synthetic probability-map SEQUENCES for the tool-tracking parity tests (TEST INFRASTRUCTURE).

Built from oracle/localize_cases._prob (no transcendental functions -> bit-identical on every box).  Classes as the video
script uses them (scripts/test_multiframe_segmentation_on_videos_v3.py:62-88): 1 right base, 2 right tip, 3 left base, 4 left tip.
"""
import numpy as np

from .localize_cases import _prob


def _ring(p, cls, cy, cx, r_out, r_in):
    """Turn class `cls` into the winner on an annulus (a component with a hole) of a (1,5,H,W) map, renormalised."""
    H, W = p.shape[2:]
    y = np.arange(H)[:, None]
    x = np.arange(W)[None, :]
    d2 = (y - cy) ** 2 + (x - cx) ** 2
    ring = (d2 <= r_out * r_out) & (d2 >= r_in * r_in)
    q = p[0].astype(np.float64)
    q[:, ring] *= 0.05
    q[cls, ring] += 0.9
    q /= q.sum(0, keepdims=True)
    return q.astype(np.float32)[None]


def sequences():
    s = {}
    H, W = 240, 320
    # two left tips and two right tips moving and crossing over (exercises the previous-frame assignment)
    seq = []
    for t in range(6):
        seq.append(_prob(H, W, [(3, 120, 80, 18), (4, 100 + 6 * t, 100, 7), (4, 130 - 6 * t, 104, 7),
                                (1, 120, 240, 18), (2, 100 + 5 * t, 215, 7), (2, 128 - 5 * t, 218, 6)]))
    s["crossing_240x320"] = seq
    # tips appear / vanish, a tip beyond dist_threshold of its base, blobs under the area threshold, no base at all
    s["appear_240x320"] = [
        _prob(H, W, [(3, 60, 60, 15), (4, 60, 90, 7)]),
        _prob(H, W, [(3, 60, 60, 15), (4, 60, 90, 7), (4, 70, 40, 6), (1, 180, 250, 14)]),
        _prob(H, W, [(3, 60, 60, 15), (4, 60, 90, 7), (4, 200, 300, 8), (1, 180, 250, 14), (2, 170, 220, 6)]),
        _prob(H, W, [(4, 60, 90, 7), (2, 170, 220, 6)]),
        _prob(H, W, [(3, 60, 60, 2.5), (4, 60, 90, 7), (1, 180, 250, 14), (2, 170, 220, 1.8), (2, 190, 225, 7)]),
        _prob(H, W, [(3, 60, 60, 15), (4, 62, 92, 7), (4, 72, 42, 6), (4, 40, 60, 5), (1, 180, 250, 14), (2, 172, 222, 6)]),
        _prob(H, W, [(3, 60, 60, 15), (4, 200, 300, 8), (4, 210, 20, 8), (1, 180, 250, 14), (2, 10, 10, 6), (2, 172, 222, 6)]),
    ]
    # a ring-shaped tip component with a second tip blob inside its hole (filled contours keep the nested blob), with noise
    a = _ring(_prob(H, W, [(3, 120, 110, 16), (4, 120, 160, 6), (1, 60, 260, 14), (2, 60, 230, 7)], noise=0.05, tag="r0"), 4, 120, 160, 22, 16)
    b = _ring(_prob(H, W, [(3, 120, 110, 16), (4, 124, 158, 6), (4, 30, 30, 9), (1, 60, 260, 14), (2, 64, 232, 7)], noise=0.05, tag="r1"), 4, 120, 160, 22, 16)
    c = _ring(_prob(H, W, [(3, 120, 110, 16), (4, 30, 30, 9), (4, 200, 60, 12), (1, 60, 260, 14)], noise=0.05, tag="r2"), 2, 60, 225, 20, 15)
    s["nested_240x320"] = [a, b, c]
    # full resolution, noisy
    s["noisy_480x640"] = [
        _prob(480, 640, [(1, 100, 100, 25), (2, 140 + 3 * t, 120, 10), (2, 90, 140 + 2 * t, 8), (3, 380, 520, 30),
                         (4, 350 - 4 * t, 500, 10), (4, 400, 490 + 3 * t, 7)], noise=0.3, tag="tn%d" % t) for t in range(3)]
    return s


PARAMS = {  # (area_threshold, dist_threshold, score_detection_threshold): the script's defaults (:82-87) and a thresholded variant
    "default": (10, 40, 0.0),
    "score": (10, 40, 0.6),
}
