"""Synthetic 5-class probability maps for the keypoint-extraction parity tests.

TEST INFRASTRUCTURE.  Built without transcendental functions (rational bumps,
one IEEE division) so the maps are bit-identical in the authoring container and
on the GPU box.  Class convention of `centroid_error`
(utils/localization_utils_v2.py:193-212): 1 right base, 2 right tip,
3 left base, 4 left tip.
"""
import numpy as np

from . import synth


def _bump(H, W, cy, cx, r, flat=0.0):
    """(1 - d^2/r^2)^2 inside radius r, 0 outside; `flat` clips the top to make ties."""
    y = np.arange(H, dtype=np.float64)[:, None]
    x = np.arange(W, dtype=np.float64)[None, :]
    d2 = ((y - cy) ** 2 + (x - cx) ** 2) / float(r * r)
    b = np.where(d2 < 1.0, (1.0 - d2) ** 2, 0.0)
    if flat > 0:
        b = np.minimum(b, 1.0 - flat)
    return b


def _prob(H, W, blobs, noise=0.0, tag="x", amp=60.0):
    """blobs: list of (cls, cy, cx, r[, flat]).  Returns (1,5,H,W) float32 that sums to 1."""
    s = np.zeros((5, H, W), dtype=np.float64)
    s[0] = 1.0
    for b in blobs:
        cls, cy, cx, r = b[:4]
        flat = b[4] if len(b) > 4 else 0.0
        s[cls] += amp * _bump(H, W, cy, cx, r, flat)
    if noise > 0:
        s += noise * synth.uniform("locnoise/" + tag, (5, H, W), 7).astype(np.float64)
    p = (s / s.sum(0, keepdims=True)).astype(np.float32)
    return p[None]


def cases():
    c = {}
    c["clean_480x640"] = _prob(480, 640, [(1, 300, 500, 30), (2, 200, 420, 9), (2, 230, 460, 7),
                                          (3, 310, 140, 28), (4, 190, 230, 8), (4, 215, 200, 10)])
    c["empty_120x160"] = _prob(120, 160, [])
    c["noisy_480x640"] = _prob(480, 640, [(1, 100, 100, 25), (2, 140, 180, 12), (3, 380, 520, 33),
                                          (4, 300, 430, 11), (4, 330, 470, 6)], noise=0.35, tag="n1")
    c["border_160x200"] = _prob(160, 200, [(1, 0, 0, 14), (2, 159, 199, 9), (3, 80, 199, 12), (4, 0, 100, 7)])
    c["tiny_120x160"] = _prob(120, 160, [(1, 30, 30, 1.2), (2, 60, 60, 1.1), (3, 90, 100, 2.2), (4, 20, 140, 1.6)])
    c["plateau_200x240"] = _prob(200, 240, [(1, 60, 60, 20, 0.5), (2, 100, 150, 12, 0.3), (3, 150, 60, 18, 0.6),
                                            (4, 50, 200, 10, 0.2), (4, 150, 200, 10, 0.2)])
    c["three_tips_240x320"] = _prob(240, 320, [(2, 50, 50, 9), (2, 120, 160, 12), (2, 200, 280, 6),
                                               (4, 60, 250, 8), (4, 180, 60, 8), (4, 120, 300, 8), (1, 200, 100, 20)])
    c["touching_160x200"] = _prob(160, 200, [(2, 80, 80, 12), (2, 80, 100, 12), (4, 40, 150, 9), (4, 52, 158, 9),
                                             (1, 120, 40, 15), (3, 120, 160, 15)], noise=0.05, tag="t1")
    return c
