"""Key-point extraction oracle: the reference's own scipy / OpenCV call sequence.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  scipy.ndimage and cv2 are third-party
dependencies of the reference that are present in this image (scipy 1.18.1, OpenCV 4.13.0); the
functions below restate utils/localization_utils_v2.py:5-40 and the prediction half of
`centroid_error` (:193-212, :247-272) on top of them, so the device kernels are checked against
the exact library semantics the reference relies on.  Pinned by tests/golden/localize_centroids.json
(outputs of the reference's own `centroid_error`, written by oracle/make_golden.py).
"""
import cv2
import numpy as np
from scipy import ndimage


def create_circular_mask(h, w, center=None, radius=None):
    """utils/localization_utils_v2.py:5-13."""
    if center is None:
        center = (int(w / 2), int(h / 2))
    if radius is None:
        radius = min(center[0], center[1], w - center[0], h - center[1])
    Y, X = np.ogrid[:h, :w]
    return np.sqrt((X - center[0]) ** 2 + (Y - center[1]) ** 2) <= radius


def contour_records(mask):
    """(area, m00, m10, m01, first_x, first_y) per external contour, in findContours order."""
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    out = []
    for c in contours:
        M = cv2.moments(c)
        out.append((cv2.contourArea(c), M["m00"], M["m10"], M["m01"], int(c[0][0][0]), int(c[0][0][1])))
    return out


def calc_centroids(mask):
    """utils/localization_utils_v2.py:15-33."""
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cnts = sorted(contours, key=cv2.contourArea, reverse=True)[:2]
    cX, cY = [], []
    for c in cnts:
        M = cv2.moments(c)
        if M["m00"] == 0:
            cX.append(int(c[0][0][0]))
            cY.append(int(c[0][0][1]))
        else:
            cX.append(int(M["m10"] / M["m00"]))
            cY.append(int(M["m01"] / M["m00"]))
    return cX, cY


def smoothed(heatmap):
    return ndimage.gaussian_filter(heatmap, 4)


def localmax_mask(heatmap, blob, mask):
    """utils/localization_utils_v2.py:35-39 up to the uint8 mask handed to calc_centroids."""
    sm = ndimage.gaussian_filter(heatmap, 4)
    localmax = ndimage.maximum_filter(sm, footprint=mask) == sm
    return 255 * (blob & localmax).astype(np.uint8)


def determine_local_maxima_and_estimate_centroids(heatmap, blob, mask):
    return calc_centroids(localmax_mask(heatmap, blob, mask))


def predicted_keypoints(prob):
    """c_pred of `centroid_error` (5 classes) for a (1,5,H,W) float32 probability map."""
    mask = create_circular_mask(10, 10).astype(np.float64)
    pred = prob.argmax(axis=1).squeeze()
    c_lb = calc_centroids(255 * (pred == 3).astype(np.uint8))
    c_lt = determine_local_maxima_and_estimate_centroids(prob[0, 4], pred == 4, mask)
    c_rb = calc_centroids(255 * (pred == 1).astype(np.uint8))
    c_rt = determine_local_maxima_and_estimate_centroids(prob[0, 2], pred == 2, mask)

    def tips(c):
        xs, ys = c
        if len(xs) == 0:
            return [np.nan, np.nan], [np.nan, np.nan]
        if len(xs) == 1:
            return [xs[0], xs[0]], [ys[0], ys[0]]
        return xs, ys

    def base(c):
        xs, ys = c
        return ([np.nan], [np.nan]) if len(xs) == 0 else (xs, ys)

    (ltx, lty), (rtx, rty), (lbx, lby), (rbx, rby) = tips(c_lt), tips(c_rt), base(c_lb), base(c_rb)
    return [rtx, rty, rbx, rby, ltx, lty, lbx, lby]
