"""Per-frame tool tracking oracle: the video script's post-processing on scipy / OpenCV.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, on top of the same third-party calls
the reference makes (cv2 4.13, scipy 1.18.1, both in this image),
scripts/test_multiframe_segmentation_on_videos_v3.py:
  * the class map of a frame                      :281-289  -> `class_map`
  * `refine_tip_segmentation`                     :32-42
  * `calc_base_centroid`                          :45-59
  * `compute_centroids_and_store` (without the drawing calls)  :61-194  -> `side_keypoints`
  * the per-video state and the two calls per frame            :219-227, :296-303 -> `Tracker`
Pinned by tests/golden/track_rows.json: rows written by the reference's OWN functions, executed
from the script's source by oracle/make_golden_track.py.
"""
import cv2
import numpy as np

from . import localize_oracle as LO


def class_map(prob, score_detection_threshold=0.0):
    """:281-289.  prob (1,5,H,W) float32 probabilities -> (H,W) class ids."""
    if score_detection_threshold > 0:
        out = np.zeros(prob.shape[2:])
        for c in (1, 2, 3, 4):
            out[np.where(prob[0, c] > score_detection_threshold)] = c
        return out
    return prob.argmax(axis=1).squeeze()


def refine_tip_segmentation(mask, area_threshold):
    """:32-42.  Keeps the mask inside the (filled) two largest external contours of area >= threshold."""
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    contours = sorted(contours, key=cv2.contourArea, reverse=True)[:2]
    sel = np.zeros_like(mask)
    for c in contours:
        if cv2.contourArea(c) < area_threshold:
            continue
        cv2.drawContours(sel, [c], 0, (255), thickness=cv2.FILLED)
    return cv2.bitwise_and(mask, mask, mask=sel)


def calc_base_centroid(mask, area_threshold):
    """:45-59.  Centroid of the largest external contour, if its area reaches the threshold."""
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cX, cY = [], []
    for c in sorted(contours, key=cv2.contourArea, reverse=True)[:1]:
        if cv2.contourArea(c) < area_threshold:
            continue
        M = cv2.moments(c)
        if M["m00"] == 0:
            continue
        cX.append(int(M["m10"] / M["m00"]))
        cY.append(int(M["m01"] / M["m00"]))
    return cX, cY


def _dist(x1, y1, x2, y2):
    return np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)


def associate(side, row, base, tips, dist_threshold, prev_detected, cX_prev, cY_prev):
    """The decision tree of `compute_centroids_and_store` (:108-192) given the base centroid lists and,
    when a base exists, the tip centroid lists.  `row` is the frame's 12-vector (written in place)."""
    t1, t2, t3, t4, b1, b2 = (0, 1, 2, 3, 8, 9) if side == "left" else (4, 5, 6, 7, 10, 11)
    iX, iY = base
    if len(iX) == 0:
        return 0, cX_prev, cY_prev
    row[b1], row[b2] = iX[0], iY[0]
    cX, cY = tips

    def one(k):
        row[t1], row[t2], row[t3], row[t4] = cX[k], cY[k], cX[k], cY[k]

    if len(cX) == 0:
        prev_detected = 0
    elif len(cX) == 1:
        if _dist(iX[0], iY[0], cX[0], cY[0]) < dist_threshold:
            prev_detected = 1
            one(0)
        else:
            prev_detected = 0
    else:
        d01 = _dist(iX[0], iY[0], cX[0], cY[0])
        d02 = _dist(iX[0], iY[0], cX[1], cY[1])
        if d01 < dist_threshold and d02 < dist_threshold:
            prev_detected = 2
            d11 = _dist(cX[0], cY[0], cX_prev[0], cY_prev[0])
            d12 = _dist(cX[0], cY[0], cX_prev[1], cY_prev[1])
            d21 = _dist(cX[1], cY[1], cX_prev[0], cY_prev[0])
            d22 = _dist(cX[1], cY[1], cX_prev[1], cY_prev[1])
            a, b = (0, 1) if d11 + d22 < d12 + d21 else (1, 0)
            row[t1], row[t2], row[t3], row[t4] = cX[a], cY[a], cX[b], cY[b]
        elif d01 < dist_threshold:
            prev_detected = 1
            one(0)
        elif d02 < dist_threshold:
            prev_detected = 1
            one(1)
        else:
            prev_detected = 0
    return prev_detected, row[t1:t1 + 4:2], row[t2:t2 + 4:2]


def side_keypoints(side, mask_array, prob, area_threshold):
    """The image half of `compute_centroids_and_store` (:62-90): (base lists, tip lists or None)."""
    base_cls, tip_cls = (3, 4) if side == "left" else (1, 2)
    base = 255 * (mask_array == base_cls).astype(np.uint8)
    tip = 255 * (mask_array == tip_cls).astype(np.uint8)
    heat = prob[0, tip_cls].copy()
    fmask = LO.create_circular_mask(10, 10).astype(np.float64)
    heat[tip == 0] = 0
    b = calc_base_centroid(base, area_threshold)
    if len(b[0]) == 0:
        return b, None
    tip = refine_tip_segmentation(tip, area_threshold)
    return b, LO.determine_local_maxima_and_estimate_centroids(heat, tip > 0, fmask)


class Tracker:
    """Per-video state of `track_on_video` (:219-227) and its two calls per frame (:296-303).  The left call
    unpacks its 3rd and 4th results into the same name (`cX_prev_left, cX_prev_left = …`, :297), so the left
    side carries the previous *y* pair as `cX_prev_left` and its `cY_prev_left` stays zero; kept as is."""

    def __init__(self, area_threshold=10, dist_threshold=40, score_detection_threshold=0.0):
        self.area_threshold, self.dist_threshold, self.score = area_threshold, dist_threshold, score_detection_threshold
        self.prev = {"left": 0, "right": 0}
        self.cX_prev = {"left": np.zeros(2), "right": np.zeros(2)}
        self.cY_prev = {"left": np.zeros(2), "right": np.zeros(2)}

    def step(self, prob):
        row = np.full(12, np.nan)
        mask_array = class_map(prob, self.score)
        for side in ("left", "right"):
            base, tips = side_keypoints(side, mask_array, prob, self.area_threshold)
            d, cx, cy = associate(side, row, base, tips, self.dist_threshold, self.prev[side], self.cX_prev[side], self.cY_prev[side])
            self.prev[side] = d
            if side == "left":
                self.cX_prev[side] = cy
            else:
                self.cX_prev[side], self.cY_prev[side] = cx, cy
        return row
