"""Writes tests/golden/track_rows.json with the reference's OWN tracking functions.

Runs only in the authoring container (needs /root/reference).  The video script cannot be imported (configargparse,
tensorboardX, natsort … are absent), so the four functions it defines for the per-frame post-processing
(`compute_distance`, `refine_tip_segmentation`, `calc_base_centroid`, `compute_centroids_and_store`,
scripts/test_multiframe_segmentation_on_videos_v3.py:29-194) are cut out of its source with `ast` and executed unmodified,
next to the reference's utils/localization_utils_v2.py; the per-frame driver below repeats :219-227 and :281-303 literally
(including the left call's `cX_prev_left, cX_prev_left` unpacking).

    python -m oracle.make_golden_track
"""
import ast
import json
import os
import sys
from types import SimpleNamespace

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Dev:
    """Stands in for the CUDA tensor `output`: `output[0,c,:,:].cpu().numpy()` is a fresh host copy there."""

    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _Dev(self.a[idx])

    def cpu(self):
        return self

    def numpy(self):
        return np.array(self.a, copy=True)


def reference_namespace():
    sys.path[:0] = [REF]
    from utils import localization_utils_v2 as L
    path = os.path.join(REF, "scripts", "test_multiframe_segmentation_on_videos_v3.py")
    src = open(path).read()
    want = {"compute_distance", "refine_tip_segmentation", "calc_base_centroid", "compute_centroids_and_store"}
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    ns = {"cv2": cv2, "np": np, "create_circular_mask": L.create_circular_mask,
          "determine_local_maxima_and_estimate_centroids": L.determine_local_maxima_and_estimate_centroids}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def run_sequence(ns, seq, area_threshold, dist_threshold, score):
    args = SimpleNamespace(area_threshold=area_threshold, dist_threshold=dist_threshold, score_detection_threshold=score)
    N = len(seq)
    H, W = seq[0].shape[2:]
    args.input_height, args.input_width = H, W
    f = ns["compute_centroids_and_store"]
    centroid_locations = np.zeros((N, 12))
    centroid_locations[:, :] = np.nan
    prev_left_pose_detected_tips = 0
    prev_right_pose_detected_tips = 0
    cX_prev_left = np.zeros(2)
    cY_prev_left = np.zeros(2)
    cX_prev_right = np.zeros(2)
    cY_prev_right = np.zeros(2)
    for count, p in enumerate(seq):
        output = _Dev(p)
        if args.score_detection_threshold > 0:
            output_classes = np.zeros((args.input_height, args.input_width))
            output_classes[np.where(output[0, 1, :, :].cpu().numpy() > args.score_detection_threshold)] = 1
            output_classes[np.where(output[0, 2, :, :].cpu().numpy() > args.score_detection_threshold)] = 2
            output_classes[np.where(output[0, 3, :, :].cpu().numpy() > args.score_detection_threshold)] = 3
            output_classes[np.where(output[0, 4, :, :].cpu().numpy() > args.score_detection_threshold)] = 4
        else:
            output_classes = p.argmax(axis=1).squeeze()
        mask_array = output_classes
        disp_image = np.zeros((H, W, 3), np.uint8)
        centroid_locations, prev_left_pose_detected_tips, cX_prev_left, cX_prev_left, disp_image = f(
            'left', mask_array, output, centroid_locations, count, args, disp_image, prev_left_pose_detected_tips, cX_prev_left, cY_prev_left)
        centroid_locations, prev_right_pose_detected_tips, cX_prev_right, cY_prev_right, disp_image = f(
            'right', mask_array, output, centroid_locations, count, args, disp_image, prev_right_pose_detected_tips, cX_prev_right, cY_prev_right)
    return centroid_locations


def main():
    from . import track_cases
    ns = reference_namespace()
    res = {}
    for pname, (a, d, s) in track_cases.PARAMS.items():
        for name, seq in track_cases.sequences().items():
            rows = run_sequence(ns, seq, a, d, s)
            res["%s/%s" % (pname, name)] = [[None if np.isnan(v) else float(v) for v in r] for r in rows]
    with open(os.path.join(OUT, "track_rows.json"), "w") as f:
        json.dump(res, f, indent=0)
    for k, v in res.items():
        print(k, v[-1])


if __name__ == "__main__":
    main()
