"""CPU tests of the tool-tracking path: the oracle (oracle/track_oracle.py) against rows written by the reference's own
functions (tests/golden/track_rows.json, oracle/make_golden_track.py), and the product's host-side association logic
(mfcnet_tracker_b200.tracking.ToolTracker._associate) against the oracle's restatement of
scripts/test_multiframe_segmentation_on_videos_v3.py:104-192 on random centroid lists."""
import json
import os

import numpy as np

from oracle import track_cases, track_oracle as TO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gold():
    with open(os.path.join(ROOT, "tests", "golden", "track_rows.json")) as f:
        g = json.load(f)
    return {k: np.array([[np.nan if v is None else v for v in r] for r in rows]) for k, rows in g.items()}


def test_track_oracle_matches_reference_golden():
    gold = _gold()
    seqs = track_cases.sequences()
    assert set(gold) == {"%s/%s" % (p, n) for p in track_cases.PARAMS for n in seqs}
    for pname, (a, d, s) in track_cases.PARAMS.items():
        for name, seq in seqs.items():
            tr = TO.Tracker(a, d, s)
            for t, p in enumerate(seq):
                assert np.array_equal(tr.step(p), gold["%s/%s" % (pname, name)][t], equal_nan=True), (pname, name, t)


def test_golden_rows_cover_the_decision_tree():
    """The fixtures exercise: no base, base without tips, one tip, two tips in both assignment orders, a tip out of range."""
    rows = np.concatenate(list(_gold().values()))
    left = rows[:, [0, 1, 2, 3, 8, 9]]
    assert np.isnan(left[:, 4]).any()                                            # no base
    assert (~np.isnan(left[:, 4]) & np.isnan(left[:, 0])).any()                  # base, no accepted tip
    assert (~np.isnan(left[:, 0]) & (left[:, 0] == left[:, 2]) & (left[:, 1] == left[:, 3])).any()   # one tip (duplicated)
    assert (~np.isnan(left[:, 0]) & ((left[:, 0] != left[:, 2]) | (left[:, 1] != left[:, 3]))).any()  # two tips


def test_association_logic_matches_oracle():
    from mfcnet_tracker_b200.tracking import ToolTracker
    rng = np.random.default_rng(5)
    for trial in range(400):
        thr = int(rng.integers(5, 60))
        side = "left" if trial % 2 else "right"
        tr = ToolTracker(10, thr)
        px = rng.choice([0.0, 30.0, 60.0, np.nan], 2)
        py = rng.choice([0.0, 30.0, 60.0, np.nan], 2)
        tr._px[side], tr._py[side] = px.copy(), py.copy()
        nb, nt = int(rng.integers(0, 2)), int(rng.integers(0, 3))
        base = ([int(v) for v in rng.integers(0, 80, nb)], [int(v) for v in rng.integers(0, 80, nb)])
        tips = ([int(v) for v in rng.integers(0, 80, nt)], [int(v) for v in rng.integers(0, 80, nt)])
        row_a, row_b = np.full(12, np.nan), np.full(12, np.nan)
        got = tr._associate(side, row_a, base, tips)
        want = TO.associate(side, row_b, base, tips, thr, 7, px.copy(), py.copy())
        assert np.array_equal(row_a, row_b, equal_nan=True), trial
        assert got[0] == want[0] and np.array_equal(got[1], want[1], equal_nan=True) and np.array_equal(got[2], want[2], equal_nan=True), trial


def test_tracker_refuses_cpu_tensors():
    import pytest
    import torch
    from mfcnet_tracker_b200.tracking import ToolTracker, class_map
    with pytest.raises(RuntimeError):
        ToolTracker().step(torch.zeros(1, 5, 8, 8))
    with pytest.raises(RuntimeError):
        class_map(torch.zeros(1, 5, 8, 8))
    with pytest.raises(ValueError):
        ToolTracker().step(torch.zeros(2, 5, 8, 8))
    with pytest.raises(RuntimeError):
        ToolTracker().collect()


def test_top_subset_keeps_the_sorted_prefix():
    """heatmap.top_subset + the consumers' stable sort == the stable sort over all records (ties in area included)."""
    from mfcnet_tracker_b200.heatmap import _centroids_from_records, contour_records, top_subset
    from mfcnet_tracker_b200.tracking import base_centroid_from_records
    rng = np.random.default_rng(9)
    W = 50
    for trial in range(200):
        n = int(rng.integers(0, 40))
        raw = np.zeros((n, 6))
        raw[:, 0] = rng.integers(0, 6, n) * 2            # few distinct areas: many ties
        raw[:, 1] = rng.integers(0, 1000, n)
        raw[:, 2] = rng.integers(0, 1000, n)
        pos = rng.permutation(W * 40)[:n]
        raw[:, 3], raw[:, 4] = pos % W, pos // W
        full = contour_records(raw, W)
        sub = contour_records(top_subset(raw, W), W)
        assert _centroids_from_records(sub) == _centroids_from_records(full), trial
        assert base_centroid_from_records(sub, 3) == base_centroid_from_records(full, 3), trial
