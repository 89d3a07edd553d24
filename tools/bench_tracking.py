"""Per-frame tool tracking (the post-processing of scripts/test_multiframe_segmentation_on_videos_v3.py:281-303), device path
(`ToolTracker.step`: class map, base / tip masks, masked blur, contour refinement, local maxima, centroids, association) vs the
reference's scipy / OpenCV sequence on the host (oracle/track_oracle.py, single-threaded as upstream, plus the D2H copy of the
probability maps it pays).  480x640, 5 classes, batch 1.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402
from oracle import track_cases, track_oracle  # noqa: E402  (bench-only CPU baseline)


def main():
    seq = track_cases.sequences()["noisy_480x640"]
    dev_seq = [torch.from_numpy(p).cuda() for p in seq]
    res = {}
    for score in (0.0, 0.6):
        tr = M.ToolTracker(10, 40, score)
        for _ in range(3):
            rows = [tr.step(p) for p in dev_seq]
        torch.cuda.synchronize()
        n = 20
        t0 = time.perf_counter()
        for _ in range(n):
            for p in dev_seq:
                tr.step(p)
        torch.cuda.synchronize()
        ms_dev = (time.perf_counter() - t0) * 1e3 / (n * len(dev_seq))
        ref = track_oracle.Tracker(10, 40, score)
        t0 = time.perf_counter()
        for p in dev_seq:
            ref.step(p.cpu().numpy())
        ms_cpu = (time.perf_counter() - t0) * 1e3 / len(dev_seq)
        tr2, ref2 = M.ToolTracker(10, 40, score), track_oracle.Tracker(10, 40, score)
        same = all(np.array_equal(tr2.step(p), ref2.step(p.cpu().numpy()), equal_nan=True) for p in dev_seq)
        res["score_threshold_%g" % score] = {"ms_per_frame_device_path_wall": round(ms_dev, 3), "ms_per_frame_reference_cpu_path": round(ms_cpu, 1),
                                            "rows_identical": bool(same)}
    print(json.dumps({"what": "per-frame tool tracking, 480x640, 5 classes, batch 1", "cases": res}))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_tracking.json", "w"), indent=1)


if __name__ == "__main__":
    main()
