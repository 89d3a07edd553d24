"""Heat-map head + key-point extraction, device path vs the reference's CPU path.
Device: fp32 logits (1,5,480,640) -> mfc_heatmap_head (log-softmax, probs, uint8 argmax) -> predicted_keypoints (scipy-exact
gaussian blur, footprint local maxima, contour tracing with exact integer moments) with only a few scalars per contour crossing
to the host.  CPU: the reference's own sequence (softmax -> D2H of the maps -> scipy.ndimage + OpenCV, oracle/localize_oracle.py),
single-threaded as upstream.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402
from oracle import localize_cases, localize_oracle  # noqa: E402  (bench-only CPU baseline)


def main():
    cases = localize_cases.cases()
    res = {}
    for name in ("clean_480x640", "noisy_480x640"):
        prob = torch.from_numpy(cases[name]).cuda()
        logits = prob.clamp_min(1e-30).log().contiguous()
        # device path
        def dev():
            logp, p, amax = M.heatmap_head(logits)
            return M.predicted_keypoints(p)
        for _ in range(3):
            kp = dev()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 20
        for _ in range(n):
            kp = dev()
        torch.cuda.synchronize()
        ms_dev = (time.perf_counter() - t0) * 1e3 / n
        # head kernel alone (device-timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            M.heatmap_head(logits)
        e1.record()
        torch.cuda.synchronize()
        us_head = e0.elapsed_time(e1) * 1000 / 50
        # reference CPU path (includes the D2H copy the reference pays)
        t0 = time.perf_counter()
        m = 3
        for _ in range(m):
            p_host = torch.softmax(logits, dim=1).cpu().numpy()
            ref = localize_oracle.predicted_keypoints(p_host)
        ms_cpu = (time.perf_counter() - t0) * 1e3 / m
        same = [list(map(lambda v: None if v is None or (isinstance(v, float) and np.isnan(v)) else int(v), a)) for a in kp] == \
               [list(map(lambda v: None if v is None or (isinstance(v, float) and np.isnan(v)) else int(v), a)) for a in ref]
        res[name] = {"ms_device_path_wall": round(ms_dev, 3), "us_heatmap_head_kernel": round(us_head, 2),
                     "heatmap_head_GBs": round((5 * 4 * 3 + 1) * 480 * 640 / us_head * 1e-3, 1),
                     "ms_reference_cpu_path": round(ms_cpu, 1), "keypoints_identical": bool(same)}
    print(json.dumps({"what": "heat-map head + key-point extraction, 480x640, 5 classes, batch 1", "cases": res}))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_keypoints.json", "w"), indent=1)


if __name__ == "__main__":
    main()
