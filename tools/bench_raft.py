"""RAFT-large on the engine: milliseconds per flow field and per frame pair batch (CUDA events), beside torchvision's own
module on the host CPU (fp32, all threads) on the same inputs.  Usage: python tools/bench_raft.py [H W] [batches...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as m  # noqa: E402
from oracle import raft_oracle as RO  # noqa: E402


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (240, 320)
    batches = [int(x) for x in sys.argv[3:]] or [1, 2, 8]
    tv = RO.build(0)
    mine = m.raft_large()
    mine.load_state_dict(tv.state_dict())
    mine = mine.cuda().eval()
    out = []
    for B in batches:
        a, b = RO.frames(B, H, W)
        x, y = a.cuda(), b.cuda()
        with torch.no_grad():
            for _ in range(3):
                mine(x, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                mine(x, y)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        rec = {"what": "RAFT-large, 12 updates, engine (fp16 storage)", "H": H, "W": W, "batch": B, "ms_per_call": round(ms, 3),
               "flows_per_s": round(B / ms * 1e3, 1)}
        if B == batches[0]:
            torch.set_num_threads(os.cpu_count() or 1)
            RO.flow(tv, a, b)
            t = time.time()
            ref = RO.flow(tv, a, b)
            rec["torchvision_cpu_ms"] = round((time.time() - t) * 1e3, 1)
            rec["cpu_threads"] = torch.get_num_threads()
            with torch.no_grad():
                got = mine(x, y)[-1].cpu()
            rec["max_err_px"] = round(float((got - ref).abs().max()), 4)
            rec["mean_err_px"] = round(float((got - ref).abs().mean()), 5)
            rec["max_flow_px"] = round(float(ref.abs().max()), 2)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_raft.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
