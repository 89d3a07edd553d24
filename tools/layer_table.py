"""Per-layer device-time table of one forward of a model (CUDA events around every launch).
Usage: python tools/layer_table.py {resunet|hrnet|ternaus} [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "hrnet"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    H, W, N = 480, 640, 5
    net = {"resunet": lambda: M.ResUnet_VB(3, 16, out_dim=N), "hrnet": lambda: M.HighResolutionNet(N),
           "ternaus": lambda: M.TernausNet16(N, 64)}[which]().cuda().eval()
    x = torch.randn(B, 3, H, W, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            net(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        prog = net._plans[(B, H, W)][0]
        per = prog.run_timed()
        per = prog.run_timed()
    shapes = {}
    for c in per:
        k = c["kind"] + " " + c.get("shape", "")
        s = shapes.setdefault(k, {"n": 0, "ms": 0.0, "bytes": 0, "flops": 0})
        s["n"] += 1
        s["ms"] += c["ms"]
        s["bytes"] += c["bytes"]
        s["flops"] += c["flops"]
    tot = sum(v["ms"] for v in shapes.values())
    print(json.dumps({"model": which, "B": B, "ms_forward": round(ms, 3), "frames_per_s": round(B * 1000 / ms, 1), "launches": prog.n_kernels,
                      "ms_sum_timed": round(tot, 3)}))
    for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:45]:
        print("%-46s n=%3d each %8.1f us  %6.0f GB/s %7.1f TF  share %.3f" % (k, v["n"], 1e3 * v["ms"] / v["n"], v["bytes"] / 1e6 / max(v["ms"], 1e-9),
                                                                             v["flops"] / 1e9 / max(v["ms"], 1e-9), v["ms"] / tot))


if __name__ == "__main__":
    main()
