"""GPU micro-benchmark of single fused-conv layers through libmfcnet_b200.so (CUDA-event timing).
Usage: python tools/conv_bench.py [case-index ...] [--iters N]   (no index = all cases)
Prints one JSON line per case: device microseconds per launch, algorithmic GB/s and TFLOP/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402,F401
from mfcnet_tracker_b200 import engine  # noqa: E402
from mfcnet_tracker_b200.engine import Act  # noqa: E402

CASES = [
    dict(name="16->16 k3 plain", B=4, H=480, W=640, cins=[16], Cout=16, k=3),
    dict(name="16->16 k3 stats", B=4, H=480, W=640, cins=[16], Cout=16, k=3, stats=True),
    dict(name="16->16 k3 aff+stats", B=4, H=480, W=640, cins=[16], Cout=16, k=3, stats=True, aff=True),
    dict(name="32->16 k3 stats", B=4, H=480, W=640, cins=[16, 16], Cout=16, k=3, stats=True),
    dict(name="32->16 k1 res", B=4, H=480, W=640, cins=[16, 16], Cout=16, k=1, res=True),
    dict(name="3->16 k7", B=4, H=480, W=640, cins=[3], Cout=16, k=7),
    dict(name="128->128 k3 stats", B=4, H=60, W=80, cins=[128], Cout=128, k=3, stats=True),
    dict(name="64->64 k3", B=4, H=120, W=160, cins=[64], Cout=64, k=3),
    dict(name="32->16 k3 up2", B=4, H=240, W=320, cins=[32], Cout=16, k=3, ups=2),
    dict(name="fusion 11x11", B=8, H=480, W=640, cins=[5, 5, 5, 7], Cout=15, k=11, act=1),
    dict(name="16->5 k1 nchw", B=4, H=480, W=640, cins=[16], Cout=5, k=1, nchw=True),
    # the headline step runs the SFC over 24 frames at once: tensors (236 MB) no longer fit the 126 MB L2
    dict(name="B24 16->16 k3 plain", B=24, H=480, W=640, cins=[16], Cout=16, k=3),
    dict(name="B24 16->16 k3 aff+stats", B=24, H=480, W=640, cins=[16], Cout=16, k=3, stats=True, aff=True),
    dict(name="B24 32->16 k3 aff+stats", B=24, H=480, W=640, cins=[16, 16], Cout=16, k=3, stats=True, aff=True),
    dict(name="B24 32->16 k1 res", B=24, H=480, W=640, cins=[16, 16], Cout=16, k=1, res=True),
    dict(name="B24 128->128 k3 aff+stats", B=24, H=60, W=80, cins=[128], Cout=128, k=3, stats=True, aff=True),
    dict(name="B24 64->64 k3 aff+stats", B=24, H=120, W=160, cins=[64], Cout=64, k=3, stats=True, aff=True),
    dict(name="B24 32->32 k3 aff+stats", B=24, H=240, W=320, cins=[32], Cout=32, k=3, stats=True, aff=True),
    # latency-bound shapes of the batch-1 streaming paths
    dict(name="B1 hrnet 48->48 k3 relu 120x160", B=1, H=120, W=160, cins=[48], Cout=48, k=3, act=1),
    dict(name="B1 hrnet 192->192 k3 relu 30x40", B=1, H=30, W=40, cins=[192], Cout=192, k=3, act=1),
    dict(name="B1 16->16 k3 aff+stats 480x640", B=1, H=480, W=640, cins=[16], Cout=16, k=3, stats=True, aff=True),
    # round 2: residual-as-source 1x1 (h2 enters as a third source with GroupNorm+SiLU on load), statistics without transform
    dict(name="B24 48->16 k1 (x, skip, silu(GN(h2)))", B=24, H=480, W=640, cins=[16, 16, 16], Cout=16, k=1, aff_last=True),
    dict(name="B24 16->16 k3 stats", B=24, H=480, W=640, cins=[16], Cout=16, k=3, stats=True),
    # tensor-bound shapes (TernausNet16's widest layers): the ncu tensor-pipe captures of profiles/r02_ncu_kernels_summary.json
    dict(name="B2 ternaus 768->512 k3 relu 120x160", B=2, H=120, W=160, cins=[512, 256], Cout=512, k=3, act=1),
    dict(name="B2 ternaus 256->256 k3 relu 240x320", B=2, H=240, W=320, cins=[256], Cout=256, k=3, act=1),
    # the low-resolution GroupNorm-on-load layers WITHOUT the transform (direct mode): what materialising silu(GN(.)) first would buy
    dict(name="B24 128->128 k3 stats (direct)", B=24, H=60, W=80, cins=[128], Cout=128, k=3, stats=True),
    dict(name="B24 64->64 k3 stats (direct)", B=24, H=120, W=160, cins=[64], Cout=64, k=3, stats=True),
    dict(name="B24 32->32 k3 stats (direct)", B=24, H=240, W=320, cins=[32], Cout=32, k=3, stats=True),
]


def run(case, iters):
    dev = torch.device("cuda")
    dt = "fp16"
    tdtype = engine._DTYPES[dt][0]
    c = dict(case)
    name = c.pop("name")
    B, H, W, cins, Cout, k = c["B"], c["H"], c["W"], c["cins"], c["Cout"], c["k"]
    ups, act = c.get("ups", 1), c.get("act", 0)
    packer = engine.WeightPacker(dev, dt)
    arena = engine.Arena(dev)
    bld = engine.Builder(dev, dt, packer, arena)
    acts = []
    for ci in cins:
        t = torch.randn(B, (ci + 7) // 8, H, W, 8, device=dev).to(tdtype)
        a = Act(t, ci)
        if c.get("aff") or (c.get("aff_last") and len(acts) == len(cins) - 1):
            aff = torch.zeros(B, ((ci + 7) // 8) * 8, 2, device=dev)
            aff[..., 0] = 1.0
            a = a.with_affine(aff)
        acts.append(a)
    cin = sum(cins)
    w = torch.randn(Cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    Ho, Wo = H * ups, W * ups
    res = None
    if c.get("res"):
        res = Act(torch.randn(B, (Cout + 7) // 8, Ho, Wo, 8, device=dev).to(tdtype), Cout)
        aff = torch.zeros(B, ((Cout + 7) // 8) * 8, 2, device=dev)
        aff[..., 0] = 1.0
        res = res.with_affine(aff)
    out_nchw = torch.empty(B, Cout, Ho, Wo, device=dev) if c.get("nchw") else None
    out, st, info, io = bld.conv("c", acts, w, k, bias=torch.zeros(Cout, device=dev), pad=k // 2, upsample=ups, act=act,
                                 residual=res, want_stats=bool(c.get("stats")), out_nchw=out_nchw, out_c8=out_nchw is None)
    prog = bld.prog
    if "--graph" in sys.argv:      # 20 back-to-back launches of the layer as ONE CUDA-graph launch: no host work in between
        for _ in range(3):
            prog.run()
        rep = engine.Program(dev, dt)
        for _ in range(20):
            rep.extend(prog)
        g = rep.capture()
        g.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.launch()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000.0 / (iters * 20)
        print(json.dumps({"case": name + " (graph x20)", "us": round(us, 2)}), flush=True)
        return {"case": name, "us": us}
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        prog.run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000.0 / iters
    m = prog.meta[0]
    r = {"case": name, "us": round(us, 2), "gbs": round(m["bytes"] / us * 1e-3, 1), "tflops": round(m["flops"] / us * 1e-6, 2),
         "tile": [info.tile_h, info.tile_w], "R": info.runs, "kst": info.kstages, "nst": info.nstages, "grid": info.grid,
         "smem": info.smem_bytes}
    print(json.dumps(r), flush=True)
    return r


def main():
    args = [a for a in sys.argv[1:] if a != "--graph"]
    iters = 20
    if "--iters" in args:
        i = args.index("--iters")
        iters = int(args[i + 1])
        del args[i:i + 2]
    idx = [int(a) for a in args] or range(len(CASES))
    out = [run(CASES[i], iters) for i in idx]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_bench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
