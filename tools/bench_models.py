"""Secondary GPU measurements (not the bench.py headline): frames/s of the SFC networks and of the
streaming MFCNet runner (BASELINE configs 1 and 4 shapes).  Usage: python tools/bench_models.py [--iters N]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 10
    H, W, N = 480, 640, 5
    out = []
    torch.manual_seed(0)
    with torch.no_grad():
        for name, ctor, B in (("ResUNet-16 SFC", lambda: M.ResUnet_VB(3, 16, out_dim=N), 4),
                              ("HRNet-W48 SFC", lambda: M.HighResolutionNet(N), 4),
                              ("HRNet-W48 SFC", lambda: M.HighResolutionNet(N), 1),
                              ("TernausNet16 SFC (762 GFLOP/frame)", lambda: M.TernausNet16(num_classes=N, num_filters=64, pretrained=False), 2)):
            net = ctor().cuda().eval()
            x = torch.randn(B, 3, H, W, device="cuda")
            ms = timeit(lambda: net(x), iters)
            prog = net._plans[(B, H, W)][0]
            flops = sum(m.get("flops", 0) for m in prog.meta)
            r = {"what": name, "batch": B, "ms": round(ms, 3), "frames_per_s": round(B * 1000.0 / ms, 1), "launches": prog.n_kernels,
                 "conv_TFLOPs": round(flops / ms * 1e-9, 1)}
            print(json.dumps(r), flush=True)
            out.append(r)
            del net
        # config 4 shape: HRNet MFCNet, K=5, streaming with the feature ring (one new frame per step)
        for name, cls, K in (("HRNetMulti-Large K=5 streaming", M.HRNetMultiLarge, 5), ("ResUNetMulti-Large K=3 streaming", M.ResUNetMultiLarge, 3)):
            net = cls(N, K, optflow_inputs=True, depth_inputs=True).cuda().eval()
            from mfcnet_tracker_b200.stream import StreamingMFCNet
            run = StreamingMFCNet(net, H, W)
            frame = torch.randn(1, 3, H, W, device="cuda")
            flows = [torch.randn(1, 2, H, W, device="cuda") for _ in range(K - 1)]
            depths = [torch.rand(1, 1, H, W, device="cuda") for _ in range(K)]
            for _ in range(K):
                run.step(frame, flows, depths)
            ms = timeit(lambda: run.step(frame, flows, depths), iters * 4)
            r = {"what": name, "ms_per_frame": round(ms, 3), "frames_per_s": round(1000.0 / ms, 1), "launches_per_frame": run.launches_per_frame}
            print(json.dumps(r), flush=True)
            out.append(r)
            del net, run
        # UnFlow (FlowNetC + 2 x FlowNetS) at the reference's operating point: 384x1280 frames -> 48x160 cost volume
        net = M.UnFlow().cuda().eval()
        for p_ in net.parameters():                       # random init with a sane scale (LeakyReLU stack)
            if p_.dim() == 4:
                torch.nn.init.kaiming_normal_(p_, a=0.1)
        for n in range(3):
            net.moduleFlownets[n].moduleUpconv.moduleTwoOut.weight.data.mul_(0.05)
        a, b = torch.rand(1, 3, 384, 1280, device="cuda"), torch.rand(1, 3, 384, 1280, device="cuda")
        ms = timeit(lambda: net(a, b), iters)
        P = net._plans[(1, 384, 1280)]
        flops = sum(m.get("flops", 0) for pr in P["progs"] for m in pr.meta)
        r = {"what": "UnFlow 384x1280 (3 stacked nets + correlation)", "batch": 1, "ms": round(ms, 3), "pairs_per_s": round(1000.0 / ms, 1),
             "launches": sum(pr.n_kernels for pr in P["progs"]) + 10, "conv_TFLOPs": round(flops / ms * 1e-9, 1)}
        try:   # the same network written with stock torch ops (cuDNN fp32) on this GPU, as an informative library baseline
            from oracle import torch_oracle as TO
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            sd = {k: v.detach() for k, v in net.state_dict().items()}
            ms_t = timeit(lambda: TO.unflow_forward(sd, a, b, corr=M.correlation), 3)
            ref = TO.unflow_forward(sd, a, b, corr=M.correlation)
            r["ms_torch_ops_fp32_same_gpu"] = round(ms_t, 2)
            r["max_abs_flow_diff_px"] = float((net(a, b) - ref).abs().max())
            r["ref_flow_absmax_px"] = float(ref.abs().max())
        except Exception as e:   # noqa: BLE001
            r["torch_baseline_error"] = repr(e)[:200]
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_models.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
