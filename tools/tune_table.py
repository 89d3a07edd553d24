"""Writes the conv tuning table (mfcnet-tracker_b200/tuning/b200.tbl) on a B200.

Every conv geometry of the workloads below is measured once with mfc_conv2d_autotune (MFC_CONV_TUNE=1: the planner's
shortlisted tilings are timed on the device with the layer's real buffers) and the winners are exported as text.  The
committed table makes tilings -- and with them the fp32 summation order -- a pure function of the geometry: every later
process plans from the table (else the cost model) and computes identical bits.

    gpurun -- python tools/tune_table.py [--fresh] [--only resunet,hrnet,...]     # -> gpurun_out/b200.tbl
"""
import os
import sys
import time

os.environ["MFC_CONV_TUNE"] = "1"
os.environ.setdefault("MFC_CONV_TUNE_REPS", "10")
if "--fresh" in sys.argv:
    os.environ["MFC_CONV_TABLE"] = "0"

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfcnet_tracker_b200 as M  # noqa: E402

H, W, N = 480, 640, 5


def inputs(B, K, h=H, w=W):
    xs = [torch.randn(B, 3, h, w, device="cuda") for _ in range(K)]
    fl = [torch.randn(B, 2, h, w, device="cuda") for _ in range(K - 1)]
    dp = [torch.rand(B, 1, h, w, device="cuda") for _ in range(K)]
    return xs, fl, dp


def window(cls, K, B, h=H, w=W):
    net = cls(N, K, optflow_inputs=True, depth_inputs=True).cuda().eval()
    xs, fl, dp = inputs(B, K, h, w)
    with torch.no_grad():
        net(xs, optflow=fl, depth=dp)
    torch.cuda.synchronize()


def stream(cls, K, B):
    net = cls(N, K, optflow_inputs=True, depth_inputs=True).cuda().eval()
    run = M.StreamingMFCNet(net, H, W, batch=B)
    xs, fl, dp = inputs(B, K)
    with torch.no_grad():
        for _ in range(K + 1):
            run.step(xs[0], fl, dp)
    torch.cuda.synchronize()


def sfc(net, B, h=H, w=W):
    net = net.cuda().eval()
    with torch.no_grad():
        net(torch.randn(B, 3, h, w, device="cuda"))
    torch.cuda.synchronize()


def raft(h, w, B):
    net = M.raft_large().cuda().eval()
    with torch.no_grad():
        net(torch.randn(B, 3, h, w, device="cuda").clamp_(-2, 2), torch.randn(B, 3, h, w, device="cuda").clamp_(-2, 2))
    torch.cuda.synchronize()


def sflow(K, B):
    net = M.raft_large().cuda().eval()
    sf = M.StreamingFlow(net, K, H, W, batch=B)
    with torch.no_grad():
        for _ in range(2):
            sf.step(torch.randn(B, 3, H, W, device="cuda").clamp_(-2, 2))
    torch.cuda.synchronize()


WORK = {
    # BASELINE configs[1] (bench.py headline) and its Basic variant: 24-frame SFC sub-batch + 8 fusion windows
    "bench": lambda: [window(M.ResUNetMultiLarge, 3, 8), window(M.ResUNetMultiBasic, 3, 8)],
    # streaming with the feature ring, B clips in lock step (bench.py's `streaming` record, tools/bench_stream.py)
    "stream": lambda: [stream(M.ResUNetMultiLarge, 3, b) for b in (1, 4, 8)] + [stream(M.ResUNetMultiBasic, 3, 8)],
    # BASELINE configs[0]: single-frame ResUNet
    "resunet": lambda: [sfc(M.ResUnet_VB(3, 16, out_dim=N), b) for b in (1, 4)],
    # BASELINE configs[3]: HRNet MFCNet K=5 streaming
    "hrnet": lambda: [stream(M.HRNetMultiLarge, 5, b) for b in (1, 4)] + [sfc(M.HighResolutionNet(num_classes=N), 1)],
    "ternaus": lambda: [sfc(M.TernausNet16(num_classes=N, num_filters=64), b) for b in (1, 2)],
    # RAFT-large at the video script's operating point (half-size frames), pair batches of bench.py / tools/bench_raft.py, the
    # streaming form (encoders on B new frames, updates on B (K-1) pairs) and the shapes of tests/test_gpu_raft.py
    "raft": lambda: [raft(240, 320, b) for b in (1, 2, 8)] + [sflow(3, b) for b in (1, 8, 16)] + [raft(128, 160, b) for b in (1, 2)] +
                    [raft(480, 640, 1)],
    # the shapes of smoke() and of the full-size parity tests
    "tests": lambda: [window(M.ResUNetMultiBasic, 3, 1, 96, 128), window(M.ResUNetMultiLarge, 3, 1), window(M.ResUNetMultiLarge, 3, 2, 96, 128),
                      window(M.HRNetMultiLarge, 5, 1)],
}


def main():
    only = None
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1].split(",")
    for name, fn in WORK.items():
        if only and name not in only:
            continue
        t0 = time.time()
        fn()
        print("%-8s tuned in %.1f s" % (name, time.time() - t0), flush=True)
    text = M.abi.export_table()
    out = os.path.join(ROOT, "gpurun_out", "b200.tbl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    head = ("# conv tuning table, measured on %s by tools/tune_table.py (mfc_conv2d_autotune)\n"
            "# B Hin Win Hout Wout Cout kh kw stride pad upsample chunks affine flags dtype out_stride : TH TW slide CBc NB nstages\n"
            % torch.cuda.get_device_name(0))
    with open(out, "w") as f:
        f.write(head + text)
    print("%d entries -> %s" % (text.count("\n"), out))


if __name__ == "__main__":
    main()
