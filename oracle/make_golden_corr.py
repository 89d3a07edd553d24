"""Golden vectors of the correlation cost volume, written by the REFERENCE's own CUDA kernels.

TEST INFRASTRUCTURE ONLY.  Runs on a GPU box (the reference kernels are CUDA-only, models/unflow_correlation.py:331-332):
the kernel strings extracted by oracle/make_corr_ref.py are compiled with NVRTC and launched exactly as
`_FunctionCorrelation.forward/backward` launch them (oracle/corr_ref_nvrtc.py) on inputs regenerated bit-identically
from oracle/synth.py.  Output: gpurun_out/corr_ref_golden.npz, committed as tests/golden/corr_ref.npz.

    gpurun -- python -m oracle.make_golden_corr
"""
import os

import numpy as np
import torch

from oracle import corr_ref_nvrtc as R, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (tag, B, C, H, W): channel counts below / not a multiple of / above the 32 lanes of the reference's reduction
FWD_CASES = [("fwd_c40", 2, 40, 12, 20), ("fwd_c7", 1, 7, 9, 11), ("fwd_c256", 1, 256, 6, 10)]
BWD_CASES = [("bwd_c8", 1, 8, 10, 12), ("bwd_c35", 2, 35, 7, 9)]


def inputs(tag, B, Cc, H, W):
    return synth.normal(tag + "/a", (B, Cc, H, W), 3), synth.normal(tag + "/b", (B, Cc, H, W), 4)


def main():
    assert R.available(), "needs a GPU, cuda-python and baseline/_ref/unflow_correlation_kernels.json"
    out = {}
    for tag, B, Cc, H, W in FWD_CASES:
        a, b = inputs(tag, B, Cc, H, W)
        y = R.forward(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        out[tag] = y.cpu().numpy()
    for tag, B, Cc, H, W in BWD_CASES:
        a, b = inputs(tag, B, Cc, H, W)
        g = synth.normal(tag + "/g", (B, 441, H, W), 5)
        g1, g2 = R.backward(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(g).cuda())
        out[tag + "/grad_first"], out[tag + "/grad_second"] = g1.cpu().numpy(), g2.cpu().numpy()
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "corr_ref_golden.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
