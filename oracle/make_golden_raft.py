"""Writes tests/golden/raft_128x160.npz: torchvision raft_large (seeded init, oracle/raft_oracle.py) on the synthetic frame pair
of raft_oracle.frames -- the final flow of 12 updates, the flow of the video script's call site on 256x320 frames, and a few
intermediate statistics.  Run in the authoring container: python oracle/make_golden_raft.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import raft_oracle as RO  # noqa: E402


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    net = RO.build(0)
    a, b = RO.frames(2, 128, 160)
    f12 = RO.flow(net, a, b)
    f1 = RO.flow(net, a, b, num_flow_updates=1)
    va, vb = RO.frames(1, 256, 320, seed=1)
    vf = RO.video_flow(net, va, vb)
    n_params = sum(p.numel() for p in net.parameters())
    wsum = float(sum(p.double().sum() for p in net.parameters()))
    out = os.path.join(ROOT, "tests", "golden", "raft_128x160.npz")
    np.savez_compressed(out, flow12=f12.numpy().astype(np.float16), flow1=f1.numpy().astype(np.float16),
                        video=vf[:, :, ::4, ::4].numpy().astype(np.float16), n_params=np.int64(n_params), wsum=np.float64(wsum),
                        torchvision=np.bytes_(__import__("torchvision").__version__))
    print(out, os.path.getsize(out), "bytes; params", n_params, "weight sum", wsum, "max |flow|", float(f12.abs().max()))


if __name__ == "__main__":
    main()
