"""GPU bring-up diagnostic of the RAFT port: compares the engine's intermediate tensors with torchvision's (CPU, fp32)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as m  # noqa: E402
from oracle import raft_oracle as RO  # noqa: E402


def from_c8(t, C_):
    B, ch, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(B, ch * 8, H, W)[:, :C_].cpu()


def rel(a, b):
    return float((a - b).abs().max()), float(b.abs().max())


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 160)
    B = 1
    tv = RO.build(0)
    a, b = RO.frames(B, H, W)
    mine = m.raft_large()
    mine.load_state_dict(tv.state_dict())
    mine = mine.cuda().eval()
    with torch.no_grad():
        fm = tv.feature_encoder(torch.cat([a, b], 0))
        ctx = tv.context_encoder(a)
        hid, con = torch.tanh(ctx[:, :128]), torch.relu(ctx[:, 128:])
        tv.corr_block.build_pyramid(fm[:B], fm[B:])
        from torchvision.models.optical_flow._utils import make_coords_grid
        c0 = make_coords_grid(B, H // 8, W // 8)
        look0 = tv.corr_block.index_pyramid(c0)
        ref1 = tv(a, b, num_flow_updates=1)[-1]
        ref12 = tv(a, b, num_flow_updates=12)[-1]
        out1 = mine(a.cuda(), b.cuda(), num_flow_updates=1)[-1].cpu()
        P = mine._plans[(B, H, W)]
        torch.cuda.synchronize()
        print("fmaps      err %.3e of %.3e" % rel(P["fmaps"].cpu(), fm))
        print("context    err %.3e of %.3e" % rel(from_c8(P["ctx"], 128), con))
        hw = (H // 8) * (W // 8)
        for l in range(4):
            print("pyramid %d  err %.3e of %.3e" % ((l,) + rel(P["vol"][l].cpu().reshape(tv.corr_block.corr_pyramid[l].shape), tv.corr_block.corr_pyramid[l])))
        print("lookup0    err %.3e of %.3e" % rel(from_c8(P["corr"], 324), look0))
        print("flow lowres after 1 it: max |flow| %.3e" % float(P["flow"].abs().max()))
        print("flow(1 it) err %.3e of %.3e  mean err %.3e" % (rel(out1, ref1) + (float((out1 - ref1).abs().mean()),)))
        out12 = mine(a.cuda(), b.cuda(), num_flow_updates=12)[-1].cpu()
        print("flow(12)   err %.3e of %.3e  mean err %.3e" % (rel(out12, ref12) + (float((out12 - ref12).abs().mean()),)))
        for n in (2, 4, 8):
            o = mine(a.cuda(), b.cuda(), num_flow_updates=n)[-1].cpu()
            r = tv(a, b, num_flow_updates=n)[-1]
            print("flow(%d)    err %.3e of %.3e  mean err %.3e" % ((n,) + rel(o, r) + (float((o - r).abs().mean()),)))
        # timing
        x, y = a.cuda(), b.cuda()
        for _ in range(3):
            mine(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            mine(x, y)
        e1.record()
        torch.cuda.synchronize()
        print("engine: %.3f ms per flow (B=%d, %dx%d, 12 updates)" % (e0.elapsed_time(e1) / 10, B, H, W))


if __name__ == "__main__":
    main()
