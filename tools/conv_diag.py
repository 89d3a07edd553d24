"""GPU bring-up diagnostic: runs single fused-conv cases through libmfcnet_b200.so and prints
max-abs error against torch fp32 conv2d on inputs rounded to the storage dtype (so that the
number isolates indexing / protocol bugs from quantisation).  Usage: python tools/conv_diag.py
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402
from mfcnet_tracker_b200 import engine  # noqa: E402
from mfcnet_tracker_b200.engine import Act  # noqa: E402


def to_c8(x, tdtype):
    B, Cc, H, W = x.shape
    ch = (Cc + 7) // 8
    t = torch.zeros(B, ch * 8, H, W, device=x.device, dtype=torch.float32)
    t[:, :Cc] = x
    return t.view(B, ch, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(tdtype)


def from_c8(t, Cc):
    B, ch, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(B, ch * 8, H, W)[:, :Cc]


def run_case(dt, B, H, W, cins, Cout, k, stride=1, pad=None, ups=1, act=0, bias=True, scale=False, residual=False,
             res_affine=False, src_affine=False, stats=False, head=0, lo=False, seed=0, kw=None, pad_yx=None):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    tdtype = engine._DTYPES[dt][0]
    pad = k // 2 if pad is None else pad
    packer = engine.WeightPacker(dev, dt)
    arena = engine.Arena(dev)
    bld = engine.Builder(dev, dt, packer, arena)
    xs, acts, ref_in = [], [], []
    for ci in cins:
        x = torch.randn(B, ci, H, W, generator=g).to(dev)
        xq = x.to(tdtype).float()
        a = Act(to_c8(x, tdtype), ci)
        r = xq
        if src_affine:
            aff = torch.zeros(B, ((ci + 7) // 8) * 8, 2, device=dev)
            aff[:, :ci, 0] = (1.0 + 0.2 * torch.randn(B, ci, generator=g)).to(dev)
            aff[:, :ci, 1] = (0.3 * torch.randn(B, ci, generator=g)).to(dev)
            a = a.with_affine(aff)
            r = F.silu(xq * aff[:, :ci, 0, None, None] + aff[:, :ci, 1, None, None]).to(tdtype).float()
        acts.append(a)
        ref_in.append(r)
    cin = sum(cins)
    kw_ = k if kw is None else kw
    w = (torch.randn(Cout, cin, k, kw_, generator=g) / (cin * k * kw_) ** 0.5).to(dev)
    b = (0.1 * torch.randn(Cout, generator=g)).to(dev) if bias else None
    sc = (1.0 + 0.2 * torch.randn(Cout, generator=g)).to(dev) if scale else None
    xin = torch.cat(ref_in, 1)
    if ups == 2:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    wq = w.to(tdtype).float()
    ref = F.conv2d(xin, wq, None, stride=stride, padding=pad if pad_yx is None else pad_yx)
    if sc is not None:
        ref = ref * sc[None, :, None, None]
    if b is not None:
        ref = ref + b[None, :, None, None]
    res_act = None
    Ho, Wo = ref.shape[-2:]
    if residual:
        r = torch.randn(B, Cout, Ho, Wo, generator=g).to(dev)
        res_act = Act(to_c8(r, tdtype), Cout)
        rq = r.to(tdtype).float()
        if res_affine:
            aff = torch.zeros(B, ((Cout + 7) // 8) * 8, 2, device=dev)
            aff[:, :Cout, 0] = (1.0 + 0.2 * torch.randn(B, Cout, generator=g)).to(dev)
            aff[:, :Cout, 1] = (0.3 * torch.randn(B, Cout, generator=g)).to(dev)
            res_act = res_act.with_affine(aff)
            rq = F.silu(rq * aff[:, :Cout, 0, None, None] + aff[:, :Cout, 1, None, None])
        ref = ref + rq
    if act:
        ref = F.relu(ref)
    hd = None
    if head:   # fused fp32 1x1 head: out_nchw holds `head` channels = hw @ v + hb
        hw = (torch.randn(head, Cout, generator=g) / Cout ** 0.5).to(dev)
        hb = (0.1 * torch.randn(head, generator=g)).to(dev)
        hd = (hw, hb)
    out_nchw = torch.full((B, head or Cout, Ho, Wo), float("nan"), device=dev)
    out, st, info, io = bld.conv("c", acts, w, k, bias=b, scale=sc, stride=stride, pad=pad, upsample=ups, act=act,
                                 residual=res_act, want_stats=stats, out_nchw=out_nchw, head=hd, want_lo=lo, kw=kw, pad_yx=pad_yx)
    bld.prog.run()
    torch.cuda.synchronize()
    e_c8 = (from_c8(out.t, Cout) - ref).abs().max().item()
    e_hilo = (from_c8(out.t, Cout) + from_c8(out.lo.t, Cout) - out_nchw).abs().max().item() if (lo and not head) else None
    if head:
        ref = torch.einsum("nc,bchw->bnhw", hw, ref) + hb[None, :, None, None]
    e_nchw = (out_nchw - ref).abs().max().item()
    res = {"dt": dt, "B": B, "H": H, "W": W, "cins": cins, "Cout": Cout, "k": k, "s": stride, "ups": ups,
           "tile": [info.tile_h, info.tile_w], "R": info.runs, "kst": info.kstages, "nb": info.nb, "nblk": info.nblk,
           "err_nchw": e_nchw, "err_c8": e_c8, "ref_absmax": ref.abs().max().item()}
    if e_hilo is not None:     # hi + lo planes against the kernel's own fp32 output: ~2^-22 relative
        res["err_hilo"] = e_hilo
    if stats:
        s = st.double().sum(1)  # [B, cpad, 2]
        rs = torch.stack([ref.double().sum((2, 3)), (ref.double() ** 2).sum((2, 3))], -1)
        res["err_stats_rel"] = ((s[:, :Cout] - rs).abs() / (rs.abs() + 1.0)).max().item()
    return res


CASES = [
    dict(B=1, H=16, W=16, cins=[16], Cout=16, k=1),
    dict(B=1, H=16, W=16, cins=[16], Cout=16, k=3),
    dict(B=2, H=32, W=48, cins=[16], Cout=16, k=3, stats=True),
    dict(B=2, H=32, W=48, cins=[3], Cout=16, k=7),
    dict(B=1, H=40, W=56, cins=[32], Cout=16, k=3, src_affine=True),
    dict(B=2, H=32, W=48, cins=[16, 16], Cout=16, k=3, stats=True),
    dict(B=2, H=32, W=48, cins=[16, 16], Cout=16, k=1, residual=True, res_affine=True),
    dict(B=2, H=32, W=48, cins=[16], Cout=32, k=2, stride=2, pad=0),
    dict(B=2, H=32, W=48, cins=[32], Cout=64, k=3, stride=2),
    dict(B=2, H=16, W=24, cins=[64], Cout=32, k=3, ups=2),
    dict(B=1, H=8, W=12, cins=[128, 64], Cout=128, k=3, stats=True),
    dict(B=1, H=8, W=12, cins=[128], Cout=256, k=3, act=1, scale=True),
    dict(B=2, H=48, W=64, cins=[5, 5, 5, 7], Cout=15, k=11, act=1, scale=True, bias=False),
    dict(B=1, H=48, W=64, cins=[5, 5, 5, 5, 5, 13], Cout=25, k=11, act=1, scale=True, bias=False),
    dict(B=1, H=64, W=96, cins=[16], Cout=5, k=1),
    dict(B=1, H=480, W=640, cins=[16], Cout=16, k=3, stats=True),
    dict(B=2, H=480, W=640, cins=[16, 16], Cout=16, k=3, src_affine=True),
    dict(B=1, H=480, W=640, cins=[5, 5, 5, 7], Cout=15, k=11, act=1, scale=True, bias=False),
    # fast epilogue paths: full-width tiles of halo-free convs (residual-as-source 1x1, fused output layer, 2x2 s2),
    # sliding mode with shift-initialised accumulators (+ReLU, + statistics, odd sizes)
    dict(B=2, H=32, W=48, cins=[16, 16, 16], Cout=16, k=1, src_affine=True),
    dict(B=3, H=60, W=80, cins=[16, 16, 16], Cout=5, k=1),
    dict(B=2, H=480, W=640, cins=[16, 16, 16], Cout=16, k=1),
    dict(B=2, H=480, W=640, cins=[16, 16, 16], Cout=5, k=1),
    dict(B=2, H=240, W=320, cins=[16], Cout=16, k=2, stride=2, pad=0),
    dict(B=2, H=33, W=47, cins=[15], Cout=15, k=3, act=1, scale=True),
    dict(B=2, H=37, W=131, cins=[16], Cout=16, k=3, stats=True, src_affine=True),
    dict(B=1, H=5, W=7, cins=[16], Cout=16, k=3, stats=True),
    dict(B=2, H=480, W=640, cins=[15], Cout=5, k=1, bias=False),
    dict(B=2, H=33, W=47, cins=[15], Cout=15, k=3, act=1, scale=True, head=5),
    dict(B=2, H=480, W=640, cins=[15], Cout=15, k=3, act=1, scale=True, head=5),
    dict(B=1, H=64, W=96, cins=[16, 16], Cout=16, k=1, head=3),
    dict(B=2, H=33, W=47, cins=[25], Cout=25, k=3, act=1, scale=True, lo=True),
    dict(B=2, H=40, W=56, cins=[16], Cout=16, k=3, lo=True),
    # rectangular kernels with per-axis padding (RAFT's ConvGRU: 1x5 / 5x1 over a four-source concat)
    dict(B=2, H=30, W=40, cins=[128, 128, 126, 2], Cout=256, k=1, kw=5, pad_yx=(0, 2)),
    dict(B=2, H=30, W=40, cins=[128, 128, 126, 2], Cout=128, k=5, kw=1, pad_yx=(2, 0)),
    dict(B=1, H=17, W=23, cins=[16], Cout=16, k=1, kw=5, pad_yx=(0, 2), act=1),
    dict(B=1, H=17, W=23, cins=[24], Cout=40, k=5, kw=1, pad_yx=(2, 0)),
    dict(B=1, H=16, W=20, cins=[128], Cout=64, k=1, stride=2, pad=0),
]


def main():
    out = []
    dts = sys.argv[1:] or ["fp16", "bf16"]
    for dt in dts:
        for c in CASES:
            try:
                r = run_case(dt, **c)
            except Exception as e:  # keep going: the table is the diagnostic
                r = dict(c, dt=dt, error=repr(e)[:300])
                try:
                    torch.cuda.synchronize()
                except Exception as e2:
                    r["sync_error"] = repr(e2)[:200]
                    print(json.dumps(r), flush=True)
                    out.append(r)
                    break
            print(json.dumps(r), flush=True)
            out.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_diag.json", "w") as f:
        json.dump(out, f, indent=1)
    bad = [r for r in out if "error" in r or r.get("err_nchw", 1) > 5e-3 * max(1.0, r.get("ref_absmax", 1.0))
           or r.get("err_hilo", 0.0) > (2e-6 if r.get("dt") == "fp16" else 4e-5) * max(1.0, r.get("ref_absmax", 1.0))]
    print("conv_diag: %d cases, %d bad" % (len(out), len(bad)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
