"""RAFT-large on the engine vs torchvision's own module (the reference's flow provider, a third-party dependency:
scripts/test_multiframe_segmentation_on_videos_v3.py:264-271,342-350) run on the CPU in fp32 -- through the C ABI.
Tolerances: the network stores activations in fp16 and iterates its GRU 12 times; flows of up to 13 px reproduce to a few
hundredths of a pixel (measured 2.3e-2 max / 6.5e-3 mean on the 128x160 pair), gates below."""
import numpy as np
import pytest
import torch

import mfcnet_tracker_b200 as m
from oracle import raft_oracle as RO

pytestmark = pytest.mark.gpu


def _pair(B, H, W, seed=0):
    tv = RO.build(0)
    mine = m.raft_large()
    mine.load_state_dict(tv.state_dict())
    return tv, mine.cuda().eval(), RO.frames(B, H, W, seed=seed)


def _from_c8(t, C_):
    B, ch, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(B, ch * 8, H, W)[:, :C_].cpu()


def test_flow_matches_torchvision_12_updates():
    tv, mine, (a, b) = _pair(2, 128, 160)
    ref = RO.flow(tv, a, b)
    with torch.no_grad():
        out = mine(a.cuda(), b.cuda())[-1].cpu()
    err = (out - ref).abs()
    assert out.shape == ref.shape == (2, 2, 128, 160)
    assert float(ref.abs().max()) > 5.0                       # a real displacement field, not zeros
    assert float(err.max()) <= 6e-2 and float(err.mean()) <= 1.5e-2, (float(err.max()), float(err.mean()))
    # identical bits from a second call (CUDA-graph replay of the update iteration)
    with torch.no_grad():
        again = mine(a.cuda(), b.cuda())[-1].cpu()
    assert torch.equal(out, again)


def test_pieces_match_torchvision():
    """Encoders, correlation pyramid, lookup, one update and the convex upsampling, each against torchvision's own function."""
    from torchvision.models.optical_flow._utils import make_coords_grid, upsample_flow
    tv, mine, (a, b) = _pair(1, 128, 160)
    with torch.no_grad():
        out1 = mine(a.cuda(), b.cuda(), num_flow_updates=1)[-1].cpu()
        P = mine._plans[(1, 128, 160)]
        torch.cuda.synchronize()
        fm = tv.feature_encoder(torch.cat([a, b], 0))
        ctx = torch.relu(tv.context_encoder(a)[:, 128:])
        got_fm = P["fmaps"].cpu()
        assert float((got_fm - fm).abs().max()) <= 4e-3 * float(fm.abs().max())
        assert float((_from_c8(P["ctx"], 128) - ctx).abs().max()) <= 4e-3 * float(ctx.abs().max())
        # pyramid and lookup from the ENGINE's feature maps: isolates the fp32 kernels (sum order only)
        tv.corr_block.build_pyramid(got_fm[:1], got_fm[1:])
        for l in range(4):
            ref = tv.corr_block.corr_pyramid[l]
            got = P["vol"][l].cpu().reshape(ref.shape)
            assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), l
        look = tv.corr_block.index_pyramid(make_coords_grid(1, 16, 20))
        assert float((_from_c8(P["corr"], 324) - look).abs().max()) <= 2e-3 * float(look.abs().max())   # fp16 storage of the planes
        # convex upsampling of the engine's own low-resolution flow and mask
        up = upsample_flow(P["flow"].cpu(), RAFT_MULT * P["mask"].cpu())
        assert float((out1 - up).abs().max()) <= 1e-5 * max(1.0, float(up.abs().max()))
        ref1 = tv(a, b, num_flow_updates=1)[-1]
        assert float((out1 - ref1).abs().max()) <= 1e-2


RAFT_MULT = 0.25


def test_video_call_site():
    tv, mine, (a, b) = _pair(1, 256, 320, seed=1)
    ref = RO.video_flow(tv, a, b)
    with torch.no_grad():
        out = m.video_flow(mine, a.cuda(), b.cuda()).cpu()
    err = (out - ref).abs()
    assert out.shape == (1, 2, 256, 320)
    assert float(err.max()) <= 1.2e-1 and float(err.mean()) <= 3e-2, (float(err.max()), float(err.mean()))


def test_golden_fixture_on_gpu():
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "raft_128x160.npz"))
    _, mine, (a, b) = _pair(2, 128, 160)
    with torch.no_grad():
        out = mine(a.cuda(), b.cuda())[-1].cpu().numpy()
    assert float(np.abs(out - g["flow12"].astype(np.float32)).max()) <= 8e-2


def test_rejects_bad_sizes():
    _, mine, _ = _pair(1, 128, 160)
    with pytest.raises(ValueError):
        mine(torch.zeros(1, 3, 100, 160, device="cuda"), torch.zeros(1, 3, 100, 160, device="cuda"))
    with pytest.raises(ValueError):
        mine(torch.zeros(1, 3, 64, 64, device="cuda"), torch.zeros(1, 3, 64, 64, device="cuda"))


def test_streaming_flow_matches_the_per_pair_calls():
    """StreamingFlow (encoders once per new frame, feature-map ring) against video_flow on every (current, earlier) pair."""
    tv, mine, _ = _pair(1, 128, 160)
    B, K, H, W = 2, 3, 256, 320
    seq = [RO.frames(B, H, W, seed=10 + t)[0].cuda() for t in range(4)]
    sf = m.StreamingFlow(mine, K, H, W, batch=B)
    with torch.no_grad():
        for t, x in enumerate(seq):
            got = sf.step(x)
            assert len(got) == K - 1 and got[0].shape == (B, 2, H, W)
            for j in range(K - 1):
                prev = seq[max(0, t - 1 - j)] if t > 0 else x
                ref = m.video_flow(mine, x, prev)
                err = (got[j] - ref).abs()
                assert float(err.max()) <= 8e-2 and float(err.mean()) <= 1e-2, (t, j, float(err.max()), float(err.mean()))
        ref_tv = RO.video_flow(tv, seq[3].cpu(), seq[1].cpu())        # and against torchvision itself for one pair
        err = (got[1].cpu() - ref_tv).abs()
        assert float(err.max()) <= 1.5e-1 and float(err.mean()) <= 3e-2, (float(err.max()), float(err.mean()))


# ---------------------------------------------------------------- the kernels behind mfc_pointwise / mfc_raft_op, one by one
def _to_c8(x, dtype=torch.float16):
    B, C_, H, W = x.shape
    ch = (C_ + 7) // 8
    t = torch.zeros(B, ch * 8, H, W, device=x.device)
    t[:, :C_] = x
    return t.view(B, ch, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(dtype)


def _call_raft(kind, ptrs, B, h, w, **kw):
    import ctypes as C
    g = m.abi.MfcRaftArgs()
    ptrs = list(ptrs) + [None] * (6 - len(ptrs))
    g.p0, g.p1, g.p2, g.p3, g.p4, g.p5 = [None if t is None else t.data_ptr() for t in ptrs]
    g.kind, g.B, g.h, g.w, g.dtype = kind, B, h, w, m.abi.MFC_F16
    g.C, g.levels, g.radius, g.scale = kw.get("C", 0), kw.get("levels", 0), kw.get("radius", 0), kw.get("scale", 0.0)
    m.abi.check(m.abi.load().mfc_raft_op(C.byref(g), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


def test_lookup_with_large_and_out_of_range_flow():
    """CorrBlock.index_pyramid at coords0 + flow, flows up to +-25 px (windows leave the maps on every side and level)."""
    from torchvision.models.optical_flow.raft import CorrBlock
    from torchvision.models.optical_flow._utils import make_coords_grid
    torch.manual_seed(3)
    B, h, w = 2, 16, 24
    f1, f2 = torch.randn(B, 64, h, w), torch.randn(B, 64, h, w)
    flow = 25.0 * (torch.rand(B, 2, h, w) - 0.5) * torch.tensor([2.0, 1.0]).view(1, 2, 1, 1)
    cb = CorrBlock(num_levels=4, radius=4)
    cb.build_pyramid(f1, f2)
    ref = cb.index_pyramid(make_coords_grid(B, h, w) + flow)
    vol = [torch.empty(B * h * w, h >> l, w >> l, device="cuda") for l in range(4)]
    _call_raft(m.abi.RAFT_CORR_VOLUME, [f1.cuda(), f2.cuda(), vol[0]], B, h, w, C=64, scale=1.0 / 8.0)
    for l in range(3):
        _call_raft(m.abi.RAFT_POOL, [vol[l], vol[l + 1]], B * h * w, h >> l, w >> l)
    for l in range(4):
        want = cb.corr_pyramid[l].reshape(vol[l].shape)
        assert float((vol[l].cpu() - want).abs().max()) <= 2e-5 * float(want.abs().max()), l
    out = torch.zeros(B, 41, h, w, 8, dtype=torch.float16, device="cuda")
    _call_raft(m.abi.RAFT_LOOKUP, vol + [flow.cuda(), out], B, h, w, levels=4, radius=4)
    got = _from_c8(out, 324)
    assert float((got - ref).abs().max()) <= 1.5e-3 * float(ref.abs().max())       # fp16 storage of the planes
    assert float(_from_c8(out, 328)[:, 324:].abs().max()) == 0.0                    # padding channels stay zero


def test_upsample_resize_and_flow_add():
    from torchvision.models.optical_flow._utils import upsample_flow
    torch.manual_seed(4)
    B, h, w = 2, 10, 14
    flow, mask = 3.0 * torch.randn(B, 2, h, w), 4.0 * torch.randn(B, 576, h, w)
    out = torch.empty(B, 2, 8 * h, 8 * w, device="cuda")
    _call_raft(m.abi.RAFT_UPSAMPLE, [flow.cuda(), mask.cuda(), out], B, h, w, scale=0.25)
    ref = upsample_flow(flow, 0.25 * mask)
    assert float((out.cpu() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    big = torch.empty(B, 2, 37, 51, device="cuda")
    _call_raft(m.abi.RAFT_RESIZE_AC, [flow.cuda(), None, big], B, h, w, C=2, levels=37, radius=51, scale=2.0)
    ref = torch.nn.functional.interpolate(flow / 0.5, size=(37, 51), mode="bilinear", align_corners=True)
    assert float((big.cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    acc, delta = flow.clone().cuda(), torch.randn(B, 2, h, w).cuda()
    _call_raft(m.abi.RAFT_FLOW_ADD, [acc, delta], B, h, w)
    assert torch.equal(acc.cpu(), flow + delta.cpu())


def test_pointwise_kinds_match_torch():
    import ctypes as C
    torch.manual_seed(5)
    B, Cc, h, w = 2, 24, 9, 13

    def call(kind, a, out, chunks, **kw):
        g = m.abi.MfcPointwiseArgs()
        g.a, g.out, g.kind, g.B, g.chunks, g.pixels, g.dtype = a.data_ptr(), out.data_ptr(), kind, B, chunks, h * w, m.abi.MFC_F16
        for k in ("a_aff", "r", "r_aff", "out2"):
            setattr(g, k, None if kw.get(k) is None else kw[k].data_ptr())
        g.relu_a, g.relu_out = int(kw.get("relu_a", 0)), int(kw.get("relu_out", 0))
        m.abi.check(m.abi.load().mfc_pointwise(C.byref(g), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()

    q = lambda t: t.half().float()                                 # the values the kernel reads
    a, r = torch.randn(B, Cc, h, w), torch.randn(B, Cc, h, w)
    aa = torch.stack([1.0 + 0.3 * torch.randn(B, Cc), 0.2 * torch.randn(B, Cc)], -1)
    ra = torch.stack([1.0 + 0.3 * torch.randn(B, Cc), 0.2 * torch.randn(B, Cc)], -1)
    out = torch.zeros(B, 3, h, w, 8, dtype=torch.float16, device="cuda")
    call(m.abi.PW_AFFINE_ADD, _to_c8(a.cuda()), out, 3, a_aff=aa.cuda().contiguous(), r=_to_c8(r.cuda()), r_aff=ra.cuda().contiguous(),
         relu_a=1, relu_out=1)
    ref = torch.relu(torch.relu(q(a) * aa[..., 0, None, None] + aa[..., 1, None, None]) + q(r) * ra[..., 0, None, None] + ra[..., 1, None, None])
    assert float((_from_c8(out, Cc) - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))
    # context split / GRU gates (C = 16 channels per half)
    x2, hh, qq = torch.randn(B, 32, h, w), torch.randn(B, 16, h, w), torch.randn(B, 16, h, w)
    o1 = torch.zeros(B, 2, h, w, 8, dtype=torch.float16, device="cuda")
    o2 = torch.zeros_like(o1)
    call(m.abi.PW_CTX_SPLIT, _to_c8(x2.cuda()), o1, 2, out2=o2)
    assert float((_from_c8(o1, 16) - torch.tanh(q(x2[:, :16]))).abs().max()) <= 1e-3
    assert float((_from_c8(o2, 16) - torch.relu(q(x2[:, 16:]))).abs().max()) <= 2e-3
    call(m.abi.PW_GRU_RH, _to_c8(x2.cuda()), o1, 2, r=_to_c8(hh.cuda()))
    assert float((_from_c8(o1, 16) - torch.sigmoid(q(x2[:, 16:])) * q(hh)).abs().max()) <= 2e-3
    hbuf = _to_c8(hh.cuda())
    call(m.abi.PW_GRU_UPDATE, _to_c8(x2.cuda()), hbuf, 2, r=_to_c8(qq.cuda()))
    z = torch.sigmoid(q(x2[:, :16]))
    assert float((_from_c8(hbuf, 16) - ((1 - z) * q(hh) + z * torch.tanh(q(qq)))).abs().max()) <= 2e-3


@pytest.mark.parametrize("case", [dict(cins=[128, 128, 126, 2], Cout=256, k=1, kw=5, pad_yx=(0, 2)), dict(cins=[128, 128, 126, 2], Cout=128, k=5, kw=1, pad_yx=(2, 0)),
                                  dict(cins=[16], Cout=16, k=1, kw=5, pad_yx=(0, 2)), dict(cins=[24], Cout=40, k=5, kw=1, pad_yx=(2, 0))])
def test_rectangular_conv_matches_torch(case):
    """The ConvGRU's 1x5 / 5x1 kernels with per-axis padding (MfcConvDesc.in_off) over a multi-source concat, against
    F.conv2d on the fp16-rounded operands (fp32 accumulation: 5e-3 of the output range)."""
    from mfcnet_tracker_b200 import engine
    from mfcnet_tracker_b200.engine import Act
    torch.manual_seed(6)
    dev = torch.device("cuda")
    B, H, W = 2, 30, 40
    bld = engine.Builder(dev, "fp16", engine.WeightPacker(dev, "fp16"), engine.Arena(dev))
    xs = [torch.randn(B, c, H, W, device=dev) for c in case["cins"]]
    cin = sum(case["cins"])
    wgt = torch.randn(case["Cout"], cin, case["k"], case["kw"], device=dev) / (cin * case["k"] * case["kw"]) ** 0.5
    bias = 0.1 * torch.randn(case["Cout"], device=dev)
    out_nchw = torch.full((B, case["Cout"], H, W), float("nan"), device=dev)
    bld.conv("c", [Act(_to_c8(x), x.shape[1]) for x in xs], wgt, case["k"], bias=bias, kw=case["kw"], pad_yx=case["pad_yx"], out_nchw=out_nchw)
    bld.prog.run()
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(torch.cat([x.half().float() for x in xs], 1), wgt.half().float(), bias, padding=case["pad_yx"])
    assert float((out_nchw - ref).abs().max()) <= 5e-3 * max(1.0, float(ref.abs().max()))
