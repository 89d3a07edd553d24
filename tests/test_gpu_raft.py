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
