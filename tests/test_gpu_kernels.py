"""GPU parity tests of the memory-bound kernels through the C ABI: correlation cost volume, heat-map
head, key-point extraction.  Integer / index results must be bit-exact; floating-point tolerances are
written next to each comparison."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import corr, localize_cases, localize_oracle as LO, synth, track_cases, track_oracle as TO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def M():
    import mfcnet_tracker_b200 as m
    assert torch.cuda.is_available()
    m.abi.load()
    return m


# ---------------------------------------------------------------- correlation
CORR_CASES = [  # (B, C, H, W, max_disp, stride2)
    (1, 7, 9, 11, 4, 1), (2, 40, 9, 11, 4, 2), (1, 64, 13, 70, 20, 2), (2, 32, 30, 40, 4, 1),
    (1, 256, 12, 40, 20, 2), (1, 16, 8, 8, 0, 1), (1, 33, 10, 67, 6, 2), (1, 8, 6, 9, 24, 2)]


@pytest.mark.parametrize("case", CORR_CASES)
def test_correlation_matches_oracle(M, case):
    B, Cc, H, W, md, s2 = case
    a = synth.normal("gcorr/a", (B, Cc, H, W), 3)
    b = synth.normal("gcorr/b", (B, Cc, H, W), 4)
    want_exact = corr.correlation_c(a, b, md, s2)
    want64 = corr.correlation_f64(a, b, md, s2)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    fast = M.correlation(ta, tb, md, s2).cpu().numpy()
    exact = M.correlation(ta, tb, md, s2, exact_order=True).cpu().numpy()
    # exact_order reproduces the reference kernel's fp32 summation order: bit-exact vs the C oracle
    assert np.array_equal(exact, want_exact)
    # the fast kernel sums channels in ascending order: fp32 rounding only (|out| <~ 1, C <= 256)
    assert np.abs(fast - want64).max() <= 2e-6


def test_correlation_reference_operating_point_properties(M):
    """C=256 at 48x160 (1/8 of 384x1280), max_disp 20 stride 2 (models/unflow_model.py:157-163):
    too slow for the CPU oracle at full size, so check size-independent properties."""
    B, Cc, H, W = 2, 256, 48, 160
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(B, Cc, H, W, device="cuda", generator=g)
    b = torch.randn(B, Cc, H, W, device="cuda", generator=g)
    out = M.correlation(a, b)
    assert out.shape == (B, 441, H, W)
    # centre displacement = per-pixel mean of a*b
    assert torch.allclose(out[:, 220], (a * b).mean(1), atol=2e-6)
    # linearity in the first argument
    out2 = M.correlation(2.0 * a, b)
    assert torch.equal(out2, 2.0 * out)
    # a displacement that looks outside the image is exactly zero there
    assert torch.count_nonzero(out[:, 0, :20, :]) == 0 and torch.count_nonzero(out[:, 0, :, :20]) == 0
    # swapping the arguments mirrors the displacement grid: out(a,b)[d](p) == out(b,a)[-d](p+d)
    o_ba = M.correlation(b, a)
    d_idx = (10 + 3) * 21 + (10 - 2)      # dy=+6, dx=-4
    m_idx = (10 - 3) * 21 + (10 + 2)
    lhs = out[:, d_idx, : H - 6, 4:]
    rhs = o_ba[:, m_idx, 6:, : W - 4]
    assert torch.allclose(lhs, rhs, atol=2e-6)
    # a random crop against the exact-order kernel
    ex = M.correlation(a, b, exact_order=True)
    assert (out - ex).abs().max().item() <= 2e-6


def test_correlation_matches_reference_kernel_golden(M):
    """Committed outputs of the reference's own kernels (tests/golden/corr_ref.npz, oracle/make_golden_corr.py):
    exact_order must match bit for bit, the fast kernels to fp32 rounding."""
    from oracle import make_golden_corr as MG
    gold = np.load(os.path.join(ROOT, "tests", "golden", "corr_ref.npz"))
    for tag, B, Cc, H, W in MG.FWD_CASES:
        a, b = MG.inputs(tag, B, Cc, H, W)
        ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        assert np.array_equal(M.correlation(ta, tb, exact_order=True).cpu().numpy(), gold[tag]), tag
        assert np.abs(M.correlation(ta, tb).cpu().numpy() - gold[tag]).max() <= 2e-6, tag


def test_correlation_matches_the_reference_kernel_itself(M):
    """The reference's own CUDA-C kernels (models/unflow_correlation.py:10-105), compiled with NVRTC and launched as
    `_FunctionCorrelation.forward` launches them (oracle/corr_ref_nvrtc.py), at the reference's operating point
    (C=256, 48x160, max displacement 20, stride 2: models/unflow_model.py:157-163) and on a ragged channel count."""
    from oracle import corr_ref_nvrtc as R
    if not R.available():
        pytest.skip("baseline/_ref/unflow_correlation_kernels.json (python -m oracle.make_corr_ref) or cuda-python missing")
    g = torch.Generator(device="cuda").manual_seed(7)
    for (B, Cc, H, W) in [(1, 256, 48, 160), (2, 40, 12, 20)]:
        a = torch.randn(B, Cc, H, W, device="cuda", generator=g)
        b = torch.randn(B, Cc, H, W, device="cuda", generator=g)
        ref = R.forward(a, b)
        exact = M.correlation(a, b, exact_order=True)
        fast = M.correlation(a, b)
        torch.cuda.synchronize()
        assert torch.equal(exact, ref), float((exact - ref).abs().max())      # same fp32 summation order: same bits
        assert float((fast - ref).abs().max()) <= 2e-6                        # fp32 rounding only (|out| <~ 1)
    # and the plain-C oracle agrees with the reference kernel bit for bit (pins oracle/corr_oracle.c)
    a = synth.normal("gcorr/a", (1, 40, 9, 11), 3)
    b = synth.normal("gcorr/b", (1, 40, 9, 11), 4)
    ref = R.forward(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.array_equal(corr.correlation_c(a, b, 20, 2), ref)


def test_correlation_backward_matches_reference_kernel_golden(M):
    """gradFirst / gradSecond of the reference's own backward kernels (models/unflow_correlation.py:107-235, run through
    NVRTC on a B200: tests/golden/corr_ref.npz) -- bit for bit, and through torch.autograd (FunctionCorrelation.backward)."""
    from oracle import make_golden_corr as MG
    gold = np.load(os.path.join(ROOT, "tests", "golden", "corr_ref.npz"))
    for tag, B, Cc, H, W in MG.BWD_CASES:
        a, b = MG.inputs(tag, B, Cc, H, W)
        g = synth.normal(tag + "/g", (B, 441, H, W), 5)
        ta = torch.from_numpy(a).cuda().requires_grad_(True)
        tb = torch.from_numpy(b).cuda().requires_grad_(True)
        out = M.correlation(ta, tb)
        out.backward(torch.from_numpy(g).cuda())
        assert np.array_equal(ta.grad.cpu().numpy(), gold[tag + "/grad_first"]), tag
        assert np.array_equal(tb.grad.cpu().numpy(), gold[tag + "/grad_second"]), tag
    # other displacement grids: against autograd of the plain-torch restatement (fp32 rounding only)
    from oracle import torch_oracle as TO
    for (B, Cc, H, W, md, s2) in [(2, 16, 12, 14, 4, 1), (1, 8, 9, 20, 6, 2)]:
        a = torch.randn(B, Cc, H, W, device="cuda")
        b = torch.randn(B, Cc, H, W, device="cuda")
        D = 2 * (md // s2) + 1
        g = torch.randn(B, D * D, H, W, device="cuda")
        ra, rb = a.clone().double().requires_grad_(True), b.clone().double().requires_grad_(True)
        TO.correlation(ra, rb, md, s2).backward(g.double())
        g1, g2 = M.correlation_backward(a, b, g, md, s2)
        assert float((g1 - ra.grad.float()).abs().max()) <= 2e-6 and float((g2 - rb.grad.float()).abs().max()) <= 2e-6


def test_correlation_rejects_cpu(M):
    with pytest.raises(NotImplementedError):
        M.correlation(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8))


# ---------------------------------------------------------------- heat-map head
def test_heatmap_head_matches_torch(M):
    x = torch.from_numpy(synth.normal("hm", (2, 5, 120, 160), 3, std=3.0)).cuda()
    x[0, :, 5, 7] = 1.25          # an exact tie: the first class must win
    logp, prob, amax = M.heatmap_head(x)
    ref_lp = torch.log_softmax(x, 1)
    assert (logp - ref_lp).abs().max().item() <= 2e-6      # fp32 log-softmax, different exp/log ulps
    assert (prob - ref_lp.exp()).abs().max().item() <= 2e-6
    assert np.array_equal(amax.cpu().numpy(), prob.cpu().numpy().argmax(1))     # first-max, bit-exact on our own probs
    assert amax[0, 5, 7].item() == 0
    agree = (amax.cpu().numpy() == ref_lp.exp().cpu().numpy().argmax(1)).mean()
    assert agree >= 0.9999


# ---------------------------------------------------------------- key-point extraction
def test_gaussian_blur_bit_exact_with_scipy(M):
    for shape in [(40, 56), (17, 23), (480, 640)]:
        img = synth.uniform("ggauss", shape, 9)
        got = M.gaussian_blur(torch.from_numpy(img).cuda(), 4).cpu().numpy()
        assert np.array_equal(got, LO.smoothed(img)), shape


def test_contours_match_cv2(M):
    from mfcnet_tracker_b200 import heatmap as HM
    rng = np.random.RandomState(0)
    masks = [255 * (rng.rand(37, 53) < p).astype(np.uint8) for p in (0.05, 0.3, 0.5, 0.6, 0.8)]
    masks += [255 * (rng.rand(480, 640) < p).astype(np.uint8) for p in (0.02, 0.45)]
    ring = np.zeros((20, 20), np.uint8)
    ring[2:18, 2:18] = 255
    ring[4:16, 4:16] = 0
    ring[8:12, 8:12] = 255
    masks += [ring, np.full((9, 9), 255, np.uint8), np.zeros((9, 9), np.uint8)]
    for mi, mask in enumerate(masks):
        want = [tuple(r) for r in LO.contour_records(mask)]
        got = HM.trace_contours(torch.from_numpy(mask).cuda())
        assert got == want, mi
        assert HM.calc_centroids(torch.from_numpy(mask).cuda()) == LO.calc_centroids(mask), mi


def test_keypoints_bit_exact_on_identical_heatmaps(M):
    """North-star criterion: tool-tip coordinates bit-exact given identical heat maps -- against the
    reference's own outputs (tests/golden/localize_centroids.json) and the scipy/cv2 oracle."""
    with open(os.path.join(ROOT, "tests", "golden", "localize_centroids.json")) as f:
        gold = json.load(f)
    for name, prob in localize_cases.cases().items():
        got = M.predicted_keypoints(torch.from_numpy(prob).cuda())
        norm = [[None if (isinstance(v, float) and np.isnan(v)) else int(v) for v in lst] for lst in got]
        assert norm == gold[name], name
    # noisy random maps: thousands of contours, ties and border effects
    for seed in range(3):
        p = synth.uniform("kp/%d" % seed, (1, 5, 120, 160), seed)
        p = (p / p.sum(1, keepdims=True)).astype(np.float32)
        got = M.predicted_keypoints(torch.from_numpy(p).cuda())
        want = LO.predicted_keypoints(p)
        assert str(got) == str(want), seed


# ---------------------------------------------------------------- video-script tracking
def _blob_masks():
    """0/255 masks with rings, blobs nested in holes (two levels), ties in area, small blobs and salt noise."""
    rng = np.random.default_rng(11)
    H, W = 96, 128
    y, x = np.mgrid[:H, :W]
    out = []
    for k in range(8):
        m = np.zeros((H, W), bool)
        for _ in range(int(rng.integers(1, 6))):
            cy, cx, r = int(rng.integers(0, H)), int(rng.integers(0, W)), int(rng.integers(2, 22))
            d2 = (y - cy) ** 2 + (x - cx) ** 2
            m |= d2 <= r * r
            if r > 8 and rng.random() < 0.8:
                m &= ~(d2 <= (r - 3) ** 2)                      # ring
                if rng.random() < 0.8:
                    m |= d2 <= (r - 6) ** 2                     # blob in the hole
                    if r > 14:
                        m &= ~(d2 <= (r - 9) ** 2)              # ... itself a ring
                        m |= d2 <= 4                            # ... with a blob inside
        if k % 2:
            m ^= rng.random((H, W)) < 0.02
        out.append(255 * m.astype(np.uint8))
    out.append(np.zeros((H, W), np.uint8))
    out.append(np.full((H, W), 255, np.uint8))
    eq = np.zeros((H, W), np.uint8)                             # three equal-area squares: order decided by the stable sort
    for cx in (10, 50, 90):
        eq[20:30, cx:cx + 10] = 255
    out.append(eq)
    out.append((255 * (rng.random((H, W)) < 0.5)).astype(np.uint8))
    return out


def test_refine_tip_segmentation_matches_cv2(M):
    """mask & (filled two largest contours over the area threshold), bit-exact with the cv2 sequence of
    scripts/test_multiframe_segmentation_on_videos_v3.py:32-42 -- including blobs nested in the holes of a kept contour."""
    nested_kept = 0
    for mi, mask in enumerate(_blob_masks()):
        for thr in (0, 10, 150):
            want = TO.refine_tip_segmentation(mask, thr)
            got = M.refine_tip_segmentation(torch.from_numpy(mask).cuda(), thr).cpu().numpy()
            assert np.array_equal(got, want), (mi, thr)
        import cv2
        n_all = cv2.connectedComponents((want > 0).astype(np.uint8), connectivity=8)[0] - 1
        nested_kept += n_all > 2
    assert nested_kept > 0      # the cases really contain kept nested components


def test_contour_labels_under_contention(M):
    """Full-resolution masks near the percolation threshold (components and holes with long union-find chains, heavy
    contention in the label kernels): contour records and the refined mask must still equal cv2's, run after run."""
    from mfcnet_tracker_b200 import heatmap as HM
    rng = np.random.default_rng(3)
    for dens in (0.4, 0.55, 0.62):
        mask = (255 * (rng.random((480, 640)) < dens)).astype(np.uint8)
        want_rec = [tuple(r) for r in LO.contour_records(mask)]
        want_ref = TO.refine_tip_segmentation(mask, 10)
        dm = torch.from_numpy(mask).cuda()
        for rep in range(4):
            assert HM.trace_contours(dm) == want_rec, (dens, rep)
            assert np.array_equal(M.refine_tip_segmentation(dm, 10).cpu().numpy(), want_ref), (dens, rep)


def test_contour_labels_odd_sizes(M):
    """Widths that are not a multiple of the 32-pixel warp segment (runs and the unions skipped as implied straddle rows)."""
    from mfcnet_tracker_b200 import heatmap as HM
    rng = np.random.default_rng(17)
    for (H, W) in ((37, 53), (50, 67), (1, 40), (33, 1), (64, 31), (9, 33)):
        for dens in (0.3, 0.55, 0.8):
            mask = (255 * (rng.random((H, W)) < dens)).astype(np.uint8)
            dm = torch.from_numpy(mask).cuda()
            assert HM.trace_contours(dm) == [tuple(r) for r in LO.contour_records(mask)], (H, W, dens)
            assert np.array_equal(M.refine_tip_segmentation(dm, 3).cpu().numpy(), TO.refine_tip_segmentation(mask, 3)), (H, W, dens)


def test_class_map_matches_script(M):
    for seed, thr in ((0, 0.0), (1, 0.3), (2, 0.5), (3, 0.21)):
        p = synth.uniform("cm/%d" % seed, (1, 5, 60, 80), seed)
        p = (p / p.sum(1, keepdims=True)).astype(np.float32)
        got = M.class_map(torch.from_numpy(p).cuda(), thr).cpu().numpy()[0]
        assert np.array_equal(got, TO.class_map(p, thr).astype(np.uint8)), seed


def test_tool_tracker_rows_bit_exact(M):
    """The 12 tracked coordinates per frame, over frame sequences: identical to the rows the reference's own functions wrote
    (tests/golden/track_rows.json) and to the cv2 / scipy oracle on noisy random maps."""
    with open(os.path.join(ROOT, "tests", "golden", "track_rows.json")) as f:
        gold = json.load(f)
    for pname, (a, d, s) in track_cases.PARAMS.items():
        for name, seq in track_cases.sequences().items():
            tr = M.ToolTracker(a, d, s)
            for t, p in enumerate(seq):
                want = np.array([np.nan if v is None else v for v in gold["%s/%s" % (pname, name)][t]])
                got = tr.step(torch.from_numpy(p).cuda())
                assert np.array_equal(got, want, equal_nan=True), (pname, name, t, got, want)
    for thr in (0.0, 0.3):
        tr, ref = M.ToolTracker(2, 60, thr), TO.Tracker(2, 60, thr)
        for seed in range(4):
            p = synth.uniform("trk/%d" % seed, (1, 5, 120, 160), seed)
            p[0, 1:] *= 0.9
            p = (p / p.sum(1, keepdims=True)).astype(np.float32)
            assert np.array_equal(tr.step(torch.from_numpy(p).cuda()), ref.step(p), equal_nan=True), (thr, seed)


# ---------------------------------------------------------------- training loss (forward)
def test_segmentation_loss_matches_reference(M):
    """mfc_segmentation_loss vs the reference's get_loss values (golden) and the fp64 oracle; the fp32 per-thread
    partial sums give ~1e-6 relative error, asserted at 2e-5."""
    from oracle import torch_oracle as TO
    with open(os.path.join(ROOT, "tests", "golden", "loss_cases.json")) as f:
        j = json.load(f)
    for tag, c in j["cases"].items():
        o, t = synth.loss_case(tag, c["B"], j["N"], c["H"], c["W"], seed=c["seed"], fg=c["fg"])
        total, d = M.segmentation_loss(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda(), j["class_weights"],
                                       ("nll", "soft_jaccard"), j["loss_wts"])
        want = TO.segmentation_loss(torch.from_numpy(o), torch.from_numpy(t), j["class_weights"], *j["loss_wts"])
        for got, ref_v, orc in ((d["loss_total"], c["total"], want[0]), (d["loss_nll"], c["nll"], want[1]),
                                (d["loss_soft_jaccard"], c["soft_jaccard"], want[2])):
            assert abs(got - ref_v) <= 2e-5 * abs(ref_v), (tag, got, ref_v)
            assert abs(got - orc) <= 2e-5 * abs(orc), (tag, got, orc)
        total2, _ = M.segmentation_loss(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda(), j["class_weights"])
        assert float(total2) == float(total)          # deterministic reduction
    with pytest.raises(ValueError):
        M.segmentation_loss(torch.zeros(1, 5, 8, 8, device="cuda"), torch.zeros(1, 8, 8, dtype=torch.int64, device="cuda"),
                            loss_fns=("mse",), loss_wts=(1.0,))


def test_resample_kernels_match_torch(M):
    """fuse_sum (bilinear align_corners=False upsampling on read) and bilinear_resize vs F.interpolate."""
    import torch.nn.functional as F
    from mfcnet_tracker_b200 import engine
    from mfcnet_tracker_b200.engine import Act
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    B, Cc, H, W = 2, 24, 24, 40
    prog = engine.Program(dev, "fp16")

    def c8(x):
        return x.view(B, Cc // 8, 8, x.shape[2], x.shape[3]).permute(0, 1, 3, 4, 2).contiguous().half()
    xs = [torch.randn(B, Cc, H >> k, W >> k, generator=g).to(dev) for k in range(4)]
    out = torch.empty(B, Cc // 8, H, W, 8, dtype=torch.float16, device=dev)
    sc = (1.0 + 0.1 * torch.randn(Cc, generator=g)).to(dev)
    sh = (0.1 * torch.randn(Cc, generator=g)).to(dev)
    prog.fuse_sum([Act(c8(x), Cc) for x in xs], out, Cc, scale=sc, shift=sh, act=1)
    low = torch.randn(B, 5, 12, 20, generator=g).to(dev)
    up = torch.empty(B, 5, 48, 80, device=dev)
    up8 = torch.zeros(B, 1, 48, 80, 8, dtype=torch.float16, device=dev)
    prog.resize(low, 48, 80, dst_nchw=up, dst_c8=up8)
    pooled = torch.empty(B, Cc // 8, H // 2, W // 2, 8, dtype=torch.float16, device=dev)
    prog.maxpool2(Act(c8(xs[0]), Cc), pooled)
    prog.run()
    torch.cuda.synchronize()
    ref = sum(F.interpolate(x.half().float(), size=(H, W), mode="bilinear", align_corners=False) if x.shape[2] != H else x.half().float()
              for x in xs)
    ref = F.relu(ref * sc[None, :, None, None] + sh[None, :, None, None])
    got = out.float().permute(0, 1, 4, 2, 3).reshape(B, Cc, H, W)
    assert (got - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())      # fp16 output rounding
    ref_up = F.interpolate(low, size=(48, 80), mode="bilinear", align_corners=False)
    assert (up - ref_up).abs().max().item() <= 1e-6
    assert (up8.float()[..., :5].permute(0, 1, 4, 2, 3).reshape(B, 5, 48, 80) - ref_up).abs().max().item() <= 2e-3 * ref_up.abs().max().item()
    ref_pool = F.max_pool2d(xs[0].half().float(), 2, 2)
    assert torch.equal(pooled.float().permute(0, 1, 4, 2, 3).reshape(B, Cc, H // 2, W // 2), ref_pool)   # bit-exact
