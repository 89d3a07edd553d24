"""The torch oracle restatement must reproduce the outputs the REAL reference
modules produced (tests/golden, written by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import synth, torch_oracle as TO
from tests import golden_util as G

TOL = 2e-5  # same fp32 ops; allows a different CPU conv kernel choice on another host


@pytest.mark.parametrize("tag", ["resunet16_64x96", "resunet8_32x48"])
def test_resunet_matches_reference_output(tag):
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"])
    x = torch.from_numpy(synth.frames(tag, meta["B"], meta["H"], meta["W"], meta["seed"]))
    with torch.no_grad():
        y = TO.resunet_forward(sd, x)
    assert y.shape == arr["logits"].shape
    assert np.abs(y.numpy() - arr["logits"]).max() < TOL


@pytest.mark.parametrize("K", [3, 5])
@pytest.mark.parametrize("variant", ["large", "basic"])
def test_fusion_matches_reference_output(variant, K):
    tag = f"fusion_{variant}_k{K}_48x64"
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"])
    x = G.fusion_input(tag, meta)
    with torch.no_grad():
        if variant == "large":
            y = TO.fusion_large_forward(sd, x)
        else:
            w = TO.warp_seg_and_depth(x, sd["grid"], meta["N"], K, True)
            assert np.abs(w.numpy() - arr["warped"]).max() < TOL
            y = TO.fusion_basic_forward(sd, x, meta["N"], K, True, True)
    assert np.abs(y.numpy() - arr["out"]).max() < TOL


@pytest.mark.parametrize("variant", ["large", "basic"])
def test_mfcnet_resunet_matches_reference_output(variant):
    tag = f"mfcnet_resunet16_{variant}_k3_64x96"
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"])
    xs, fl, dp = G.mfcnet_inputs(tag, meta)
    with torch.no_grad():
        y = TO.mfcnet_forward(sd, xs, fl, dp, base=TO.resunet_forward, variant=variant, N=meta["N"])
    assert np.abs(y.numpy() - arr["out"]).max() < TOL


def test_hrnet_matches_reference_output():
    """HighResolutionNet (HRNet-W48): the restatement vs the real models/hrnet.py output."""
    tag = "hrnet_w48_64x96"
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"])
    x = torch.from_numpy(synth.frames(tag, meta["B"], meta["H"], meta["W"], meta["seed"]))
    with torch.no_grad():
        y = TO.hrnet_forward(sd, x)
    assert y.shape == arr["logits"].shape
    assert np.abs(y.numpy() - arr["logits"]).max() < TOL * max(1.0, float(np.abs(arr["logits"]).max()))


def test_mfcnet_hrnet_matches_reference_output():
    tag = "mfcnet_hrnet_large_k3_64x96"
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"])
    xs, fl, dp = G.mfcnet_inputs(tag, meta)
    with torch.no_grad():
        y = TO.mfcnet_forward(sd, xs, fl, dp, base=TO.hrnet_forward, variant="large", N=meta["N"])
    assert np.abs(y.numpy() - arr["out"]).max() < TOL * max(1.0, float(np.abs(arr["out"]).max()))


def test_ternaus16_matches_reference_output():
    """TernausNet16 (VGG16 encoder, ConvTranspose2d(4,2,1) decoder, log_softmax head) vs models/ternausnet.py."""
    tag = "ternaus16_64x96"
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"])
    x = torch.from_numpy(synth.frames(tag, meta["B"], meta["H"], meta["W"], meta["seed"]))
    with torch.no_grad():
        y = TO.ternaus_forward(sd, x)
    assert np.abs(y.numpy() - arr["logp"]).max() < TOL * max(1.0, float(np.abs(arr["logp"]).max()))


def test_mfcnet_ternaus_matches_reference_output():
    tag = "mfcnet_ternaus16_basic_k3_64x96"
    meta, man, arr = G.load(tag)
    sd = G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"])
    xs, fl, dp = G.mfcnet_inputs(tag, meta)
    with torch.no_grad():
        y = TO.mfcnet_forward(sd, xs, fl, dp, base=TO.ternaus_probs, variant="basic", N=meta["N"])
    assert np.abs(y.numpy() - arr["out"]).max() < TOL * max(1.0, float(np.abs(arr["out"]).max()))


def test_loss_matches_reference_get_loss():
    """oracle segmentation_loss vs the reference's own get_loss outputs (tests/golden/loss_cases.json)."""
    import json, os
    with open(os.path.join(G.GOLDEN, "loss_cases.json")) as f:
        j = json.load(f)
    for tag, c in j["cases"].items():
        o, t = synth.loss_case(tag, c["B"], j["N"], c["H"], c["W"], seed=c["seed"], fg=c["fg"])
        total, nll, jac = TO.segmentation_loss(torch.from_numpy(o), torch.from_numpy(t), j["class_weights"], *j["loss_wts"])
        # the reference accumulates in fp32, the oracle in fp64: 1e-5 relative
        assert abs(total - c["total"]) <= 1e-5 * abs(c["total"]), (tag, total, c["total"])
        assert abs(nll - c["nll"]) <= 1e-5 * abs(c["nll"]) and abs(jac - c["soft_jaccard"]) <= 1e-5 * abs(c["soft_jaccard"])


def test_synth_is_stable():
    """Known-answer check of the platform-independent generator itself."""
    a = synth.normal("kat", (4,), seed=5)
    b = synth.uniform("kat", (3,), seed=5)
    assert a.dtype == np.float32 and b.dtype == np.float32
    assert np.array_equal(a, synth.normal("kat", (4,), seed=5))
    assert abs(float(synth.normal("stat", (200000,), 1).std()) - 1.0) < 0.01
    assert abs(float(synth.uniform("stat", (200000,), 1).mean()) - 0.5) < 0.01


def test_unflow_oracle_matches_reference_golden():
    """oracle/torch_oracle.py::unflow_forward against the flow written by the reference's own UnFlow module."""
    import torch
    from oracle import synth, torch_oracle as TO
    from tests import golden_util as G
    meta, man, arr = G.load("unflow_64x128")
    sd = G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"])
    a = torch.from_numpy(synth.uniform("unflow_64x128/first", (meta["B"], 3, meta["H"], meta["W"]), meta["seed"]))
    b = torch.from_numpy(synth.uniform("unflow_64x128/second", (meta["B"], 3, meta["H"], meta["W"]), meta["seed"]))
    with torch.no_grad():
        flow = TO.unflow_forward(sd, a, b)
    assert float((flow - torch.from_numpy(arr["flow"])).abs().max()) <= 1e-4
