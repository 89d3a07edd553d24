"""Generate tests/golden/* by running the REAL reference modules (authoring container).

TEST INFRASTRUCTURE.  Usage:  python -m oracle.make_golden
Each fixture = a JSON manifest (state-dict keys/shapes/dtypes of the reference
module = the checkpoint-compatibility contract) + an .npz with the reference's
fp32 outputs.  Weights and inputs are NOT stored: they are regenerated from
oracle/synth.py (platform-independent), so fixtures stay small.
"""
import json
import os
import sys

import numpy as np
import torch

from . import refload, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load_sd(module, seed, scale_keys=None):
    """scale_keys {key: factor}: post-scaling of some tensors, recorded in the fixture's meta (HRNet: the
    random-init 300-layer stack grows activations to ~6e3, the head is scaled down so that logits are O(5)
    like a trained net's and the absolute 2e-2 logit tolerance of BASELINE.json means what it says)."""
    man = refload.manifest_of(module)
    sd = synth.apply_fixture_rules(synth.fill_state_dict(man, seed), scale_keys)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    module.eval()
    return man


def _save(name, manifest, meta, **arrays):
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump({"meta": meta, "manifest": manifest}, f, indent=0, separators=(",", ":"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items()})


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = refload.load()
    torch.set_grad_enabled(False)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- SFC: ResUnet_VB(dim=16) (BASELINE config 1 shape, reduced resolution)
    for tag, dim, H, W, B in (("resunet16_64x96", 16, 64, 96, 2), ("resunet8_32x48", 8, 32, 48, 1)):
        m = ref.resunet.ResUnet_VB(channels=3, dim=dim, out_dim=5)
        man = _load_sd(m, seed=1)
        x = synth.frames(tag, B, H, W, seed=1)
        _save(tag, man, {"kind": "resunet", "dim": dim, "B": B, "H": H, "W": W, "seed": 1, "classes": 5},
              logits=m(_t(x)).numpy())

    # ---- fusion heads alone
    N = 5
    for K in (3, 5):
        for variant, cls in (("large", ref.multiframe.MultiFrameNetLarge), ("basic", ref.multiframe.MultiFrameNetBasic)):
            tag = f"fusion_{variant}_k{K}_48x64"
            m = cls(N, K, False, with_optflow=True, with_depth=True)
            man = _load_sd(m, seed=2)
            B, H, W = 2, 48, 64
            x = np.concatenate([synth.normal(tag + "/seg", (B, N * K, H, W), 2, std=2.0),
                                synth.flow(tag, B, H, W, 2).repeat(K - 1, axis=1) * np.linspace(1, 2, 2 * (K - 1), dtype=np.float32)[None, :, None, None],
                                synth.uniform(tag + "/dep", (B, K, H, W), 2)], axis=1).astype(np.float32)
            arrays = {"out": m(_t(x)).numpy()}
            if variant == "basic":
                arrays["warped"] = m.warp_segmentation_and_depth(_t(x)).numpy()
            _save(tag, man, {"kind": "fusion", "variant": variant, "K": K, "N": N, "B": B, "H": H, "W": W, "seed": 2}, **arrays)

    # ---- full MFCNet with ResUNet-16 base: the reference has no ResUNetMulti class
    # (SURVEY.md D2); oracle = reference HRNetMulti{Large,Basic} with base_model swapped.
    for variant, cls in (("large", ref.multiframe.HRNetMultiLarge), ("basic", ref.multiframe.HRNetMultiBasic)):
        K, B, H, W = 3, 1, 64, 96
        tag = f"mfcnet_resunet16_{variant}_k{K}_64x96"
        m = cls(num_classes=N, num_frames=K, pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True)
        m.base_model = ref.resunet.ResUnet_VB(channels=3, dim=16, out_dim=N)
        man = _load_sd(m, seed=3)
        xs = [synth.frames(f"{tag}/{i}", B, H, W, 3) for i in range(K)]
        fl = [synth.flow(f"{tag}/{i}", B, H, W, 3) for i in range(K - 1)]
        dp = [synth.depth(f"{tag}/{i}", B, H, W, 3) for i in range(K)]
        out = m([_t(a) for a in xs], optflow=[_t(a) for a in fl], depth=[_t(a) for a in dp])
        _save(tag, man, {"kind": "mfcnet", "base": "resunet16", "variant": variant, "K": K, "N": N, "B": B, "H": H, "W": W, "seed": 3},
              out=out.numpy())

    # ---- SFC: HighResolutionNet (HRNet-W48, BASELINE config 4 base) at reduced resolution
    tag = "hrnet_w48_64x96"
    m = ref.hrnet.HighResolutionNet(num_classes=N)
    sk = {"last_layer.3.weight": 1e-3, "last_layer.3.bias": 1e-3}
    man = _load_sd(m, seed=4, scale_keys=sk)
    x = synth.frames(tag, 1, 64, 96, seed=4)
    _save(tag, man, {"kind": "hrnet", "B": 1, "H": 64, "W": 96, "seed": 4, "classes": N, "scale_keys": sk},
          logits=m(_t(x)).numpy())

    # ---- full MFCNet with the HRNet base (reference HRNetMultiLarge, K=3)
    K, B, H, W = 3, 1, 64, 96
    tag = f"mfcnet_hrnet_large_k{K}_64x96"
    m = ref.multiframe.HRNetMultiLarge(num_classes=N, num_frames=K, pretrained=False, loadpath=None, optflow_inputs=True,
                                       depth_inputs=True)
    sk = {"base_model.last_layer.3.weight": 1e-3, "base_model.last_layer.3.bias": 1e-3}
    man = _load_sd(m, seed=5, scale_keys=sk)
    xs = [synth.frames(f"{tag}/{i}", B, H, W, 5) for i in range(K)]
    fl = [synth.flow(f"{tag}/{i}", B, H, W, 5) for i in range(K - 1)]
    dp = [synth.depth(f"{tag}/{i}", B, H, W, 5) for i in range(K)]
    out = m([_t(a) for a in xs], optflow=[_t(a) for a in fl], depth=[_t(a) for a in dp])
    _save(tag, man, {"kind": "mfcnet", "base": "hrnet", "variant": "large", "K": K, "N": N, "B": B, "H": H, "W": W, "seed": 5,
                      "scale_keys": sk},
          out=out.numpy())

    # ---- SFC: TernausNet16 (VGG16 encoder, k4 s2 transposed-conv decoder, log_softmax head), num_filters=64 as
    # at the reference call sites (models/__init__.py:27, models/multiframe_model.py:216).  Fixture rules: all conv
    # weights x sqrt(2) (He gain: the plain ReLU stack would otherwise shrink activations to ~1e-3 and every
    # pixel would be a near-tie) and the aliased convK.* entries copy encoder.* (the same Parameter upstream).
    tag = "ternaus16_64x96"
    m = ref.ternaus.TernausNet16(num_classes=N, num_filters=64, pretrained=False)
    sk = {"__all_4d__": 1.4142135, "__alias__": "ternaus"}
    man = _load_sd(m, seed=6, scale_keys=sk)
    x = synth.frames(tag, 1, 64, 96, seed=6)
    _save(tag, man, {"kind": "ternaus16", "B": 1, "H": 64, "W": 96, "seed": 6, "classes": N, "scale_keys": sk},
          logp=m(_t(x)).numpy())

    tag = "mfcnet_ternaus16_basic_k3_64x96"
    m = ref.multiframe.TernausNetMultiBasic(num_classes=N, num_frames=3, pretrained=False, loadpath=None, optflow_inputs=True,
                                            depth_inputs=True)
    sk = {"__all_4d__": 1.4142135, "__alias__": "ternaus"}
    man = _load_sd(m, seed=7, scale_keys=sk)
    K, B, H, W = 3, 1, 64, 96
    xs = [synth.frames(f"{tag}/{i}", B, H, W, 7) for i in range(K)]
    fl = [synth.flow(f"{tag}/{i}", B, H, W, 7) for i in range(K - 1)]
    dp = [synth.depth(f"{tag}/{i}", B, H, W, 7) for i in range(K)]
    out = m([_t(a) for a in xs], optflow=[_t(a) for a in fl], depth=[_t(a) for a in dp])
    _save(tag, man, {"kind": "mfcnet", "base": "ternaus16", "variant": "basic", "K": K, "N": N, "B": B, "H": H, "W": W, "seed": 7,
                     "scale_keys": sk}, out=out.numpy())

    # ---- training loss forward: the reference's own get_loss (src/loss.py) on seeded cases
    class LA:
        num_classes = N
        class_weights = np.array([1.0, 1000.0, 1000.0, 1000.0, 1000.0])
    lres = {}
    for tag, (B, H, W, fg) in {"small": (2, 48, 64, 0.05), "sparse": (1, 96, 128, 0.01), "dense": (2, 32, 32, 0.6)}.items():
        o, t = synth.loss_case(tag, B, N, H, W, seed=8, fg=fg)
        logp = torch.nn.functional.log_softmax(_t(o), dim=1)               # src/engine.py:65
        total, d = ref.loss.get_loss(logp, _t(t), ["nll", "soft_jaccard"], [0.7, 0.3], LA)
        lres[tag] = {"B": B, "H": H, "W": W, "fg": fg, "seed": 8, "total": float(total), "nll": d["loss_nll"],
                     "soft_jaccard": d["loss_soft_jaccard"]}
    with open(os.path.join(OUT, "loss_cases.json"), "w") as f:
        json.dump({"class_weights": [1.0, 1000.0, 1000.0, 1000.0, 1000.0], "loss_wts": [0.7, 0.3], "N": N, "cases": lres}, f, indent=1)
    print("wrote loss_cases", lres)

    # ---- the 576x720 grid buffer itself
    g = ref.multiframe.MultiFrameNetBasic(N, 3, False, True, True).grid.numpy()
    assert np.array_equal(g, synth.mesh_grid_576x720()), "synth grid != reference grid"

    # ---- localisation: reference centroid_error on synthetic blob heatmaps
    from . import localize_cases
    loc = ref.localization
    cases = localize_cases.cases()
    res = {}
    for name, prob in cases.items():
        class A:  # the reference reads args.num_classes only
            num_classes = 5
        gt = torch.zeros(1, prob.shape[2], prob.shape[3], dtype=torch.int64)
        *_, c_pred = loc.centroid_error(_t(prob), gt, A)
        res[name] = [[(None if (isinstance(v, float) and np.isnan(v)) else int(v)) for v in lst] for lst in c_pred]
    with open(os.path.join(OUT, "localize_centroids.json"), "w") as f:
        json.dump(res, f)
    print("wrote localize_centroids", {k: v for k, v in list(res.items())[:3]})


if __name__ == "__main__":
    sys.exit(main())
