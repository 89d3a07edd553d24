"""Helpers shared by the parity tests: load a golden fixture and rebuild its
weights/inputs from oracle/synth.py (see oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(tag):
    with open(os.path.join(GOLDEN, tag + ".json")) as f:
        j = json.load(f)
    arrays = dict(np.load(os.path.join(GOLDEN, tag + ".npz")))
    man = [(k, tuple(s), d) for k, s, d in j["manifest"]]
    return j["meta"], man, arrays


def state_dict(man, seed, device="cpu", scale_keys=None):
    """scale_keys: the fixture's meta["scale_keys"] (see oracle/make_golden.py::_load_sd)."""
    sd = synth.apply_fixture_rules(synth.fill_state_dict(man, seed), scale_keys)
    return {k: torch.from_numpy(v).to(device) for k, v in sd.items()}


def fusion_input(tag, meta):
    """Mirror of the input construction in oracle/make_golden.py (fusion cases)."""
    B, H, W, K, N = meta["B"], meta["H"], meta["W"], meta["K"], meta["N"]
    x = np.concatenate([synth.normal(tag + "/seg", (B, N * K, H, W), 2, std=2.0),
                        synth.flow(tag, B, H, W, 2).repeat(K - 1, axis=1)
                        * np.linspace(1, 2, 2 * (K - 1), dtype=np.float32)[None, :, None, None],
                        synth.uniform(tag + "/dep", (B, K, H, W), 2)], axis=1).astype(np.float32)
    return torch.from_numpy(x)


def mfcnet_inputs(tag, meta):
    B, H, W, K = meta["B"], meta["H"], meta["W"], meta["K"]
    s = meta["seed"]
    xs = [torch.from_numpy(synth.frames(f"{tag}/{i}", B, H, W, s)) for i in range(K)]
    fl = [torch.from_numpy(synth.flow(f"{tag}/{i}", B, H, W, s)) for i in range(K - 1)]
    dp = [torch.from_numpy(synth.depth(f"{tag}/{i}", B, H, W, s)) for i in range(K)]
    return xs, fl, dp
