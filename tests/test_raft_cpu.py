"""CPU side of the RAFT port: checkpoint compatibility with torchvision's module and the oracle pinned by its golden fixture."""
import os

import numpy as np
import torch

import mfcnet_tracker_b200 as m
from oracle import raft_oracle as RO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "raft_128x160.npz")


def test_state_dict_is_torchvisions():
    from torchvision.models.optical_flow import raft_large
    ref = {k: (tuple(v.shape), v.dtype) for k, v in raft_large(weights=None).state_dict().items()}
    mine = m.raft_large()
    assert {k: (tuple(v.shape), v.dtype) for k, v in mine.state_dict().items()} == ref
    mine.load_state_dict(raft_large(weights=None).state_dict())      # strict


def test_oracle_matches_golden():
    """The seeded torchvision module reproduces the committed fixture (guards the weights / inputs the GPU tests regenerate)."""
    g = np.load(GOLD)
    net = RO.build(0)
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"])
    assert abs(float(sum(p.double().sum() for p in net.parameters())) - float(g["wsum"])) <= 1e-6 * max(1.0, abs(float(g["wsum"])))
    a, b = RO.frames(2, 128, 160)
    f1 = RO.flow(net, a, b, num_flow_updates=1).numpy()
    assert float(np.abs(f1 - g["flow1"].astype(np.float32)).max()) <= 5e-3
    f12 = RO.flow(net, a, b).numpy()
    assert float(np.abs(f12 - g["flow12"].astype(np.float32)).max()) <= 2e-2


def test_cpu_call_fails_loudly(monkeypatch):
    monkeypatch.delenv("MFC_B200_PLAN_ONLY", raising=False)    # (the host-logic tests' plan-only mode records plans on the CPU)
    net = m.raft_large().eval()
    x = torch.zeros(1, 3, 128, 160)
    try:
        net(x, x)
    except (RuntimeError, ValueError, NotImplementedError):
        return
    raise AssertionError("RAFT.forward on CPU tensors must raise: the product path has no CPU fallback")
