"""Golden vector of the UnFlow network, written by the REFERENCE module itself (models/unflow_model.py).

TEST INFRASTRUCTURE ONLY.  Runs in the authoring container: the reference `UnFlow` is imported from /root/reference with
  * `models.unflow_correlation` replaced by a stub whose ModuleCorrelation calls the correlation oracle (the reference file
    needs CuPy + a GPU at import time; its kernels are pinned separately: tests/golden/corr_ref.npz), and
  * `Tensor.cuda()` made a no-op (models/unflow_model.py:11 moves its sampling grid to the GPU unconditionally),
then run on CPU in fp32 on inputs / weights regenerated from oracle/synth.py.  Output: tests/golden/unflow_64x128.{json,npz}.

    python -m oracle.make_golden_unflow
"""
import json
import os
import sys
import types

import numpy as np
import torch

from oracle import synth, torch_oracle as TO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG, B, H, W, SEED = "unflow_64x128", 1, 64, 128, 3
# conv weights x sqrt(2): a LeakyReLU stack halves the variance per layer otherwise (same rule as the TernausNet fixtures);
# moduleUpscale (bias-free 2->2 transposed conv applied twice, then x20) is scaled down so that the flow stays O(1)
RULES = {"__all_4d__": 2 ** 0.5}
for _n in range(3):   # flow2 heads scaled down: the x20 of Upconv.forward (:78) then gives flows of a few pixels, not hundreds
    RULES["moduleFlownets.%d.moduleUpconv.moduleTwoOut.weight" % _n] = 0.1
    RULES["moduleFlownets.%d.moduleUpconv.moduleTwoOut.bias" % _n] = 0.1


def load_reference():
    stub = types.ModuleType("models.unflow_correlation")

    class ModuleCorrelation(torch.nn.Module):
        def forward(self, a, b):
            return TO.correlation(a, b)
    stub.ModuleCorrelation = ModuleCorrelation
    pkg = types.ModuleType("models")
    pkg.__path__ = ["/root/reference/models"]
    sys.modules["models"] = pkg
    sys.modules["models.unflow_correlation"] = stub
    import importlib
    return importlib.import_module("models.unflow_model")


def main():
    um = load_reference()
    torch.Tensor.cuda = lambda self, *a, **k: self           # CPU run of a module that calls .cuda() on its grid
    net = um.UnFlow().eval()
    man = [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in net.state_dict().items()]
    sd = synth.apply_fixture_rules(synth.fill_state_dict([(k, tuple(s), d) for k, s, d in man], SEED), RULES)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    a = torch.from_numpy(synth.uniform(TAG + "/first", (B, 3, H, W), SEED))
    b = torch.from_numpy(synth.uniform(TAG + "/second", (B, 3, H, W), SEED))
    with torch.no_grad():
        flow = net(a.clone(), b.clone())
        mine = TO.unflow_forward({k: torch.from_numpy(v) for k, v in sd.items()}, a, b)
    print("reference flow", tuple(flow.shape), "absmax %.4f" % float(flow.abs().max()), "| restatement max diff %.3e" % float((flow - mine).abs().max()))
    assert torch.allclose(flow, mine, atol=1e-5, rtol=1e-5)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", TAG + ".npz"), flow=flow.numpy())
    with open(os.path.join(ROOT, "tests", "golden", TAG + ".json"), "w") as f:
        json.dump({"meta": {"kind": "unflow", "B": B, "H": H, "W": W, "seed": SEED, "scale_keys": RULES}, "manifest": man}, f)


if __name__ == "__main__":
    main()
