"""Generate tests/golden/train_step.json by running the REAL reference training step (authoring container).

TEST INFRASTRUCTURE.  Usage:  python -m oracle.make_golden_train
The reference step (src/engine.py:56-71 with scripts/train_multiframe_detection.py:128-151): HRNetMultiLarge / -Basic with
its base_model swapped for ResUnet_VB (the reference has no ResUNetMulti class) in train() mode,
F.log_softmax -> get_loss(nll + soft_jaccard, 0.7/0.3, class weights 1/1000x4) -> backward -> Adam with the two
parameter groups (base lr/K, multiframe_net lr).  Two steps on seeded inputs; stored: the three loss values of each step,
and after the last step a few elements + the L2 norm of selected parameters and of the BatchNorm running statistics.
"""
import json
import os
import sys

import numpy as np
import torch

from . import refload, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PROBE = ["base_model.init_conv.weight", "base_model.downs.0.0.block1.proj.weight", "base_model.mid_block.block2.norm.weight",
         "base_model.output_layer.bias", "multiframe_net.multiframe_net.0.weight", "multiframe_net.multiframe_net.4.bias",
         "multiframe_net.multiframe_net.9.weight", "multiframe_net.multiframe_net.1.running_mean",
         "multiframe_net.multiframe_net.7.running_var"]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def main():
    ref = refload.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    N, K, B, H, W, lr, steps = 5, 3, 2, 32, 48, 1e-3, 2

    class LA:
        num_classes = N
        class_weights = np.array([1.0, 1000.0, 1000.0, 1000.0, 1000.0])
    res = {"N": N, "K": K, "B": B, "H": H, "W": W, "lr": lr, "steps": steps, "seed": 11, "cases": {}}
    for variant, cls in (("large", ref.multiframe.HRNetMultiLarge), ("basic", ref.multiframe.HRNetMultiBasic)):
        m = cls(num_classes=N, num_frames=K, pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True)
        m.base_model = ref.resunet.ResUnet_VB(channels=3, dim=16, out_dim=N)
        man = refload.manifest_of(m)
        sd = synth.fill_state_dict(man, 11)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        m.train()
        opt = torch.optim.Adam([{"params": m.base_model.parameters(), "lr": lr / K}, {"params": m.multiframe_net.parameters()}], lr=lr)
        losses = []
        for s in range(steps):
            tag = "train/%s/%d" % (variant, s)
            xs = [_t(synth.frames("%s/%d" % (tag, i), B, H, W, 11)) for i in range(K)]
            fl = [_t(synth.flow("%s/%d" % (tag, i), B, H, W, 11, scale=2.0)) for i in range(K - 1)]
            dp = [_t(synth.depth("%s/%d" % (tag, i), B, H, W, 11)) for i in range(K)]
            _, tg = synth.loss_case(tag, B, N, H, W, seed=11, fg=0.3)
            opt.zero_grad()
            out = torch.nn.functional.log_softmax(m(xs, optflow=fl, depth=dp), dim=1)
            total, d = ref.loss.get_loss(out, _t(tg), ["nll", "soft_jaccard"], [0.7, 0.3], LA)
            total.backward()
            opt.step()
            losses.append([float(total), d["loss_nll"], d["loss_soft_jaccard"]])
        after = m.state_dict()
        probes = {k: {"head": after[k].reshape(-1)[:6].double().tolist(), "norm": float(after[k].double().norm())} for k in PROBE}
        res["cases"][variant] = {"losses": losses, "probes": probes, "manifest": man}
        print(variant, losses)
    with open(os.path.join(OUT, "train_step.json"), "w") as f:
        json.dump(res, f, separators=(",", ":"))
    print("wrote train_step.json")


if __name__ == "__main__":
    sys.exit(main())
