"""CPU study: how much of HRNet's fp16-storage error comes from the residual stream?

Emulates the engine's storage rounding inside the fp32 torch oracle (weights and every stored activation rounded to
fp16, fp32 accumulation) and compares variants against the unrounded fp32 oracle:
  all   : every conv / block / fuse output stored in fp16 (what the engine does)
  res32 : block outputs (the residual stream relu(h + res)) kept in fp32 for the NEXT residual add; convs still read fp16
  python tools/precision_study.py [H W]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, torch_oracle as TO  # noqa: E402
from tests import golden_util as G  # noqa: E402


def q(x):
    return x.half().float()


def run(mode, sd, x):
    orig_conv_bn, orig_block, orig_module = TO._conv_bn, TO._hr_block, TO._hr_module

    def conv_bn(sd_, pc, pb, x_, stride=1, relu=False):
        x_ = getattr(x_, "lo16", x_)
        w = sd_[pc + "weight"]
        g = sd_[pb + "weight"] / torch.sqrt(sd_[pb + "running_var"] + 1e-5)
        wq = q(w * g[:, None, None, None])
        sh = sd_[pb + "bias"] - sd_[pb + "running_mean"] * g
        y = F.conv2d(q(x_), wq, None, stride=stride, padding=w.shape[-1] // 2) + sh[None, :, None, None]
        y = F.relu(y) if relu else y
        return y  # caller rounds (so that residual adds see the fp32 accumulator, as the epilogue does)

    def block(sd_, p, x_):
        res = x_
        if p + "downsample.0.weight" in sd_:
            res = q(conv_bn(sd_, p + "downsample.0.", p + "downsample.1.", x_))
        h = q(conv_bn(sd_, p + "conv1.", p + "bn1.", x_, relu=True))
        if p + "conv3.weight" in sd_:
            h = q(conv_bn(sd_, p + "conv2.", p + "bn2.", h, relu=True))
            h = conv_bn(sd_, p + "conv3.", p + "bn3.", h)
        else:
            h = conv_bn(sd_, p + "conv2.", p + "bn2.", h)
        out = F.relu(h + res)
        return out if mode == "res32" else q(out)

    TO._conv_bn = lambda *a, **k: q(conv_bn(*a, **k))
    TO._hr_block = block
    orig_cat, orig_conv2d = torch.cat, F.conv2d
    state = {}
    if mode.startswith("commute"):
        # the engine's head: last_layer.0 applied per branch at the branch's resolution, partial results stored
        # (fp16, or fp32 in "commute32"), then upsampled + summed in fp32 (fuse_sum)
        def cat(ts, dim=0):
            if dim == 1 and len(ts) == 4 and sum(t.shape[1] for t in ts) == 720:
                state["branches"] = True
            return orig_cat(ts, dim)
        torch.cat = cat
    try:
        if not mode.startswith("commute"):
            return TO.hrnet_forward(sd, x)
        return _forward_commuted(sd, x, mode)
    finally:
        TO._conv_bn, TO._hr_block, TO._hr_module = orig_conv_bn, orig_block, orig_module
        torch.cat = orig_cat


def _forward_commuted(sd, x, mode):
    """hrnet_forward with the head evaluated as the engine does (copy of the tail of TO.hrnet_forward)."""
    x = TO._conv_bn(sd, "conv1.", "bn1.", x, stride=2, relu=True)
    x = TO._conv_bn(sd, "conv2.", "bn2.", x, stride=2, relu=True)
    for k in range(TO._count(sd, "layer1.")):
        x = TO._hr_block(sd, "layer1.%d." % k, x)
    xs = [TO._conv_bn(sd, "transition1.0.0.", "transition1.0.1.", x, relu=True), TO._hr_chain(sd, "transition1.1.", x, relu_last=True)]
    for m in range(TO._count(sd, "stage2.")):
        xs = TO._hr_module(sd, "stage2.%d." % m, xs)
    xs = xs + [TO._hr_chain(sd, "transition2.2.", xs[-1], relu_last=True)]
    for m in range(TO._count(sd, "stage3.")):
        xs = TO._hr_module(sd, "stage3.%d." % m, xs)
    xs = xs + [TO._hr_chain(sd, "transition3.3.", xs[-1], relu_last=True)]
    for m in range(TO._count(sd, "stage4.")):
        xs = TO._hr_module(sd, "stage4.%d." % m, xs)
    size = xs[0].shape[-2:]
    w0 = sd["last_layer.0.weight"]
    off, acc = 0, None
    for t in xs:
        part = F.conv2d(q(t), q(w0[:, off:off + t.shape[1]]))
        if mode == "commute":
            part = q(part)
        elif mode == "commute_scaled":   # BN scale folded into the branch weights: partial sums are in post-BN units
            g = sd["last_layer.1.weight"] / torch.sqrt(sd["last_layer.1.running_var"] + 1e-5)
            part = q(F.conv2d(q(t), q(w0[:, off:off + t.shape[1]] * g[:, None, None, None])))
        off += t.shape[1]
        part = part if part.shape[-2:] == size else F.interpolate(part, size=size, mode="bilinear", align_corners=False)
        acc = part if acc is None else acc + part
    g = sd["last_layer.1.weight"] / torch.sqrt(sd["last_layer.1.running_var"] + 1e-5)
    if mode == "commute_scaled":
        y = acc + (sd["last_layer.1.bias"] + (sd["last_layer.0.bias"] - sd["last_layer.1.running_mean"]) * g)[None, :, None, None]
    else:
        y = (acc + sd["last_layer.0.bias"][None, :, None, None] - sd["last_layer.1.running_mean"][None, :, None, None]) * g[None, :, None, None] \
            + sd["last_layer.1.bias"][None, :, None, None]
    if mode == "commute_zw32":     # last activation and last weights kept in fp32 (hi/lo split through the tensor core)
        y = F.conv2d(F.relu(y), sd["last_layer.3.weight"], sd["last_layer.3.bias"])
    else:
        y = q(F.relu(y))
        y = F.conv2d(y, q(sd["last_layer.3.weight"]), sd["last_layer.3.bias"])
    return F.interpolate(y, size=(size[0] * 4, size[1] * 4), mode="bilinear", align_corners=False)


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 96)
    meta, man, _ = G.load("hrnet_w48_64x96")
    sd = G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"])
    x = torch.from_numpy(synth.frames("hrnet_w48_64x96", 1, H, W, meta["seed"]))
    with torch.no_grad():
        ref = TO.hrnet_forward(sd, x)
        for mode in ("all", "commute", "commute_zw32"):
            y = run(mode, sd, x)
            err = float((y - ref).abs().max())
            agree = float((y.argmax(1) == ref.argmax(1)).float().mean())
            print("%-6s %dx%d: max-abs err %.3e  argmax agreement %.5f  (ref absmax %.2f)" % (mode, H, W, err, agree, float(ref.abs().max())))


if __name__ == "__main__":
    main()
