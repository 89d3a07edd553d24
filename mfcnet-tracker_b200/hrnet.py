"""HRNet-W48 segmentation network as a drop-in nn.Module on the B200 engine.

Mirrors the constructor, forward signature and state_dict keys of the reference
``HighResolutionNet(num_classes)`` (models/hrnet.py:271-476): ``conv1/bn1/conv2/bn2`` stem,
``layer1`` (4 Bottlenecks), ``transition{1,2,3}``, ``stage{2,3,4}`` (1/4/3 HighResolutionModules
with ``branches`` and ``fuse_layers``), ``last_layer.{0,1,3}``; the reference aliases
``BatchNorm2d = torch.nn.SyncBatchNorm`` (models/hrnet.py:31), whose state-dict entries are those
of ``nn.BatchNorm2d``.  The modules below only HOLD parameters; the arithmetic runs in
libmfcnet_b200.so:

  * every conv + eval-mode BN (+ ReLU, + residual add) is ONE fused tcgen05 conv launch: BN is folded
    into the epilogue's per-channel scale/shift, the BasicBlock / Bottleneck residual
    (models/hrnet.py:58-74, 97-115) is added in the epilogue before the ReLU;
  * the fuse step of HighResolutionModule.forward (:237-260) is one ``fuse_sum`` launch per output
    branch: lower-resolution terms are bilinearly upsampled (align_corners=False) while being read;
  * the head (:464-474) upsamples three branches, concatenates 720 channels and applies a 1x1 conv:
    a 1x1 conv commutes with bilinear interpolation, so the 720x720 conv is applied per branch at the
    branch's own resolution (8x fewer FLOPs, no 720-channel concat), the four results are upsampled and
    summed by ``fuse_sum`` together with BN + ReLU, then the 720->N 1x1 conv and the final x4
    bilinear upsampling of the logits.
"""
import os

import torch
from torch import nn

from . import abi, engine
from .engine import Act, Ext

BN_MOMENTUM = 0.1


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = _bn(planes)
        self.downsample = downsample
        self.stride = stride


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4)
        self.downsample = downsample
        self.stride = stride


def _conv_bn(cin, cout, k, stride, relu):
    layers = [nn.Conv2d(cin, cout, k, stride, k // 2, bias=False), _bn(cout)]
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class _HRModule(nn.Module):
    """Parameter holder with the child names of HighResolutionModule (models/hrnet.py:118-262)."""

    def __init__(self, num_branches, block, num_blocks, channels):
        super().__init__()
        self.num_branches = num_branches
        self.channels = list(channels)
        self.branches = nn.ModuleList([
            nn.Sequential(*[block(channels[i], channels[i]) for _ in range(num_blocks[i])]) for i in range(num_branches)])
        fuse = []
        for i in range(num_branches):
            row = []
            for j in range(num_branches):
                if j > i:
                    row.append(_conv_bn(channels[j], channels[i], 1, 1, relu=False))
                elif j == i:
                    row.append(None)
                else:
                    chain = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        chain.append(_conv_bn(channels[j], channels[i] if last else channels[j], 3, 2, relu=not last))
                    row.append(nn.Sequential(*chain))
            fuse.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(fuse)


class HighResolutionNet(nn.Module):
    """Drop-in for models/hrnet.py:271 ``HighResolutionNet``; returns raw logits (B, num_classes, H, W) fp32."""

    def __init__(self, num_classes=19):
        super().__init__()
        self.channels = 3
        self.out_dim = num_classes
        self.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False)
        self.bn1 = _bn(64)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = _bn(64)
        ds = nn.Sequential(nn.Conv2d(64, 256, 1, 1, bias=False), _bn(256))
        self.layer1 = nn.Sequential(_Bottleneck(64, 64, 1, ds), *[_Bottleneck(256, 64) for _ in range(3)])
        self.transition1 = self._transition([256], [48, 96])
        self.stage2 = nn.Sequential(_HRModule(2, _BasicBlock, [4, 4], [48, 96]))
        self.transition2 = self._transition([48, 96], [48, 96, 192])
        self.stage3 = nn.Sequential(*[_HRModule(3, _BasicBlock, [4, 4, 4], [48, 96, 192]) for _ in range(4)])
        self.transition3 = self._transition([48, 96, 192], [48, 96, 192, 384])
        self.stage4 = nn.Sequential(*[_HRModule(4, _BasicBlock, [4, 4, 4, 4], [48, 96, 192, 384]) for _ in range(3)])
        c = 48 + 96 + 192 + 384
        self.last_layer = nn.Sequential(nn.Conv2d(c, c, 1), _bn(c), nn.ReLU(inplace=True), nn.Conv2d(c, num_classes, 1))
        self._plans = {}
        self._packer = None
        self._fingerprint = None
        self.dtype_name = None

    @staticmethod
    def _transition(pre, cur):
        """models/hrnet.py:353-389."""
        layers = []
        for i, c in enumerate(cur):
            if i < len(pre):
                layers.append(_conv_bn(pre[i], c, 3, 1, relu=True) if c != pre[i] else None)
            else:
                chain = []
                for j in range(i + 1 - len(pre)):
                    chain.append(_conv_bn(pre[-1], c if j == i - len(pre) else pre[-1], 3, 2, relu=True))
                layers.append(nn.Sequential(*chain))
        return nn.ModuleList(layers)

    # ---- plan recording ----------------------------------------------------------------------------
    def _cb(self, bld, name, conv, bn, srcs, act, residual=None):
        """conv (+bias) -> eval BN -> (+residual) -> ReLU?, one launch."""
        scale, shift = bld.packer.bn_affine(name, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, conv.bias)
        out, _, _, _ = bld.conv(name, srcs, conv.weight, conv.kernel_size[0], scale=scale, shift=shift, stride=conv.stride[0],
                                pad=conv.padding[0], act=act, residual=residual)
        return out

    def _seq(self, bld, name, seq, x):
        """nn.Sequential of (conv, bn[, relu]) or of such Sequentials."""
        if isinstance(seq[0], nn.Conv2d):
            return self._cb(bld, name, seq[0], seq[1], [x], act=1 if len(seq) > 2 else 0)
        for k, sub in enumerate(seq):
            x = self._seq(bld, "%s.%d" % (name, k), sub, x)
        return x

    def _block(self, bld, name, blk, x):
        res = x if blk.downsample is None else self._cb(bld, name + ".downsample", blk.downsample[0], blk.downsample[1], [x], 0)
        h = self._cb(bld, name + ".conv1", blk.conv1, blk.bn1, [x], 1)
        if isinstance(blk, _Bottleneck):
            h = self._cb(bld, name + ".conv2", blk.conv2, blk.bn2, [h], 1)
            return self._cb(bld, name + ".conv3", blk.conv3, blk.bn3, [h], 1, residual=res)
        return self._cb(bld, name + ".conv2", blk.conv2, blk.bn2, [h], 1, residual=res)

    def _module(self, bld, name, mod, xs):
        # The branches of a module are independent until the fuse step and each of their convs fills only part of the GPU
        # (72..120 work items at batch 1): they are recorded on separate lanes = concurrent streams / parallel graph branches.
        lanes = mod.num_branches > 1 and mod.num_branches <= abi.MFC_MAX_LANES and os.environ.get("MFC_LANES", "1") != "0"
        ys = []
        if lanes:
            bld.prog.fork()
        for i in range(mod.num_branches):
            bld.prog.lane = i if lanes else 0
            x = xs[i]
            for k, blk in enumerate(mod.branches[i]):
                x = self._block(bld, "%s.branches.%d.%d" % (name, i, k), blk, x)
            ys.append(x)
        if lanes:
            bld.prog.join()
            bld.prog.fork()
        outs = []
        for i in range(mod.num_branches):
            bld.prog.lane = i if lanes else 0      # the fuse step of output i only reads the branch results
            terms = []
            for j in range(mod.num_branches):
                if j == i:
                    terms.append(ys[j])
                else:
                    terms.append(self._seq(bld, "%s.fuse_layers.%d.%d" % (name, i, j), mod.fuse_layers[i][j], ys[j]))
            o = bld.arena.alloc(tuple(ys[i].t.shape), bld.tdtype)
            outs.append(bld.prog.fuse_sum(terms, o, ys[i].C, act=1))
        if lanes:
            bld.prog.join()
        return outs

    def record(self, bld, x_act, logits_nchw=None, maps_c8=None):
        """Record the whole forward on `x_act` (C8 input).  Writes fp32 NCHW logits into
        `logits_nchw` and/or C8 class maps into `maps_c8`.  Returns (None, resize-args struct)."""
        B, H, W = x_act.B, x_act.H, x_act.W
        x = self._cb(bld, "conv1", self.conv1, self.bn1, [x_act], 1)
        x = self._cb(bld, "conv2", self.conv2, self.bn2, [x], 1)
        for k, blk in enumerate(self.layer1):
            x = self._block(bld, "layer1.%d" % k, blk, x)
        xs = [self._seq(bld, "transition1.%d" % i, t, x) if t is not None else x for i, t in enumerate(self.transition1)]
        for m, mod in enumerate(self.stage2):
            xs = self._module(bld, "stage2.%d" % m, mod, xs)
        xs = [self._seq(bld, "transition2.%d" % i, t, xs[min(i, len(xs) - 1)] if i < 2 else xs[-1]) if t is not None else xs[i]
              for i, t in enumerate(self.transition2)]
        for m, mod in enumerate(self.stage3):
            xs = self._module(bld, "stage3.%d" % m, mod, xs)
        xs = [self._seq(bld, "transition3.%d" % i, t, xs[i] if i < 3 else xs[-1]) if t is not None else xs[i]
              for i, t in enumerate(self.transition3)]
        for m, mod in enumerate(self.stage4):
            xs = self._module(bld, "stage4.%d" % m, mod, xs)
        # ---- head: W*cat(x0, up(x1), up(x2), up(x3)) + b == W0*x0 + up(W1*x1) + up(W2*x2) + up(W3*x3) + b
        c0, bn, c3 = self.last_layer[0], self.last_layer[1], self.last_layer[3]
        scale, shift = bld.packer.bn_affine("last_layer.1", bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, c0.bias)
        terms, off = [], 0
        for i, xi in enumerate(xs):
            w = c0.weight.detach()[:, off:off + xi.C].contiguous()
            t, _, _, _ = bld.conv("last_layer.0.branch%d" % i, [xi], w, 1)
            terms.append(t)
            off += xi.C
        ctot = c0.weight.shape[0]
        cpad = ((ctot + 7) // 8) * 8
        sc = torch.zeros(cpad, dtype=torch.float32, device=bld.device)
        sh = torch.zeros(cpad, dtype=torch.float32, device=bld.device)
        sc[:ctot], sh[:ctot] = scale, shift
        zshape = tuple(terms[0].t.shape)
        low = bld.arena.alloc((B, self.out_dim, terms[0].H, terms[0].W), torch.float32)
        if engine.hilo_enabled():
            z_lo = bld.arena.alloc(zshape, bld.tdtype)
            z = bld.prog.fuse_sum(terms, bld.arena.alloc(zshape, bld.tdtype), ctot, scale=sc, shift=sh, act=1, out_lo=z_lo)
            engine.hilo_last_conv(bld, "last_layer.3", z, Act(z_lo, ctot), c3.weight, c3.bias, out_c8=False, out_nchw=low)
        else:
            z = bld.prog.fuse_sum(terms, bld.arena.alloc(zshape, bld.tdtype), ctot, scale=sc, shift=sh, act=1)
            bld.conv("last_layer.3", [z], c3.weight, 1, bias=c3.bias, out_c8=False, out_nchw=low)
        rz = bld.prog.resize(low, z.H * 4, z.W * 4, dst_nchw=logits_nchw, dst_c8=maps_c8)
        return None, rz

    # ---- engine plumbing (same contract as ResUnet_VB) ------------------------------------------------
    def _check_weights(self, device):
        dt = self.dtype_name or engine.default_dtype()
        fp = (engine.params_fingerprint(self), str(device), dt)
        if fp != self._fingerprint:
            self._plans = {}
            self._packer = engine.WeightPacker(device, dt)
            self._fingerprint = fp
        return dt

    def _plan(self, B, H, W, device, dt):
        key = (B, H, W)
        if key not in self._plans:
            if H % 32 or W % 32:
                raise ValueError("HighResolutionNet: H and W must be divisible by 32")
            arena = engine.Arena(device)
            bld = engine.Builder(device, dt, self._packer, arena)
            x_c8 = arena.alloc((B, 1, H, W, 8), bld.tdtype)
            dummy_in = torch.zeros((B, 3, H, W), dtype=torch.float32, device=device)
            dummy_out = torch.empty((B, self.out_dim, H, W), dtype=torch.float32, device=device)
            ext = Ext("x", dummy_in)
            bld.prog.gather([(ext, c) for c in range(3)], x_c8[:, 0], B, H, W)
            _, rz = self.record(bld, Act(x_c8, 3), logits_nchw=dummy_out)
            bld.prog.finalize()
            self._plans[key] = (bld.prog, rz, arena)
        return self._plans[key]

    def forward(self, x):
        engine.require_cuda(x, "HighResolutionNet.forward")
        if self.training:
            raise RuntimeError("HighResolutionNet (B200 engine) implements inference only: call .eval()")
        x = x.contiguous().float()
        B, Cc, H, W = x.shape
        if Cc != 3:
            raise ValueError("expected 3 input channels, got %d" % Cc)
        dt = self._check_weights(x.device)
        prog, rz, _ = self._plan(B, H, W, x.device, dt)
        out = torch.empty((B, self.out_dim, H, W), dtype=torch.float32, device=x.device)
        prog.call({"x": x}, lambda: setattr(rz, "dst_nchw", out.data_ptr()))
        engine.record_stream(x)
        return out
