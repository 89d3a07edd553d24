"""ResUnet_VB single-frame-context (SFC) network as a drop-in nn.Module on the B200 engine.

Mirrors the constructor, forward signature and state_dict keys of the reference
``ResUnet_VB`` (models/resunet.py:97-180) so reference checkpoints load unchanged:
``init_conv``, ``downs.{i}.0.{block1,block2}.{proj,norm}``, ``downs.{i}.0.res_conv``,
``downs.{i}.1.1`` / ``downs.3.1``, ``mid_block``, ``ups.{i}.0``, ``ups.{i}.1.1`` / ``ups.3.1``,
``final_res_block``, ``output_layer``.  The modules below only HOLD parameters; the arithmetic
runs in libmfcnet_b200.so:

  * WeightStandardizedConv2d (models/resunet.py:51-64): weights standardised once per
    checkpoint (they are constants in eval) and packed for the tensor-core conv kernel.
  * Block = WS-conv3x3 -> GroupNorm -> SiLU (models/resunet.py:68-80): the conv's epilogue
    emits per-tile sum / sum-of-squares, a tiny finalise kernel turns them into a per-(sample,
    channel) affine, and the CONSUMER applies silu(affine(x)) while staging its input tile, so
    the normalised tensor is never written to HBM.
  * ResnetBlock (models/resunet.py:82-95): the 1x1 res_conv adds silu(affine(h2)) in its
    epilogue; skip-concats (forward :168,:174) are extra sources of the next conv (zero-copy).
  * Downsample (models/resunet.py:45-49) = pixel-unshuffle + 1x1 conv, expressed as one 2x2
    stride-2 conv (the weight reshape [Cout, C*4] -> [Cout, C, 2, 2] is exact);
    Upsample (:39-43) = nearest x2 fused into the 3x3 conv's tile loader.
"""
import torch
from torch import nn

from . import engine
from .engine import Act, Ext


class _WSConv(nn.Conv2d):
    """Parameter holder for WeightStandardizedConv2d (weights are standardised by the engine)."""


class _Block(nn.Module):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.proj = _WSConv(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)


class _ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.block1 = _Block(dim, dim_out, groups)
        self.block2 = _Block(dim_out, dim_out, groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out, self.groups = dim, dim_out, groups


def _downsample(dim, dim_out):
    # index 0 is the parameter-less pixel-unshuffle of the reference; the conv must be index 1
    return nn.Sequential(nn.Identity(), nn.Conv2d(dim * 4, dim_out, 1))


def _upsample(dim, dim_out):
    return nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim_out, 3, padding=1))


def _residual_as_source():
    """MFC_RES_AS_SOURCE=0: keep the 1x1 res_conv's `+ silu(GN(h2))` in the conv epilogue (residual operand) instead of feeding
    h2 through the conv as extra input channels with identity weights (measurement switch)."""
    import os
    return os.environ.get("MFC_RES_AS_SOURCE", "1") != "0"


def record_resnet_block(bld, name, blk, srcs, tail=None, tail_kw=None):
    """Records a ResnetBlock over the channel-concat `srcs`; returns the materialised output Act.

    `tail` = a 1x1 nn.Conv2d applied to the block's output (the network's output_layer): both are linear after the SiLU, so
    it is composed into the block's own 1x1 conv -- tail(res_conv(x) + s) = (Wt Wr) x + Wt s + (Wt br + bt) -- and the block
    output is never written.  Returns the result of Builder.conv for that fused layer (tail_kw: its output arguments)."""
    H, W = srcs[0].H, srcs[0].W
    h1, st1, info1, _ = bld.conv(name + ".block1.proj", srcs, bld.packer.standardized(name + ".block1.proj", blk.block1.proj.weight),
                                 3, bias=blk.block1.proj.bias, pad=1, want_stats=True)
    aff1 = bld.group_norm_affine(st1, info1, blk.block1.norm.weight, blk.block1.norm.bias, blk.dim_out, blk.groups, H * W,
                                 blk.block1.norm.eps)
    h2, st2, info2, _ = bld.conv(name + ".block2.proj", [h1.with_affine(aff1)],
                                 bld.packer.standardized(name + ".block2.proj", blk.block2.proj.weight), 3,
                                 bias=blk.block2.proj.bias, pad=1, want_stats=True)
    aff2 = bld.group_norm_affine(st2, info2, blk.block2.norm.weight, blk.block2.norm.bias, blk.dim_out, blk.groups, H * W,
                                 blk.block2.norm.eps)
    h2 = h2.with_affine(aff2)
    if isinstance(blk.res_conv, nn.Conv2d) and _residual_as_source():
        # res_conv(x) + silu(GN(h2)) = ONE 1x1 conv over the concat [x..., silu(GN(h2))] with the weights [Wr | I]: the GroupNorm
        # affine + SiLU of h2 runs in the tile loader like for any other source, the identity columns pass it through the tensor
        # core exactly (fp16 x 1.0, fp32 accumulate), and the epilogue is the plain one (no residual operand to fetch).
        C_ = blk.dim_out
        wr = blk.res_conv.weight.detach().float().reshape(C_, -1)
        w = torch.cat([wr, torch.eye(C_, dtype=wr.dtype, device=wr.device)], 1)
        b = blk.res_conv.bias.detach().float()
        if tail is not None:
            wt = tail.weight.detach().float().reshape(tail.out_channels, C_)
            w = wt @ w
            b = wt @ b + (tail.bias.detach().float() if tail.bias is not None else 0.0)
            return bld.conv(name + ".res_conv+tail", list(srcs) + [h2], w.reshape(tail.out_channels, -1, 1, 1), 1, bias=b, **(tail_kw or {}))
        out, _, _, _ = bld.conv(name + ".res_conv+h2", list(srcs) + [h2], w.reshape(C_, -1, 1, 1), 1, bias=b)
        return out
    if isinstance(blk.res_conv, nn.Conv2d):
        out, _, _, _ = bld.conv(name + ".res_conv", srcs, blk.res_conv.weight, 1, bias=blk.res_conv.bias, residual=h2)
    else:
        if len(srcs) != 1 or srcs[0].affine is not None:
            raise RuntimeError("identity residual needs one materialised source")
        out_t = bld.arena.alloc(tuple(h2.t.shape), bld.tdtype)
        out = bld.prog.affine_silu_add(h2, srcs[0], out_t)
    if tail is not None:
        return bld.conv(name + ".tail", [out], tail.weight, 1, bias=tail.bias, **(tail_kw or {}))
    return out


class ResUnet_VB(nn.Module):
    """Drop-in for models/resunet.py:97 ``ResUnet_VB``; returns raw logits (B, out_dim, H, W) fp32."""

    def __init__(self, channels, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), resnet_block_groups=8):
        super().__init__()
        self.channels = channels
        init_dim = dim if init_dim is None else init_dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim] + [dim * m for m in dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        g = resnet_block_groups
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        n = len(in_out)
        for i, (din, dout) in enumerate(in_out):
            last = i >= n - 1
            self.downs.append(nn.ModuleList([
                _ResnetBlock(din, din, g),
                _downsample(din, dout) if not last else nn.Conv2d(din, dout, 3, padding=1)]))
        self.mid_block = _ResnetBlock(dims[-1], dims[-1], g)
        for i, (din, dout) in enumerate(reversed(in_out)):
            last = i == n - 1
            self.ups.append(nn.ModuleList([
                _ResnetBlock(dout + din, dout, g),
                _upsample(dout, din) if not last else nn.Conv2d(dout, din, 3, padding=1)]))
        self.final_res_block = _ResnetBlock(dim * 2, dim, g)
        self.out_dim = channels if out_dim is None else out_dim
        self.output_layer = nn.Conv2d(dim, self.out_dim, 1, bias=True)
        self._plans = {}
        self._packer = None
        self._fingerprint = None
        self.dtype_name = None  # None -> engine.default_dtype() at first use

    # ---- engine plumbing -------------------------------------------------------------------------
    def _check_weights(self, device):
        dt = self.dtype_name or engine.default_dtype()
        fp = (engine.params_fingerprint(self), str(device), dt)
        if fp != self._fingerprint:
            self._plans = {}
            self._packer = engine.WeightPacker(device, dt)
            self._fingerprint = fp
        return dt

    def record(self, bld, x_act, logits_nchw=None, maps_c8=None):
        """Record the whole SFC forward on `x_act` (C8 input).  Writes fp32 NCHW logits into
        `logits_nchw` and/or C8 class maps into `maps_c8` ([B,chunks,H,W,8])."""
        x, _, _, _ = bld.conv("init_conv", [x_act], self.init_conv.weight, 7, bias=self.init_conv.bias, pad=3)
        stem = x
        skips = []
        for i, (blk, down) in enumerate(self.downs):
            x = record_resnet_block(bld, "downs.%d.0" % i, blk, [x])
            skips.append(x)
            if isinstance(down, nn.Sequential):
                cv = down[1]
                w2 = cv.weight.detach().reshape(cv.weight.shape[0], cv.weight.shape[1] // 4, 2, 2)
                x, _, _, _ = bld.conv("downs.%d.1.1" % i, [x], w2, 2, bias=cv.bias, stride=2, pad=0)
            else:
                x, _, _, _ = bld.conv("downs.%d.1" % i, [x], down.weight, 3, bias=down.bias, pad=1)
        x = record_resnet_block(bld, "mid_block", self.mid_block, [x])
        for i, (blk, up) in enumerate(self.ups):
            x = record_resnet_block(bld, "ups.%d.0" % i, blk, [x, skips.pop()])
            if isinstance(up, nn.Sequential):
                cv = up[1]
                x, _, _, _ = bld.conv("ups.%d.1.1" % i, [x], cv.weight, 3, bias=cv.bias, pad=1, upsample=2)
            else:
                x, _, _, _ = bld.conv("ups.%d.1" % i, [x], up.weight, 3, bias=up.bias, pad=1)
        out, _, _, io = record_resnet_block(bld, "final_res_block", self.final_res_block, [x, stem], tail=self.output_layer,
                                            tail_kw=dict(out_c8=maps_c8 is not None, y_c8=maps_c8, out_nchw=logits_nchw))
        return out, io

    def _plan(self, B, H, W, device, dt):
        key = (B, H, W)
        if key not in self._plans:
            if H % (2 ** (len(self.downs) - 1)) or W % (2 ** (len(self.downs) - 1)):
                raise ValueError("ResUnet_VB: H and W must be divisible by %d" % (2 ** (len(self.downs) - 1)))
            arena = engine.Arena(device)
            bld = engine.Builder(device, dt, self._packer, arena)
            x_c8 = arena.alloc((B, (self.channels + 7) // 8, H, W, 8), bld.tdtype)
            dummy_in = torch.zeros((B, self.channels, H, W), dtype=torch.float32, device=device)
            dummy_out = torch.empty((B, self.out_dim, H, W), dtype=torch.float32, device=device)
            ext = Ext("x", dummy_in)
            for c0 in range(0, self.channels, 8):
                bld.prog.gather([(ext, c) for c in range(c0, min(c0 + 8, self.channels))], x_c8[:, c0 // 8], B, H, W)
            _, io = self.record(bld, Act(x_c8, self.channels), logits_nchw=dummy_out)
            bld.prog.finalize()
            self._plans[key] = (bld.prog, io, arena)
        return self._plans[key]

    def forward(self, captimgs, *args, **kwargs):
        engine.require_cuda(captimgs, "ResUnet_VB.forward")
        if self.training:
            raise RuntimeError("ResUnet_VB (B200 engine) implements inference only: call .eval() and run under torch.no_grad()")
        x = captimgs.contiguous().float()
        B, Cc, H, W = x.shape
        if Cc != self.channels:
            raise ValueError("expected %d input channels, got %d" % (self.channels, Cc))
        dt = self._check_weights(x.device)
        prog, io, _ = self._plan(B, H, W, x.device, dt)
        out = torch.empty((B, self.out_dim, H, W), dtype=torch.float32, device=x.device)
        prog.call({"x": x}, lambda: setattr(io, "y_nchw", out.data_ptr()))
        # x and out are referenced by raw pointer until the kernels run: keep them alive on this stream
        engine.record_stream(x)
        return out
