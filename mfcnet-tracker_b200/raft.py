"""RAFT-large optical flow on the B200 engine -- the online flow provider of the video loop.

The reference obtains its flow inputs from ``torchvision.models.optical_flow.raft_large`` (a third-party dependency:
scripts/test_multiframe_segmentation_on_videos_v3.py:342-350 builds it, :264-271 calls it on the nearest-neighbour half-size
frames and resizes ``flow / 0.5`` back with bilinear / align_corners=True; src/engine.py:39-53 calls it at full size).  This is
a drop-in for that module: same constructor result (``raft_large()``), same ``forward(image1, image2, num_flow_updates=12)``,
same state_dict keys as torchvision 0.26 ``models/optical_flow/raft.py`` (``feature_encoder.*``, ``context_encoder.*``,
``update_block.{motion_encoder,recurrent_block,flow_head}.*``, ``mask_predictor.*``; the parameter-less norm / ReLU children
keep the Sequential indices).  The modules below only HOLD parameters; the arithmetic runs in libmfcnet_b200.so:

  * every convolution = one fused tensor-core conv (`mfc_conv2d_fwd`): eval BatchNorm folded into scale / shift (context
    encoder), InstanceNorm as per-channel GroupNorm statistics in the conv epilogue + `mfc_gn_finalize` (feature encoder), the
    z and r gates of a ConvGRU as ONE conv with 256 output channels, channel concats as multi-source convs (never
    materialised), the 1x5 / 5x1 GRU kernels as rectangular convs with per-axis padding (MfcConvDesc.in_off);
  * the element-wise glue (norm + ReLU + residual add, tanh / ReLU split of the context, GRU gates) = `mfc_pointwise`;
  * all-pairs correlation volume, its 4-level pyramid, the 9x9 x 4 bilinear lookup (written straight into the C8 planes the
    motion encoder reads), the flow update and the convex upsampling = `mfc_raft_op`.
The whole forward (encoders, `num_flow_updates` update iterations, upsampling) is captured into ONE CUDA graph per input shape.  Only the last
prediction is computed (torchvision returns one per iteration; the reference takes ``[-1]``): forward returns ``[flow]``.
Inference only.
"""
import ctypes as C
import os
import threading

import torch
from torch import nn

from . import abi, engine
from .engine import Act, Ext


_FORWARD_LOCK = threading.Lock()   # a plan's buffers are shared state: one RAFT forward at a time per process


def _cna(cin, cout, k, stride=1, norm=None, act=True):
    """torchvision Conv2dNormActivation: Sequential(conv, [norm], [ReLU]), padding (k-1)//2, bias=True at every RAFT call site."""
    layers = [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=True)]
    if norm == "instance":
        layers.append(nn.InstanceNorm2d(cout))
    elif norm == "batch":
        layers.append(nn.BatchNorm2d(cout))
    if act:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class _ResidualBlock(nn.Module):
    def __init__(self, cin, cout, norm, stride):
        super().__init__()
        self.convnormrelu1 = _cna(cin, cout, 3, stride, norm)
        self.convnormrelu2 = _cna(cout, cout, 3, 1, norm)
        self.downsample = nn.Identity() if stride == 1 else _cna(cin, cout, 1, stride, norm, act=False)
        self.relu = nn.ReLU(inplace=True)


class _Encoder(nn.Module):
    def __init__(self, norm, layers=(64, 64, 96, 128, 256)):
        super().__init__()
        self.norm = norm
        self.convnormrelu = _cna(3, layers[0], 7, 2, norm)
        mk = lambda ci, co, s: nn.Sequential(_ResidualBlock(ci, co, norm, s), _ResidualBlock(co, co, norm, 1))
        self.layer1, self.layer2, self.layer3 = mk(layers[0], layers[1], 1), mk(layers[1], layers[2], 2), mk(layers[2], layers[3], 2)
        self.conv = nn.Conv2d(layers[3], layers[4], 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _MotionEncoder(nn.Module):
    def __init__(self, cin_corr):
        super().__init__()
        self.convcorr1, self.convcorr2 = _cna(cin_corr, 256, 1), _cna(256, 192, 3)
        self.convflow1, self.convflow2 = _cna(2, 128, 7), _cna(128, 64, 3)
        self.conv = _cna(192 + 64, 126, 3)


class _ConvGRU(nn.Module):
    def __init__(self, k, pad):
        super().__init__()
        self.convz, self.convr, self.convq = (nn.Conv2d(384, 128, k, padding=pad) for _ in range(3))


class _RecurrentBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.convgru1, self.convgru2 = _ConvGRU((1, 5), (0, 2)), _ConvGRU((5, 1), (2, 0))


class _FlowHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.conv2, self.relu = nn.Conv2d(128, 256, 3, padding=1), nn.Conv2d(256, 2, 3, padding=1), nn.ReLU(inplace=True)


class _UpdateBlock(nn.Module):
    def __init__(self, cin_corr):
        super().__init__()
        self.motion_encoder, self.recurrent_block, self.flow_head = _MotionEncoder(cin_corr), _RecurrentBlock(), _FlowHead()


class _MaskPredictor(nn.Module):
    def __init__(self):
        super().__init__()
        self.convrelu, self.conv = _cna(128, 256, 3), nn.Conv2d(256, 8 * 8 * 9, 1)


class RAFT(nn.Module):
    """Drop-in for torchvision's ``raft_large()`` module; ``forward(image1, image2, num_flow_updates=12) -> [flow]`` with
    flow (B, 2, H, W) fp32 = the LAST of torchvision's per-iteration predictions."""
    LEVELS, RADIUS, HIDDEN, MULT = 4, 4, 128, 0.25

    def __init__(self):
        super().__init__()
        self.feature_encoder, self.context_encoder = _Encoder("instance"), _Encoder("batch")
        self.corr_block = nn.Module()          # parameter-less in torchvision too (CorrBlock)
        self.update_block = _UpdateBlock(self.LEVELS * (2 * self.RADIUS + 1) ** 2)
        self.mask_predictor = _MaskPredictor()
        self._plans, self._packer, self._fingerprint, self.dtype_name = {}, None, None, None

    def _check_weights(self, device):
        dt = self.dtype_name or engine.default_dtype()
        fp = (engine.params_fingerprint(self), str(device), dt)
        if fp != self._fingerprint:
            self._plans = {}
            self._packer = engine.WeightPacker(device, dt)
            self._fingerprint = fp
        return dt

    # ---- plan recording ----------------------------------------------------------------------------
    def _cnr(self, bld, name, seq, srcs, norm, ident, act=True):
        """Conv2dNormActivation.  BatchNorm: folded, returns the activated Act.  InstanceNorm: returns (raw Act, affine)."""
        conv = seq[0]
        k, s = conv.kernel_size[0], conv.stride[0]
        if norm == "batch":
            bn = seq[1]
            scale, shift = bld.packer.bn_affine(name + ".1", bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, conv.bias)
            return bld.conv(name + ".0", srcs, conv.weight, k, scale=scale, shift=shift, stride=s, pad=conv.padding[0], act=1 if act else 0)[0]
        out, st, info, _ = bld.conv(name + ".0", srcs, conv.weight, k, bias=conv.bias, stride=s, pad=conv.padding[0], want_stats=True)
        C_ = conv.out_channels
        ones, zeros = ident
        aff = bld.group_norm_affine(st, info, ones[:C_], zeros[:C_], C_, C_, out.H * out.W, eps=seq[1].eps)
        return out, aff

    def _materialise(self, bld, raw, aff, relu=True):
        y = bld.arena.alloc(tuple(raw.t.shape), bld.tdtype)
        bld.prog.pointwise(abi.PW_AFFINE_ADD, raw.t, y, a_aff=aff, relu_a=relu)
        return Act(y, raw.C)

    def _block(self, bld, name, blk, x, norm, ident):
        has_down = not isinstance(blk.downsample, nn.Identity)
        out_t = None
        if norm == "batch":
            y = self._cnr(bld, name + ".convnormrelu1", blk.convnormrelu1, [x], norm, ident)
            y = self._cnr(bld, name + ".convnormrelu2", blk.convnormrelu2, [y], norm, ident)
            r = self._cnr(bld, name + ".downsample", blk.downsample, [x], norm, ident, act=False) if has_down else x
            out_t = bld.arena.alloc(tuple(y.t.shape), bld.tdtype)
            bld.prog.pointwise(abi.PW_AFFINE_ADD, y.t, out_t, r=r.t, relu_out=True)
            return Act(out_t, y.C)
        c1, a1 = self._cnr(bld, name + ".convnormrelu1", blk.convnormrelu1, [x], norm, ident)
        y1 = self._materialise(bld, c1, a1)
        c2, a2 = self._cnr(bld, name + ".convnormrelu2", blk.convnormrelu2, [y1], norm, ident)
        r, ra = (self._cnr(bld, name + ".downsample", blk.downsample, [x], norm, ident, act=False) if has_down else (x, None))
        out_t = bld.arena.alloc(tuple(c2.t.shape), bld.tdtype)
        bld.prog.pointwise(abi.PW_AFFINE_ADD, c2.t, out_t, a_aff=a2, relu_a=True, r=r.t, r_aff=ra, relu_out=True)
        return Act(out_t, c2.C)

    def _encoder(self, bld, name, enc, x, ident, **final_kw):
        norm = enc.norm
        y = self._cnr(bld, name + ".convnormrelu", enc.convnormrelu, [x], norm, ident)
        if norm == "instance":
            y = self._materialise(bld, *y)
        for ln, layer in (("layer1", enc.layer1), ("layer2", enc.layer2), ("layer3", enc.layer3)):
            for i, blk in enumerate(layer):
                y = self._block(bld, "%s.%s.%d" % (name, ln, i), blk, y, norm, ident)
        return bld.conv(name + ".conv", [y], enc.conv.weight, 1, bias=enc.conv.bias, **final_kw)[0]

    def _record_update(self, P, B, h, w, dev, dt, arena):
        """Records program U (one update iteration on B frame pairs) and program M (mask predictor + convex upsampling) over
        the buffers of P: vol (pyramid), flow, delta, corr, h, ctx, mask, out."""
        tdt = engine._DTYPES[dt][0]
        # ---- program U: one update iteration (lookup, motion encoder, two ConvGRUs, flow head, flow += delta)
        bu = engine.Builder(dev, dt, self._packer, arena)
        ub, me, rb, fh = self.update_block, self.update_block.motion_encoder, self.update_block.recurrent_block, self.update_block.flow_head
        bu.prog.raft(abi.RAFT_LOOKUP, P["vol"] + [P["flow"], P["corr"]], B, h, w, levels=self.LEVELS, radius=self.RADIUS)
        hA, cA = Act(P["h"], 128), Act(P["ctx"], 128)
        fl = bu.gather_channels([Ext("flow", P["flow"])], B, h, w)
        cv = lambda nm, seq, srcs: bu.conv("update_block.motion_encoder." + nm, srcs, seq[0].weight, seq[0].kernel_size[0], bias=seq[0].bias,
                                           pad=seq[0].padding[0], act=1)[0]
        lanes = os.environ.get("MFC_RAFT_LANES", "1") == "1"   # corr and flow branches of the motion encoder as parallel graph branches (-7 % at one pair)
        if lanes:
            bu.prog.fork()
            bu.prog.lane = 1
        flo = cv("convflow2", me.convflow2, [cv("convflow1", me.convflow1, [fl])])
        bu.prog.lane = 0
        corr = cv("convcorr2", me.convcorr2, [cv("convcorr1", me.convcorr1, [Act(P["corr"], self.LEVELS * (2 * self.RADIUS + 1) ** 2)])])
        if lanes:
            bu.prog.join()
        mot = cv("conv", me.conv, [corr, flo])
        for gn, gru in (("convgru1", rb.convgru1), ("convgru2", rb.convgru2)):
            p = "update_block.recurrent_block.%s." % gn
            kh, kw = gru.convz.kernel_size
            rect = dict(kw=kw, pad_yx=tuple(gru.convz.padding))      # 1x5 / 5x1 kernels, padding (0, 2) / (2, 0)
            wzr = torch.cat([gru.convz.weight.detach(), gru.convr.weight.detach()], 0).float()
            bzr = torch.cat([gru.convz.bias.detach(), gru.convr.bias.detach()], 0).float()
            zr = bu.conv(p + "convzr", [hA, cA, mot, fl], wzr, kh, bias=bzr, **rect)[0]
            rh = bu.arena.alloc(tuple(P["h"].shape), tdt)
            bu.prog.pointwise(abi.PW_GRU_RH, zr.t, rh, r=P["h"], chunks=16)
            q = bu.conv(p + "convq", [Act(rh, 128), cA, mot, fl], gru.convq.weight, kh, bias=gru.convq.bias, **rect)[0]
            bu.prog.pointwise(abi.PW_GRU_UPDATE, zr.t, P["h"], r=q.t, chunks=16)
        f1 = bu.conv("update_block.flow_head.conv1", [hA], fh.conv1.weight, 3, bias=fh.conv1.bias, pad=1, act=1)[0]
        bu.conv("update_block.flow_head.conv2", [f1], fh.conv2.weight, 3, bias=fh.conv2.bias, pad=1, out_c8=False, out_nchw=P["delta"])
        bu.prog.raft(abi.RAFT_FLOW_ADD, [P["flow"], P["delta"]], B, h, w)
        bu.prog.finalize()
        # ---- program M: mask predictor on the final hidden state + convex upsampling
        bm = engine.Builder(dev, dt, self._packer, arena)
        mp = self.mask_predictor
        m1 = bm.conv("mask_predictor.convrelu", [hA], mp.convrelu[0].weight, 3, bias=mp.convrelu[0].bias, pad=1, act=1)[0]
        bm.conv("mask_predictor.conv", [m1], mp.conv.weight, 1, bias=mp.conv.bias, out_c8=False, out_nchw=P["mask"])
        bm.prog.raft(abi.RAFT_UPSAMPLE, [P["flow"], P["mask"], P["out"]], B, h, w, scale=self.MULT)
        bm.prog.finalize()
        return bu.prog, bm.prog

    def _build(self, B, H, W, dev, dt):
        if H % 8 or W % 8:
            raise ValueError("input image H and W should be divisible by 8, instead got %d (h) and %d (w)" % (H, W))
        h, w = H // 8, W // 8
        if (h >> (self.LEVELS - 1)) < 2 or (w >> (self.LEVELS - 1)) < 2:
            raise ValueError("Feature maps are too small to be down-sampled by the correlation pyramid: H and W must be at least 128")
        tdt = engine._DTYPES[dt][0]
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        c8 = lambda b, c, hh, ww: torch.zeros((b, (c + 7) // 8, hh, ww, 8), dtype=tdt, device=dev)
        hw = h * w
        P = {"img": f32(2 * B, 3, H, W), "fmaps": f32(2 * B, 256, h, w), "flow": f32(B, 2, h, w), "delta": f32(B, 2, h, w),
             "mask": f32(B, 576, h, w), "out": f32(B, 2, H, W), "h": c8(B, 128, h, w), "ctx": c8(B, 128, h, w),
             "corr": c8(B, self.LEVELS * (2 * self.RADIUS + 1) ** 2, h, w),
             "vol": [f32(B * hw, h >> l, w >> l) for l in range(self.LEVELS)]}
        ident = (torch.ones(256, dtype=torch.float32, device=dev), torch.zeros(256, dtype=torch.float32, device=dev))
        arena = engine.Arena(dev)
        # ---- program E: both encoders, context split, correlation pyramid
        be = engine.Builder(dev, dt, self._packer, arena)
        lanes = os.environ.get("MFC_RAFT_GRAPH", "1") == "0" and os.environ.get("MFC_LANES", "1") != "0"   # eager mode only: the
        if lanes:                                            # two encoders on concurrent streams (measured: no gain, the host issues)
            be.prog.fork()
            be.prog.lane = 1
        x2 = be.gather_channels([Ext("img", P["img"])], 2 * B, H, W)
        self._encoder(be, "feature_encoder", self.feature_encoder, x2, ident, out_c8=False, out_nchw=P["fmaps"])
        be.prog.lane = 0
        x1 = be.gather_channels([Ext("img1", P["img"][:B])], B, H, W)
        ctx = self._encoder(be, "context_encoder", self.context_encoder, x1, ident)
        be.prog.pointwise(abi.PW_CTX_SPLIT, ctx.t, P["h"], out2=P["ctx"], chunks=16)
        if lanes:
            be.prog.join()
        be.prog.raft(abi.RAFT_CORR_VOLUME, [P["fmaps"][:B], P["fmaps"][B:], P["vol"][0]], B, h, w, C_=256, scale=1.0 / 16.0)
        for l in range(self.LEVELS - 1):
            be.prog.raft(abi.RAFT_POOL, [P["vol"][l], P["vol"][l + 1]], B * hw, h >> l, w >> l)
        be.prog.finalize()
        U, M = self._record_update(P, B, h, w, dev, dt, arena)
        P.update(E=be.prog, U=U, M=M, graphs={}, arena=arena, ident=ident, dt=dt)
        return P

    def forward(self, image1, image2, num_flow_updates=12):
        engine.require_cuda(image1, "RAFT.forward")
        if self.training:
            raise RuntimeError("RAFT (B200 engine) implements inference only: call .eval()")
        if image1.shape != image2.shape or image1.dim() != 4 or image1.shape[1] != 3:
            raise ValueError("input images should have the same shape (B, 3, H, W), instead got %s and %s" % (tuple(image1.shape), tuple(image2.shape)))
        B, _, H, W = image1.shape
        dev = image1.device
        with _FORWARD_LOCK, engine.device_guard(dev):
            dt = self._check_weights(dev)
            key = (B, H, W)
            if key not in self._plans:
                self._plans[key] = self._build(B, H, W, dev, dt)
            P = self._plans[key]
            P["img"][:B].copy_(image1)
            P["img"][B:].copy_(image2)
            P["flow"].zero_()
            n = int(num_flow_updates)
            if os.environ.get("MFC_RAFT_GRAPH", "1") == "0":
                P["E"].run()
                for _ in range(n):
                    P["U"].run()
                P["M"].run()
            elif n in P["graphs"]:
                P["graphs"][n].launch()               # the whole forward -- encoders, n updates, upsampling -- is ONE graph launch
            else:
                P["E"].run()                          # first call eagerly (settles the packed weights), then capture
                for _ in range(n):
                    P["U"].run()
                P["M"].run()
                whole = engine.Program(dev, P["dt"])
                whole.extend(P["E"])
                for _ in range(n):
                    whole.extend(P["U"])
                whole.extend(P["M"])
                P["graphs"][n] = whole.capture()
            out = P["out"].clone()
        engine.record_stream(image1)
        engine.record_stream(image2)
        return [out]


def video_flow(net, frame0, frame_i):
    """The video script's flow provider call (scripts/test_multiframe_segmentation_on_videos_v3.py:266-270): RAFT on the
    nearest-neighbour half-size frames, ``flow / 0.5`` resized to the frame size (bilinear, align_corners=True).
    frame0 / frame_i: (B, 3, H, W) fp32 on the GPU (H, W multiples of 16); returns (B, 2, H, W) fp32."""
    B, _, H, W = frame0.shape
    a, b = frame0[:, :, ::2, ::2], frame_i[:, :, ::2, ::2]   # F.interpolate(scale_factor=0.5, mode='nearest') for even sizes
    low = net(a, b)[-1]
    out = torch.empty((B, 2, H, W), dtype=torch.float32, device=frame0.device)
    g = abi.MfcRaftArgs()
    g.p0, g.p2 = low.data_ptr(), out.data_ptr()
    g.kind, g.B, g.C, g.h, g.w, g.levels, g.radius, g.scale = abi.RAFT_RESIZE_AC, B, 2, H // 2, W // 2, H, W, 2.0
    with engine.device_guard(frame0.device):
        abi.check(abi.load().mfc_raft_op(C.byref(g), torch.cuda.current_stream(frame0.device).cuda_stream))
    return out


class StreamingFlow:
    """The video loop's flow provider with the per-frame work done once (scripts/test_multiframe_segmentation_on_videos_v3.py:
    264-271 calls RAFT on (current, i frames earlier) for i = 1..K-1, every frame): B clips advance in lock step; per step the
    feature and context encoders run on the B NEW half-size frames only -- the feature maps of the K-1 earlier frames are kept
    in a ring -- then the correlation pyramids of the B (K-1) pairs are built and the update iterations run on all pairs at
    once.  `step(frames)` -> list of K-1 tensors (B, 2, H, W): flow(current -> i frames earlier), i = 1..K-1, resized as the
    script does (flow / 0.5, bilinear, align_corners=True).  A clip's first steps see its own frame for the missing history."""

    def __init__(self, net, num_frames, H, W, batch=1, num_flow_updates=12, device=None):
        if H % 16 or W % 16:
            raise ValueError("StreamingFlow: H and W must be multiples of 16 (half-size frames, 1/8-resolution feature maps)")
        self.net, self.K, self.H, self.W, self.B, self.n = net, int(num_frames), H, W, int(batch), int(num_flow_updates)
        self.dev = engine.canonical_device(device if device is not None else next(net.parameters()).device)
        self.P, self.t = None, 0

    def reset(self):
        self.t = 0

    def _build(self):
        net, B, K, dev = self.net, self.B, self.K, self.dev
        dt = net._check_weights(dev)
        Hh, Wh = self.H // 2, self.W // 2
        h, w = Hh // 8, Wh // 8
        hw, NP = h * w, B * (K - 1)
        tdt = engine._DTYPES[dt][0]
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        c8 = lambda b, c: torch.zeros((b, (c + 7) // 8, h, w, 8), dtype=tdt, device=dev)
        P = {"img": f32(B, 3, Hh, Wh), "fm_new": f32(B, 256, h, w), "ring": [f32(B, 256, h, w) for _ in range(K - 1)],
             "h0": c8(B, 128), "ctx0": c8(B, 128), "h": c8(NP, 128), "ctx": c8(NP, 128),
             "flow": f32(NP, 2, h, w), "delta": f32(NP, 2, h, w), "mask": f32(NP, 576, h, w), "out": f32(NP, 2, Hh, Wh),
             "full": f32(NP, 2, self.H, self.W), "corr": c8(NP, net.LEVELS * (2 * net.RADIUS + 1) ** 2),
             "vol": [f32(NP * hw, h >> l, w >> l) for l in range(net.LEVELS)]}
        ident = (torch.ones(256, dtype=torch.float32, device=dev), torch.zeros(256, dtype=torch.float32, device=dev))
        arena = engine.Arena(dev)
        b1 = engine.Builder(dev, dt, net._packer, arena)          # E1: the two encoders on the new frames
        x = b1.gather_channels([Ext("img", P["img"])], B, Hh, Wh)
        net._encoder(b1, "feature_encoder", net.feature_encoder, x, ident, out_c8=False, out_nchw=P["fm_new"])
        ctx = net._encoder(b1, "context_encoder", net.context_encoder, x, ident)
        b1.prog.pointwise(abi.PW_CTX_SPLIT, ctx.t, P["h0"], out2=P["ctx0"], chunks=16)
        b1.prog.finalize()
        b2 = engine.Builder(dev, dt, net._packer, arena)          # E2: pyramids of the B (K-1) pairs
        for j in range(K - 1):
            b2.prog.raft(abi.RAFT_CORR_VOLUME, [P["fm_new"], P["ring"][j], P["vol"][0][j * B * hw:(j + 1) * B * hw]], B, h, w, C_=256,
                         scale=1.0 / 16.0)
        for l in range(net.LEVELS - 1):
            b2.prog.raft(abi.RAFT_POOL, [P["vol"][l], P["vol"][l + 1]], NP * hw, h >> l, w >> l)
        U, M = net._record_update(P, NP, h, w, dev, dt, arena)
        M.raft(abi.RAFT_RESIZE_AC, [P["out"], None, P["full"]], NP, Hh, Wh, C_=2, levels=self.H, radius=self.W, scale=2.0)
        M.finalize()
        P.update(E1=b1.prog, E2=b2.prog, U=U, M=M, graph1=None, graph2=None, arena=arena, ident=ident, dt=dt, fp=net._fingerprint)
        return P

    def step(self, frames):
        engine.require_cuda(frames, "StreamingFlow.step")
        B, K, net = self.B, self.K, self.net
        if tuple(frames.shape) != (B, 3, self.H, self.W):
            raise ValueError("StreamingFlow.step: expected frames of shape %s, got %s" % ((B, 3, self.H, self.W), tuple(frames.shape)))
        with _FORWARD_LOCK, engine.device_guard(self.dev):
            net._check_weights(self.dev)
            if self.P is None or self.P["fp"] != net._fingerprint:
                self.P, self.t = self._build(), 0
            P = self.P
            P["img"].copy_(frames[:, :, ::2, ::2])                  # F.interpolate(scale_factor=0.5, mode='nearest')
            graphs = os.environ.get("MFC_RAFT_GRAPH", "1") != "0"
            if P["graph1"] is not None:
                P["graph1"].launch()
            else:
                P["E1"].run()
                if graphs:
                    P["graph1"] = P["E1"].capture()
            if self.t == 0:                                         # no history yet: every earlier frame = this frame
                for r in P["ring"]:
                    r.copy_(P["fm_new"])
            for j in range(K - 1):                                  # hidden state / context of pair (j, b) = those of frame b
                P["h"][j * B:(j + 1) * B].copy_(P["h0"])
                P["ctx"][j * B:(j + 1) * B].copy_(P["ctx0"])
            P["flow"].zero_()
            if P["graph2"] is not None:
                P["graph2"].launch()
            else:
                P["E2"].run()
                for _ in range(self.n):
                    P["U"].run()
                P["M"].run()
                if graphs:
                    whole = engine.Program(self.dev, P["dt"])
                    whole.extend(P["E2"])
                    for _ in range(self.n):
                        whole.extend(P["U"])
                    whole.extend(P["M"])
                    P["graph2"] = whole.capture()
            for j in range(K - 2, 0, -1):                           # age the ring: slot j <- slot j-1, slot 0 <- the new frame
                P["ring"][j].copy_(P["ring"][j - 1])
            P["ring"][0].copy_(P["fm_new"])
            self.t += 1
            out = P["full"].clone()
        engine.record_stream(frames)
        return [out[j * B:(j + 1) * B] for j in range(K - 1)]


def raft_large(**_ignored):
    """torchvision.models.optical_flow.raft_large() without the weight download: load a state_dict into the result."""
    return RAFT()
