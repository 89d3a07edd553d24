"""Temporal context fusion heads ``MultiFrameNetBasic`` / ``MultiFrameNetLarge`` on the B200 engine.

Drop-in for models/multiframe_model.py:14-205: same constructors, ``forward(x)`` on the
channel-concatenated tensor, same state_dict keys (``multiframe_net.{0,3,6,9}.weight``, BatchNorm
at ``{1,4,7}``, and the persistent ``grid`` buffer of the Basic variant).  Arithmetic:

  * the 4-conv stack (:62-73, :191-202): 11x11 / 3x3 / 3x3 / 1x1 tensor-core convs with eval-mode
    BatchNorm folded into the epilogue scale/shift and ReLU fused; the 11x11's 121 taps reuse one
    staged halo tile (no im2col).
  * Basic's flow warp (:89-170): one kernel warps all class maps and the depth map of every
    earlier frame (the reference issues (K-1)(N+1) single-channel grid_samples plus
    stack/permute/cat copies), reproducing the cropped 576x720 grid quirk.
  * the channel concat (:424-436 in the wrappers) is never materialised: per-frame class maps stay
    in their own C8 planes and enter the first conv as separate sources.
"""
import torch
from torch import nn

from . import abi, engine
from .engine import Act, Ext


def _make_grid():
    """Normalised 576x720 mesh of `_create_mesh_grid` (models/multiframe_model.py:172-185): fp32,
    channel 0 = x, channel 1 = y."""
    H, W = 576, 720
    ys = 2.0 * torch.arange(0, H) / (H - 1) - 1.0
    xs = 2.0 * torch.arange(0, W) / (W - 1) - 1.0
    gy = ys.view(H, 1).expand(H, W)
    gx = xs.view(1, W).expand(H, W)
    return torch.stack((gx, gy), dim=0).float().unsqueeze(0).contiguous()


def _stack(in_ch, N, K):
    NK = N * K
    return nn.Sequential(
        nn.Conv2d(in_ch, NK, kernel_size=11, stride=1, padding=5, bias=False), nn.BatchNorm2d(NK), nn.ReLU(),
        nn.Conv2d(NK, NK, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(NK), nn.ReLU(),
        nn.Conv2d(NK, NK, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(NK), nn.ReLU(),
        nn.Conv2d(NK, N, kernel_size=1, stride=1, padding=0, bias=False))


def record_stack(bld, name, seq, srcs, first_weight_channel, out_nchw):
    """Records the 4-conv fusion stack over concat sources `srcs`; the result is written as fp32 NCHW.
    With at most 16 stack channels (N*K <= 16: the 3-frame, 5-class configuration) the final bias-free 1x1 conv is
    evaluated in fp32 inside the third conv's epilogue (Builder.conv(head=...)): one launch and one tensor round trip less,
    and the last activation / weights are never rounded to fp16."""
    x = srcs
    fwc = first_weight_channel
    layers = ((0, 1), (3, 4), (6, 7))
    for li, (ci, bi) in enumerate(layers):
        conv, bn = seq[ci], seq[bi]
        sc, sh = bld.packer.bn_affine("%s.%d" % (name, bi), bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
        k = conv.kernel_size[0]
        last = li == len(layers) - 1
        if last and conv.out_channels <= 16 and seq[9].bias is None and engine.head_fusion_enabled():
            _, _, _, io = bld.conv("%s.%d+9" % (name, ci), x, conv.weight, k, scale=sc, shift=sh, pad=conv.padding[0], act=1,
                                   first_weight_channel=fwc, out_c8=False, out_nchw=out_nchw,
                                   head=(seq[9].weight.detach().reshape(seq[9].out_channels, -1), None))
            return io
        y, _, _, _ = bld.conv("%s.%d" % (name, ci), x, conv.weight, k, scale=sc, shift=sh, pad=conv.padding[0], act=1,
                              first_weight_channel=fwc, want_lo=last and engine.hilo_enabled())
        x, fwc = [y], None
    if x[0].lo is not None:     # wider stacks (N*K > 16): the final 1x1 conv sees its input and weights as (hi, lo) pairs
        _, _, _, io = engine.hilo_last_conv(bld, "%s.9" % name, x[0], x[0].lo, seq[9].weight, seq[9].bias, out_c8=False, out_nchw=out_nchw)
        return io
    _, _, _, io = bld.conv("%s.9" % name, x, seq[9].weight, 1, out_c8=False, out_nchw=out_nchw)
    return io


class MultiFrameNetBase(nn.Module):
    """Channel arithmetic of models/multiframe_model.py:14-32."""

    def __init__(self, num_classes, num_frames, has_base_perframe_model_trained=False, with_optflow=False, with_depth=False):
        super().__init__()
        self.num_classes, self.num_frames = num_classes, num_frames
        self.with_optflow, self.with_depth = with_optflow, with_depth
        self.in_channels = num_frames * num_classes
        if with_optflow:
            self.in_channels += 2 * (num_frames - 1)
        if with_depth:
            self.in_channels += num_frames
        self._plans = {}
        self._packer = None
        self._fingerprint = None
        self.dtype_name = None

    def _check_weights(self, device):
        dt = self.dtype_name or engine.default_dtype()
        fp = (engine.params_fingerprint(self), str(device), dt)
        if fp != self._fingerprint:
            self._plans = {}
            self._packer = engine.WeightPacker(device, dt)
            self._fingerprint = fp
        return dt

    def expected_input_channels(self):
        return self.in_channels

    def _split(self, x):
        """Channel views of the concatenated input, keyed like the wrappers' per-frame inputs."""
        N, K = self.num_classes, self.num_frames
        t = {("maps", i): x[:, i * N:(i + 1) * N] for i in range(K)}
        nf = 2 * (K - 1) if self.with_optflow else 0
        for i in range(nf // 2):
            t[("flow", i)] = x[:, N * K + 2 * i:N * K + 2 * i + 2]
        if self.with_depth:
            for i in range(K):
                t[("depth", i)] = x[:, N * K + nf + i:N * K + nf + i + 1]
        return t

    # subclasses: record(bld, maps, flows, depths, out_nchw) with per-frame inputs
    def forward(self, x):
        engine.require_cuda(x, type(self).__name__ + ".forward")
        if self.training:
            raise RuntimeError("%s (B200 engine) implements inference only: call .eval()" % type(self).__name__)
        x = x.contiguous().float()
        B, Cc, H, W = x.shape
        cin = self.expected_input_channels()
        if Cc != cin:
            raise ValueError("expected %d input channels, got %d" % (cin, Cc))
        dt = self._check_weights(x.device)
        key = (B, H, W)
        K = self.num_frames
        if key not in self._plans:
            arena = engine.Arena(x.device)
            bld = engine.Builder(x.device, dt, self._packer, arena)
            views = self._split(torch.zeros((B, Cc, H, W), dtype=torch.float32, device=x.device))
            out = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=x.device)
            maps = [bld.gather_channels([Ext(("maps", i), views[("maps", i)])], B, H, W) for i in range(K)]
            flows = [Ext(("flow", i), views[("flow", i)]) for i in range(K - 1)] if self.with_optflow else None
            depths = [Ext(("depth", i), views[("depth", i)]) for i in range(K)] if self.with_depth else None
            io = self.record(bld, maps, flows, depths, out)
            bld.prog.finalize()
            self._plans[key] = (bld.prog, io, arena)
        prog, io, _ = self._plans[key]
        out = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=x.device)
        prog.call(self._split(x), lambda: setattr(io, "y_nchw", out.data_ptr()))
        engine.record_stream(x)
        return out


class MultiFrameNetLarge(MultiFrameNetBase):
    """Drop-in for models/multiframe_model.py:187-205."""

    def __init__(self, num_classes, num_frames, has_base_perframe_model_trained=False, with_optflow=False, with_depth=False):
        super().__init__(num_classes, num_frames, has_base_perframe_model_trained, with_optflow, with_depth)
        self.multiframe_net = _stack(self.in_channels, num_classes, num_frames)

    def _input_has_flow(self):
        return self.with_optflow

    def record(self, bld, maps, flows, depths, out_nchw):
        """maps: K Acts (N real channels each); flows: K-1 Ext (B,2,H,W) or None; depths: K Ext
        (B,1,H,W) or None.  Returns the io struct of the last conv (its y_nchw is the output)."""
        B, H, W = maps[0].B, maps[0].H, maps[0].W
        N, K = self.num_classes, self.num_frames
        if (flows is not None) != self.with_optflow or (depths is not None) != self.with_depth:
            raise ValueError("optflow / depth inputs do not match how the fusion head was constructed")
        srcs = list(maps)
        fwc = [i * N for i in range(K)]
        aux = (list(flows) if flows else []) + (list(depths) if depths else [])
        if aux:
            srcs.append(bld.gather_channels(aux, B, H, W))
            fwc.append(N * K)
        return record_stack(bld, "multiframe_net", self.multiframe_net, srcs, fwc, out_nchw)


class MultiFrameNetBasic(MultiFrameNetBase):
    """Drop-in for models/multiframe_model.py:51-185 (flow-warping variant)."""

    def __init__(self, num_classes, num_frames, has_base_perframe_model_trained=False, with_optflow=False, with_depth=False):
        super().__init__(num_classes, num_frames, has_base_perframe_model_trained, with_optflow, with_depth)
        self.in_channels = num_classes * num_frames + (num_frames if with_depth else 0)
        self.multiframe_net = _stack(self.in_channels, num_classes, num_frames)
        self.register_buffer("grid", _make_grid())

    def expected_input_channels(self):
        N, K = self.num_classes, self.num_frames
        return N * K + (2 * (K - 1) if self.with_optflow else 0) + (K if self.with_depth else 0)

    def _input_has_flow(self):
        return self.with_optflow

    def record(self, bld, maps, flows, depths, out_nchw):
        B, H, W = maps[0].B, maps[0].H, maps[0].W
        N, K = self.num_classes, self.num_frames
        if (flows is not None) != self.with_optflow or (depths is not None) != self.with_depth:
            raise ValueError("optflow / depth inputs do not match how the fusion head was constructed")
        fwc = [i * N for i in range(K)]
        if not self.with_optflow:
            srcs = list(maps)
            if depths:
                srcs.append(bld.gather_channels(list(depths), B, H, W))
                fwc.append(N * K)
            return record_stack(bld, "multiframe_net", self.multiframe_net, srcs, fwc, out_nchw)
        if K > abi.MFC_MAX_SRC:
            raise ValueError("flow warp supports at most %d frames" % abi.MFC_MAX_SRC)
        if H > self.grid.shape[2] or W > self.grid.shape[3]:
            raise ValueError("input %dx%d exceeds the stored %dx%d grid" % (H, W, self.grid.shape[2], self.grid.shape[3]))
        prog = bld.prog
        a = abi.MfcWarpArgs()
        a.B, a.H, a.W, a.K = B, H, W, K
        a.seg_chunks = maps[0].chunks
        a.grid_h, a.grid_w = self.grid.shape[2], self.grid.shape[3]
        a.grid = self.grid.data_ptr()
        a.dtype = prog.cdtype
        prog.keep.append(self.grid)
        warped = [maps[0]]
        for f in range(1, K):
            o = bld.arena.alloc(tuple(maps[f].t.shape), bld.tdtype)
            a.seg[f], a.seg_bstride[f] = maps[f].t.data_ptr(), maps[f].bstride
            a.seg_out[f], a.seg_out_bstride[f] = o.data_ptr(), o.stride(0) * o.element_size()
            warped.append(Act(o, N))
            prog.keep += [maps[f].t, o]

            def set_flow(t, a=a, i=f - 1):
                a.flow[i] = t.data_ptr()
                a.flow_bstride[i] = t.stride(0)
            set_flow(flows[f - 1].t)
            prog.bind(flows[f - 1].key, set_flow)
        srcs = warped
        if depths:
            dplane = bld.arena.alloc((B, 1, H, W, 8), bld.tdtype)
            for f in range(K):
                def set_depth(t, a=a, f=f):
                    a.depth[f] = t.data_ptr()
                    a.depth_bstride[f] = t.stride(0)
                set_depth(depths[f].t)
                prog.bind(depths[f].key, set_depth)
            a.depth_out, a.depth_out_bstride = dplane.data_ptr(), dplane.stride(0) * dplane.element_size()
            prog.keep.append(dplane)
            srcs = srcs + [Act(dplane, K)]
            fwc.append(N * K)
        nb = B * H * W * ((K - 1) * (2 * maps[0].chunks * 16 + 8) + (16 + 4 * K if depths else 0))
        prog.warp(a, nb)
        return record_stack(bld, "multiframe_net", self.multiframe_net, srcs, fwc, out_nchw)
