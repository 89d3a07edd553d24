"""MFCNet wrappers ``XMulti{Basic,Large}``: per-frame SFC network + temporal fusion head.

Drop-in for the wrapper pattern of models/multiframe_model.py:408-471 (``HRNetMultiBasic`` /
``HRNetMultiLarge``): constructor ``(num_classes, num_frames, pretrained=True, loadpath=None,
optflow_inputs=False, depth_inputs=False)``, children named ``base_model`` and
``multiframe_net``, and ``forward(x, optflow=None, depth=None)`` with ``x`` a list of K
(B,3,H,W) frames (index 0 = current frame), ``optflow`` K-1 (B,2,H,W) flows and ``depth`` K
(B,1,H,W) maps.  The reference has no ResUNet wrapper; ``ResUNetMultiBasic`` /
``ResUNetMultiLarge`` follow the HRNet wrapper exactly (raw logits into the fusion head).

One forward = one native command list: the K*B frames are gathered to C8, the SFC network runs
over them in large sub-batches (see `_sub_batch`), its head writes each frame's class maps into
a per-frame C8 plane, and the fusion head reads those planes (plus flow / depth) as separate
concat sources -- the (B, N*K+..., H, W) tensor of models/multiframe_model.py:429-436 is never
materialised.
"""
import os

import torch
from torch import nn

from . import engine
from .engine import Act, Ext
from .fusion import MultiFrameNetBasic, MultiFrameNetLarge
from .hrnet import HighResolutionNet
from .resunet import ResUnet_VB
from .ternausnet import TernausNet16


def _sub_batch(n_frames, H, W):
    env = os.environ.get("MFC_B200_SUBBATCH")
    if env:
        sb = max(1, min(n_frames, int(env)))
        while n_frames % sb:      # sub-batches replay one arena: they must all have the same size
            sb -= 1
        return sb
    # pixels per SFC pass.  Measured on B200 (ResUNet-16, 24 frames of 480x640): sub-batches of 4 / 8 / 12 / 24
    # frames give 910 / 1080 / 1227 / 1331 frames/s -- amortising the ~61 launches of a pass over more frames
    # beats keeping the 39 MB (4-frame) layer outputs L2-resident, so the cap is only a memory bound.
    budget = 32 * 480 * 640
    sb = max(1, min(n_frames, budget // (H * W)))
    while n_frames % sb:
        sb -= 1
    return sb


class _MultiFrame(nn.Module):
    head = "logits"  # what the fusion head is fed: raw logits (HRNet/ResUNet convention)

    def __init__(self, base_model, fusion_cls, num_classes, num_frames, optflow_inputs, depth_inputs):
        super().__init__()
        self.base_model = base_model
        self.multiframe_net = fusion_cls(num_classes=num_classes, num_frames=num_frames, has_base_perframe_model_trained=True,
                                         with_optflow=optflow_inputs, with_depth=depth_inputs)
        self.num_classes, self.num_frames = num_classes, num_frames
        self.optflow_inputs, self.depth_inputs = optflow_inputs, depth_inputs
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

    def build_plan(self, B, H, W, device, dt, packer, want_base_logits=False):
        """Returns a dict with the program, its output io struct and the static buffers."""
        K, N = self.num_frames, self.num_classes
        F = K * B
        arena_sfc = engine.Arena(device)
        arena = engine.Arena(device)
        main = engine.Builder(device, dt, packer, arena)
        cin = self.base_model.channels
        x_c8 = torch.empty((F, (cin + 7) // 8, H, W, 8), dtype=main.tdtype, device=device)
        maps_c8 = torch.empty((F, (N + 7) // 8, H, W, 8), dtype=main.tdtype, device=device)
        base_logits = torch.empty((F, N, H, W), dtype=torch.float32, device=device) if want_base_logits else None
        ph = torch.zeros((B, cin, H, W), dtype=torch.float32, device=device)
        for k in range(K):
            ext = Ext(("frame", k), ph)
            for c0 in range(0, cin, 8):
                main.prog.gather([(ext, c) for c in range(c0, min(c0 + 8, cin))], x_c8[k * B:(k + 1) * B, c0 // 8], B, H, W)
        sb = _sub_batch(F, H, W)
        for lo in range(0, F, sb):
            arena_sfc.reset()
            sub = engine.Builder(device, dt, packer, arena_sfc)
            self.base_model.record(sub, Act(x_c8[lo:lo + sb], cin), maps_c8=maps_c8[lo:lo + sb],
                                   logits_nchw=None if base_logits is None else base_logits[lo:lo + sb])
            main.prog.extend(sub.prog)
        maps = [Act(maps_c8[k * B:(k + 1) * B], N) for k in range(K)]
        flows = depths = None
        if self.optflow_inputs:
            phf = torch.zeros((B, 2, H, W), dtype=torch.float32, device=device)
            flows = [Ext(("flow", i), phf) for i in range(K - 1)]
        if self.depth_inputs:
            phd = torch.zeros((B, 1, H, W), dtype=torch.float32, device=device)
            depths = [Ext(("depth", i), phd) for i in range(K)]
        out = torch.empty((B, N, H, W), dtype=torch.float32, device=device)
        io = self.multiframe_net.record(main, maps, flows, depths, out)
        main.prog.finalize()
        return {"prog": main.prog, "io": io, "out": out, "x_c8": x_c8, "maps_c8": maps_c8, "base_logits": base_logits,
                "arenas": (arena, arena_sfc), "sub_batch": sb}

    def forward(self, x, optflow=None, depth=None):
        if self.training:
            raise RuntimeError("%s (B200 engine) implements inference only: call .eval()" % type(self).__name__)
        K = self.num_frames
        if len(x) != K:
            raise ValueError("expected %d frames, got %d" % (K, len(x)))
        if (optflow is not None) != self.optflow_inputs or (depth is not None) != self.depth_inputs:
            raise ValueError("optflow / depth arguments do not match optflow_inputs / depth_inputs of the constructor")
        engine.require_cuda(x[0], type(self).__name__ + ".forward")
        B, _, H, W = x[0].shape
        dev = x[0].device
        dt = self._check_weights(dev)
        key = (B, H, W)
        if key not in self._plans:
            self._plans[key] = self.build_plan(B, H, W, dev, dt, self._packer)
        plan = self._plans[key]
        tensors = {("frame", k): x[k].contiguous().float() for k in range(K)}
        if optflow is not None:
            tensors.update({("flow", i): optflow[i].contiguous().float() for i in range(K - 1)})
        if depth is not None:
            tensors.update({("depth", i): depth[i].contiguous().float() for i in range(K)})
        out = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=dev)
        plan["prog"].call(tensors, lambda: setattr(plan["io"], "y_nchw", out.data_ptr()))
        for t in tensors.values():
            engine.record_stream(t)
        return out


# NOTE: like the reference wrappers (models/multiframe_model.py:408-423) the constructors do not
# read `loadpath`; it only records that a trained SFC checkpoint exists.  Scripts load the base
# weights themselves via `model.base_model.load_state_dict(...)`
# (scripts/train_multiframe_detection.py:115-118).


class ResUNetMultiBasic(_MultiFrame):
    def __init__(self, num_classes, num_frames, pretrained=True, loadpath=None, optflow_inputs=False, depth_inputs=False, dim=16):
        base = ResUnet_VB(channels=3, dim=dim, out_dim=num_classes)
        super().__init__(base, MultiFrameNetBasic, num_classes, num_frames, optflow_inputs, depth_inputs)


class ResUNetMultiLarge(_MultiFrame):
    def __init__(self, num_classes, num_frames, pretrained=True, loadpath=None, optflow_inputs=False, depth_inputs=False, dim=16):
        base = ResUnet_VB(channels=3, dim=dim, out_dim=num_classes)
        super().__init__(base, MultiFrameNetLarge, num_classes, num_frames, optflow_inputs, depth_inputs)


class HRNetMultiBasic(_MultiFrame):
    """Drop-in for models/multiframe_model.py:408-438 (HRNet-W48 base, raw logits into the fusion head)."""

    def __init__(self, num_classes=2, num_frames=1, pretrained=True, loadpath=None, optflow_inputs=False, depth_inputs=False):
        super().__init__(HighResolutionNet(num_classes=num_classes), MultiFrameNetBasic, num_classes, num_frames, optflow_inputs,
                         depth_inputs)


class HRNetMultiLarge(_MultiFrame):
    """Drop-in for models/multiframe_model.py:441-471."""

    def __init__(self, num_classes=2, num_frames=1, pretrained=True, loadpath=None, optflow_inputs=False, depth_inputs=False):
        super().__init__(HighResolutionNet(num_classes=num_classes), MultiFrameNetLarge, num_classes, num_frames, optflow_inputs,
                         depth_inputs)


class TernausNetMultiBasic(_MultiFrame):
    """Drop-in for models/multiframe_model.py:207-238: TernausNet16(num_filters=64) base whose softmax
    probabilities (`base_model(x).exp()`, :227) feed the fusion head."""
    head = "probs"

    def __init__(self, num_classes, num_frames, pretrained=True, loadpath=None, optflow_inputs=False, depth_inputs=False):
        super().__init__(TernausNet16(num_classes=num_classes, num_filters=64, pretrained=False), MultiFrameNetBasic, num_classes,
                         num_frames, optflow_inputs, depth_inputs)


class TernausNetMultiLarge(_MultiFrame):
    """Drop-in for models/multiframe_model.py:240-271."""
    head = "probs"

    def __init__(self, num_classes, num_frames, pretrained=True, loadpath=None, optflow_inputs=False, depth_inputs=False):
        super().__init__(TernausNet16(num_classes=num_classes, num_filters=64, pretrained=False), MultiFrameNetLarge, num_classes,
                         num_frames, optflow_inputs, depth_inputs)
