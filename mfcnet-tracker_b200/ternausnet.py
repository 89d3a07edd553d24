"""TernausNet11 / TernausNet16 (VGG encoder + transposed-conv decoder) as drop-in nn.Modules on the B200 engine.

Mirrors the constructors, forward signatures and state_dict keys of models/ternausnet.py:45-149: ``encoder``
(the torchvision ``vgg11`` / ``vgg16`` ``features`` Sequential, conv layers at the torchvision indices),
``conv1``..``conv5`` (Sequentials that RE-REGISTER the encoder convs, so their parameters appear twice in the
state dict exactly as upstream), ``center``, ``dec5``..``dec2`` (``DecoderBlock.block`` = ConvRelu,
ConvTranspose2d(4, 2, 1), ReLU), ``dec1`` (ConvRelu) and ``final`` (1x1).  With ``num_classes > 1`` the forward
returns ``log_softmax`` over channels (models/ternausnet.py:145-148).

Arithmetic in libmfcnet_b200.so:
  * conv3x3 + bias + ReLU = one fused tcgen05 conv launch; the decoder's skip concats are extra conv sources;
  * MaxPool2d(2,2) = ``maxpool2`` on C8 planes;
  * ConvTranspose2d(k=4, s=2, p=1) + ReLU = FOUR 2x2 convolutions, one per output parity (py, px):
      out[2y+py, 2x+px] = sum_{a,b in {0,1}} in[y+py-1+a, x+px-1+b] * Wt[:, :, 3-py-2a, 3-px-2b]
    each a tensor-core conv whose epilogue stores with stride 2 into the [2H][2W] output (no zero-insertion,
    no scatter, every MAC of the transposed conv is done exactly once);
  * the head: 1x1 conv -> fp32 NCHW logits -> ``heatmap_head`` (log-softmax, and softmax for the MFCNet
    wrappers which feed ``base(x).exp()`` to the fusion head, models/multiframe_model.py:227,260).
"""
import torch
from torch import nn

from . import engine
from .engine import Act, Ext

# torchvision vgg `features` layouts: ints = conv output channels, "M" = MaxPool2d(2, 2); a ReLU follows every conv
_VGG = {11: [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
        16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]}
# encoder indices of the convs of conv1..conv5 (models/ternausnet.py:63-67, 114-118)
_STAGES = {11: [[0], [3], [6, 8], [11, 13], [16, 18]],
           16: [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]]}


def _vgg_features(depth):
    layers, cin = [], 3
    for v in _VGG[depth]:
        if v == "M":
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return nn.Sequential(*layers)


class _ConvRelu(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1)
        self.activation = nn.ReLU(inplace=True)


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cmid, cout):
        super().__init__()
        self.in_channels = cin
        self.block = nn.Sequential(_ConvRelu(cin, cmid), nn.ConvTranspose2d(cmid, cout, kernel_size=4, stride=2, padding=1),
                                   nn.ReLU(inplace=True))


class _TernausNet(nn.Module):
    depth = 16

    def __init__(self, num_classes=1, num_filters=32, pretrained=False):
        super().__init__()
        if pretrained:
            raise RuntimeError("pretrained VGG weights cannot be downloaded here; load a checkpoint with load_state_dict")
        self.num_classes = num_classes
        self.channels = 3
        self.out_dim = num_classes
        nf = num_filters
        self.pool = nn.MaxPool2d(2, 2)
        self.encoder = _vgg_features(self.depth)
        self.relu = nn.ReLU(inplace=True)
        for i, idxs in enumerate(_STAGES[self.depth]):
            mods = []
            for j in idxs:
                mods += [self.encoder[j], self.relu]
            setattr(self, "conv%d" % (i + 1), nn.Sequential(*mods))
        if self.depth == 11:
            self.center = _DecoderBlock(256 + nf * 8, nf * 8 * 2, nf * 8)
            self.dec5 = _DecoderBlock(512 + nf * 8, nf * 8 * 2, nf * 8)
            self.dec4 = _DecoderBlock(512 + nf * 8, nf * 8 * 2, nf * 4)
            self.dec3 = _DecoderBlock(256 + nf * 4, nf * 4 * 2, nf * 2)
        else:
            self.center = _DecoderBlock(512, nf * 8 * 2, nf * 8)
            self.dec5 = _DecoderBlock(512 + nf * 8, nf * 8 * 2, nf * 8)
            self.dec4 = _DecoderBlock(512 + nf * 8, nf * 8 * 2, nf * 8)
            self.dec3 = _DecoderBlock(256 + nf * 8, nf * 4 * 2, nf * 2)
        self.dec2 = _DecoderBlock(128 + nf * 2, nf * 2 * 2, nf)
        self.dec1 = _ConvRelu(64 + nf, nf)
        self.final = nn.Conv2d(nf, num_classes, kernel_size=1)
        self._plans = {}
        self._packer = None
        self._fingerprint = None
        self.dtype_name = None

    # ---- plan recording ----------------------------------------------------------------------------
    @staticmethod
    def _conv_relu(bld, name, conv, srcs):
        out, _, _, _ = bld.conv(name, srcs, conv.weight, 3, bias=conv.bias, pad=1, act=1)
        return out

    @staticmethod
    def _deconv_relu(bld, name, ct, x):
        """ConvTranspose2d(4,2,1) + ReLU as four parity 2x2 convs writing one [2H][2W] tensor."""
        wt = ct.weight.detach()  # [Cin][Cout][4][4]
        cout = wt.shape[1]
        y = bld.arena.alloc((x.B, (cout + 7) // 8, 2 * x.H, 2 * x.W, 8), bld.tdtype)
        for py in (0, 1):
            for px in (0, 1):
                ky = [3 - py, 1 - py]     # a = 0, 1  ->  ky = 3-py-2a
                kx = [3 - px, 1 - px]
                w2 = wt[:, :, ky][:, :, :, kx].permute(1, 0, 2, 3).contiguous()  # -> [Cout][Cin][2][2]
                bld.conv("%s.p%d%d" % (name, py, px), [x], w2, 2, bias=ct.bias, pad=1, act=1, y_c8=y, parity=(py, px))
        return Act(y, cout)

    def _decoder(self, bld, name, blk, srcs):
        h = self._conv_relu(bld, name + ".block.0.conv", blk.block[0].conv, srcs)
        return self._deconv_relu(bld, name + ".block.1", blk.block[1], h)

    def _pool(self, bld, x):
        return bld.prog.maxpool2(x, bld.arena.alloc((x.B, x.chunks, x.H // 2, x.W // 2, 8), bld.tdtype))

    def record(self, bld, x_act, logits_nchw=None, maps_c8=None, out_nchw=None):
        """Record the forward on `x_act`.  `out_nchw`: the module output (log-probs when num_classes > 1, else
        logits); `logits_nchw`: raw logits if wanted; `maps_c8`: C8 planes of exp(output) = softmax
        probabilities, what the MFCNet wrappers feed to the fusion head.  Returns (None, heatmap/conv struct)."""
        B, H, W = x_act.B, x_act.H, x_act.W
        feats, x = [], x_act
        for i, idxs in enumerate(_STAGES[self.depth]):
            if i > 0:
                x = self._pool(bld, x)
            for j in idxs:
                x = self._conv_relu(bld, "encoder.%d" % j, self.encoder[j], [x])
            feats.append(x)
        c1, c2, c3, c4, c5 = feats
        d = self._decoder(bld, "center", self.center, [self._pool(bld, c5)])
        d = self._decoder(bld, "dec5", self.dec5, [d, c5])
        d = self._decoder(bld, "dec4", self.dec4, [d, c4])
        d = self._decoder(bld, "dec3", self.dec3, [d, c3])
        d = self._decoder(bld, "dec2", self.dec2, [d, c2])
        d = self._conv_relu(bld, "dec1.conv", self.dec1.conv, [d, c1])
        N = self.num_classes
        if N > 1:
            logits = logits_nchw if logits_nchw is not None else bld.arena.alloc((B, N, H, W), torch.float32)
            bld.conv("final", [d], self.final.weight, 1, bias=self.final.bias, out_c8=False, out_nchw=logits)
            prob = None
            if maps_c8 is not None:
                prob = bld.arena.alloc((B, N, H, W), torch.float32)
            hm = bld.prog.heatmap(logits, logp=out_nchw, prob=prob)
            if maps_c8 is not None:
                ext = Ext(("_prob", id(prob)), prob)
                for c0 in range(0, N, 8):
                    bld.prog.gather([(ext, c) for c in range(c0, min(c0 + 8, N))], maps_c8[:, c0 // 8], B, H, W)
                bld.prog.bindings.pop(ext.key, None)   # internal buffer, not a caller input
            return None, hm
        target = out_nchw if out_nchw is not None else logits_nchw
        _, _, _, io = bld.conv("final", [d], self.final.weight, 1, bias=self.final.bias, out_c8=maps_c8 is not None, y_c8=maps_c8,
                               out_nchw=target)
        return None, io

    # ---- engine plumbing ------------------------------------------------------------------------------
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
                raise ValueError("TernausNet: H and W must be divisible by 32")
            arena = engine.Arena(device)
            bld = engine.Builder(device, dt, self._packer, arena)
            x_c8 = arena.alloc((B, 1, H, W, 8), bld.tdtype)
            dummy_in = torch.zeros((B, 3, H, W), dtype=torch.float32, device=device)
            dummy_out = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=device)
            ext = Ext("x", dummy_in)
            bld.prog.gather([(ext, c) for c in range(3)], x_c8[:, 0], B, H, W)
            _, last = self.record(bld, Act(x_c8, 3), out_nchw=dummy_out)
            bld.prog.finalize()
            self._plans[key] = (bld.prog, last, arena)
        return self._plans[key]

    def forward(self, x):
        engine.require_cuda(x, type(self).__name__ + ".forward")
        if self.training:
            raise RuntimeError("%s (B200 engine) implements inference only: call .eval()" % type(self).__name__)
        x = x.contiguous().float()
        B, Cc, H, W = x.shape
        if Cc != 3:
            raise ValueError("expected 3 input channels, got %d" % Cc)
        dt = self._check_weights(x.device)
        prog, last, _ = self._plan(B, H, W, x.device, dt)
        out = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=x.device)
        prog.call({"x": x}, lambda: setattr(last, "logp" if self.num_classes > 1 else "y_nchw", out.data_ptr()))
        engine.record_stream(x)
        return out


class TernausNet11(_TernausNet):
    """Drop-in for models/ternausnet.py:45."""
    depth = 11


class TernausNet16(_TernausNet):
    """Drop-in for models/ternausnet.py:98."""
    depth = 16
