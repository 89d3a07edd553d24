"""UnFlow optical-flow network (FlowNetC + 2 x FlowNetS around the correlation cost volume) on the B200 engine.

Drop-in for ``UnFlow`` (models/unflow_model.py:19-270): same constructor, ``forward(tensorFirst, tensorSecond)`` on RGB frames
in [0, 1], same state_dict keys (``moduleFlownets.{0,1,2}.module{One,Two,Thr,Redir,Combined,Fou,Fiv,Six}.<idx>`` and
``moduleUpconv.module*``; the parameter-less ZeroPad2d / LeakyReLU / ReplicationPad2d children keep the Sequential indices).
The modules below only HOLD parameters; the arithmetic runs in libmfcnet_b200.so:

  * ZeroPad2d([l, l+e, t, t+e]) + Conv2d(stride 2, padding 0) + LeakyReLU(0.1) (:85-130) = ONE fused tensor-core conv
    (`pad` = l, `pad_br` = e: the extra bottom / right padding is the zero fill of the TMA box; `act` = 2);
  * ConvTranspose2d(k4, s2, p1) (+ LeakyReLU) (:30-56) = four parity 2x2 convs writing one [2H][2W] tensor, over the
    multi-source concat [conv_k, next_k, up(flow_k+1)] (:66-75) that is never materialised;
  * the correlation (:113,163) = `mfc_correlation_fwd` on the fp32 conv3 features of both frames, its 441-channel volume
    re-laid as C8 planes by `mfc_nchw_to_c8` and consumed together with moduleRedir's 32 channels by moduleCombined;
  * `backward()` warp + |first - warp| (:6-17,224-226) = `mfc_unflow_warp`; moduleUpscale x 2 (x 20) = `mfc_unflow_upscale`;
  * BGR flip and mean subtraction (:253-262) = `mfc_unflow_preprocess`.
Inference only (the reference uses UnFlow frozen, as a flow provider).
"""
import torch
from torch import nn

from . import abi, engine
from .engine import Act, Ext


def _padconv(cin, cout, k, pad):
    return nn.Sequential(nn.ZeroPad2d(pad), nn.Conv2d(cin, cout, k, stride=2, padding=0), nn.LeakyReLU(0.1))


def _pad2conv(cin, cout):
    return nn.Sequential(nn.ZeroPad2d([0, 2, 0, 2]), nn.Conv2d(cin, cout, 3, stride=2, padding=0), nn.LeakyReLU(0.1),
                         nn.Conv2d(cout, cout, 3, stride=1, padding=1), nn.LeakyReLU(0.1))


class _Upconv(nn.Module):
    def __init__(self):
        super().__init__()
        ct = lambda ci, co: nn.ConvTranspose2d(ci, co, 4, stride=2, padding=1)
        nxt = lambda ci, co: nn.Sequential(ct(ci, co), nn.LeakyReLU(0.1))
        self.moduleSixOut = nn.Conv2d(1024, 2, 3, padding=1)
        self.moduleSixUp = ct(2, 2)
        self.moduleFivNext = nxt(1024, 512)
        self.moduleFivOut = nn.Conv2d(1026, 2, 3, padding=1)
        self.moduleFivUp = ct(2, 2)
        self.moduleFouNext = nxt(1026, 256)
        self.moduleFouOut = nn.Conv2d(770, 2, 3, padding=1)
        self.moduleFouUp = ct(2, 2)
        self.moduleThrNext = nxt(770, 128)
        self.moduleThrOut = nn.Conv2d(386, 2, 3, padding=1)
        self.moduleThrUp = ct(2, 2)
        self.moduleTwoNext = nxt(386, 64)
        self.moduleTwoOut = nn.Conv2d(194, 2, 3, padding=1)
        self.moduleUpscale = nn.Sequential(nn.ConvTranspose2d(2, 2, 3, stride=2, padding=1, bias=False), nn.ReplicationPad2d([0, 1, 0, 1]))


class _Complex(nn.Module):
    def __init__(self):
        super().__init__()
        self.moduleOne = _padconv(3, 64, 7, [2, 4, 2, 4])
        self.moduleTwo = _padconv(64, 128, 5, [1, 3, 1, 3])
        self.moduleThr = _padconv(128, 256, 5, [1, 3, 1, 3])
        self.moduleRedir = nn.Sequential(nn.Conv2d(256, 32, 1), nn.LeakyReLU(0.1))
        self.moduleCorrelation = nn.Identity()
        self.moduleCombined = nn.Sequential(nn.Conv2d(473, 256, 3, padding=1), nn.LeakyReLU(0.1))
        self.moduleFou = _pad2conv(256, 512)
        self.moduleFiv = _pad2conv(512, 512)
        self.moduleSix = _pad2conv(512, 1024)
        self.moduleUpconv = _Upconv()


class _Simple(nn.Module):
    def __init__(self):
        super().__init__()
        self.moduleOne = _padconv(14, 64, 7, [2, 4, 2, 4])
        self.moduleTwo = _padconv(64, 128, 5, [1, 3, 1, 3])
        self.moduleThr = nn.Sequential(nn.ZeroPad2d([1, 3, 1, 3]), nn.Conv2d(128, 256, 5, stride=2, padding=0), nn.LeakyReLU(0.1),
                                       nn.Conv2d(256, 256, 3, stride=1, padding=1), nn.LeakyReLU(0.1))
        self.moduleFou = _pad2conv(256, 512)
        self.moduleFiv = _pad2conv(512, 512)
        self.moduleSix = _pad2conv(512, 1024)
        self.moduleUpconv = _Upconv()


LEAKY = 2


def _conv(bld, name, c, srcs, **kw):
    out, _, _, _ = bld.conv(name, srcs, c.weight, c.kernel_size[0], bias=c.bias, **kw)
    return out


def _deconv(bld, name, ct, srcs, act):
    """ConvTranspose2d(4, 2, 1) over the concat `srcs` as four parity 2x2 convs writing one [2H][2W] tensor."""
    wt = ct.weight.detach()                       # [Cin][Cout][4][4]
    cout = wt.shape[1]
    x0 = srcs[0]
    y = bld.arena.alloc((x0.B, (cout + 7) // 8, 2 * x0.H, 2 * x0.W, 8), bld.tdtype)
    for py in (0, 1):
        for px in (0, 1):
            ky, kx = [3 - py, 1 - py], [3 - px, 1 - px]
            w2 = wt[:, :, ky][:, :, :, kx].permute(1, 0, 2, 3).contiguous()
            bld.conv("%s.p%d%d" % (name, py, px), srcs, w2, 2, bias=ct.bias, pad=1, act=act, y_c8=y, parity=(py, px))
    return Act(y, cout)


def _tail(bld, p, net, o):
    for src, dst, seq, name in (("conv3", "conv4", net.moduleFou, "moduleFou"), ("conv4", "conv5", net.moduleFiv, "moduleFiv"),
                                ("conv5", "conv6", net.moduleSix, "moduleSix")):
        x = _conv(bld, "%s%s.1" % (p, name), seq[1], [o[src]], stride=2, pad=0, pad_br=2, act=LEAKY)
        o[dst] = _conv(bld, "%s%s.3" % (p, name), seq[3], [x], pad=1, act=LEAKY)
    return o


def _upconv(bld, p, up, o, flow2_nchw):
    """Upconv.forward (:64-78) up to flow2 (fp32 NCHW into `flow2_nchw`); the two moduleUpscale steps follow outside."""
    x = [o["conv6"]]
    f = _conv(bld, p + "moduleSixOut", up.moduleSixOut, x, pad=1)
    for lvl, skip, nxt, out_c, up_c in (("Fiv", "conv5", up.moduleFivNext, up.moduleFivOut, up.moduleSixUp),
                                        ("Fou", "conv4", up.moduleFouNext, up.moduleFouOut, up.moduleFivUp),
                                        ("Thr", "conv3", up.moduleThrNext, up.moduleThrOut, up.moduleFouUp),
                                        ("Two", "conv2", up.moduleTwoNext, up.moduleTwoOut, up.moduleThrUp)):
        x = [o[skip], _deconv(bld, "%smodule%sNext.0" % (p, lvl), nxt[0], x, LEAKY), _deconv(bld, "%sup_into_%s" % (p, lvl), up_c, [f], 0)]
        if lvl == "Two":
            bld.conv(p + "moduleTwoOut", x, out_c.weight, 3, bias=out_c.bias, pad=1, out_c8=False, out_nchw=flow2_nchw)
        else:
            f = _conv(bld, "%smodule%sOut" % (p, lvl), out_c, x, pad=1)


class UnFlow(nn.Module):
    """Drop-in for models/unflow_model.py:19 ``UnFlow``; forward(tensorFirst, tensorSecond) -> flow (B, 2, H, W) fp32."""

    def __init__(self):
        super().__init__()
        self.moduleFlownets = nn.ModuleList([_Complex(), _Simple(), _Simple()])
        self._plans, self._packer, self._fingerprint, self.dtype_name = {}, None, None, None

    def _check_weights(self, device):
        dt = self.dtype_name or engine.default_dtype()
        fp = (engine.params_fingerprint(self), str(device), dt)
        if fp != self._fingerprint:
            self._plans = {}
            self._packer = engine.WeightPacker(device, dt)
            self._fingerprint = fp
        return dt

    def _build(self, B, H, W, dev, dt):
        if H % 64 or W % 64:
            raise ValueError("UnFlow: H and W must be multiples of 64 (six stride-2 stages)")
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        P = {"first": f32(B, 3, H, W), "second": f32(B, 3, H, W), "flows": [f32(B, 2, H, W) for _ in range(3)],
             "flow2": [f32(B, 2, H // 4, W // 4) for _ in range(3)], "half": f32(B, 2, H // 2, W // 2),
             "warp": f32(B, 3, H, W), "diff": f32(B, 3, H, W), "feat1": f32(B, 256, H // 8, W // 8), "feat2": f32(B, 256, H // 8, W // 8),
             "cv": f32(B, 441, H // 8, W // 8)}
        arena = engine.Arena(dev)
        tdt = engine._DTYPES[dt][0]
        # ---- Complex, part 1: both frames through moduleOne / Two / Thr, moduleRedir on the first
        net = self.moduleFlownets[0]
        p = "moduleFlownets.0."
        b1 = engine.Builder(dev, dt, self._packer, arena)
        o = {}
        for which, tensor in (("first", P["first"]), ("second", P["second"])):
            x = b1.gather_channels([Ext(which, tensor)], B, H, W)
            c1 = _conv(b1, p + "moduleOne.1", net.moduleOne[1], [x], stride=2, pad=2, pad_br=2, act=LEAKY)
            c2 = _conv(b1, p + "moduleTwo.1", net.moduleTwo[1], [c1], stride=2, pad=1, pad_br=2, act=LEAKY)
            c3, _, _, _ = b1.conv(p + "moduleThr.1", [c2], net.moduleThr[1].weight, 5, bias=net.moduleThr[1].bias, stride=2, pad=1, pad_br=2,
                                  act=LEAKY, out_nchw=P["feat1"] if which == "first" else P["feat2"], out_c8=which == "first")
            if which == "first":
                o["conv1"], o["conv2"], o["conv3_pre"] = c1, c2, c3
        redir = _conv(b1, p + "moduleRedir.0", net.moduleRedir[0], [o["conv3_pre"]], act=LEAKY)
        b1.prog.finalize()
        # ---- Complex, part 2: moduleCombined over [redir, cost volume] ... flow2
        cv_c8 = torch.empty((B, (441 + 7) // 8, H // 8, W // 8, 8), dtype=tdt, device=dev)
        b2 = engine.Builder(dev, dt, self._packer, arena)
        o["conv3"] = _conv(b2, p + "moduleCombined.0", net.moduleCombined[0], [redir, Act(cv_c8, 441)], pad=1, act=LEAKY)
        _upconv(b2, p + "moduleUpconv.", net.moduleUpconv, _tail(b2, p, net, o), P["flow2"][0])
        b2.prog.finalize()
        progs = [b1.prog, b2.prog]
        # ---- Simple x 2: [first, second, flow, warp, |first - warp|] (14 channels) -> flow2
        for n in (1, 2):
            net = self.moduleFlownets[n]
            p = "moduleFlownets.%d." % n
            b = engine.Builder(dev, dt, self._packer, arena)
            x = b.gather_channels([Ext("first", P["first"]), Ext("second", P["second"]), Ext("flow", P["flows"][n - 1]),
                                   Ext("warp", P["warp"]), Ext("diff", P["diff"])], B, H, W)
            o = {}
            o["conv1"] = _conv(b, p + "moduleOne.1", net.moduleOne[1], [x], stride=2, pad=2, pad_br=2, act=LEAKY)
            o["conv2"] = _conv(b, p + "moduleTwo.1", net.moduleTwo[1], [o["conv1"]], stride=2, pad=1, pad_br=2, act=LEAKY)
            t = _conv(b, p + "moduleThr.1", net.moduleThr[1], [o["conv2"]], stride=2, pad=1, pad_br=2, act=LEAKY)
            o["conv3"] = _conv(b, p + "moduleThr.3", net.moduleThr[3], [t], pad=1, act=LEAKY)
            _upconv(b, p + "moduleUpconv.", net.moduleUpconv, _tail(b, p, net, o), P["flow2"][n])
            b.prog.finalize()
            progs.append(b.prog)
        P.update(progs=progs, cv_c8=cv_c8, arena=arena,
                 upw=[self.moduleFlownets[n].moduleUpconv.moduleUpscale[0].weight.detach().float().contiguous() for n in range(3)])
        return P

    def forward(self, tensorFirst, tensorSecond):
        engine.require_cuda(tensorFirst, "UnFlow.forward")
        if self.training:
            raise RuntimeError("UnFlow (B200 engine) implements inference only: call .eval()")
        a, b = tensorFirst.contiguous().float(), tensorSecond.contiguous().float()
        B, Cc, H, W = a.shape
        if Cc != 3 or b.shape != a.shape:
            raise ValueError("UnFlow: two (B, 3, H, W) frames expected")
        dev = a.device
        dt = self._check_weights(dev)
        key = (B, H, W)
        if key not in self._plans:
            self._plans[key] = self._build(B, H, W, dev, dt)
        P = self._plans[key]
        lib = abi.load()
        cd = engine._DTYPES[dt][1]
        with engine.device_guard(dev):
            st = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
            abi.check(lib.mfc_unflow_preprocess(a.data_ptr(), P["first"].data_ptr(), B, H, W, st))
            abi.check(lib.mfc_unflow_preprocess(b.data_ptr(), P["second"].data_ptr(), B, H, W, st))
            P["progs"][0].run()
            h8, w8 = H // 8, W // 8
            abi.check(lib.mfc_correlation_fwd(P["feat1"].data_ptr(), P["feat2"].data_ptr(), P["cv"].data_ptr(), B, 256, h8, w8, 20, 2, 0, st))
            abi.check(lib.mfc_nchw_to_c8(P["cv"].data_ptr(), P["cv_c8"].data_ptr(), P["cv_c8"].stride(0) * P["cv_c8"].element_size(), B, 441,
                                         h8, w8, cd, st))
            for n in range(3):
                if n > 0:
                    abi.check(lib.mfc_unflow_warp(P["second"].data_ptr(), P["flows"][n - 1].data_ptr(), P["first"].data_ptr(),
                                                  P["warp"].data_ptr(), P["diff"].data_ptr(), B, 3, H, W, st))
                P["progs"][n + 1].run()
                up = P["upw"][n]
                abi.check(lib.mfc_unflow_upscale(P["flow2"][n].data_ptr(), up.data_ptr(), P["half"].data_ptr(), B, H // 4, W // 4, 1.0, st))
                abi.check(lib.mfc_unflow_upscale(P["half"].data_ptr(), up.data_ptr(), P["flows"][n].data_ptr(), B, H // 2, W // 2, 20.0, st))
            out = P["flows"][2].clone()
        engine.record_stream(a)
        engine.record_stream(b)
        return out
