"""fp32 functional restatement of the reference modules on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function takes a plain
``state_dict`` (the reference's own key names) and tensors, and recomputes
what the reference ``nn.Module`` computes, in fp32 on whatever device the
tensors live on (CPU for the oracle / CPU baseline).  It is written against
the state-dict *keys*, not as a module tree, so it doubles as a check that a
checkpoint alone determines the result.

Pinned by tests/test_oracle_vs_reference.py (live reference, authoring
container) and tests/test_oracle_golden.py (committed reference outputs).
"""
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# ResUnet_VB  (models/resunet.py:97-180)
# ----------------------------------------------------------------------------
def weight_standardize(w, eps=1e-5):
    """models/resunet.py:56-62: per-out-channel (w-mean)*rsqrt(biased var+eps);
    eps is 1e-5 for fp32 inputs (1e-3 otherwise)."""
    flat = w.reshape(w.shape[0], -1)
    mu = flat.mean(dim=1).reshape(-1, 1, 1, 1)
    var = flat.var(dim=1, unbiased=False).reshape(-1, 1, 1, 1)
    return (w - mu) * torch.rsqrt(var + eps)


def _ws_gn_silu(sd, p, x, groups):
    """`Block.forward` (models/resunet.py:75-80): WS-conv3x3 -> GroupNorm -> SiLU."""
    y = F.conv2d(x, weight_standardize(sd[p + "proj.weight"]), sd[p + "proj.bias"], padding=1)
    y = F.group_norm(y, groups, sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-5)
    return F.silu(y)


def _resnet_block(sd, p, x, groups):
    """`ResnetBlock.forward` (models/resunet.py:90-95)."""
    h = _ws_gn_silu(sd, p + "block1.", x, groups)
    h = _ws_gn_silu(sd, p + "block2.", h, groups)
    if p + "res_conv.weight" in sd:
        x = F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def _pixel_unshuffle2(x):
    """einops 'b c (h p1) (w p2) -> b (c p1 p2) h w', p1=p2=2 (models/resunet.py:47)."""
    b, c, h, w = x.shape
    x = x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(b, c * 4, h // 2, w // 2)


def resunet_forward(sd, x, groups=8, prefix=""):
    """`ResUnet_VB.forward` (models/resunet.py:153-180) -> raw logits."""
    g = lambda k: sd[prefix + k]
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    n_levels = 1 + max(int(k.split(".")[1]) for k in sub if k.startswith("downs."))
    x = F.conv2d(x, sub["init_conv.weight"], sub["init_conv.bias"], padding=3)
    stem = x
    skips = []
    for i in range(n_levels):
        x = _resnet_block(sub, f"downs.{i}.0.", x, groups)
        skips.append(x)
        if f"downs.{i}.1.1.weight" in sub:      # Downsample = unshuffle + 1x1 (models/resunet.py:45-49)
            x = F.conv2d(_pixel_unshuffle2(x), sub[f"downs.{i}.1.1.weight"], sub[f"downs.{i}.1.1.bias"])
        else:                                    # last level: plain 3x3 (models/resunet.py:134)
            x = F.conv2d(x, sub[f"downs.{i}.1.weight"], sub[f"downs.{i}.1.bias"], padding=1)
    x = _resnet_block(sub, "mid_block.", x, groups)
    for i in range(n_levels):
        x = torch.cat((x, skips.pop()), dim=1)
        x = _resnet_block(sub, f"ups.{i}.0.", x, groups)
        if f"ups.{i}.1.1.weight" in sub:        # Upsample = nearest x2 + 3x3 (models/resunet.py:39-43)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sub[f"ups.{i}.1.1.weight"], sub[f"ups.{i}.1.1.bias"], padding=1)
        else:
            x = F.conv2d(x, sub[f"ups.{i}.1.weight"], sub[f"ups.{i}.1.bias"], padding=1)
    x = torch.cat((x, stem), dim=1)
    x = _resnet_block(sub, "final_res_block.", x, groups)
    return F.conv2d(x, sub["output_layer.weight"], sub["output_layer.bias"])


# ----------------------------------------------------------------------------
# MultiFrameNet{Basic,Large}  (models/multiframe_model.py:51-205)
# ----------------------------------------------------------------------------
def _bn_eval(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, eps=eps)


def fusion_stack(sd, x, prefix="multiframe_net."):
    """The 4-conv Sequential shared by Basic and Large
    (models/multiframe_model.py:62-73 and :191-202): 11x11(p5) BN ReLU 3x3 BN ReLU 3x3 BN ReLU 1x1."""
    p = prefix
    x = F.relu(_bn_eval(sd, p + "1.", F.conv2d(x, sd[p + "0.weight"], padding=5)))
    x = F.relu(_bn_eval(sd, p + "4.", F.conv2d(x, sd[p + "3.weight"], padding=1)))
    x = F.relu(_bn_eval(sd, p + "7.", F.conv2d(x, sd[p + "6.weight"], padding=1)))
    return F.conv2d(x, sd[p + "9.weight"])


def flow_warp_single(m, flow, grid):
    """`_warp_single_map` (models/multiframe_model.py:141-170): the stored 576x720
    normalised grid is *cropped* to HxW, the flow is normalised by (W-1)/2,(H-1)/2,
    then grid_sample(bilinear, zeros, align_corners=True)."""
    H, W = m.shape[-2:]
    g = grid[:, :, :H, :W]
    fx = flow[:, 0] / ((W - 1) / 2.0)
    fy = flow[:, 1] / ((H - 1) / 2.0)
    new = (g + torch.stack((fx, fy), dim=1)).permute(0, 2, 3, 1)
    return F.grid_sample(m, new, mode="bilinear", padding_mode="zeros", align_corners=True)


def warp_seg_and_depth(x, grid, N, K, with_depth):
    """`warp_segmentation_and_depth` (models/multiframe_model.py:89-139)."""
    seg = x[:, : N * K]
    flo = x[:, N * K : N * K + 2 * K - 2]
    dep = x[:, N * K + 2 * K - 2 :] if with_depth else None
    segs = [seg[:, :N]]
    deps = [dep[:, 0:1]] if with_depth else []
    for i in range(1, K):
        f = flo[:, 2 * (i - 1) : 2 * i]
        for j in range(N):
            segs.append(flow_warp_single(seg[:, i * N + j : i * N + j + 1], f, grid))
        if with_depth:
            deps.append(flow_warp_single(dep[:, i : i + 1], f, grid))
    return torch.cat(segs + deps, dim=1)


def fusion_large_forward(sd, x, prefix=""):
    """`MultiFrameNetLarge.forward` (models/multiframe_model.py:204-205)."""
    return fusion_stack(sd, x, prefix + "multiframe_net.")


def fusion_basic_forward(sd, x, N, K, with_optflow, with_depth, prefix=""):
    """`MultiFrameNetBasic.forward` (models/multiframe_model.py:84-87)."""
    if with_optflow:
        x = warp_seg_and_depth(x, sd[prefix + "grid"], N, K, with_depth)
    return fusion_stack(sd, x, prefix + "multiframe_net.")


def mfcnet_forward(sd, frames, optflow, depth, *, base, variant, N, head="logits"):
    """`XMulti{Basic,Large}.forward` (models/multiframe_model.py:424-438 for HRNet;
    :224-239 for Ternaus which feeds exp(log-probs)).  `base` is a callable
    (sub_state_dict, frame) -> per-frame class maps; `head` selects the family's
    convention: 'logits' (HRNet/ResUNet) or 'exp' (Ternaus)."""
    K = len(frames)
    bsd = {k[len("base_model."):]: v for k, v in sd.items() if k.startswith("base_model.")}
    fsd = {k[len("multiframe_net."):]: v for k, v in sd.items() if k.startswith("multiframe_net.")}
    maps = []
    for f in frames:
        y = base(bsd, f)
        maps.append(y.exp() if head == "exp" else y)
    if optflow is not None:
        maps += list(optflow)
    if depth is not None:
        maps += list(depth)
    x = torch.cat(maps, dim=1)
    if variant == "large":
        return fusion_large_forward(fsd, x)
    return fusion_basic_forward(fsd, x, N, K, optflow is not None, depth is not None)


# ----------------------------------------------------------------------------
# heatmap head  (call sites: src/engine.py:65,141;
# scripts/test_multiframe_segmentation_on_videos_v3.py:281,289)
# ----------------------------------------------------------------------------
def heatmap_head(logits):
    """log-softmax over classes, its exp (the 'probabilities' the video script
    uses), and numpy-first-max argmax of the probabilities."""
    logp = F.log_softmax(logits, dim=1)
    prob = torch.exp(logp)
    return logp, prob, prob.cpu().numpy().argmax(axis=1)


# ----------------------------------------------------------------------------
# correlation  (models/unflow_correlation.py:10-105, 282-337) -- fast torch form;
# the summation-order-exact restatement is oracle/corr_oracle.c
# ----------------------------------------------------------------------------
def correlation(first, second, max_disp=20, stride2=2):
    B, C, H, W = first.shape
    D = 2 * (max_disp // stride2) + 1
    pad = F.pad(second, (max_disp,) * 4)
    out = first.new_zeros(B, D * D, H, W)
    for iy in range(D):
        for ix in range(D):
            dy = (iy - D // 2) * stride2 + max_disp
            dx = (ix - D // 2) * stride2 + max_disp
            out[:, iy * D + ix] = (first * pad[:, :, dy : dy + H, dx : dx + W]).sum(1) / C
    return out


# ----------------------------------------------------------------------------
# HighResolutionNet  (models/hrnet.py:271-476)
# ----------------------------------------------------------------------------
def _bn_eval(sd, p, x, eps=1e-5):
    """Eval-mode BatchNorm2d / SyncBatchNorm (models/hrnet.py:31): running statistics."""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, eps)


def _conv_bn(sd, p_conv, p_bn, x, stride=1, relu=False):
    w = sd[p_conv + "weight"]
    y = F.conv2d(x, w, sd.get(p_conv + "bias"), stride=stride, padding=w.shape[-1] // 2)
    y = _bn_eval(sd, p_bn, y)
    return F.relu(y) if relu else y


def _hr_block(sd, p, x):
    """BasicBlock (models/hrnet.py:58-74) or Bottleneck (:97-115), told apart by the presence of conv3."""
    res = x
    if p + "downsample.0.weight" in sd:
        res = _conv_bn(sd, p + "downsample.0.", p + "downsample.1.", x)
    h = _conv_bn(sd, p + "conv1.", p + "bn1.", x, relu=True)
    if p + "conv3.weight" in sd:
        h = _conv_bn(sd, p + "conv2.", p + "bn2.", h, relu=True)
        h = _conv_bn(sd, p + "conv3.", p + "bn3.", h)
    else:
        h = _conv_bn(sd, p + "conv2.", p + "bn2.", h)
    return F.relu(h + res)


def _count(sd, prefix):
    """number of consecutive integer children `prefix<i>.` present in the state dict"""
    n = 0
    while any(k.startswith("%s%d." % (prefix, n)) for k in sd):
        n += 1
    return n


def _hr_chain(sd, p, x, relu_last):
    """Sequential of stride-2 (conv3x3, bn[, relu]) units `p<k>.{0,1}` (fuse down-paths :213-231, transitions :371-387)."""
    n = _count(sd, p)
    for k in range(n):
        x = _conv_bn(sd, "%s%d.0." % (p, k), "%s%d.1." % (p, k), x, stride=2, relu=relu_last or k < n - 1)
    return x


def _hr_module(sd, p, xs):
    """HighResolutionModule.forward (models/hrnet.py:237-260)."""
    nb = len(xs)
    ys = []
    for i in range(nb):
        x = xs[i]
        for k in range(_count(sd, "%sbranches.%d." % (p, i))):
            x = _hr_block(sd, "%sbranches.%d.%d." % (p, i, k), x)
        ys.append(x)
    out = []
    for i in range(nb):
        acc = None
        for j in range(nb):
            q = "%sfuse_layers.%d.%d." % (p, i, j)
            if j == i:
                t = ys[j]
            elif j > i:
                t = _conv_bn(sd, q + "0.", q + "1.", ys[j])
                t = F.interpolate(t, size=ys[i].shape[-2:], mode="bilinear", align_corners=False)
            else:
                t = _hr_chain(sd, q, ys[j], relu_last=False)
            acc = t if acc is None else acc + t
        out.append(F.relu(acc))
    return out


def hrnet_forward(sd, x, prefix=""):
    """HighResolutionNet.forward (models/hrnet.py:424-476) -> raw logits at the input resolution."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = _conv_bn(sd, "conv1.", "bn1.", x, stride=2, relu=True)
    x = _conv_bn(sd, "conv2.", "bn2.", x, stride=2, relu=True)
    for k in range(_count(sd, "layer1.")):
        x = _hr_block(sd, "layer1.%d." % k, x)
    xs = [_conv_bn(sd, "transition1.0.0.", "transition1.0.1.", x, relu=True), _hr_chain(sd, "transition1.1.", x, relu_last=True)]
    for m in range(_count(sd, "stage2.")):
        xs = _hr_module(sd, "stage2.%d." % m, xs)
    xs = xs + [_hr_chain(sd, "transition2.2.", xs[-1], relu_last=True)]
    for m in range(_count(sd, "stage3.")):
        xs = _hr_module(sd, "stage3.%d." % m, xs)
    xs = xs + [_hr_chain(sd, "transition3.3.", xs[-1], relu_last=True)]
    for m in range(_count(sd, "stage4.")):
        xs = _hr_module(sd, "stage4.%d." % m, xs)
    size = xs[0].shape[-2:]
    cat = torch.cat([xs[0]] + [F.interpolate(t, size=size, mode="bilinear", align_corners=False) for t in xs[1:]], 1)
    y = F.relu(_bn_eval(sd, "last_layer.1.", F.conv2d(cat, sd["last_layer.0.weight"], sd["last_layer.0.bias"])))
    y = F.conv2d(y, sd["last_layer.3.weight"], sd["last_layer.3.bias"])
    return F.interpolate(y, size=(size[0] * 4, size[1] * 4), mode="bilinear", align_corners=False)


# ----------------------------------------------------------------------------
# TernausNet11 / TernausNet16  (models/ternausnet.py:45-149)
# ----------------------------------------------------------------------------
_TERNAUS_STAGES = {11: [[0], [3], [6, 8], [11, 13], [16, 18]],
                   16: [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]]}


def _decoder_block(sd, p, x):
    """DecoderBlock (models/ternausnet.py:25-43): ConvRelu -> ConvTranspose2d(4, 2, 1) -> ReLU."""
    x = F.relu(F.conv2d(x, sd[p + "block.0.conv.weight"], sd[p + "block.0.conv.bias"], padding=1))
    return F.relu(F.conv_transpose2d(x, sd[p + "block.1.weight"], sd[p + "block.1.bias"], stride=2, padding=1))


def ternaus_forward(sd, x, prefix=""):
    """TernausNet{11,16}.forward (models/ternausnet.py:77-95, 127-148): log_softmax output when num_classes > 1.
    The depth is read off the state dict (vgg16 has encoder.28)."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    depth = 16 if "encoder.28.weight" in sd else 11
    feats = []
    for i, idxs in enumerate(_TERNAUS_STAGES[depth]):
        if i > 0:
            x = F.max_pool2d(x, 2, 2)
        for j in idxs:
            x = F.relu(F.conv2d(x, sd["encoder.%d.weight" % j], sd["encoder.%d.bias" % j], padding=1))
        feats.append(x)
    c1, c2, c3, c4, c5 = feats
    d = _decoder_block(sd, "center.", F.max_pool2d(c5, 2, 2))
    d = _decoder_block(sd, "dec5.", torch.cat([d, c5], 1))
    d = _decoder_block(sd, "dec4.", torch.cat([d, c4], 1))
    d = _decoder_block(sd, "dec3.", torch.cat([d, c3], 1))
    d = _decoder_block(sd, "dec2.", torch.cat([d, c2], 1))
    d = F.relu(F.conv2d(torch.cat([d, c1], 1), sd["dec1.conv.weight"], sd["dec1.conv.bias"], padding=1))
    y = F.conv2d(d, sd["final.weight"], sd["final.bias"])
    return F.log_softmax(y, dim=1) if y.shape[1] > 1 else y


def ternaus_probs(sd, x, prefix=""):
    """What TernausNetMulti* feed to the fusion head: base_model(x).exp() (models/multiframe_model.py:227,260)."""
    return ternaus_forward(sd, x, prefix).exp()


# ----------------------------------------------------------------------------
# Training loss, forward  (src/loss.py:6-63, src/engine.py:65-66)
# ----------------------------------------------------------------------------
def segmentation_loss(output, targets, class_weights=None, w_nll=0.7, w_jaccard=0.3):
    """total = w_nll * NLLLoss(weight)(log_softmax(output), t) + w_jaccard * LossSoftJaccard (src/loss.py:31-63).
    Returns (total, nll, jaccard) as python floats, accumulated in float64."""
    logp = F.log_softmax(output.double(), dim=1)
    N = output.shape[1]
    w = torch.ones(N, dtype=torch.float64) if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float64)
    nll = F.nll_loss(logp, targets, weight=w)
    jac = 0.0
    for c in range(1, N):
        tgt = (targets == c).double()
        p = logp[:, c].exp()
        inter = (p * tgt).sum()
        union = p.sum() + tgt.sum() - inter
        jac = jac - torch.log((inter + 1e-15) / (union + 1e-15))
    jac = jac / N
    return float(w_nll * nll + w_jaccard * jac), float(nll), float(jac)


# ----------------------------------------------------------------------------
# UnFlow network around the correlation  (models/unflow_model.py:6-270)
# ----------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, 0.1)


def unflow_backward_warp(inp, flow):
    """`backward()` (models/unflow_model.py:6-17): grid_sample(bilinear, border) at linspace(-1,1) + flow/((size-1)/2),
    align_corners left at torch's default (False)."""
    B, _, H, W = flow.shape
    gx = torch.linspace(-1.0, 1.0, W, dtype=flow.dtype, device=flow.device).view(1, 1, 1, W).expand(B, -1, H, -1)
    gy = torch.linspace(-1.0, 1.0, H, dtype=flow.dtype, device=flow.device).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([gx, gy], 1)
    fl = torch.cat([flow[:, 0:1] / ((inp.shape[3] - 1.0) / 2.0), flow[:, 1:2] / ((inp.shape[2] - 1.0) / 2.0)], 1)
    return F.grid_sample(inp, (grid + fl).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=False)


def _unflow_upconv(sd, p, o):
    """Upconv.forward (models/unflow_model.py:64-78)."""
    def conv(name, x):
        return F.conv2d(x, sd[p + name + ".weight"], sd[p + name + ".bias"], padding=1)

    def up(name, x):
        return F.conv_transpose2d(x, sd[p + name + ".weight"], sd[p + name + ".bias"], stride=2, padding=1)

    def nxt(name, x):
        return _lrelu(F.conv_transpose2d(x, sd[p + name + ".0.weight"], sd[p + name + ".0.bias"], stride=2, padding=1))

    x = o["conv6"]
    f6 = conv("moduleSixOut", x)
    x = torch.cat([o["conv5"], nxt("moduleFivNext", x), up("moduleSixUp", f6)], 1)
    f5 = conv("moduleFivOut", x)
    x = torch.cat([o["conv4"], nxt("moduleFouNext", x), up("moduleFivUp", f5)], 1)
    f4 = conv("moduleFouOut", x)
    x = torch.cat([o["conv3"], nxt("moduleThrNext", x), up("moduleFouUp", f4)], 1)
    f3 = conv("moduleThrOut", x)
    x = torch.cat([o["conv2"], nxt("moduleTwoNext", x), up("moduleThrUp", f3)], 1)
    f2 = conv("moduleTwoOut", x)

    def upscale(t):
        t = F.conv_transpose2d(t, sd[p + "moduleUpscale.0.weight"], None, stride=2, padding=1)
        return F.pad(t, [0, 1, 0, 1], mode="replicate")
    return upscale(upscale(f2)) * 20.0


def _unflow_padconv(sd, name, x, pad, k_stride=2):
    """nn.Sequential(ZeroPad2d(pad), Conv2d(stride 2, padding 0), LeakyReLU(0.1)) -- keys `<name>.1.*`."""
    return _lrelu(F.conv2d(F.pad(x, pad), sd[name + ".1.weight"], sd[name + ".1.bias"], stride=k_stride))


def _unflow_tail(sd, p, o):
    """moduleFou / moduleFiv / moduleSix (shared by Complex and Simple): pad [0,2,0,2] + conv3 s2 + lrelu + conv3 + lrelu."""
    for src, dst, name in (("conv3", "conv4", "moduleFou"), ("conv4", "conv5", "moduleFiv"), ("conv5", "conv6", "moduleSix")):
        x = _unflow_padconv(sd, p + name, o[src], [0, 2, 0, 2])
        o[dst] = _lrelu(F.conv2d(x, sd[p + name + ".3.weight"], sd[p + name + ".3.bias"], padding=1))
    return o


def unflow_forward(sd, first, second, corr=None):
    """UnFlow.forward (models/unflow_model.py:253-270) = Complex (FlowNetC, :81-175) then two Simple nets (FlowNetS, :177-245).
    `corr(a, b)` = the cost volume (default: the torch restatement `correlation`)."""
    corr = corr or correlation
    mean = torch.tensor([104.920005 / 255.0, 110.175300 / 255.0, 114.785955 / 255.0], dtype=first.dtype, device=first.device)
    first = first[:, [2, 1, 0]] - mean.view(1, 3, 1, 1)
    second = second[:, [2, 1, 0]] - mean.view(1, 3, 1, 1)
    # ---- Complex
    p = "moduleFlownets.0."

    def enc(x):
        c1 = _unflow_padconv(sd, p + "moduleOne", x, [2, 4, 2, 4])
        c2 = _unflow_padconv(sd, p + "moduleTwo", c1, [1, 3, 1, 3])
        c3 = _unflow_padconv(sd, p + "moduleThr", c2, [1, 3, 1, 3])
        return c1, c2, c3
    o = {}
    o["conv1"], o["conv2"], o["conv3"] = enc(first)
    redir = _lrelu(F.conv2d(o["conv3"], sd[p + "moduleRedir.0.weight"], sd[p + "moduleRedir.0.bias"]))
    other = enc(second)[2]
    cv = corr(o["conv3"].contiguous(), other.contiguous())
    o["conv3"] = _lrelu(F.conv2d(torch.cat([redir, cv], 1), sd[p + "moduleCombined.0.weight"], sd[p + "moduleCombined.0.bias"], padding=1))
    flow = _unflow_upconv(sd, p + "moduleUpconv.", _unflow_tail(sd, p, o))
    # ---- Simple x 2
    for n in (1, 2):
        p = "moduleFlownets.%d." % n
        warp = unflow_backward_warp(second, flow)
        x = torch.cat([first, second, flow, warp, (first - warp).abs()], 1)
        o = {}
        o["conv1"] = _unflow_padconv(sd, p + "moduleOne", x, [2, 4, 2, 4])
        o["conv2"] = _unflow_padconv(sd, p + "moduleTwo", o["conv1"], [1, 3, 1, 3])
        x = _unflow_padconv(sd, p + "moduleThr", o["conv2"], [1, 3, 1, 3])
        o["conv3"] = _lrelu(F.conv2d(x, sd[p + "moduleThr.3.weight"], sd[p + "moduleThr.3.bias"], padding=1))
        flow = _unflow_upconv(sd, p + "moduleUpconv.", _unflow_tail(sd, p, o))
    return flow
