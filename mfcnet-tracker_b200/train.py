"""Data-parallel training step of the MFCNet wrappers (BASELINE config 5).

Reference: scripts/train_multiframe_detection.py:106-157 (``nn.DataParallel``, Adam with two parameter groups:
``base_model`` at lr/K -- or lr/(100 K) when pre-trained base weights are loaded -- and ``multiframe_net`` at lr) and
src/engine.py:56-71 (forward -> ``F.log_softmax`` -> ``get_loss`` -> ``loss.backward()`` -> ``optimizer.step()``).

What runs where in ``DataParallelTrainer.step``:

* one process per GPU (``torch.distributed``), each with its own shard of the batch -- instead of ``nn.DataParallel``'s
  scatter / replicate / gather-to-GPU0 threads;
* forward and backward of the networks: **torch autograd over the reference's own math** (``autograd_forward`` below: WS-conv
  + GroupNorm + SiLU blocks, train-mode BatchNorm in the fusion head, the flow warp).  The B200 conv engine of this
  package is inference-only; hand-written dgrad / wgrad kernels are the next row.  This is a library (ATen /
  cuDNN) path and is reported as such;
* the loss: ``mfc_segmentation_loss_sums`` (one pass: the additive loss statistics of the shard) -> a 14-double
  ``all_reduce`` -> ``mfc_segmentation_loss_from_sums`` + ``mfc_segmentation_loss_bwd`` (d total / d logits of the shard under
  the GLOBAL-batch loss), hand-written -- autograd starts from the model output with that gradient.  This reproduces what
  ``nn.DataParallel`` computes (outputs gathered to GPU 0, one loss over the whole batch) without gathering anything;
* the exchange step: all parameters and all gradients live in flat fp32 buckets (``param.data`` / ``param.grad`` are views),
  so the gradient reduction is ONE ``all_reduce(SUM)`` per parameter group over NCCL (NVLink / NVSwitch) on the flat
  gradient bucket (4.6 MB for ResUNet-16 MFCNet: latency-bound, no bucketing needed); the shard gradients of the global
  loss simply add, there is no 1/world factor;
* the optimiser: ``mfc_adam_step`` (torch.optim.Adam semantics) on the flat bucket, one launch per parameter group.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import abi, engine

# README.md:62-66 of the reference: nll + soft_jaccard, weights 0.7 / 0.3, class weights 1 / 1000 x (N-1)
DEFAULT_LOSS_FNS = ("nll", "soft_jaccard")
DEFAULT_LOSS_WTS = (0.7, 0.3)


def default_class_weights(num_classes):
    return [1.0] + [1000.0] * (num_classes - 1)


# ------------------------------------------------------------------------------------------------
# differentiable forward (reference math, torch ops) over the drop-in modules' own parameters
# ------------------------------------------------------------------------------------------------
def _weight_standardize(w, eps=1e-5):
    """models/resunet.py:56-62 (fp32: eps 1e-5, biased variance)."""
    flat = w.reshape(w.shape[0], -1)
    mu = flat.mean(dim=1).reshape(-1, 1, 1, 1)
    var = flat.var(dim=1, unbiased=False).reshape(-1, 1, 1, 1)
    return (w - mu) * torch.rsqrt(var + eps)


def _block(sd, p, x, groups):
    y = F.conv2d(x, _weight_standardize(sd[p + "proj.weight"]), sd[p + "proj.bias"], padding=1)
    return F.silu(F.group_norm(y, groups, sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-5))


def _resnet_block(sd, p, x, groups):
    h = _block(sd, p + "block2.", _block(sd, p + "block1.", x, groups), groups)
    if p + "res_conv.weight" in sd:
        x = F.conv2d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def _pixel_unshuffle2(x):
    b, c, h, w = x.shape
    return x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, h // 2, w // 2)


def resunet_autograd_forward(sd, x, groups=8):
    """ResUnet_VB.forward (models/resunet.py:153-180) over a {name: tensor} dict of its parameters."""
    n_levels = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    stem, skips = x, []
    for i in range(n_levels):
        x = _resnet_block(sd, "downs.%d.0." % i, x, groups)
        skips.append(x)
        if "downs.%d.1.1.weight" % i in sd:
            x = F.conv2d(_pixel_unshuffle2(x), sd["downs.%d.1.1.weight" % i], sd["downs.%d.1.1.bias" % i])
        else:
            x = F.conv2d(x, sd["downs.%d.1.weight" % i], sd["downs.%d.1.bias" % i], padding=1)
    x = _resnet_block(sd, "mid_block.", x, groups)
    for i in range(n_levels):
        x = _resnet_block(sd, "ups.%d.0." % i, torch.cat((x, skips.pop()), dim=1), groups)
        if "ups.%d.1.1.weight" % i in sd:
            x = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), sd["ups.%d.1.1.weight" % i], sd["ups.%d.1.1.bias" % i], padding=1)
        else:
            x = F.conv2d(x, sd["ups.%d.1.weight" % i], sd["ups.%d.1.bias" % i], padding=1)
    x = _resnet_block(sd, "final_res_block.", torch.cat((x, stem), dim=1), groups)
    return F.conv2d(x, sd["output_layer.weight"], sd["output_layer.bias"])


def _fusion_stack(seq, x, training):
    """The conv-BN-ReLU x3 + 1x1 Sequential (models/multiframe_model.py:62-73, :191-202); BatchNorm in batch-statistics
    mode when training (running statistics are updated in place, momentum 0.1, as nn.BatchNorm2d does)."""
    for ci, bi in ((0, 1), (3, 4), (6, 7)):
        conv, bn = seq[ci], seq[bi]
        x = F.conv2d(x, conv.weight, None, padding=conv.padding)
        x = F.relu(F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=training,
                                momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps))
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return F.conv2d(x, seq[9].weight, None)


def _warp_single(m, flow, grid):
    """_warp_single_map (models/multiframe_model.py:141-170): cropped 576x720 grid + normalised flow, bilinear, zeros."""
    H, W = m.shape[-2:]
    g = grid[:, :, :H, :W]
    new = (g + torch.stack((flow[:, 0] / ((W - 1) / 2.0), flow[:, 1] / ((H - 1) / 2.0)), dim=1)).permute(0, 2, 3, 1)
    return F.grid_sample(m, new, mode="bilinear", padding_mode="zeros", align_corners=True)


def autograd_forward(model, frames, optflow=None, depth=None):
    """Differentiable forward of a ResUNetMulti{Basic,Large} wrapper (models/multiframe_model.py:424-438 pattern: SFC net on
    every frame, raw logits + flows + depths concatenated, fusion head), honouring model.training for BatchNorm."""
    from .fusion import MultiFrameNetBasic
    from .resunet import ResUnet_VB
    if not isinstance(model.base_model, ResUnet_VB):
        raise NotImplementedError("the training step covers the ResUNet MFCNet wrappers (BASELINE config 5)")
    K, N = model.num_frames, model.num_classes
    sd = dict(model.base_model.named_parameters())
    groups = next((m.num_groups for m in model.base_model.modules() if isinstance(m, torch.nn.GroupNorm)), 8)
    B = frames[0].shape[0]
    logits = resunet_autograd_forward(sd, torch.cat(list(frames), dim=0), groups)      # frames are independent: one batched pass
    maps = list(logits.split(B, dim=0))
    head = model.multiframe_net
    seq = head.multiframe_net
    if isinstance(head, MultiFrameNetBasic) and model.optflow_inputs:
        segs, deps = [maps[0]], ([depth[0]] if model.depth_inputs else [])
        for i in range(1, K):
            f = optflow[i - 1]
            segs.append(torch.cat([_warp_single(maps[i][:, j:j + 1], f, head.grid) for j in range(N)], dim=1))
            if model.depth_inputs:
                deps.append(_warp_single(depth[i], f, head.grid))
        x = torch.cat(segs + deps, dim=1)
    else:
        x = torch.cat(maps + (list(optflow) if model.optflow_inputs else []) + (list(depth) if model.depth_inputs else []), dim=1)
    return _fusion_stack(seq, x, model.training)


# ------------------------------------------------------------------------------------------------
# loss with a hand-written backward
# ------------------------------------------------------------------------------------------------
def loss_and_grad(output, targets, class_weights=None, loss_fns=DEFAULT_LOSS_FNS, loss_wts=DEFAULT_LOSS_WTS, group=None, world=1):
    """(losses[3], dlogits): get_loss(F.log_softmax(output), targets) of src/loss.py over the GLOBAL batch of all ranks and
    d total / d output for this rank's shard, from libmfcnet_b200.so.  The loss statistics (weighted NLL sums, per-class
    intersection / sum / count) are additive, so the ranks add their 14-double records with one all-reduce and each
    evaluates the same global loss -- the quantity nn.DataParallel computes on GPU 0 after gathering the outputs."""
    engine.require_cuda(output, "loss_and_grad")
    w = {"nll": 0.0, "soft_jaccard": 0.0}
    for fn, wt in zip(loss_fns, loss_wts):
        if fn not in w:
            raise ValueError(f"Loss function {fn} not implemented")
        w[fn] += float(wt)
    lib = abi.load()
    x = output.detach().contiguous().float()
    t = targets.contiguous().to(torch.int64)
    B, N, H, W = x.shape
    cw = None if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32, device=x.device).contiguous()
    ws = torch.empty(max(8, int(lib.mfc_segmentation_loss_workspace(B, N, H * W))), dtype=torch.uint8, device=x.device)
    sums = torch.empty(2 + 3 * (N - 1), dtype=torch.float64, device=x.device)
    out = torch.empty(3, dtype=torch.float32, device=x.device)
    coef = torch.empty(64, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x)
    with engine.device_guard(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        abi.check(lib.mfc_segmentation_loss_sums(x.data_ptr(), t.data_ptr(), abi.ptr(cw), B, N, H * W, ws.data_ptr(), sums.data_ptr(), st))
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        abi.check(lib.mfc_segmentation_loss_from_sums(sums.data_ptr(), N, w["nll"], w["soft_jaccard"], out.data_ptr(), st))
        abi.check(lib.mfc_segmentation_loss_bwd(x.data_ptr(), t.data_ptr(), abi.ptr(cw), B, N, H * W, w["nll"], w["soft_jaccard"],
                                                1.0, sums.data_ptr(), 1, coef.data_ptr(), grad.data_ptr(), st))
    return out, grad


# ------------------------------------------------------------------------------------------------
# flat buckets, exchange step, optimiser
# ------------------------------------------------------------------------------------------------
class FlatBucket:
    """All parameters of `params` re-homed into one flat fp32 tensor (param.data becomes a view, so state_dict keys, shapes and
    values are unchanged) with a parallel flat gradient tensor (param.grad views) and Adam moments."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            off += k
        self.numel = n


def adam_step_torch(bucket, lr, betas, eps, step, grad_scale):
    """Reference-math Adam on a flat bucket in plain torch ops: the CPU stand-in used by the gloo tests of the host logic
    (the product path on a GPU is `mfc_adam_step`)."""
    g = bucket.grad * grad_scale
    bucket.exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    bucket.exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    bucket.flat.addcdiv_(bucket.exp_avg, bucket.exp_avg_sq.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)


class DataParallelTrainer:
    """One rank of the data-parallel training job.  `model` is a ResUNetMulti{Basic,Large} wrapper on this rank's GPU.

    lr groups follow scripts/train_multiframe_detection.py:128-151: base_model at lr/K (lr/(100 K) with
    `pretrained_base=True`), multiframe_net at lr; `train_base_model=False` freezes the base network (:142-148)."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, train_base_model=True, pretrained_base=False,
                 class_weights=None, loss_fns=DEFAULT_LOSS_FNS, loss_wts=DEFAULT_LOSS_WTS, process_group=None):
        self.model = model
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.group = process_group
        K = model.num_frames
        self.train_base_model = train_base_model
        if not train_base_model:
            for p in model.base_model.parameters():
                p.requires_grad = False
        self.buckets = []   # (FlatBucket, lr)
        if train_base_model:
            self.buckets.append((FlatBucket(model.base_model.parameters()), lr / (100.0 * K) if pretrained_base else lr / K))
        self.buckets.append((FlatBucket(model.multiframe_net.parameters()), lr))
        self.betas, self.eps = betas, eps
        self.class_weights = class_weights if class_weights is not None else default_class_weights(model.num_classes)
        self.loss_fns, self.loss_wts = loss_fns, loss_wts
        self.steps = 0
        self.on_gpu = self.buckets[0][0].flat.is_cuda
        if self.world > 1:   # every rank starts from rank 0's weights, as DataParallel's replicate() does each step
            for b, _ in self.buckets:
                dist.broadcast(b.flat, src=0, group=self.group)
            for buf in model.buffers():
                dist.broadcast(buf, src=0, group=self.group)

    def zero_grad(self):
        for b, _ in self.buckets:
            b.grad.zero_()

    def exchange(self):
        """The one collective of the step: SUM of the flat gradient buckets over all ranks (NCCL over NVLink on GPUs)."""
        if self.world > 1:
            for b, _ in self.buckets:
                dist.all_reduce(b.grad, op=dist.ReduceOp.SUM, group=self.group)

    def optimizer_step(self):
        self.steps += 1
        scale = 1.0   # the loss is already normalised over the global batch: the summed shard gradients ARE its gradient
        for b, lr in self.buckets:
            if self.on_gpu:
                lib = abi.load()
                with engine.device_guard(b.flat.device):
                    abi.check(lib.mfc_adam_step(b.flat.data_ptr(), b.grad.data_ptr(), b.exp_avg.data_ptr(), b.exp_avg_sq.data_ptr(),
                                                b.numel, lr, self.betas[0], self.betas[1], self.eps, 0.0, self.steps, scale,
                                                torch.cuda.current_stream(b.flat.device).cuda_stream))
            else:
                adam_step_torch(b, lr, self.betas, self.eps, self.steps, scale)
        # the kernels update the weights through raw pointers: drop the inference engine's packed / folded weight caches
        for m in self.model.modules():
            if hasattr(m, "_fingerprint"):
                m._fingerprint = None

    def step(self, frames, targets, optflow=None, depth=None):
        """One training step on this rank's shard; returns the 3 loss values (total, nll, jaccard) as a device tensor."""
        model = self.model
        model.train()
        if not self.train_base_model:
            model.base_model.eval()
        self.zero_grad()
        out = autograd_forward(model, frames, optflow, depth)
        losses, dlogits = loss_and_grad(out, targets, self.class_weights, self.loss_fns, self.loss_wts, self.group, self.world)
        out.backward(dlogits)
        self.exchange()
        self.optimizer_step()
        return losses
