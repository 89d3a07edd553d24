"""CPU tests of the training-step host logic (mfcnet_tracker_b200/train.py): the differentiable reference-math forward
against the REAL reference's outputs (committed fixtures), the whole step (forward, loss, backward, Adam with the two
parameter groups) against the reference's own training step (tests/golden/train_step.json, oracle/make_golden_train.py)
with a torch stand-in for the two CUDA kernels, and the world_size-2 gloo exchange (flat gradient bucket all-reduce +
all-reduced loss statistics = the gradient of the global-batch loss)."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

if not torch.cuda.is_available():      # CPU box: the C-ABI library loads in plan-only mode (no device calls)
    os.environ.setdefault("MFC_B200_PLAN_ONLY", "1")
from oracle import synth  # noqa: E402
from tests import golden_util as G  # noqa: E402

CW = [1.0, 1000.0, 1000.0, 1000.0, 1000.0]


def _torch_loss(out, tgt):
    """src/loss.py get_loss(nll + soft_jaccard, 0.7 / 0.3) in differentiable torch ops (float32, as the reference)."""
    logp = F.log_softmax(out, dim=1)
    nll = F.nll_loss(logp, tgt, weight=torch.tensor(CW))
    jac = 0.0
    N = out.shape[1]
    for c in range(1, N):
        t = (tgt == c).float()
        p = logp[:, c].exp()
        inter = (p * t).sum()
        jac = jac - torch.log((inter + 1e-15) / (p.sum() + t.sum() - inter + 1e-15))
    jac = jac / N
    return 0.7 * nll + 0.3 * jac, nll, jac


@pytest.mark.parametrize("variant", ["large", "basic"])
def test_autograd_forward_matches_reference_outputs(variant):
    import mfcnet_tracker_b200 as M
    tag = f"mfcnet_resunet16_{variant}_k3_64x96"
    meta, man, arr = G.load(tag)
    cls = M.ResUNetMultiLarge if variant == "large" else M.ResUNetMultiBasic
    net = cls(num_classes=meta["N"], num_frames=meta["K"], pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True)
    net.load_state_dict(G.state_dict(man, meta["seed"]), strict=True)
    net.eval()
    xs, fl, dp = G.mfcnet_inputs(tag, meta)
    with torch.no_grad():
        y = M.autograd_forward(net, xs, fl, dp)
    assert float((y - torch.from_numpy(arr["out"])).abs().max()) < 2e-4


def _inputs(variant, s, c):
    tag = "train/%s/%d" % (variant, s)
    B, H, W, K, N = c["B"], c["H"], c["W"], c["K"], c["N"]
    xs = [torch.from_numpy(synth.frames("%s/%d" % (tag, i), B, H, W, 11)) for i in range(K)]
    fl = [torch.from_numpy(synth.flow("%s/%d" % (tag, i), B, H, W, 11, scale=2.0)) for i in range(K - 1)]
    dp = [torch.from_numpy(synth.depth("%s/%d" % (tag, i), B, H, W, 11)) for i in range(K)]
    _, tg = synth.loss_case(tag, B, N, H, W, seed=11, fg=0.3)
    return xs, fl, dp, torch.from_numpy(tg)


def _make(variant, c):
    import mfcnet_tracker_b200 as M
    cls = M.ResUNetMultiLarge if variant == "large" else M.ResUNetMultiBasic
    net = cls(num_classes=c["N"], num_frames=c["K"], pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True)
    man = [(k, tuple(s), d) for k, s, d in c["cases"][variant]["manifest"]]
    net.load_state_dict(G.state_dict(man, 11), strict=True)
    return net


@pytest.mark.parametrize("variant", ["large", "basic"])
def test_training_step_matches_reference(variant):
    """Two steps of DataParallelTrainer (CPU: torch stand-ins for the loss and Adam kernels) against the reference run."""
    import mfcnet_tracker_b200 as M
    with open(os.path.join(G.GOLDEN, "train_step.json")) as f:
        c = json.load(f)
    net = _make(variant, c)
    tr = M.DataParallelTrainer(net, lr=c["lr"])
    keys_before = list(net.state_dict().keys())
    for s in range(c["steps"]):
        xs, fl, dp, tg = _inputs(variant, s, c)
        net.train()
        tr.zero_grad()
        out = M.autograd_forward(net, xs, fl, dp)
        total, nll, jac = _torch_loss(out, tg)
        total.backward()
        tr.exchange()
        tr.optimizer_step()
        ref = c["cases"][variant]["losses"][s]
        assert abs(float(total) - ref[0]) < 2e-4 * max(1.0, abs(ref[0])), (s, float(total), ref)
        assert abs(float(nll) - ref[1]) < 2e-4 * max(1.0, abs(ref[1])) and abs(float(jac) - ref[2]) < 2e-4 * max(1.0, abs(ref[2]))
    sd = net.state_dict()
    assert list(sd.keys()) == keys_before          # re-homing the parameters into flat buckets keeps the checkpoint layout
    for k, pr in c["cases"][variant]["probes"].items():
        got = sd[k].reshape(-1)[:6].double()
        want = torch.tensor(pr["head"], dtype=torch.float64)
        assert float((got - want).abs().max()) < 5e-5 + 2e-3 * float(want.abs().max()), (k, got, want)
        assert abs(float(sd[k].double().norm()) - pr["norm"]) < 1e-3 * max(1.0, pr["norm"]), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _loss_sums(out, tgt):
    """The additive loss statistics of mfc_segmentation_loss_sums, in torch (float64)."""
    logp = F.log_softmax(out.double(), dim=1)
    w = torch.tensor(CW, dtype=torch.float64)[tgt]
    s = [-(w * logp.gather(1, tgt[:, None]).squeeze(1)).sum(), w.sum()]
    for c in range(1, out.shape[1]):
        t = (tgt == c).double()
        p = logp[:, c].exp()
        s += [(p * t).sum(), p.sum(), t.sum()]
    return torch.stack(s)


def _total_from_sums(s, N):
    nll = s[0] / s[1]
    jac = 0.0
    for c in range(1, N):
        i, sm, t = s[2 + 3 * (c - 1)], s[3 + 3 * (c - 1)], s[4 + 3 * (c - 1)]
        jac = jac - torch.log((i + 1e-15) / (sm + t - i + 1e-15))
    return 0.7 * nll + 0.3 * jac / N


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import mfcnet_tracker_b200 as M
    with open(os.path.join(G.GOLDEN, "train_step.json")) as f:
        c = json.load(f)
    net = _make("large", c)
    if rank == 1:                      # a rank that starts from different weights must be overwritten by rank 0's
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.1)
    tr = M.DataParallelTrainer(net, lr=c["lr"])
    xs, fl, dp, tg = _inputs("large", 0, c)
    sl = slice(rank, rank + 1)         # the batch of 2 is sharded 1 + 1
    net.train()
    tr.zero_grad()
    out = M.autograd_forward(net, [x[sl] for x in xs], [x[sl] for x in fl], [x[sl] for x in dp])
    # global-batch loss: local statistics, all-reduced with everything but this rank's own record detached
    local = _loss_sums(out, tg[sl])
    glob = local.detach().clone()
    dist.all_reduce(glob)
    total = _total_from_sums(glob - local.detach() + local, c["N"])
    total.backward()
    tr.exchange()
    tr.optimizer_step()
    flat = torch.cat([b.flat for b, _ in tr.buckets])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((float(total), [g.numpy() for g in gathered]))
    dist.destroy_process_group()


def test_data_parallel_step_world2_gloo():
    """Two ranks, one sample each: identical weights on both ranks afterwards, and the loss each rank evaluates is the loss of
    the global batch (no BatchNorm coupling is expected: like nn.DataParallel, BN statistics are per replica)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, flats = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(flats[0], flats[1])
    assert np.isfinite(total) and 0.5 < total < 10.0
