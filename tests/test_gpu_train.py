"""GPU tests of the training-step kernels through the C ABI (mfc_segmentation_loss_sums / _from_sums / _bwd, mfc_adam_step)
and of DataParallelTrainer.step against the reference's own training step (tests/golden/train_step.json)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from tests import golden_util as G
from tests.test_train_cpu import CW, _inputs, _make

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mfcnet_tracker_b200 as m
    return m


@pytest.mark.parametrize("case", [(2, 5, 48, 64, 0.05), (1, 5, 96, 128, 0.01), (3, 4, 33, 47, 0.6)])
def test_loss_grad_matches_autograd(M, case):
    from oracle import synth
    B, N, H, W, fg = case
    o, t = synth.loss_case("g%d" % H, B, N, H, W, seed=5, fg=fg)
    cw = CW[:N]
    x = torch.from_numpy(o).double().requires_grad_(True)
    tt = torch.from_numpy(t)
    logp = F.log_softmax(x, dim=1)
    nll = F.nll_loss(logp, tt, weight=torch.tensor(cw, dtype=torch.float64))
    jac = 0.0
    for c in range(1, N):
        m = (tt == c).double()
        p = logp[:, c].exp()
        inter = (p * m).sum()
        jac = jac - torch.log((inter + 1e-15) / (p.sum() + m.sum() - inter + 1e-15))
    total = 0.7 * nll + 0.3 * jac / N
    total.backward()
    losses, grad = M.loss_and_grad(torch.from_numpy(o).cuda(), tt.cuda(), cw)
    assert abs(float(losses[0]) - float(total)) < 1e-5 * max(1.0, abs(float(total)))
    ref = x.grad.float()
    err = float((grad.cpu() - ref).abs().max())
    assert err < 1e-5 * float(ref.abs().max()) + 1e-9, (err, float(ref.abs().max()))


def test_loss_ignores_out_of_range_labels(M):
    """nn.NLLLoss's ignore_index = -100 (src/loss.py:38 uses the default): such pixels add nothing to the NLL term nor to the
    per-class target masks, and other out-of-range labels (255 = void) must not read outside the class-weight table."""
    from oracle import synth
    B, N, H, W = 2, 5, 40, 56
    o, t = synth.loss_case("ign", B, N, H, W, seed=2, fg=0.2)
    t[:, :5, :] = -100
    cw = CW[:N]
    x = torch.from_numpy(o).double().requires_grad_(True)
    tt = torch.from_numpy(t)
    logp = F.log_softmax(x, dim=1)
    nll = F.nll_loss(logp, tt, weight=torch.tensor(cw, dtype=torch.float64))       # ignore_index=-100
    jac = 0.0
    for c in range(1, N):
        m = (tt == c).double()
        p = logp[:, c].exp()
        inter = (p * m).sum()
        jac = jac - torch.log((inter + 1e-15) / (p.sum() + m.sum() - inter + 1e-15))
    total = 0.7 * nll + 0.3 * jac / N
    total.backward()
    losses, grad = M.loss_and_grad(torch.from_numpy(o).cuda(), tt.cuda(), cw)
    assert abs(float(losses[0]) - float(total)) < 1e-5 * max(1.0, abs(float(total)))
    ref = x.grad.float()
    assert float((grad.cpu() - ref).abs().max()) < 1e-5 * float(ref.abs().max()) + 1e-9
    t2 = t.copy()
    t2[t2 == -100] = 255                                                           # any other out-of-range value: same result
    losses2, grad2 = M.loss_and_grad(torch.from_numpy(o).cuda(), torch.from_numpy(t2).cuda(), cw)
    assert float(losses2[0]) == float(losses[0]) and torch.equal(grad2, grad)


def test_adam_matches_torch(M):
    from mfcnet_tracker_b200 import abi
    lib = abi.load()
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=3e-3)
    p = p0.clone().cuda()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n) * (0.1 if step % 2 else 3.0)
        p_ref.grad = g.clone()
        opt.step()
        gd = (2.0 * g).cuda()      # exercised with grad_scale = 0.5
        abi.check(lib.mfc_adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 3e-3, 0.9, 0.999, 1e-8, 0.0, step, 0.5,
                                    torch.cuda.current_stream().cuda_stream))
        assert float((p.cpu() - p_ref.data).abs().max()) < 2e-6, step


@pytest.mark.parametrize("variant", ["large", "basic"])
def test_trainer_step_matches_reference(M, variant):
    with open(os.path.join(G.GOLDEN, "train_step.json")) as f:
        c = json.load(f)
    torch.backends.cudnn.allow_tf32 = False      # the reference run behind the fixture is fp32 on the CPU
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _make(variant, c).cuda()
    tr = M.DataParallelTrainer(net, lr=c["lr"])
    for s in range(c["steps"]):
        xs, fl, dp, tg = _inputs(variant, s, c)
        losses = tr.step([x.cuda() for x in xs], tg.cuda(), optflow=[x.cuda() for x in fl], depth=[x.cuda() for x in dp]).tolist()
        ref = c["cases"][variant]["losses"][s]
        for a, b in zip(losses, ref):
            assert abs(a - b) < 5e-4 * max(1.0, abs(b)), (s, losses, ref)
    sd = net.state_dict()
    for k, pr in c["cases"][variant]["probes"].items():
        got = sd[k].reshape(-1)[:6].double().cpu()
        want = torch.tensor(pr["head"], dtype=torch.float64)
        # Adam's first steps move every element by ~lr whatever the gradient's size: an element whose gradient is ~0 may
        # legitimately end up a full step apart under a different summation order; the norm pins the bulk
        assert float((got - want).abs().max()) <= 1.05 * c["steps"] * c["lr"] + 1e-5, (k, got, want)
        assert abs(float(sd[k].double().norm()) - pr["norm"]) < 2e-3 * max(1.0, pr["norm"]), k
    # the trained weights are picked up by the inference engine (packed-weight caches are dropped by the optimiser step)
    net.eval()
    xs, fl, dp, _ = _inputs(variant, 0, c)
    with torch.no_grad():
        y = net([x.cuda() for x in xs], optflow=[x.cuda() for x in fl], depth=[x.cuda() for x in dp])
        y_ref = M.autograd_forward(net, [x.cuda() for x in xs], [x.cuda() for x in fl], [x.cuda() for x in dp])
    assert float((y - y_ref).abs().max()) < 2e-2
