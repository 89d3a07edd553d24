"""Training-step throughput (BASELINE config 5): ResUNetMulti-Large K=3 with flow + depth inputs, 480x640, batch 8 per GPU,
nll + soft_jaccard loss, Adam with the reference's two parameter groups, one process per GPU.
  python tools/bench_train.py [--batch 8] [--steps 10] [--warmup 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py
Prints one JSON line on rank 0: samples/s over all ranks (device-timed, max over ranks), and the per-phase split of one step."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, steps, warmup = arg("--batch", 8), arg("--steps", 10), arg("--warmup", 3)
    H, W, N, K = 480, 640, 5, 3
    torch.manual_seed(0)
    net = M.ResUNetMultiLarge(N, K, optflow_inputs=True, depth_inputs=True).cuda()
    tr = M.DataParallelTrainer(net, lr=1e-4)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    frames = [torch.randn(B, 3, H, W, device="cuda", generator=g) for _ in range(K)]
    flows = [4 * torch.randn(B, 2, H, W, device="cuda", generator=g) for _ in range(K - 1)]
    depths = [torch.rand(B, 1, H, W, device="cuda", generator=g) for _ in range(K)]
    u = torch.rand(B, H, W, device="cuda", generator=g)
    tgt = torch.zeros(B, H, W, dtype=torch.int64, device="cuda")
    for c in range(1, N):
        tgt[(u >= (c - 1) * 0.01 / (N - 1)) & (u < c * 0.01 / (N - 1))] = c      # ~1 % foreground
    for _ in range(warmup):
        losses = tr.step(frames, tgt, optflow=flows, depth=depths)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses = tr.step(frames, tgt, optflow=flows, depth=depths)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # phase split of one more step (this rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    net.train()
    tr.zero_grad()
    ev[0].record()
    out = M.autograd_forward(net, frames, flows, depths)
    ev[1].record()
    l2, dl = M.loss_and_grad(out, tgt, tr.class_weights, group=None, world=world)
    ev[2].record()
    out.backward(dl)
    ev[3].record()
    tr.exchange()
    tr.optimizer_step()
    ev[4].record()
    torch.cuda.synchronize()
    if rank == 0:
        nparam = sum(b.numel for b, _ in tr.buckets)
        print(json.dumps({"metric": "training samples/sec, MFCNet K=3 (ResUNet-16 + Large fusion) 480x640, nll+soft_jaccard, Adam",
                          "value": world * B * 1000.0 / float(ms), "unit": "samples/s", "n_gpus": world, "batch_per_gpu": B,
                          "ms_per_step": float(ms), "loss": [round(v, 5) for v in losses.tolist()],
                          "phases_ms": {"forward_autograd(ATen)": ev[0].elapsed_time(ev[1]), "loss_fwd+bwd(kernels, +14-double allreduce)": ev[1].elapsed_time(ev[2]),
                                        "backward_autograd(ATen)": ev[2].elapsed_time(ev[3]),
                                        "allreduce(NCCL flat bucket)+adam(kernel)": ev[3].elapsed_time(ev[4])},
                          "grad_bucket_bytes": 4 * nparam, "scaling": "weak"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
