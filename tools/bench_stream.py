"""BASELINE config 4: HRNet-base MFCNet, 5-frame sliding window over a synthetic video, clips sharded across the GPUs.
  python tools/bench_stream.py [--frames 9000] [--model hrnet|resunet] [--k 5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_stream.py
Every rank takes one contiguous clip (stream.shard_frames) plus a K-1-frame halo that only fills its feature ring; frames are
generated on the device from (seed, frame index), so shards are reproducible; no collective on the data path (NCCL is used
for the final max-over-ranks of the device time only).  Prints one JSON line on rank 0."""
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
    F, K, model = arg("--frames", 9000), arg("--k", 5), arg("--model", "hrnet")
    H, W, N = 480, 640, 5
    torch.manual_seed(0)
    cls = M.HRNetMultiLarge if model == "hrnet" else M.ResUNetMultiLarge
    net = cls(N, K, optflow_inputs=True, depth_inputs=True).cuda().eval()
    run = M.StreamingMFCNet(net, H, W)
    sh = M.shard_frames(F, world, rank, K)
    g = torch.Generator(device="cuda")
    pool = 16   # distinct synthetic frames cycled through (frame t uses entry t % pool: reproducible per frame index)
    frames, flows, depths = [], [], []
    for t in range(pool):
        g.manual_seed(1000 + t)
        frames.append(torch.randn(1, 3, H, W, device="cuda", generator=g))
        flows.append([4 * torch.randn(1, 2, H, W, device="cuda", generator=g) for _ in range(K - 1)])
        depths.append([torch.rand(1, 1, H, W, device="cuda", generator=g) for _ in range(K)])
    out = torch.empty(1, N, H, W, device="cuda")
    amax = torch.empty(1, H, W, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        for t in range(sh["enc_lo"], min(sh["enc_lo"] + 3 * K, sh["hi"])):   # warm-up (plans, autotuning, graphs)
            run.step(frames[t % pool], flows[t % pool], depths[t % pool], out=out)
        run.reset()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_out = 0
        for t in range(sh["enc_lo"], sh["hi"]):
            y = run.step(frames[t % pool], flows[t % pool], depths[t % pool], out=out)
            if y is not None and t >= sh["lo"]:
                n_out += 1
        e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    tot = torch.tensor([n_out], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    if rank == 0:
        print(json.dumps({"metric": "output frames/sec, %s MFCNet K=%d sliding window, 480x640, clips sharded across GPUs" % (model, K),
                          "value": float(tot) * 1000.0 / float(ms), "unit": "frames/s", "n_gpus": world, "frames": F,
                          "outputs": int(tot), "halo_frames_per_rank": K - 1, "ms_total": float(ms), "scaling": "strong",
                          "launches_per_frame": run.launches_per_frame}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
