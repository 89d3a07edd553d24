"""BASELINE config 4: HRNet-base MFCNet, 5-frame sliding window over a synthetic video, clips sharded across the GPUs.
  python tools/bench_stream.py [--frames 9000] [--model hrnet|resunet] [--k 5] [--clips 4]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_stream.py
Every rank takes `--clips` contiguous sub-clips of its share of the video (stream.shard_clips), each preceded by a K-1-frame
halo that only fills its feature ring, and advances them in lock step: one SFC pass over the B new frames + one fusion pass
over the B windows per step (StreamingMFCNet(batch=B)).  Frames are generated on the device from (seed, frame index), so
shards are reproducible; no collective on the data path (NCCL is used for the final max-over-ranks of the device time only).
Prints one JSON line on rank 0.  `run()` is also what bench.py calls for its secondary records."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def run(model="hrnet", K=5, F=9000, B=4, H=480, W=640, N=5, world=1, rank=0, net=None, variant="large", online_flow=False):
    """Streams the rank's clips; returns (device ms, outputs produced, launches per step).  All ranks must call it.
    online_flow: the K-1 flow fields of every clip and step come from RAFT-large on the engine (the video script's call site,
    scripts/test_multiframe_segmentation_on_videos_v3.py:264-271: half-size frames), B x (K-1) frame pairs in one call."""
    dev = torch.device("cuda", torch.cuda.current_device())
    if net is None:
        torch.manual_seed(0)
        if model == "hrnet":
            cls = M.HRNetMultiLarge if variant == "large" else M.HRNetMultiBasic
        else:
            cls = M.ResUNetMultiLarge if variant == "large" else M.ResUNetMultiBasic
        net = cls(N, K, optflow_inputs=True, depth_inputs=True).to(dev).eval()
        if model == "hrnet":     # random-init HRNet logits are O(1e3): scale the head as the fixtures do
            with torch.no_grad():
                net.base_model.last_layer[3].weight.mul_(1e-3)
                net.base_model.last_layer[3].bias.mul_(1e-3)
    runner = M.StreamingMFCNet(net, H, W, device=dev, batch=B)
    clips = M.shard_clips(F, world, rank, K, B)
    steps = max(c["hi"] - c["enc_lo"] for c in clips)
    g = torch.Generator(device=dev)
    pool = 16   # distinct synthetic frames cycled through (frame t uses entry t % pool: reproducible per frame index)
    g.manual_seed(1000)
    frames = torch.randn(pool, 3, H, W, device=dev, generator=g)
    flows = [4 * torch.randn(pool, 2, H, W, device=dev, generator=g) for _ in range(K - 1)]
    depths = [torch.rand(pool, 1, H, W, device=dev, generator=g) for _ in range(K)]
    out = torch.empty(B, N, H, W, device=dev)

    def indices(i):   # frame index of every clip at step i (finished / empty clips repeat their last frame, output discarded)
        return torch.tensor([max(0, min(c["enc_lo"] + i, c["hi"] - 1)) % pool for c in clips], device=dev)

    raft = M.raft_large().to(dev).eval() if online_flow else None
    reuse = online_flow and os.environ.get("MFC_FLOW_REUSE", "1") != "0"
    sflow = M.StreamingFlow(raft, K, H, W, batch=B) if reuse else None     # encoders once per new frame, feature-map ring
    ring = []   # (without reuse) the clips' previous frames (step i-1, i-2, ...), each (B, 3, H, W)

    def step(i):
        idx = indices(i)
        x = frames[idx]
        if raft is None:
            fl = [f[idx] for f in flows]
        elif sflow is not None:
            fl = sflow.step(x)
        else:
            prev = (ring + [x] * K)[:K - 1]            # a clip's first steps see its own frame (zero motion) for missing history
            f = M.video_flow(raft, x.repeat(K - 1, 1, 1, 1), torch.cat(prev, 0))
            fl = [f[j * B:(j + 1) * B] for j in range(K - 1)]
            ring.insert(0, x)
            del ring[K - 1:]
        return runner.step(x, fl, [d[idx] for d in depths], out=out)

    with torch.no_grad():
        for i in range(min(3 * K, steps)):   # warm-up (plans, graphs)
            step(i)
        runner.reset()
        del ring[:]
        if sflow is not None:
            sflow.reset()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_out = 0
        for i in range(steps):
            y = step(i)
            if y is not None:
                n_out += sum(1 for c in clips if c["lo"] <= c["enc_lo"] + i < c["hi"])
        e1.record()
        torch.cuda.synchronize()
    M.engine.check_overflow(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    tot = torch.tensor([n_out], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    return float(ms), int(tot), runner.launches_per_frame


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, K, model, B = arg("--frames", 9000), arg("--k", 5), arg("--model", "hrnet"), arg("--clips", 4)
    online = "--online-flow" in sys.argv
    ms, tot, launches = run(model, K, F, B, world=world, rank=rank, online_flow=online)
    if rank == 0:
        print(json.dumps({"metric": "output frames/sec, %s MFCNet K=%d sliding window%s, 480x640, clips sharded across GPUs"
                                    % (model, K, " with online RAFT-large flow" if online else ""),
                          "value": tot * 1000.0 / ms, "unit": "frames/s", "n_gpus": world, "frames": F, "clips_per_gpu": B,
                          "outputs": tot, "halo_frames_per_clip": K - 1, "ms_total": ms, "scaling": "strong",
                          "launches_per_step": launches}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
