"""CPU tests of the streaming / sharding host logic (no GPU): frame-shard arithmetic, a world_size-2
gloo run of the sharded driver loop with a fake per-frame compute, and plan construction of the
ring-buffer runner (MFC_B200_PLAN_ONLY)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_frames_cover_every_output_once():
    from mfcnet_tracker_b200.stream import shard_frames
    for F, world, K in [(9000, 8, 5), (9000, 4, 5), (100, 3, 3), (7, 8, 3), (2, 2, 3), (9000, 1, 5)]:
        seen = []
        for r in range(world):
            s = shard_frames(F, world, r, K)
            assert s["enc_lo"] == max(0, s["lo"] - (K - 1)) and s["n_out"] == s["hi"] - s["lo"]
            seen += list(range(s["lo"], s["hi"]))
        assert seen == list(range(K - 1, F)) if F >= K else seen == []


def test_shard_clips_cover_every_output_once():
    from mfcnet_tracker_b200.stream import shard_clips
    for F, world, K, B in [(9000, 8, 5, 4), (9000, 1, 5, 8), (50, 2, 3, 4), (9, 2, 5, 4)]:
        seen = []
        for r in range(world):
            clips = shard_clips(F, world, r, K, B)
            assert len(clips) == B
            for c in clips:
                assert c["enc_lo"] == max(0, c["lo"] - (K - 1))
                seen += list(range(c["lo"], c["hi"]))
        assert seen == list(range(K - 1, F))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, F, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mfcnet_tracker_b200.stream import shard_frames
    s = shard_frames(F, world, rank, K)
    # fake "video": frame t has value t; fake SFC = identity; fake fusion = sum over the window.
    ring = [None] * K
    outs = []
    for t in range(s["enc_lo"], s["hi"]):
        ring[t % K] = float(t)                      # encode (halo frames only fill the ring)
        if t >= s["lo"]:
            outs.append((t, sum(ring[(t - i) % K] for i in range(K))))
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)          # host-side merge of the per-clip results (no data-path collective)
    n = torch.tensor([len(outs)], dtype=torch.int64)
    dist.all_reduce(n)
    if rank == 0:
        q.put(([x for g in gathered for x in g], int(n)))
    dist.destroy_process_group()


def test_sharded_stream_world2_gloo():
    F, K, world = 41, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, F, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, n = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n == F - (K - 1)
    assert [t for t, _ in merged] == list(range(K - 1, F))
    for t, v in merged:                              # identical to the unsharded sliding window
        assert v == sum(float(t - i) for i in range(K))


def test_streaming_runner_plans(monkeypatch):
    monkeypatch.setenv("MFC_B200_PLAN_ONLY", "1")
    import mfcnet_tracker_b200 as m
    from mfcnet_tracker_b200.stream import StreamingMFCNet
    monkeypatch.setattr(m.abi, "_lib", None)
    try:
        for cls in (m.ResUNetMultiLarge, m.ResUNetMultiBasic):
            net = cls(5, 3, optflow_inputs=True, depth_inputs=True).eval()
            run = StreamingMFCNet(net, 64, 96, device="cpu")
            fl = [torch.zeros(1, 2, 64, 96)] * 2
            dp = [torch.zeros(1, 1, 64, 96)] * 3
            outs = [run.step(torch.zeros(1, 3, 64, 96), fl, dp) for _ in range(5)]
            assert outs[0] is None and outs[1] is None and all(o.shape == (1, 5, 64, 96) for o in outs[2:])
            assert run.launches_per_frame == 60 + 4   # 1 gather + 59 SFC launches, aux gather/warp + 3 fusion convs
            # B clips in lock step: same launches, batch-B buffers
            run4 = StreamingMFCNet(net, 64, 96, device="cpu", batch=4)
            fl4, dp4 = [torch.zeros(4, 2, 64, 96)] * 2, [torch.zeros(4, 1, 64, 96)] * 3
            outs = [run4.step(torch.zeros(4, 3, 64, 96), fl4, dp4) for _ in range(4)]
            assert outs[1] is None and outs[2].shape == (4, 5, 64, 96) and run4.launches_per_frame == 60 + 4
    finally:
        m.abi._lib = None
