"""The video script's per-frame loop end to end (scripts/test_multiframe_segmentation_on_videos_v3.py:233-303) on one GPU:
uint8 frame ingest -> StreamingMFCNet.step (K-frame window, feature ring, one graph replay) -> heat-map head (probabilities)
-> ToolTracker.step (class map, contours, tip maxima, association; one synchronisation) -> the frame's 12 coordinates on the host.
  python tools/bench_video.py [--frames 300] [--model resunet|hrnet] [--k 3]
Random-init weights give salt-and-pepper class maps (thousands of contours per mask): the tracking share measured here is a
worst case; `tools/bench_tracking.py` has the blob-like maps of a trained network.  Wall-clock timing (the loop synchronises
every frame by construction).  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    F, K, model = arg("--frames", 300), arg("--k", 3), arg("--model", "resunet")
    H, W, N = 480, 640, 5
    torch.manual_seed(0)
    cls = M.HRNetMultiLarge if model == "hrnet" else M.ResUNetMultiLarge
    net = cls(N, K, optflow_inputs=True, depth_inputs=True).cuda().eval()
    run = M.StreamingMFCNet(net, H, W)
    g = torch.Generator(device="cuda")
    pool = 8
    bgr, gray, flows = [], [], []
    for t in range(pool):
        g.manual_seed(2000 + t)
        bgr.append(torch.randint(0, 256, (H, W, 3), device="cuda", generator=g, dtype=torch.uint8))
        gray.append(torch.randint(0, 256, (H, W, 3), device="cuda", generator=g, dtype=torch.uint8))
        flows.append([4 * torch.randn(1, 2, H, W, device="cuda", generator=g) for _ in range(K - 1)])
    out = torch.empty(1, N, H, W, device="cuda")
    depth_ring = []
    # --online-flow: the K-1 flow fields of every frame come from RAFT-large on the engine, as in the video script
    # (scripts/test_multiframe_segmentation_on_videos_v3.py:264-271: half-size frames, flow / 0.5 resized back), one batched call
    online = "--online-flow" in sys.argv
    raft = M.raft_large().cuda().eval() if online else None
    frame_ring = []

    def flows_for(t, x):
        if not online:
            return flows[t % pool]
        frame_ring.insert(0, x)
        del frame_ring[K:]
        prev = (frame_ring + frame_ring[-1:] * K)[1:K]
        f = M.video_flow(raft, x.expand(K - 1, -1, -1, -1), torch.cat(prev, 0))
        return [f[i:i + 1] for i in range(K - 1)]

    def forward(t):
        x = M.ingest_rgb(bgr[t % pool])
        depth_ring.insert(0, M.ingest_depth(gray[t % pool]))
        del depth_ring[K:]
        y = run.step(x, flows_for(t, x), (depth_ring + depth_ring[-1:] * K)[:K], out=out)
        return None if y is None else M.heatmap_head(y, want_logp=False, want_argmax=False)[1]

    def frame(t, track):
        x = M.ingest_rgb(bgr[t % pool])
        depth_ring.insert(0, M.ingest_depth(gray[t % pool]))
        del depth_ring[K:]
        y = run.step(x, flows_for(t, x), (depth_ring + depth_ring[-1:] * K)[:K], out=out)
        if y is None or track is None:
            return None
        _, prob, _ = M.heatmap_head(y, want_logp=False, want_argmax=False)
        return track.step(prob)

    res = {}
    with torch.no_grad():
        for name, mk in (("model_only", lambda: None), ("model_head_tracking", lambda: M.ToolTracker(10, 40, 0.0))):
            run.reset()
            del depth_ring[:]
            del frame_ring[:]
            tr = mk()
            for t in range(3 * K):
                frame(t, tr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t in range(3 * K, 3 * K + F):
                row = frame(t, tr)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            res[name] = {"frames_per_s": round(F / dt, 1), "ms_per_frame": round(dt / F * 1e3, 3)}
        res["last_row"] = [None if v != v else float(v) for v in row]
        # one frame in flight: frame t's tracking is submitted, frame t+1's forward is launched, then frame t's row is collected
        run.reset()
        del depth_ring[:]
        del frame_ring[:]
        tr = M.ToolTracker(10, 40, 0.0)
        for t in range(3 * K):
            frame(t, tr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = []
        for t in range(3 * K, 3 * K + F):
            prob = forward(t)
            if tr._pending is not None:
                rows.append(tr.collect())
            tr.submit(prob)
        rows.append(tr.collect())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res["model_head_tracking_pipelined"] = {"frames_per_s": round(F / dt, 1), "ms_per_frame": round(dt / F * 1e3, 3),
                                                "last_row_identical": bool(all((a == b) or (a != a and b != b) for a, b in zip(rows[-1], row)))}
    line = {"what": "video loop, %s MFCNet K=%d, 480x640, batch 1, ingest + %sstreaming forward + head + tracking"
                    % (model, K, "online RAFT-large flow (K-1 fields per frame, one batched call) + " if online else ""), "frames": F, **res}
    print(json.dumps(line))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(line, open("gpurun_out/bench_video_%s%s.json" % (model, "_online_flow" if online else ""), "w"), indent=1)


if __name__ == "__main__":
    main()
