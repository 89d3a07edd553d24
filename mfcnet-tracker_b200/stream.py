"""Sliding-window (streaming) MFCNet inference with a zero-copy feature ring buffer, and clip sharding.

The reference video loop (scripts/test_multiframe_segmentation_on_videos_v3.py:233-280) re-runs the
SFC network on all K frames of the window for every output frame, although K-1 of them were already
processed for the previous outputs (models/multiframe_model.py:426-428).  The SFC network is
per-frame and deterministic in eval mode, so here every frame is encoded ONCE: its class maps are
written into slot ``t % K`` of a ring of C8 planes, and the fusion head of step t reads the slots in
window order (current frame first) simply as a different list of source pointers -- the frame-window
channel concat is a rotation of pointers, not a copy.

``shard_frames`` splits a video of F frames into contiguous clips, one per GPU, each preceded by a
K-1-frame halo that only fills the ring (no collective on the inference path).
"""
import os

import torch

from . import engine
from .engine import Act, Ext


def shard_frames(n_frames, world, rank, num_input_frames):
    """Contiguous clip of output frames [lo, hi) for `rank`, and the first frame it must encode
    (`enc_lo` = lo - (K-1), clamped at 0).  Outputs exist for t >= K-1 only, as in the reference loop."""
    K = num_input_frames
    first_out = K - 1
    n_out = max(0, n_frames - first_out)
    per = (n_out + world - 1) // world
    lo = first_out + min(n_out, rank * per)
    hi = first_out + min(n_out, (rank + 1) * per)
    return {"lo": lo, "hi": hi, "enc_lo": max(0, lo - (K - 1)), "n_out": hi - lo}


def shard_clips(n_frames, world, rank, num_input_frames, clips_per_rank):
    """The `clips_per_rank` contiguous sub-clips of `rank`'s share of the video (each with its own K-1-frame halo), for
    StreamingMFCNet(batch=clips_per_rank): the clips of one GPU advance in lock step, one frame each per step, which turns the
    batch-1 sliding window (launch / latency bound on a B200) into a batch of B.  Equivalent to shard_frames over
    world * clips_per_rank virtual ranks; trailing clips may be shorter (or empty) and idle while the others finish."""
    return [shard_frames(n_frames, world * clips_per_rank, rank * clips_per_rank + j, num_input_frames) for j in range(clips_per_rank)]


class StreamingMFCNet:
    """Feeds frames one at a time through a `ResUNetMulti*`-style wrapper with SFC feature reuse.

    step(frame, flows, depths) takes the newest frame (B,3,H,W), the K-1 flows of the current window
    (flows[i-1]: current frame -> frame i earlier, as in the reference loop :264-271) and the K depth
    maps in window order; returns fp32 logits (B,N,H,W) once K frames have been seen, else None.

    `batch` = B independent clips advancing in lock step (e.g. the B clips one GPU owns when a long video is
    sharded, or B cameras): ONE SFC pass over the B new frames and ONE fusion pass over the B windows per step, the
    ring holding K x B class-map slots.  Per output frame that is 1 SFC + 1 fusion pass (67.7 GFLOP for ResUNet-16 /
    K=3) instead of the K SFC passes the reference spends, at a batch size that fills the GPU.
    """

    def __init__(self, model, H, W, device="cuda", dtype_name=None, batch=1):
        self.model = model
        self.K, self.N = model.num_frames, model.num_classes
        self.H, self.W = H, W
        self.B = int(batch)
        self.device = engine.canonical_device(device)
        self.dt = dtype_name or model.dtype_name or engine.default_dtype()
        self.t = 0
        self._build()

    def _build(self):
        m, K, N, H, W, dev, B = self.model, self.K, self.N, self.H, self.W, self.device, self.B
        packer = engine.WeightPacker(dev, self.dt)
        self._fingerprint = engine.params_fingerprint(m)
        tdtype = engine._DTYPES[self.dt][0]
        cin = m.base_model.channels
        self.ring = torch.zeros((K, B, (N + 7) // 8, H, W, 8), dtype=tdtype, device=dev)   # slot s holds frame t with t % K == s
        self.x_c8 = torch.empty((B, (cin + 7) // 8, H, W, 8), dtype=tdtype, device=dev)
        self.out = torch.empty((B, N, H, W), dtype=torch.float32, device=dev)
        # static fp32 input buffers: every program reads these (the caller's tensors are copied in per step), so that all
        # pointers of a step are fixed and the step can be replayed as one CUDA graph
        self.in_frame = torch.zeros((B, cin, H, W), dtype=torch.float32, device=dev)
        self.in_flow = [torch.zeros((B, 2, H, W), dtype=torch.float32, device=dev) for _ in range(K - 1)] if m.optflow_inputs else None
        self.in_depth = [torch.zeros((B, 1, H, W), dtype=torch.float32, device=dev) for _ in range(K)] if m.depth_inputs else None
        arena_sfc, arena_fus = engine.Arena(dev), engine.Arena(dev)
        self.sfc, self.fus = [], []
        for s in range(K):
            arena_sfc.reset()
            b = engine.Builder(dev, self.dt, packer, arena_sfc)
            ext = Ext("frame", self.in_frame)
            for c0 in range(0, cin, 8):
                b.prog.gather([(ext, c) for c in range(c0, min(c0 + 8, cin))], self.x_c8[:, c0 // 8], B, H, W)
            m.base_model.record(b, Act(self.x_c8, cin), maps_c8=self.ring[s])
            b.prog.finalize()
            self.sfc.append(b.prog)
        for s in range(K):  # s = slot of the current frame; frame i earlier lives in slot (s - i) % K
            arena_fus.reset()
            b = engine.Builder(dev, self.dt, packer, arena_fus)
            maps = [Act(self.ring[(s - i) % K], N) for i in range(K)]
            flows = [Ext(("flow", i), self.in_flow[i]) for i in range(K - 1)] if m.optflow_inputs else None
            depths = [Ext(("depth", i), self.in_depth[i]) for i in range(K)] if m.depth_inputs else None
            m.multiframe_net.record(b, maps, flows, depths, self.out)
            b.prog.finalize()
            self.fus.append(b.prog)
        self.launches_per_frame = self.sfc[0].n_kernels + self.fus[0].n_kernels
        # measured on B200: replaying a graph with parallel branches (HRNet's lanes) is slower than issuing the lanes on real
        # streams (109 vs 206 frames/s for HRNetMulti-Large K=5), single-lane programs are the same or slightly faster as a graph
        self.use_graphs = (os.environ.get("MFC_STREAM_GRAPH", "1") != "0" and dev.type == "cuda"
                           and not any(p.has_lanes for p in self.sfc + self.fus))
        self.graphs = [None] * K      # slot -> CUDA graph of (SFC of the new frame + fusion head), captured after one eager run
        self._ran_eager = [False] * K
        self._keep = (packer, arena_sfc, arena_fus)

    def reset(self):
        self.t = 0

    def _load(self, dst, src):
        dst.copy_(src, non_blocking=True)   # fp32 cast + layout fix-up + staging in one copy on the current stream

    def _rebuild_if_stale(self):
        """New weights invalidate the plans AND the ring contents: the window is refilled before the next output."""
        if engine.params_fingerprint(self.model) != self._fingerprint:
            self._build()
            self.t = 0

    def encode(self, frame):
        """Run the SFC network on one frame into its ring slot (used alone for the shard halo)."""
        self._rebuild_if_stale()
        s = self.t % self.K
        with engine.device_guard(self.device):
            self._load(self.in_frame, frame)
            self.sfc[s].run()
        self.t += 1
        return s

    def step(self, frame, flows=None, depths=None, out=None):
        self._rebuild_if_stale()
        if self.t < self.K - 1:
            self.encode(frame)
            return None
        s = self.t % self.K
        with engine.device_guard(self.device):
            self._load(self.in_frame, frame)
            if self.model.optflow_inputs:
                for i in range(self.K - 1):
                    self._load(self.in_flow[i], flows[i])
            if self.model.depth_inputs:
                for i in range(self.K):
                    self._load(self.in_depth[i], depths[i])
            if self.use_graphs and self.graphs[s] is None and self._ran_eager[s]:
                both = engine.Program(self.device, self.dt)
                both.extend(self.sfc[s])
                both.extend(self.fus[s])
                self.graphs[s] = both.capture()
            if self.graphs[s] is not None:
                self.graphs[s].launch()      # one launch: the whole frame (SFC + fusion head)
            else:
                self.sfc[s].run()
                self.fus[s].run()
                self._ran_eager[s] = True
            y = self.out.clone() if out is None else out.copy_(self.out)
        self.t += 1
        return y


class HostPipeline:
    """Double-buffered host -> device staging for streaming inference.

    The reference loop uploads every window synchronously before the forward
    (scripts/test_multiframe_segmentation_on_videos_v3.py:256-263: `.cuda()` on the frame, flow and depth
    tensors).  Here the H2D copies of step i+1 run on a dedicated copy stream while step i computes:
    `submit(host_tensors)` enqueues the copies of one step into the next free slot, `acquire()` makes the
    compute stream wait for the oldest submitted slot and returns its device tensors, `release()` marks the
    slot reusable once the work enqueued so far has consumed it.  Host tensors should be pinned."""

    def __init__(self, like, device="cuda", slots=2):
        self.device = engine.canonical_device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [[torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in like] for _ in range(slots)]
        self.copied = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [None] * slots
        self.head = self.tail = 0      # next slot to fill / next slot to hand out
        self.in_flight = 0

    def submit(self, host_tensors):
        if self.in_flight == len(self.slots):
            raise RuntimeError("HostPipeline: all slots are in flight; acquire()/release() one first")
        s = self.head
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[s] is not None:
                self.copy_stream.wait_event(self.consumed[s])   # the previous user of this slot is done
            for d, h in zip(self.slots[s], host_tensors):
                d.copy_(h, non_blocking=True)
            self.copied[s].record(self.copy_stream)
        self.head = (s + 1) % len(self.slots)
        self.in_flight += 1

    def acquire(self):
        if self.in_flight == 0:
            raise RuntimeError("HostPipeline: nothing submitted")
        s = self.tail
        torch.cuda.current_stream(self.device).wait_event(self.copied[s])
        return self.slots[s]

    def release(self):
        s = self.tail
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[s] = ev
        self.tail = (s + 1) % len(self.slots)
        self.in_flight -= 1
