"""Host-side plan builder: turns a module's layers into a command list for libmfcnet_b200.so.

PyTorch is used for device memory and the current stream only; every arithmetic step of the
forward is a kernel of the native library, issued by ONE ``mfc_run_list`` call per program.

Activations use the "C8" layout of include/mfcnet_b200.h: ``[B][ceil(C/8)][H][W][8]`` fp16/bf16.
A channel concat is a list of sources (zero-copy); GroupNorm+SiLU is a pending per-(sample,
channel) affine that the consuming conv applies while staging its input tile.
"""
import ctypes as C
import os
import threading

import torch

from . import abi

_DTYPES = {"fp16": (torch.float16, abi.MFC_F16), "bf16": (torch.bfloat16, abi.MFC_BF16)}


def default_dtype():
    """Activation / weight STORAGE type of the conv path; accumulation is always fp32 in TMEM.
    fp16 is the default: same width and tensor-core rate as bf16, but its 11-bit significand is what
    keeps the end-to-end logit error inside the 2e-2 bound of BASELINE.json (measured on B200:
    fp16 5.2e-3 / 99.93 % argmax agreement, bf16 4.2e-2 / 99.45 % on the 480x640 K=3 MFCNet; see
    DESIGN.md).  MFC_B200_DTYPE=bf16 selects bf16 storage."""
    return os.environ.get("MFC_B200_DTYPE", "fp16")


def autotune_enabled():
    """MFC_CONV_TUNE=1 measures the conv tilings of geometries that have no plan yet on the device at plan time
    (tools/tune_table.py uses it to write the committed table).  Default off: tilings come from the committed tuning
    table, else the cost model -- both deterministic, so every process computes the same bits."""
    return os.environ.get("MFC_CONV_TUNE", "0") == "1"


def prescale_factor(wmax):
    """fp16 has a short exponent: values below 2^-14 are subnormal (and below 2^-24 vanish).  A layer whose largest weight is
    tiny (a down-scaled head) is packed times this power of two and rescaled exactly in the epilogue."""
    import math
    if 0.0 < wmax < 2.0 ** -6:
        return 2.0 ** min(24, math.floor(-math.log2(wmax)))
    return 1.0


def hilo_enabled():
    """MFC_HILO=0: do not carry a network's last hidden activation / last weights as fp16 (hi, lo) pairs (measurement switch)."""
    return os.environ.get("MFC_HILO", "1") != "0"


def hilo_last_conv(bld, key, x, x_lo, w, bias, **kw):
    """The LAST 1x1 conv of a network with both its activation and its weights carried to ~22 bits through the fp16 tensor core:
    x = x_hi + x_lo (x_lo = the rounding residue the producer emitted), w = w_hi + w_lo, and
    w x ~ w_hi x_hi + w_hi x_lo + w_lo x_hi as ONE conv over the sources [x_hi, x_lo, x_hi] with the weights [w | w | w - fp16(w)]
    (the packer rounds each block to fp16: the first two become w_hi, the third w_lo).  The last layer's rounding noise goes
    straight into the logits -- nothing downstream averages it -- which is why it is worth three K-blocks instead of one."""
    tdtype = bld.tdtype
    wf = w.detach().float()
    p2 = prescale_factor(float(wf.abs().amax())) if bld.prog.dtype_name == "fp16" else 1.0   # what Builder.conv will pack with
    w_hi = (wf * p2).to(tdtype).float() / p2
    wcat = torch.cat([wf, wf, wf - w_hi], 1)
    return bld.conv(key, [x, x_lo, x], wcat, 1, bias=bias, **kw)


def head_fusion_enabled():
    """MFC_CONV_HEAD=0: keep a network's final 1x1 conv as its own launch (measurement switch)."""
    return os.environ.get("MFC_CONV_HEAD", "1") != "0"


def snake_enabled():
    return os.environ.get("MFC_CONV_SNAKE", "1") != "0"


def autotune_reps():
    return max(1, int(os.environ.get("MFC_CONV_TUNE_REPS", "3")))


_OVF = {}


def overflow_counter(device):
    """The per-device fp16 range-guard counter every recorded conv / fuse_sum / affine_silu_add points at
    (MfcConvIO.overflow): a kernel that stores a value beyond +-65504 into an fp16 activation tensor adds to it."""
    device = canonical_device(device)
    if device.type != "cuda" or abi.plan_only():
        return None
    if device not in _OVF:
        _OVF[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return _OVF[device]


def check_overflow(device="cuda"):
    """Synchronises and raises FloatingPointError if any fp16 activation overflowed on `device` since the last check.
    Every program checks itself after its first run; MFC_B200_CHECK_OVERFLOW=1 checks after every run (one sync per
    forward), and streaming / long-running callers can call this at their own cadence."""
    c = overflow_counter(device)
    if c is None:
        return 0
    n = int(c.item())
    if n:
        c.zero_()
        raise FloatingPointError(
            "mfcnet_tracker_b200: %d epilogue warp(s) stored activations beyond the fp16 range (+-65504 -> inf) on %s; "
            "the result is invalid.  Use bf16 storage (MFC_B200_DTYPE=bf16 / model.dtype_name = 'bf16') for this checkpoint."
            % (n, canonical_device(device)))
    return 0


def _check_every_run():
    return os.environ.get("MFC_B200_CHECK_OVERFLOW", "0") == "1"


def require_cuda(t, what):
    if not t.is_cuda and not abi.plan_only():
        raise RuntimeError("%s: mfcnet_tracker_b200 runs on a B200 only; got a %s tensor (there is no CPU fallback)"
                           % (what, t.device.type))


def canonical_device(device):
    """torch.device('cuda') and torch.device('cuda:0') compare unequal: always carry the index."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is None and torch.cuda.is_available():
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def device_guard(device):
    import contextlib
    return torch.cuda.device(device) if torch.device(device).type == "cuda" else contextlib.nullcontext()


def record_stream(t):
    """Inputs are referenced by raw pointer until the kernels run: tie them to the launch stream."""
    if t.is_cuda:
        t.record_stream(torch.cuda.current_stream(t.device))


class Ext:
    """An external fp32 NCHW input of a plan: a binding key + the placeholder tensor used at build time."""
    __slots__ = ("key", "t")

    def __init__(self, key, t):
        self.key, self.t = key, t

    @property
    def channels(self):
        return self.t.shape[1]


class Act:
    """A C8 activation: tensor [B, chunks, H, W, 8] + real channel count + optional pending affine."""
    __slots__ = ("t", "C", "affine", "lo")

    def __init__(self, t, C_, affine=None):
        self.t, self.C, self.affine, self.lo = t, C_, affine, None

    @property
    def B(self):
        return self.t.shape[0]

    @property
    def chunks(self):
        return self.t.shape[1]

    @property
    def H(self):
        return self.t.shape[2]

    @property
    def W(self):
        return self.t.shape[3]

    @property
    def bstride(self):  # bytes between samples
        return self.t.stride(0) * self.t.element_size()

    def with_affine(self, aff):
        return Act(self.t, self.C, aff)

    def batch_slice(self, lo, hi):
        return Act(self.t[lo:hi], self.C, None if self.affine is None else self.affine[lo:hi])


class Arena:
    """Replay allocator: the first program built after construction allocates device tensors;
    after `reset()` the same sequence of requests returns the SAME tensors, so programs built
    with an identical allocation sequence (the sub-batches of one forward, which run one after
    another on one stream) share their intermediate buffers."""

    def __init__(self, device):
        self.device = canonical_device(device)
        self.slots = []
        self.pos = 0

    def reset(self):
        self.pos = 0

    def alloc(self, shape, dtype, zero=False):
        shape = tuple(int(s) for s in shape)
        if self.pos < len(self.slots):
            t = self.slots[self.pos]
            if tuple(t.shape) != shape or t.dtype != dtype:
                raise RuntimeError("arena replay mismatch: %s/%s vs %s/%s" % (tuple(t.shape), t.dtype, shape, dtype))
        else:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self.slots.append(t)
        self.pos += 1
        return t

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.slots)


class Program:
    """A recorded list of native commands plus everything that must stay alive for them."""

    def __init__(self, device, dtype_name):
        self.lib = abi.load()
        self.device = canonical_device(device)
        self.dtype_name = dtype_name
        self.tdtype, self.cdtype = _DTYPES[dtype_name]
        self.cmds = []       # (op, struct_a, struct_b, lane)
        self.lane = 0        # lane (stream) the next recorded commands run on; see fork() / join()
        self._last_forward = True   # direction in which the previous bulk op walked the batch (snake order, see conv())
        self.keep = []       # tensors referenced by raw pointer
        self._array = None
        self.n_kernels = 0   # kernel launches one run() issues
        self.bindings = {}   # external input key -> [setter(tensor)]
        self.meta = []       # per command: {'kind', 'name', 'bytes', 'flops'} (algorithmic, for the roofline)
        self.packer = None   # WeightPacker whose device writes must have completed before the first run (Builder sets it)
        self._ovf = overflow_counter(self.device) if dtype_name == "fp16" else None
        self._checked = False
        self._lock = threading.Lock()   # a program's pointer slots are shared state: one caller binds + issues at a time

    # ---- low-level recording -----------------------------------------------------------------
    def _push(self, op, a, b=None, launches=1, meta=None):
        if op != abi.OP_CONV and op != abi.OP_GN_FINALIZE:
            self._last_forward = True      # every other kernel walks the batch front to back
        self.cmds.append((op, a, b, self.lane))
        self.meta.append(meta or {"kind": "op%d" % op, "name": "", "bytes": 0, "flops": 0})
        self._array = None
        self.n_kernels += launches

    def finalize(self):
        arr = (abi.MfcCmd * len(self.cmds))()
        for i, (op, a, b, lane) in enumerate(self.cmds):
            arr[i].op = op
            arr[i].lane = lane
            arr[i].a = C.cast(C.pointer(a), C.c_void_p) if a is not None else None
            arr[i].b = C.cast(C.pointer(b), C.c_void_p) if b is not None else None
        self._array = arr
        return self

    @property
    def has_lanes(self):
        return any(c[3] != 0 for c in self.cmds)

    def fork(self):
        """Commands recorded on lanes 1..3 after this point run concurrently with lane 0 (and with each other); they see
        everything recorded before the fork.  Every fork needs a join()."""
        self._push(abi.OP_FORK, None, launches=0, meta={"kind": "fork", "name": "", "bytes": 0, "flops": 0})

    def join(self):
        self.lane = 0
        self._push(abi.OP_JOIN, None, launches=0, meta={"kind": "join", "name": "", "bytes": 0, "flops": 0})

    def run(self, stream=None):
        if self._array is None:
            self.finalize()
        if self.packer is not None:
            self.packer.settle()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None
        abi.check(self.lib.mfc_run_list(self._array, len(self.cmds), stream))
        if self._ovf is not None and (not self._checked or _check_every_run()):
            self._checked = True
            check_overflow(self.device)

    def call(self, tensors, set_outputs=None, stream=None):
        """rebind(tensors) + set_outputs() + run() as one critical section: two threads (or two streams of one thread) that
        share a module -- and with it this cached program -- cannot interleave their pointer updates.  The kernels themselves
        are enqueued on the caller's stream; intermediate buffers of the plan are reused in stream order."""
        with self._lock:
            self.rebind(tensors)
            if set_outputs is not None:
                set_outputs()
            with device_guard(self.device):
                self.run(stream)

    def capture(self):
        """Capture this program into a CUDA graph (all of its pointers must be static from now on)."""
        if self._array is None:
            self.finalize()
        g = C.c_void_p()
        with device_guard(self.device):
            abi.check(self.lib.mfc_graph_capture(self._array, len(self.cmds), C.byref(g)))
        return Graph(self.lib, g, self)

    def run_timed(self, stream=None):
        """Measurement only: runs the list with CUDA events around every command and returns
        the per-command meta dicts with an added 'ms' (device milliseconds).  Synchronises."""
        if self._array is None:
            self.finalize()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        ms = (C.c_float * len(self.cmds))()
        abi.check(self.lib.mfc_run_list_timed(self._array, len(self.cmds), stream, ms))
        return [dict(self.meta[i], ms=float(ms[i])) for i in range(len(self.cmds))]

    def extend(self, other):
        self.packer = self.packer or other.packer
        self._ovf = self._ovf if self._ovf is not None else other._ovf
        self.cmds += other.cmds
        self.meta += other.meta
        self.keep += other.keep
        self.n_kernels += other.n_kernels
        self._array = None

    # ---- ops ---------------------------------------------------------------------------------
    def bind(self, key, setter):
        """Register `setter(tensor)` to re-point a raw pointer at the caller's tensor `key`."""
        self.bindings.setdefault(key, []).append(setter)

    def rebind(self, tensors):
        """tensors: {key: fp32 tensor with contiguous (C,H,W) inner dims}.  Must cover every key."""
        for key, setters in self.bindings.items():
            t = tensors[key]
            if t.dtype != torch.float32 or t.device != self.device:
                raise ValueError("input %r must be a float32 tensor on %s" % (key, self.device))
            if t.dim() != 4 or t.stride(3) != 1 or t.stride(2) != t.shape[3] or t.stride(1) != t.shape[2] * t.shape[3]:
                raise ValueError("input %r must have contiguous (C,H,W) dims" % (key,))
            for fn in setters:
                fn(t)

    def gather(self, planes, dst_chunk_tensor, B, H, W):
        """planes: up to 8 entries of Ext-channel (ext, channel) or None -> one C8 chunk, where
        ext = Ext(key, placeholder fp32 [B,Cx,H,W] tensor).  The pointers are re-pointed per call
        through `rebind`."""
        a = abi.MfcGatherArgs()
        for j in range(8):
            p = planes[j] if j < len(planes) else None
            if p is None:
                a.g.plane[j] = None
                a.g.plane_bstride[j] = 0
            else:
                ext, ch = p
                if tuple(ext.t.shape[0:1] + ext.t.shape[2:]) != (B, H, W):
                    raise ValueError("input %r has shape %s, expected batch %d and %dx%d" % (ext.key, tuple(ext.t.shape), B, H, W))

                def setter(t, a=a, j=j, ch=ch, HW=H * W):
                    a.g.plane[j] = t.data_ptr() + ch * HW * 4
                    a.g.plane_bstride[j] = t.stride(0)
                setter(ext.t)
                self.bind(ext.key, setter)
        a.dst = dst_chunk_tensor.data_ptr()
        a.dst_bstride_bytes = dst_chunk_tensor.stride(0) * dst_chunk_tensor.element_size()
        a.B, a.H, a.W, a.dtype = B, H, W, self.cdtype
        self.keep.append(dst_chunk_tensor)
        nreal = sum(1 for p in planes if p is not None)
        self._push(abi.OP_GATHER, a, meta={"kind": "gather", "name": "", "flops": 0,
                                           "bytes": B * H * W * (4 * nreal + 16)})
        return a

    def conv_desc(self, srcs, Cout, k, stride, pad, upsample, act, parity=None, pad_br=0, kw=None, pad_yx=None):
        """parity=(py, px): one output parity of ConvTranspose2d(4, 2, 1), a 2x2 conv whose result lands on
        the pixels (2y+py, 2x+px) of a [2H][2W] output (include/mfcnet_b200.h, MfcConvDesc)."""
        d = abi.MfcConvDesc()
        s0 = srcs[0]
        d.B, d.Hin, d.Win = s0.B, s0.H, s0.W
        Hup, Wup = s0.H * upsample, s0.W * upsample
        kw = k if kw is None else kw          # rectangular kernel k x kw with paddings pad_yx = (py, px): pad = max, in_off = pad - p
        if pad_yx is not None:
            pad = max(pad_yx)
            d.in_off_y, d.in_off_x = pad - pad_yx[0], pad - pad_yx[1]
        py, px = pad - d.in_off_y, pad - d.in_off_x
        d.Hout = (Hup + 2 * py + pad_br - k) // stride + 1
        d.Wout = (Wup + 2 * px + pad_br - kw) // stride + 1
        d.pad_br = pad_br
        if parity is not None:
            d.Hout, d.Wout = s0.H, s0.W
            d.in_off_y, d.in_off_x, d.out_stride, d.out_off_y, d.out_off_x = parity[0], parity[1], 2, parity[0], parity[1]
        d.Cout, d.kh, d.kw, d.stride, d.pad, d.upsample, d.act = Cout, k, kw, stride, pad, upsample, act
        d.dtype = self.cdtype
        if len(srcs) > abi.MFC_MAX_SRC:
            raise ValueError("too many concat sources")
        d.nsrc = len(srcs)
        for i, s in enumerate(srcs):
            if (s.B, s.H, s.W) != (s0.B, s0.H, s0.W):
                raise ValueError("concat sources disagree in shape")
            d.src[i].ptr = s.t.data_ptr()
            d.src[i].affine = abi.ptr(s.affine)
            d.src[i].batch_stride = s.bstride
            d.src[i].nchunks = s.chunks
        return d

    def query(self, d):
        info = abi.MfcConvInfo()
        abi.check(self.lib.mfc_conv2d_query(C.byref(d), C.byref(info)))
        return info

    def conv(self, d, info, srcs, packed, residual=None, want_stats=False, out_c8=True, out_nchw=None, arena=None,
             y_c8=None, name="", want_lo=False):
        """Record one fused conv for descriptor `d` (from conv_desc) / `info` (from query).
        `packed` = PackedConv (weights + scale/shift).  Returns (Act or None, stats or None, io)."""
        io = abi.MfcConvIO()
        io.w_packed = packed.w.data_ptr()
        io.scale = abi.ptr(packed.scale)
        io.shift = abi.ptr(packed.shift)
        self.keep += [packed.w, packed.scale, packed.shift]
        if residual is not None:
            io.residual = residual.t.data_ptr()
            io.res_affine = abi.ptr(residual.affine)
            io.res_batch_stride = residual.bstride
            self.keep += [residual.t, residual.affine]
        io.overflow = abi.ptr(self._ovf)
        out = None
        B, Cout = d.B, d.Cout
        os_ = 2 if d.out_stride == 2 else 1
        if out_c8:
            if y_c8 is None:
                y_c8 = arena.alloc((B, (Cout + 7) // 8, d.Hout * os_, d.Wout * os_, 8), self.tdtype)
            out = Act(y_c8, Cout)
            io.y_c8 = y_c8.data_ptr()
            io.y_batch_stride = out.bstride
            self.keep.append(y_c8)
            if want_lo:    # rounding residue of the stored values (MfcConvIO.y_lo): same shape / strides
                self.lo = arena.alloc(tuple(y_c8.shape), self.tdtype)
                if self.lo.stride() != y_c8.stride():
                    raise ValueError("y_lo needs the strides of y_c8")
                io.y_lo = self.lo.data_ptr()
                self.keep.append(self.lo)
        if out_nchw is not None:
            io.y_nchw = out_nchw.data_ptr()
            self.keep.append(out_nchw)
        stats = None
        if want_stats:
            stats = arena.alloc((B, info.stats_per_image, info.nb * info.nblk, 2), torch.float32)
            io.stats = stats.data_ptr()
            self.keep.append(stats)
        for s in srcs:
            self.keep += [s.t, s.affine]
        # algorithmic work of this layer as the reference conv sees it (DESIGN.md "roofline"): the
        # input tensor read once (at the resolution the conv consumes), the output written once, the
        # weights once; 2 bytes per activation element (fp32 NCHW outputs count 4).
        cin = sum(s.C for s in srcs)
        hin, win = d.Hin * d.upsample, d.Win * d.upsample
        nbytes = B * cin * hin * win * 2 + Cout * cin * d.kh * d.kw * 2
        if out_c8:
            nbytes += B * Cout * d.Hout * d.Wout * 2
        if out_nchw is not None:
            nbytes += B * Cout * d.Hout * d.Wout * 4
        if residual is not None:
            nbytes += B * Cout * d.Hout * d.Wout * 2
        flops = 2 * B * Cout * d.Hout * d.Wout * cin * d.kh * d.kw
        # snake order: a conv walks the batch in the direction opposite to its producer, so it starts with the part of its
        # input that is still in the L2 (a batch of activations is larger than the cache)
        if snake_enabled() and B > 1:
            if self._last_forward:
                d.reserved |= abi.MFC_CONV_REVERSE_ORDER
            self._last_forward = not self._last_forward
        self._push(abi.OP_CONV, d, io, meta={"kind": "conv", "name": name, "bytes": nbytes, "flops": flops,
                                             "shape": "B%d %dx%d %d->%d k%d s%d u%d" % (B, d.Hout, d.Wout, cin, Cout, d.kh, d.stride, d.upsample)})
        return out, stats, io

    def gn_finalize(self, stats, info, gamma, beta, C_, groups, pixels, eps, affine):
        a = abi.MfcGnArgs()
        a.stats, a.gamma, a.beta, a.affine = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), affine.data_ptr()
        a.pixels, a.B, a.stats_per_image = pixels, stats.shape[0], info.stats_per_image
        a.cpad, a.C, a.groups, a.eps = info.nb * info.nblk, C_, groups, eps
        self.keep += [stats, gamma, beta, affine]
        self._push(abi.OP_GN_FINALIZE, a, meta={"kind": "gn_finalize", "name": "", "flops": 0, "bytes": stats.numel() * 4})
        return affine

    def affine_silu_add(self, a_act, r_act, out_t):
        a = abi.MfcAddArgs()
        a.a, a.affine, a.r, a.out = a_act.t.data_ptr(), a_act.affine.data_ptr(), r_act.t.data_ptr(), out_t.data_ptr()
        a.pixels, a.B, a.chunks, a.dtype = a_act.H * a_act.W, a_act.B, a_act.chunks, self.cdtype
        a.overflow = abi.ptr(self._ovf)
        for t in (a_act.t, r_act.t, out_t):
            if not t.is_contiguous():
                raise ValueError("affine_silu_add needs dense C8 tensors")
        self.keep += [a_act.t, a_act.affine, r_act.t, out_t]
        self._push(abi.OP_AFFINE_SILU_ADD, a, meta={"kind": "affine_silu_add", "name": "", "flops": 0,
                                                    "bytes": 3 * a_act.B * a_act.C * a_act.H * a_act.W * 2})
        return Act(out_t, a_act.C)

    def fuse_sum(self, terms, out_t, C_, scale=None, shift=None, act=1, out_lo=None):
        """out = act(scale * sum(terms) + shift); terms: Acts at the output size or lower (bilinear
        align_corners=False upsampling on read).  Returns the output Act."""
        a = abi.MfcFuseArgs()
        B, chunks, H, W = out_t.shape[0], out_t.shape[1], out_t.shape[2], out_t.shape[3]
        a.B, a.chunks, a.H, a.W, a.nterms, a.act, a.dtype = B, chunks, H, W, len(terms), act, self.cdtype
        nbytes = B * C_ * H * W * 2
        for j, t in enumerate(terms):
            if t.affine is not None or t.chunks != chunks or t.B != B:
                raise ValueError("fuse_sum terms must be materialised C8 tensors with the output's channels")
            a.term[j].ptr, a.term[j].batch_stride, a.term[j].H, a.term[j].W = t.t.data_ptr(), t.bstride, t.H, t.W
            self.keep.append(t.t)
            nbytes += B * C_ * t.H * t.W * 2
        a.scale, a.shift = abi.ptr(scale), abi.ptr(shift)
        a.out, a.out_batch_stride = out_t.data_ptr(), out_t.stride(0) * out_t.element_size()
        a.overflow = abi.ptr(self._ovf)
        if out_lo is not None:
            if out_lo.shape != out_t.shape or out_lo.stride() != out_t.stride():
                raise ValueError("fuse_sum: out_lo must match out")
            a.out_lo = out_lo.data_ptr()
            nbytes += B * C_ * H * W * 2
        self.keep += [out_t, scale, shift, out_lo]
        self._push(abi.OP_FUSE_SUM, a, meta={"kind": "fuse_sum", "name": "", "flops": 0, "bytes": nbytes})
        return Act(out_t, C_)

    def resize(self, src_nchw, Hout, Wout, dst_nchw=None, dst_c8=None):
        """Bilinear (align_corners=False) resize of an fp32 NCHW tensor into fp32 NCHW and/or C8."""
        a = abi.MfcResizeArgs()
        B, C_, Hin, Win = src_nchw.shape
        a.src, a.dst_nchw, a.dst_c8 = src_nchw.data_ptr(), abi.ptr(dst_nchw), abi.ptr(dst_c8)
        a.c8_batch_stride = 0 if dst_c8 is None else dst_c8.stride(0) * dst_c8.element_size()
        a.B, a.C, a.Hin, a.Win, a.Hout, a.Wout, a.dtype = B, C_, Hin, Win, Hout, Wout, self.cdtype
        self.keep += [src_nchw, dst_nchw, dst_c8]
        nbytes = B * C_ * (Hin * Win * 4 + Hout * Wout * ((4 if dst_nchw is not None else 0) + (2 if dst_c8 is not None else 0)))
        self._push(abi.OP_RESIZE, a, meta={"kind": "resize", "name": "", "flops": 0, "bytes": nbytes})
        return a

    def maxpool2(self, x, out_t):
        """nn.MaxPool2d(2, 2) on a materialised C8 Act."""
        if x.affine is not None:
            raise ValueError("maxpool2 needs a materialised tensor")
        a = abi.MfcPoolArgs()
        a.src, a.dst = x.t.data_ptr(), out_t.data_ptr()
        a.src_bstride_bytes, a.dst_bstride_bytes = x.bstride, out_t.stride(0) * out_t.element_size()
        a.B, a.chunks, a.H, a.W, a.dtype = x.B, x.chunks, x.H, x.W, self.cdtype
        self.keep += [x.t, out_t]
        self._push(abi.OP_MAXPOOL2, a, meta={"kind": "maxpool2", "name": "", "flops": 0, "bytes": int(x.B * x.C * x.H * x.W * 2 * 1.25)})
        return Act(out_t, x.C)

    def heatmap(self, logits, logp=None, prob=None, argmax=None):
        """log-softmax / softmax / argmax over the channel axis of fp32 NCHW logits, as a list command."""
        a = abi.MfcHeatmapArgs()
        B, N = logits.shape[0], logits.shape[1]
        a.logits, a.logp, a.prob, a.argmax = logits.data_ptr(), abi.ptr(logp), abi.ptr(prob), abi.ptr(argmax)
        a.pixels, a.B, a.N = logits.shape[2] * logits.shape[3], B, N
        self.keep += [logits, logp, prob, argmax]
        self._push(abi.OP_HEATMAP, a, meta={"kind": "heatmap_head", "name": "", "flops": 0, "bytes": B * N * a.pixels * 8})
        return a

    def pointwise(self, kind, a, out, a_aff=None, r=None, r_aff=None, out2=None, relu_a=False, relu_out=False, chunks=None):
        """Element-wise glue on dense C8 tensors (MfcPointwiseArgs; kinds abi.PW_*).  `chunks` = planes of the OUTPUT."""
        g = abi.MfcPointwiseArgs()
        g.a, g.a_aff, g.r, g.r_aff = a.data_ptr(), abi.ptr(a_aff), abi.ptr(r), abi.ptr(r_aff)
        g.out, g.out2 = out.data_ptr(), abi.ptr(out2)
        g.B, g.chunks, g.pixels = out.shape[0], (chunks or out.shape[1]), out.shape[2] * out.shape[3]
        g.kind, g.dtype, g.relu_a, g.relu_out = kind, self.cdtype, int(relu_a), int(relu_out)
        for t in (a, r, out, out2):
            if t is not None and not t.is_contiguous():
                raise ValueError("pointwise needs dense C8 tensors")
        self.keep += [t for t in (a, a_aff, r, r_aff, out, out2) if t is not None]
        self._push(abi.OP_POINTWISE, g, meta={"kind": "pointwise", "name": "", "flops": 0, "bytes": 2 * out.numel() * out.element_size()})

    def raft(self, kind, ptrs, B, h, w, C_=0, levels=0, radius=0, scale=0.0):
        """One RAFT piece (MfcRaftArgs; kinds abi.RAFT_*); `ptrs` = up to six tensors p0..p5 (None = NULL)."""
        g = abi.MfcRaftArgs()
        ptrs = list(ptrs) + [None] * (6 - len(ptrs))
        g.p0, g.p1, g.p2, g.p3, g.p4, g.p5 = [abi.ptr(t) for t in ptrs]
        g.kind, g.B, g.C, g.h, g.w, g.levels, g.radius, g.dtype, g.scale = kind, B, C_, h, w, levels, radius, self.cdtype, scale
        self.keep += [t for t in ptrs if t is not None]
        self._push(abi.OP_RAFT, g, meta={"kind": "raft%d" % kind, "name": "", "flops": 0, "bytes": 0})

    def warp(self, args, nbytes=0):
        self._push(abi.OP_WARP, args, meta={"kind": "flow_warp", "name": "", "flops": 0, "bytes": nbytes})


class Graph:
    """An instantiated CUDA graph of one Program (keeps the program, hence every buffer it points to, alive)."""

    def __init__(self, lib, handle, prog):
        self.lib, self.handle, self.prog = lib, handle, prog

    def launch(self, stream=None):
        if stream is None:
            stream = torch.cuda.current_stream(self.prog.device).cuda_stream
        abi.check(self.lib.mfc_graph_launch(self.handle, stream))

    def __del__(self):
        try:
            if self.handle:
                self.lib.mfc_graph_destroy(self.handle)
        except Exception:
            pass


class PackedConv:
    __slots__ = ("w", "scale", "shift")

    def __init__(self, w, scale, shift):
        self.w, self.scale, self.shift = w, scale, shift


def chan_map_for(srcs_channels):
    """Padded-concat channel -> weight Cin index.  srcs_channels: list of (real_channels, chunks,
    first_weight_channel)."""
    m = []
    for real, chunks, first in srcs_channels:
        for j in range(chunks * 8):
            m.append(first + j if j < real else -1)
    return m


class WeightPacker:
    """Derives the packed / folded device tensors a plan needs from a module's raw parameters.
    Everything here runs native kernels (weight standardisation, BN fold, bf16/fp16 packing) on
    the current stream; results are cached per (layer, geometry) until the weights change."""

    def __init__(self, device, dtype_name):
        self.lib = abi.load()
        self.device = canonical_device(device)
        self.tdtype, self.cdtype = _DTYPES[dtype_name]
        self.cache = {}

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def standardized(self, key, w, eps=1e-5):
        k = ("ws", key)
        if k not in self.cache:
            w = w.detach().contiguous().float()
            out = torch.empty_like(w)
            abi.check(self.lib.mfc_weight_standardize(w.data_ptr(), out.data_ptr(), w.shape[0], w[0].numel(), eps, self._stream()))
            self.cache[k] = out
        return self.cache[k]

    def bn_affine(self, key, bn_w, bn_b, mean, var, eps, conv_bias=None):
        k = ("bn", key)
        if k not in self.cache:
            Cc = bn_w.numel()
            scale = torch.empty(Cc, dtype=torch.float32, device=self.device)
            shift = torch.empty(Cc, dtype=torch.float32, device=self.device)
            f = lambda t: t.detach().contiguous().float()
            args = [f(bn_w), f(bn_b), f(mean), f(var)]
            cb = f(conv_bias) if conv_bias is not None else None
            abi.check(self.lib.mfc_bn_fold(args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(),
                                           abi.ptr(cb), eps, scale.data_ptr(), shift.data_ptr(), Cc, self._stream()))
            self.cache[k] = (scale, shift, args, cb)
        return self.cache[k][0], self.cache[k][1]

    def pack(self, key, prog, desc, info, w_oihw, cmap, scale=None, shift=None, post_scale=None):
        """w_oihw fp32 [Cout][Cin_w][kh][kw] device tensor -> PackedConv for this geometry.  `scale` is multiplied INTO the
        weights; `post_scale` (a power of two or None) is applied to the accumulators in the epilogue."""
        k = ("pk", key, info.nb, info.nblk, info.ksteps, info.weight_layout, tuple(cmap) if cmap is not None else None)
        if k not in self.cache:
            w = w_oihw.detach().contiguous().float()
            packed = torch.empty(int(info.packed_weight_bytes), dtype=torch.uint8, device=self.device)
            cm = None
            if cmap is not None:
                cm = torch.tensor(cmap, dtype=torch.int32, device=self.device)
            # a per-channel scale (folded BatchNorm) goes INTO the weights: the epilogue only adds the shift
            scf = scale.detach().contiguous().float() if scale is not None else None
            abi.check(self.lib.mfc_conv2d_pack_weights(C.byref(desc), w.data_ptr(), w.shape[1], abi.ptr(cm), abi.ptr(scf),
                                                       packed.data_ptr(), self._stream()))
            cpad = info.nb * info.nblk
            sc = sh = None
            if shift is not None:
                sh = torch.zeros(cpad, dtype=torch.float32, device=self.device)
                sh[: shift.numel()] = shift.detach().float()
            if post_scale is not None:
                sc = torch.full((cpad,), float(post_scale), dtype=torch.float32, device=self.device)
            self.cache[k] = (PackedConv(packed, sc, sh), w, cm, scf)
            # conv_tc_kernel stages weights / scale / shift before its griddepcontrol.wait (they overlap the previous
            # kernel's tail): everything written here must be complete before the first conv launch that uses it
            self._dirty = True
        return self.cache[k][0]

    def settle(self):
        """Synchronise once after (re)packing: see pack()."""
        if getattr(self, "_dirty", False) and self.device.type == "cuda" and not abi.plan_only():
            torch.cuda.current_stream(self.device).synchronize()
        self._dirty = False


def params_fingerprint(module):
    """Cheap change detector for a module's parameters and buffers (in-place updates bump
    `_version`; `.to()` / `load_state_dict(assign=True)` change `data_ptr`)."""
    h = 0
    for t in list(module.parameters()) + list(module.buffers()):
        h = (h * 1000003 + t.data_ptr() * 31 + t._version) & 0xFFFFFFFFFFFFFFF
    return h


class Builder:
    """Program + weight packer + arena: the object module plans are written against."""

    def __init__(self, device, dtype_name, packer, arena):
        self.prog = Program(device, dtype_name)
        self.prog.packer = packer
        self.packer = packer
        self.arena = arena
        self.device = self.prog.device
        self.tdtype = self.prog.tdtype

    def conv(self, key, srcs, w_oihw, k, *, bias=None, scale=None, shift=None, stride=1, pad=0, upsample=1, act=0,
             residual=None, want_stats=False, out_c8=True, out_nchw=None, y_c8=None, first_weight_channel=None, parity=None,
             head=None, want_lo=False, pad_br=0, kw=None, pad_yx=None):
        """srcs: list of Act (channel concat in order).  w_oihw: fp32 device weight whose Cin axis
        is the concat of the sources' REAL channels (or, with first_weight_channel=[...], starts
        at the given offsets).  head=(w [Nh, Cout] fp32, bias [Nh] or None): a following 1x1 conv evaluated in fp32 inside this
        conv's epilogue (MfcConvIO.head_w; Cout <= 16), `out_nchw` then has Nh channels.  Returns (Act|None, stats|None, info, io)."""
        Cout = w_oihw.shape[0]
        # A halo-free stride-1 conv (1x1) does not care about the 2-D geometry of its planes: present every plane as an
        # image 128 pixels wide (H*W/128 rows).  Tiles then span the full width, so a tile of R rows IS R MMA runs of 128
        # consecutive pixels, one contiguous TMA box per plane (R*2 KB) instead of TH short row segments, and the epilogue's
        # pixel addresses are affine in the run index.  Pure re-interpretation: same bytes, same arithmetic.
        H0, W0 = srcs[0].H, srcs[0].W
        flat = (k == 1 and kw in (None, 1) and stride == 1 and upsample == 1 and pad == 0 and parity is None and W0 != 128 and (H0 * W0) % 128 == 0
                and os.environ.get("MFC_CONV_FLAT", "1") != "0")
        if flat:
            rows = H0 * W0 // 128

            def fl(a):
                return None if a is None else Act(a.t.view(a.t.shape[0], a.t.shape[1], rows, 128, 8), a.C, a.affine)
            srcs = [fl(a) for a in srcs]
            residual = fl(residual)
            if y_c8 is not None:
                y_c8 = y_c8.view(y_c8.shape[0], y_c8.shape[1], rows, 128, 8)
            if out_nchw is not None:
                out_nchw = out_nchw.view(out_nchw.shape[0], out_nchw.shape[1], rows, 128)
        d = self.prog.conv_desc(srcs, Cout, k, stride, pad, upsample, act, parity=parity, pad_br=pad_br, kw=kw, pad_yx=pad_yx)
        if residual is not None:
            d.reserved |= abi.MFC_CONV_HAS_RESIDUAL
        if want_stats:
            d.reserved |= abi.MFC_CONV_WANT_STATS
        if head is not None:
            d.reserved |= abi.MFC_CONV_WANT_HEAD
        layout = []
        off = 0
        for i, s in enumerate(srcs):
            first = off if first_weight_channel is None else first_weight_channel[i]
            layout.append((s.C, s.chunks, first))
            off = first + s.C
        identity = len(srcs) == 1 and srcs[0].C == w_oihw.shape[1] and layout[0][2] == 0
        cmap = None if identity else chan_map_for(layout)
        if shift is None and bias is not None:
            shift = bias
        # tiny-weight layers are packed times a power of two and rescaled exactly in the epilogue (prescale_factor)
        post_scale = None
        if self.prog.dtype_name == "fp16":
            wmax = float(w_oihw.detach().abs().amax()) * (float(scale.detach().abs().amax()) if scale is not None else 1.0)
            p2 = prescale_factor(wmax)
            if p2 != 1.0:
                scale = (scale.detach().float() * p2) if scale is not None else torch.full((Cout,), p2, dtype=torch.float32, device=self.device)
                post_scale = 1.0 / p2
        if autotune_enabled() and self.device.type == "cuda" and not abi.plan_only():
            self._autotune(d, srcs, w_oihw, cmap, False, shift is not None, residual, want_stats, out_c8, out_nchw, y_c8,
                           None if head is None else head[0].shape[0])
        info = self.prog.query(d)
        packed = self.packer.pack(key, self.prog, d, info, w_oihw, cmap, scale, shift, post_scale)
        out, stats, io = self.prog.conv(d, info, srcs, packed, residual=residual, want_stats=want_stats, out_c8=out_c8,
                                        out_nchw=out_nchw, arena=self.arena, y_c8=y_c8, name=key, want_lo=want_lo)
        if want_lo:        # [out, out.lo] as two sources with the same weights = the activation to ~22 bits
            out.lo = Act(self.prog.lo.view(out.t.shape), out.C)
        if head is not None:
            hw, hb = head
            nh = hw.shape[0]
            w16 = torch.zeros((nh, 16), dtype=torch.float32, device=self.device)
            w16[:, :Cout] = hw.detach().float().reshape(nh, Cout)
            hbf = hb.detach().float().contiguous() if hb is not None else None
            io.head_w, io.head_b, io.head_n = w16.data_ptr(), abi.ptr(hbf), nh
            self.prog.keep += [w16, hbf]
            self.packer._dirty = True          # read before griddepcontrol.wait, like the packed weights
            m = self.prog.meta[-1]
            m["flops"] += 2 * d.B * d.Hout * d.Wout * Cout * nh
            m["bytes"] += d.B * d.Hout * d.Wout * (4 * nh - 4 * Cout)   # the fp32 map written has nh channels, not Cout
        if flat and out is not None:
            lo = out.lo
            out = Act(out.t.view(out.t.shape[0], out.t.shape[1], H0, W0, 8), out.C, out.affine)
            if lo is not None:
                out.lo = Act(lo.t.view(out.t.shape), lo.C)
        return out, stats, info, io

    def _autotune(self, d, srcs, w_oihw, cmap, has_scale, has_shift, residual, want_stats, out_c8, out_nchw, y_c8, head_n=None):
        """Plan-time measurement of the candidate tilings of this conv on the device, with buffers of the real sizes and
        the real epilogue mode (mfc_conv2d_autotune keeps the fastest; temporaries are released afterwards)."""
        lib = self.prog.lib
        need = int(lib.mfc_conv2d_autotune_scratch_bytes(C.byref(d)))   # NOT mfc_conv2d_query: that would freeze the plan
        dev = self.device
        cpad = d.Cout + 256      # upper bound over the N-block widths the tuner may try
        io = abi.MfcConvIO()
        keep = []

        def tmp(shape, dtype, fill=None):
            t = torch.empty(shape, dtype=dtype, device=dev) if fill is None else torch.full(shape, fill, dtype=dtype, device=dev)
            keep.append(t)
            return t

        os_ = 2 if d.out_stride == 2 else 1
        if has_scale:
            io.scale = tmp((cpad,), torch.float32, 1.0).data_ptr()
        if has_shift:
            io.shift = tmp((cpad,), torch.float32, 0.0).data_ptr()
        if residual is not None:
            io.residual = residual.t.data_ptr()
            io.res_affine = abi.ptr(residual.affine)
            io.res_batch_stride = residual.bstride
        if out_c8:
            y = y_c8 if y_c8 is not None else tmp((d.B, (d.Cout + 7) // 8, d.Hout * os_, d.Wout * os_, 8), self.tdtype)
            io.y_c8 = y.data_ptr()
            io.y_batch_stride = y.stride(0) * y.element_size()
        if out_nchw is not None:
            io.y_nchw = out_nchw.data_ptr()
        if head_n is not None:
            io.head_w, io.head_n = tmp((head_n, 16), torch.float32, 0.0).data_ptr(), head_n
        if want_stats:
            io.stats = tmp((d.B, 148, cpad, 2), torch.float32).data_ptr()
        w = w_oihw.detach().contiguous().float()
        cm = torch.tensor(cmap, dtype=torch.int32, device=dev) if cmap is not None else None
        scratch = tmp((need + 4096,), torch.uint8)
        stream = torch.cuda.current_stream(dev).cuda_stream
        abi.check(lib.mfc_conv2d_autotune(C.byref(d), C.byref(io), w.data_ptr(), w.shape[1], abi.ptr(cm), scratch.data_ptr(),
                                          scratch.numel(), autotune_reps(), stream))
        torch.cuda.current_stream(dev).synchronize()
        del keep

    def group_norm_affine(self, stats, info, gamma, beta, C_, groups, pixels, eps=1e-5):
        """Finalise GroupNorm statistics into the pending affine of the producing conv's output."""
        B = stats.shape[0]
        affine = self.arena.alloc((B, ((C_ + 7) // 8) * 8, 2), torch.float32, zero=True)
        return self.prog.gn_finalize(stats, info, gamma.detach(), beta.detach(), C_, groups, pixels, eps, affine)

    def gather_channels(self, exts, B, H, W):
        """Ext inputs (all their channels, in order) -> one tightly packed C8 Act."""
        planes = [(e, c) for e in exts for c in range(e.channels)]
        chunks = (len(planes) + 7) // 8
        dst = self.arena.alloc((B, chunks, H, W, 8), self.tdtype)
        for q in range(chunks):
            self.prog.gather(planes[q * 8:(q + 1) * 8], dst[:, q], B, H, W)
        return Act(dst, len(planes))
