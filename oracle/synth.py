"""Deterministic synthetic tensors, bit-identical on every platform.

TEST INFRASTRUCTURE (see oracle/__init__.py).  No transcendental functions and
no library RNG streams are used, so the container that writes the golden
fixtures and the GPU box that replays them produce the same bits.

Values are a 4-term Irwin-Hall approximation of a unit normal built from a
splitmix64 counter hash: exact integer arithmetic, then a handful of exactly
rounded float64 operations.
"""
import zlib

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _key(name, seed):
    return np.uint64((zlib.crc32(name.encode()) << 16) ^ (seed * 0x1000193 + 12345))


def uniform(name, shape, seed=0, lo=0.0, hi=1.0):
    """U[lo,hi) float32 tensor keyed by (name, seed)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) * np.uint64(4) + (_key(name, seed) << np.uint64(20))
        bits = _splitmix64(ctr)
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def normal(name, shape, seed=0, std=1.0, mean=0.0):
    """Approximately N(mean, std^2) float32 tensor keyed by (name, seed)."""
    n = int(np.prod(shape)) if len(shape) else 1
    acc = np.zeros(n, dtype=np.float64)
    with np.errstate(over="ignore"):
        base = np.arange(n, dtype=np.uint64) * np.uint64(4) + (_key(name, seed) << np.uint64(20))
        for j in range(4):
            bits = _splitmix64(base + np.uint64(j))
            acc += (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    z = (acc - 2.0) * 1.7320508075688772  # var of 4 uniforms = 1/3
    return (mean + std * z).astype(np.float32).reshape(shape)


def fill_state_dict(manifest, seed=0):
    """Build a state dict {name: np.ndarray} for a manifest [(name, shape, dtype)].

    Rules (chosen so every fold path is exercised):
      *.running_mean       U(-0.5, 0.5)
      *.running_var        U(0.5, 1.5)
      *.num_batches_tracked  int64 zero
      grid                 the (1,2,576,720) normalised mesh of
                           models/multiframe_model.py:172-185
      norm / bn weight     1 + 0.1 N(0,1)      (1-D '.weight' tensors)
      1-D bias             0.1 N(0,1)
      conv weight (4-D)    N(0, 1/fan_in)  scaled by 1.0 (He-like without gain)
    """
    out = {}
    for name, shape, dtype in manifest:
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = np.zeros(shape, dtype=np.int64)
        elif leaf == "running_mean":
            out[name] = uniform(name, shape, seed, -0.5, 0.5)
        elif leaf == "running_var":
            out[name] = uniform(name, shape, seed, 0.5, 1.5)
        elif leaf == "grid":
            out[name] = mesh_grid_576x720()
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            out[name] = normal(name, shape, seed, std=float(1.0 / np.sqrt(fan_in)))
        elif len(shape) == 1 and leaf == "weight":
            out[name] = normal(name, shape, seed, std=0.1, mean=1.0)
        else:
            out[name] = normal(name, shape, seed, std=0.1)
        assert out[name].shape == shape, (name, shape)
    return out


_TERNAUS_ALIAS = {16: [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]], 11: [[0], [3], [6, 8], [11, 13], [16, 18]]}


def apply_fixture_rules(sd, rules):
    """Post-processing of a generated state dict, recorded in a fixture's meta["scale_keys"]:
      {key: factor}            scale one tensor
      {"__all_4d__": factor}   scale every 4-D (conv / transposed-conv) weight
      {"__alias__": "ternaus"} TernausNet registers the VGG convs twice (encoder.N and convK.M are the SAME
                               Parameter, models/ternausnet.py:63-67,114-118): copy encoder.N over convK.M"""
    for k, f in (rules or {}).items():
        if k == "__all_4d__":
            for name, v in sd.items():
                if v.ndim == 4:
                    sd[name] = (v * np.float32(f)).astype(np.float32)
        elif k == "__alias__":
            prefixes = sorted({name[: name.index("encoder.")] for name in sd if "encoder." in name})
            for pre in prefixes:
                depth = 16 if pre + "encoder.28.weight" in sd else 11
                for si, idxs in enumerate(_TERNAUS_ALIAS[depth]):
                    for pos, j in enumerate(idxs):
                        for leaf in ("weight", "bias"):
                            sd["%sconv%d.%d.%s" % (pre, si + 1, 2 * pos, leaf)] = sd["%sencoder.%d.%s" % (pre, j, leaf)]
        else:
            sd[k] = (sd[k] * np.float32(f)).astype(np.float32)
    return sd


def mesh_grid_576x720():
    """`MultiFrameNetBasic._create_mesh_grid` (models/multiframe_model.py:172-185):
    x,y in [-1,1] for a fixed 576x720 image, stacked (x, y), float32.
    The reference computes 2.0*idx/(N-1)-1.0 on int64 tensors in float32."""
    H, W = 576, 720
    ys = (np.float32(2.0) * np.arange(H, dtype=np.float32) / np.float32(H - 1) - np.float32(1.0)).astype(np.float32)
    xs = (np.float32(2.0) * np.arange(W, dtype=np.float32) / np.float32(W - 1) - np.float32(1.0)).astype(np.float32)
    gy = np.repeat(ys[:, None], W, axis=1)
    gx = np.repeat(xs[None, :], H, axis=0)
    return np.stack([gx, gy], 0)[None].astype(np.float32)


def frames(tag, B, H, W, seed=0):
    """ImageNet-normalised-like RGB frames."""
    return normal("frame/" + tag, (B, 3, H, W), seed)


def smooth(name, shape, seed=0, cell=32, std=1.0, noise=0.02):
    """Spatially coherent field: a coarse N(0,1) grid (one value per `cell` pixels) bilinearly interpolated -- weights are
    multiples of 1/cell, all arithmetic exactly rounded float64 -- plus `noise` * N(0,1) per pixel.  Real video frames are
    dominated by low spatial frequencies; white-noise frames are the worst case for argmax agreement (every pixel is a
    potential near-tie), these are the realistic one."""
    *lead, H, W = shape
    gh, gw = H // cell + 2, W // cell + 2
    coarse = normal(name + "/coarse", tuple(lead) + (gh, gw), seed).astype(np.float64)
    ys, xs = np.arange(H), np.arange(W)
    y0, x0 = ys // cell, xs // cell
    fy = ((ys % cell).astype(np.float64) / cell)[:, None]
    fx = ((xs % cell).astype(np.float64) / cell)[None, :]
    c00 = coarse[..., y0[:, None], x0[None, :]]
    c01 = coarse[..., y0[:, None], x0[None, :] + 1]
    c10 = coarse[..., y0[:, None] + 1, x0[None, :]]
    c11 = coarse[..., y0[:, None] + 1, x0[None, :] + 1]
    f = (1 - fy) * ((1 - fx) * c00 + fx * c01) + fy * ((1 - fx) * c10 + fx * c11)
    f = std * (1.5 * f) + noise * normal(name + "/noise", shape, seed).astype(np.float64)
    return f.astype(np.float32)


def smooth_frames(tag, B, H, W, seed=0):
    """Realistic-margin variant of `frames`: same scale, low-frequency content."""
    return smooth("sframe/" + tag, (B, 3, H, W), seed)


def smooth_flow(tag, B, H, W, seed=0, scale=4.0):
    return smooth("sflow/" + tag, (B, 2, H, W), seed, cell=64, std=scale, noise=0.05)


def smooth_depth(tag, B, H, W, seed=0):
    d = 0.5 + 0.25 * smooth("sdepth/" + tag, (B, 1, H, W), seed, cell=64, noise=0.01)
    return np.clip(d, 0.0, 1.0).astype(np.float32)


def depth(tag, B, H, W, seed=0):
    return uniform("depth/" + tag, (B, 1, H, W), seed)


def flow(tag, B, H, W, seed=0, scale=4.0):
    return normal("flow/" + tag, (B, 2, H, W), seed, std=scale)


def loss_case(tag, B, N, H, W, seed=0, fg=0.05):
    """Model output (B,N,H,W) ~ 2*N(0,1) and an int64 target map with ~`fg` foreground pixels spread over the
    N-1 tool classes (SURVEY.md section 8d, config 5 inputs)."""
    out = normal("loss/out/" + tag, (B, N, H, W), seed, std=2.0)
    u = uniform("loss/tgt/" + tag, (B, H, W), seed)
    t = np.zeros((B, H, W), dtype=np.int64)
    for c in range(1, N):
        t[(u >= (c - 1) * fg / (N - 1)) & (u < c * fg / (N - 1))] = c
    return out, t
