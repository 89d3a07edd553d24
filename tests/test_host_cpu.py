"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU), the drop-in modules expose the reference's state_dict keys, and
plans are constructible (MFC_B200_PLAN_ONLY: real descriptor validation / tiling queries, no launches)."""
import ctypes as C
import os
import re

import pytest
import torch

from tests import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def M(monkeypatch):
    monkeypatch.setenv("MFC_B200_PLAN_ONLY", "1")
    import mfcnet_tracker_b200 as m
    monkeypatch.setattr(m.abi, "_lib", None)
    yield m
    m.abi._lib = None


def _declared():
    with open(os.path.join(ROOT, "include", "mfcnet_b200.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"^(?:int|long long|const char\*)\s+(mfc_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import mfcnet_tracker_b200 as m
    lib = C.CDLL(m.abi.library_path())
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(m.abi.EXPORTS) == names          # the ctypes binding covers the whole header
    lib.mfc_abi_version.restype = C.c_int
    assert lib.mfc_abi_version() == 5


def test_struct_sizes_match_header():
    """sizeof() of every ctypes mirror equals the C compiler's (checked through a tiny gcc probe)."""
    import shutil
    import subprocess
    import tempfile
    import mfcnet_tracker_b200 as m
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    names = ["MfcGather", "MfcConvInfo", "MfcSrc", "MfcConvDesc", "MfcConvIO", "MfcWarpArgs", "MfcGnArgs", "MfcAddArgs",
             "MfcFuseTerm", "MfcFuseArgs", "MfcResizeArgs", "MfcPoolArgs", "MfcHeatmapArgs",
             "MfcGatherArgs", "MfcCmd", "MfcPointwiseArgs", "MfcRaftArgs"]
    prog = '#include <stdio.h>\n#include "mfcnet_b200.h"\nint main(){' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        with open(src, "w") as f:
            f.write(prog)
        exe = os.path.join(td, "p")
        subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n in names:
        assert C.sizeof(getattr(m.abi, n)) == int(sizes[n]), n


def _desc(m, B=2, H=64, W=96, cin=16, cout=16, k=3, flags=0):
    d = m.abi.MfcConvDesc()
    d.B, d.Hin, d.Win, d.Hout, d.Wout, d.Cout = B, H, W, H, W, cout
    d.kh = d.kw = k
    d.stride, d.pad, d.upsample, d.dtype, d.nsrc, d.reserved = 1, k // 2, 1, m.abi.MFC_F16, 1, flags
    d.src[0].nchunks = (cin + 7) // 8
    return d


_PLAN_PROBE = """
import ctypes as C, json, sys
sys.path.insert(0, %r)
import mfcnet_tracker_b200 as m
from tests.test_host_cpu import _desc
lib = m.abi.load()
out = []
for (B, H, W, cin, cout, k) in [(24, 480, 640, 16, 16, 3), (8, 480, 640, 22, 15, 11), (1, 60, 80, 128, 128, 3), (3, 64, 96, 32, 16, 1)]:
    info = m.abi.MfcConvInfo()
    assert lib.mfc_conv2d_query(C.byref(_desc(m, B, H, W, cin, cout, k)), C.byref(info)) == 0
    out.append([getattr(info, f) for f, _ in info._fields_])
print(json.dumps(out))
"""


def test_plans_are_identical_across_processes():
    """Tilings decide the fp32 summation order: they must be a pure function of the geometry (committed table, else the
    cost model), not of per-process timing."""
    import subprocess
    import sys
    env = dict(os.environ, MFC_B200_PLAN_ONLY="1")
    env.pop("MFC_CONV_TUNE", None)
    outs = [subprocess.run([sys.executable, "-c", _PLAN_PROBE % ROOT], env=env, capture_output=True, text=True, check=True).stdout
            for _ in range(2)]
    assert outs[0] == outs[1] and outs[0].startswith("[[")


def test_plan_table_roundtrip_and_frozen_plans():
    """mfc_conv2d_plan_import decides the tiling of a geometry that has no plan yet; a plan that was handed out stays."""
    import mfcnet_tracker_b200 as m
    lib = m.abi.load()
    d = _desc(m, B=5, H=40, W=56, cin=24, cout=24, k=3)        # a geometry nothing else in the suite plans
    key = "5 40 56 40 56 24 3 3 1 1 1 3 0 0 0 1"
    assert lib.mfc_conv2d_plan_import(("# comment\n%s : 4 56 0 4 32 4\n" % key).encode()) == 1
    info = m.abi.MfcConvInfo()
    assert lib.mfc_conv2d_query(C.byref(d), C.byref(info)) == 0
    assert (info.tile_h, info.tile_w, info.nb, info.nstages, info.weight_layout) == (4, 56, 32, 4, 0)
    # frozen: a second import for the same geometry is ignored, the exported table still carries the first choice
    assert lib.mfc_conv2d_plan_import(("%s : 8 56 0 4 32 2\n" % key).encode()) == 0
    assert lib.mfc_conv2d_query(C.byref(d), C.byref(info)) == 0 and info.tile_h == 4
    assert (key + " : 4 56 0 4 32 4") in m.abi.export_table()
    assert lib.mfc_conv2d_plan_import(b"1 2 3\n") < 0 and b"malformed" in lib.mfc_last_error()
    # an entry that matches no candidate of the planner falls back to the cost model
    d2 = _desc(m, B=5, H=40, W=56, cin=24, cout=40, k=3)
    assert lib.mfc_conv2d_plan_import(b"5 40 56 40 56 40 3 3 1 1 1 3 0 0 0 1 : 999 56 0 4 32 3\n") == 1
    assert lib.mfc_conv2d_query(C.byref(d2), C.byref(info)) == 0 and info.tile_h != 999
    # statistics need a padded Cout <= 256: the flag is part of the key and constrains the choice
    d3 = _desc(m, B=1, H=30, W=40, cin=64, cout=272, k=3, flags=m.abi.MFC_CONV_WANT_STATS)
    assert lib.mfc_conv2d_query(C.byref(d3), C.byref(info)) != 0 or info.nb * info.nblk <= 256


def test_invalid_descriptor_is_rejected_with_message():
    import mfcnet_tracker_b200 as m
    lib = m.abi.load()
    d = m.abi.MfcConvDesc()
    info = m.abi.MfcConvInfo()
    rc = lib.mfc_conv2d_query(C.byref(d), C.byref(info))
    assert rc == -1 and b"conv" in lib.mfc_last_error()


@pytest.mark.parametrize("tag", ["resunet16_64x96", "resunet8_32x48"])
def test_resunet_state_dict_keys(M, tag):
    meta, man, _ = G.load(tag)
    net = M.ResUnet_VB(channels=3, dim=meta["dim"], out_dim=meta["classes"])
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    net.load_state_dict(G.state_dict(man, meta["seed"]), strict=True)


@pytest.mark.parametrize("variant", ["large", "basic"])
@pytest.mark.parametrize("K", [3, 5])
def test_fusion_state_dict_keys_and_plan(M, variant, K):
    meta, man, _ = G.load(f"fusion_{variant}_k{K}_48x64")
    cls = M.MultiFrameNetLarge if variant == "large" else M.MultiFrameNetBasic
    net = cls(meta["N"], K, False, with_optflow=True, with_depth=True).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    net.load_state_dict(G.state_dict(man, meta["seed"]), strict=True)
    y = net(torch.zeros(2, net.expected_input_channels(), 48, 64))
    assert y.shape == (2, meta["N"], 48, 64)
    with pytest.raises(ValueError):
        net(torch.zeros(2, net.expected_input_channels() + 1, 48, 64))


@pytest.mark.parametrize("variant", ["large", "basic"])
def test_mfcnet_wrapper_keys_and_plan(M, variant):
    meta, man, _ = G.load(f"mfcnet_resunet16_{variant}_k3_64x96")
    cls = M.ResUNetMultiLarge if variant == "large" else M.ResUNetMultiBasic
    net = cls(num_classes=5, num_frames=3, pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    assert [n for n, _ in net.named_children()] == ["base_model", "multiframe_net"]
    xs = [torch.zeros(2, 3, 64, 96)] * 3
    y = net(xs, optflow=[torch.zeros(2, 2, 64, 96)] * 2, depth=[torch.zeros(2, 1, 64, 96)] * 3)
    assert y.shape == (2, 5, 64, 96)
    plan = net._plans[(2, 64, 96)]
    # 3 frame gathers + SFC passes of 59 launches + aux gather (large) or warp (basic) + 3 fusion convs (the final 1x1 is fused into the third)
    n_sfc = (3 * 2) // plan["sub_batch"]
    assert plan["prog"].n_kernels == 3 + 59 * n_sfc + 1 + 3
    with pytest.raises(ValueError):
        net(xs)                                    # flow / depth missing
    with pytest.raises(RuntimeError):
        net.train()(xs, optflow=[torch.zeros(2, 2, 64, 96)] * 2, depth=[torch.zeros(2, 1, 64, 96)] * 3)


def test_unflow_state_dict_keys_and_plan(M):
    meta, man, _ = G.load("unflow_64x128")
    net = M.UnFlow().eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    y = net(torch.zeros(1, 3, 64, 128), torch.zeros(1, 3, 64, 128))
    assert y.shape == (1, 2, 64, 128)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 60, 128), torch.zeros(1, 3, 60, 128))


def test_hrnet_state_dict_keys_and_plan(M):
    """Checkpoint compatibility of the HRNet-W48 shell (1 839 tensors) and a plan-only forward."""
    meta, man, _ = G.load("hrnet_w48_64x96")
    net = M.HighResolutionNet(num_classes=meta["classes"]).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    y = net(torch.zeros(1, 3, 64, 96))
    assert y.shape == (1, meta["classes"], 64, 96)
    prog = net._plans[(1, 64, 96)][0]
    kinds = [m["kind"] for m in prog.meta]
    # 307 reference convs, of which the 720x720 head conv is applied per branch (4 launches): 310
    assert kinds.count("conv") == 310 and kinds.count("fuse_sum") == 2 + 4 * 3 + 3 * 4 + 1 and kinds.count("resize") == 1
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 48, 96))             # not divisible by 32


def test_hrnet_multi_wrapper_keys(M):
    meta, man, _ = G.load("mfcnet_hrnet_large_k3_64x96")
    net = M.HRNetMultiLarge(num_classes=5, num_frames=3, pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    xs = [torch.zeros(1, 3, 64, 96)] * 3
    y = net(xs, optflow=[torch.zeros(1, 2, 64, 96)] * 2, depth=[torch.zeros(1, 1, 64, 96)] * 3)
    assert y.shape == (1, 5, 64, 96)


def test_ternaus16_state_dict_keys_and_plan(M):
    meta, man, _ = G.load("ternaus16_64x96")
    net = M.TernausNet16(num_classes=meta["classes"], num_filters=64).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    net.load_state_dict(G.state_dict(man, meta["seed"], scale_keys=meta["scale_keys"]), strict=True)
    y = net(torch.zeros(1, 3, 64, 96))
    assert y.shape == (1, meta["classes"], 64, 96)
    kinds = [m["kind"] for m in net._plans[(1, 64, 96)][0].meta]
    # 13 VGG convs + 5 decoder ConvRelu + 5 x 4 parity convs of the transposed convs + dec1 + final; 5 pools
    assert kinds.count("conv") == 13 + 5 + 20 + 1 + 1 and kinds.count("maxpool2") == 5 and kinds.count("heatmap_head") == 1


def test_ternaus_multi_wrapper_keys(M):
    meta, man, _ = G.load("mfcnet_ternaus16_basic_k3_64x96")
    net = M.TernausNetMultiBasic(num_classes=5, num_frames=3, pretrained=False, loadpath=None, optflow_inputs=True, depth_inputs=True).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s, _ in man]
    xs = [torch.zeros(1, 3, 64, 96)] * 3
    y = net(xs, optflow=[torch.zeros(1, 2, 64, 96)] * 2, depth=[torch.zeros(1, 1, 64, 96)] * 3)
    assert y.shape == (1, 5, 64, 96)


def test_factories(M):
    class A:
        model_type = "ResUNetMulti-Large"
        num_classes, num_input_frames, pretrained, load_wts_base_model = 5, 3, False, None
        add_optflow_inputs, add_depth_inputs = True, True
    net = M.get_multiframe_segmentation_model(A)
    assert type(net).__name__ == "ResUNetMultiLarge" and net.multiframe_net.in_channels == 22
    A.model_type = "HRNetMulti-Basic"
    assert type(M.get_multiframe_segmentation_model(A)).__name__ == "HRNetMultiBasic"
    A.model_type = "nope"
    with pytest.raises(ValueError):
        M.get_multiframe_segmentation_model(A)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mfcnet-tracker_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            with open(os.path.join(pkg, fn)) as f:
                src = f.read()
            assert "oracle" not in src.replace("torch_oracle", "oracle") or fn == "" , fn


def test_hrnet_program_lanes_are_balanced(M):
    """The HRNet plan records module branches on lanes 0..3 between FORK / JOIN pairs; every list ends joined on lane 0."""
    net = M.HighResolutionNet(num_classes=5).eval()
    net(torch.zeros(1, 3, 64, 96))                    # plan-only forward
    prog = net._plans[(1, 64, 96)][0]
    depth, lanes = 0, set()
    for op, _, _, lane in prog.cmds:
        assert 0 <= lane < M.abi.MFC_MAX_LANES
        lanes.add(lane)
        if op == M.abi.OP_FORK:
            depth += 1
        elif op == M.abi.OP_JOIN:
            depth -= 1
            assert lane == 0
        else:
            assert lane == 0 or depth == 1
        assert depth in (0, 1)
    assert depth == 0 and prog.cmds[-1][3] == 0 and len(lanes) > 1
