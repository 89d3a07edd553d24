"""Builds libmfcnet_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

No torch / pybind involvement: the library exposes only the plain-C entry points declared in
include/mfcnet_b200.h.  nvcc cross-compiles without a GPU, so this runs in the authoring
container; the built .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
# MFC_B200_LIB_TAG / MFC_B200_NVCC_DEFS: measurement builds (e.g. TAG=hint DEFS="-DMFC_WAIT_HINT_NS=4000") live next to the product
_TAG = os.environ.get("MFC_B200_LIB_TAG", "")
LIB = os.path.join(LIBDIR, "libmfcnet_b200%s.so" % (("_" + _TAG) if _TAG else ""))
if _TAG:
    OBJDIR = OBJDIR + "_" + _TAG
UNITS = ["api", "conv_tc", "pointwise", "fusion_ops", "resample", "loss", "train_ops", "correlation", "correlation_tma", "ingest", "localize", "unflow_ops", "raft_ops"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC)")


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mfcnet_b200.h"))
    return deps


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _sources())


def build(force=False, verbose=False):
    """Compile every translation unit for sm_100a and link the shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)

    def one(u):
        obj = os.path.join(OBJDIR, u + ".o")
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("MFC_B200_NVCC_DEFS", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, u + ".cu"), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (u, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(one, UNITS))
    tmp = LIB + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
