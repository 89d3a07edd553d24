"""Compiles the oracle's C restatements with gcc into oracle/_build/ (git-ignored, travels to the
GPU box).  TEST INFRASTRUCTURE ONLY."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
CORR_SO = os.path.join(OUT, "libcorr_oracle.so")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "corr_oracle.c")
    if force or not os.path.exists(CORR_SO) or os.path.getmtime(CORR_SO) < os.path.getmtime(src):
        gcc = shutil.which("gcc")
        if gcc is None:
            if os.path.exists(CORR_SO):
                return CORR_SO
            raise RuntimeError("gcc not found and %s is missing" % CORR_SO)
        subprocess.run([gcc, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", CORR_SO, src, "-lm"], check=True)
    return CORR_SO
