"""Extracts the reference's CUDA-C correlation kernels for the NVRTC runner (oracle/corr_ref_nvrtc.py).

TEST INFRASTRUCTURE ONLY.  Runs in the authoring container, where /root/reference exists: the four kernel
strings of models/unflow_correlation.py:10-235 are read out of the module's source with `ast` (the module
itself cannot be imported: it needs cupy and a CUDA device at import time, :3,:6-8) and written, unmodified,
to the git-ignored baseline/_ref/unflow_correlation_kernels.json, which travels to the GPU box with the
snapshot.  No reference source enters the repository's history.

    python -m oracle.make_corr_ref
"""
import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/models/unflow_correlation.py"
OUT = os.path.join(ROOT, "baseline", "_ref", "unflow_correlation_kernels.json")
NAMES = ("kernel_Correlation_rearrange", "kernel_Correlation_updateOutput", "kernel_Correlation_updateGradFirst",
         "kernel_Correlation_updateGradSecond")


def main():
    import warnings
    with open(SRC) as f, warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)    # the module's regex literals use '\\(' in plain strings
        tree = ast.parse(f.read())
    kernels = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            if node.targets[0].id in NAMES and isinstance(node.value, ast.Constant):
                kernels[node.targets[0].id] = node.value.value
    missing = [n for n in NAMES if n not in kernels]
    if missing:
        raise SystemExit("kernel strings not found: %s" % missing)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump({"source": "models/unflow_correlation.py", "kernels": kernels}, f)
    print(OUT)


if __name__ == "__main__":
    main()
