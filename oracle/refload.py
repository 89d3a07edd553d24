"""Import the *real* reference modules from /root/reference (authoring container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box; callers
must check `available()` and skip.  Recipe from SURVEY.md section 8c:
  * put /root/reference and /root/reference/models on sys.path
    (models/multiframe_model.py:9 does `from hrnet import ...`)
  * stub `segmentation_models_pytorch` (imported at top level by
    models/__init__.py:6 and models/multiframe_model.py:8, not installed here)
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MFC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "multiframe_model.py"))


def load():
    """Returns a namespace with the reference classes used on the hot path."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    for p in (os.path.join(REF_ROOT, "models"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "segmentation_models_pytorch" not in sys.modules:
        stub = types.ModuleType("segmentation_models_pytorch")
        stub.Segformer = object
        sys.modules["segmentation_models_pytorch"] = stub
    import importlib
    ns = types.SimpleNamespace()
    ns.resunet = importlib.import_module("models.resunet")
    ns.multiframe = importlib.import_module("models.multiframe_model")
    ns.hrnet = importlib.import_module("models.hrnet")
    ns.ternaus = importlib.import_module("models.ternausnet")
    ns.loss = importlib.import_module("src.loss")
    loc = importlib.import_module("utils.localization_utils_v2")
    ns.localization = loc
    return ns


def manifest_of(module):
    """[(key, shape, dtype-string)] of a module's state_dict, in order."""
    return [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in module.state_dict().items()]
