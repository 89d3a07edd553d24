"""Importable alias of the product package.

The product lives in ``mfcnet-tracker_b200/`` (the directory name the project layout asks for);
a hyphen cannot appear in a Python module name, so this shim puts that directory on its own
``__path__`` and re-exports the public API.  ``import mfcnet_tracker_b200`` is the user-facing import.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mfcnet-tracker_b200"))

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
