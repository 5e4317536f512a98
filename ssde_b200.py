"""Importable alias for the package directory ``self-supervised-depth-estimation_b200/``.

The directory name is fixed by the project layout and is not a valid Python identifier, so
``import ssde_b200`` loads it through importlib and registers the package (and its submodules)
under this shorter name.  ``from ssde_b200 import layers`` / ``import ssde_b200.trainer_hooks``
both resolve to the single real module object.
"""
import importlib
import os
import sys

_REAL = "self-supervised-depth-estimation_b200"
_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)

_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
