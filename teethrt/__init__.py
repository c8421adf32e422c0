"""Alias: `import teethrt` == the package directory `multimodal-teeth-restoration-selection_b200/` (whose name is not a
valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = importlib.import_module("multimodal-teeth-restoration-selection_b200")
sys.modules[__name__] = _real
