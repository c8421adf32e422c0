"""Alias: `import teethrt` == the package directory `multimodal-teeth-restoration-selection_b200/` (whose name is not a
valid Python identifier).  Submodules are aliased too (`teethrt.calib` IS `multimodal-...-b200.calib`, one module object),
so classes such as TeethRTError have a single identity whichever name they were imported under."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_ALIAS = __name__
_REAL = "multimodal-teeth-restoration-selection_b200"
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_real = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _real
