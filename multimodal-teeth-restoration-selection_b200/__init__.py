"""teethrt — B200-native (sm_100a) hot path of ahmedmajid92/multimodal-teeth-restoration-selection.

Importable as `teethrt` (alias package at the repo root) or via importlib under its directory name.
The CUDA library is loaded on import of `._lib`; nothing here falls back to PyTorch/CPU compute.
"""
from ._lib import TeethRTError, init, lib  # noqa: F401

__all__ = ["TeethRTError", "init", "lib"]
