"""Import the UNMODIFIED reference modules for golden-vector generation.  TEST INFRASTRUCTURE ONLY.

Works only where ``/root/reference`` exists (the build container).  ``librosa`` and
``matplotlib`` are not installed, and ``utilityFunctions.py:3,5`` imports them at module
scope, so empty stub modules are placed in ``sys.modules`` first; every function that does
not touch them (``get_STFT``, ``inverse_STFT``, ``get_overlap_windows``,
``sections2spectrogram``, ``concat_stft_cqt``, ``dataloader.normalize``,
``dataloader.custom_collate_fn``) then runs exactly as shipped.  ``get_CQT`` cannot run.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AST_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utilityFunctions.py"))


def load():
    """Returns ``(utilityFunctions, dataloader)`` reference modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("librosa", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the product ships modules with the same names; make sure the reference wins here
    for name in ("utilityFunctions", "dataloader"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    uf = importlib.import_module("utilityFunctions")
    dl = importlib.import_module("dataloader")
    assert uf.__file__.startswith(REFERENCE_ROOT) and dl.__file__.startswith(REFERENCE_ROOT)
    return uf, dl
