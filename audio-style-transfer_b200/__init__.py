"""B200-native spectral front-end for Audio-Style-Transfer.

STFT + CQT real/imag feature extraction, per-bin normalisation, section layout, inverse-STFT
reconstruction and dataset statistics as hand-written sm_100a CUDA kernels behind a C-ABI
(``include/ast_frontend.h``), with Python mirrors of the reference's ``utilityFunctions.py`` /
``dataloader.py`` call signatures.  Import is cheap: the CUDA library is loaded on first use.
"""
__version__ = "0.1.0"

from . import synth  # noqa: F401  (pure numpy)


def __getattr__(name):
    import importlib

    if name in ("frontend", "utilityFunctions", "dataloader", "stats", "_lib", "build"):
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("FrontEnd", "default_frontend"):
        return getattr(importlib.import_module(f"{__name__}.frontend"), name)
    raise AttributeError(name)
