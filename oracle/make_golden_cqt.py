"""Pin the CQT on REAL librosa: writes ``tests/golden/cqt_librosa.npz``.  TEST INFRASTRUCTURE ONLY.

``get_CQT`` (``utilityFunctions.py:39-60``) is ``librosa.cqt(y, sr=22050, n_bins=84, hop_length=256)`` with every
other argument at its default (``utilityFunctions.py:52``).  Neither ``librosa`` nor ``soxr`` (its default
``res_type="soxr_hq"`` resampler) is importable in the build container or on the GPU box and there is no network, so
the CQT parity of this repository is "unpinned": ``oracle/cqt.py`` restates the documented algorithm and the CUDA
path is checked against that restatement only.

This script closes the pin the moment both packages are importable anywhere::

    pip install librosa soxr            # on any machine with a network
    python -m oracle.make_golden_cqt    # from the repository root

It runs the reference's exact call on the seeded synthetic clips the parity tests use (``synth.piano_clip`` /
``violin_clip`` / ``noise_clip`` / ``chirp_clip``; NumPy only, no GPU needed), records ``librosa.__version__`` and
``soxr.__version__`` next to the outputs, and ``tests/test_cqt_librosa_pin.py`` then compares both the oracle
restatement (CPU suite) and the CUDA path (``-m gpu``) with those vectors.  While the file is absent those tests SKIP
with the reason; nothing else changes.

If ``/root/reference`` is present the call is made through the unmodified ``utilityFunctions.get_CQT`` itself,
otherwise through the one-line call it wraps.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "cqt_librosa.npz")
sys.path.insert(0, ROOT)

# (kind, clip id, samples): one full 10 s clip of each instrument, short / odd / ragged lengths, the stress clips
CASES = [
    ("piano", 0, 220500), ("violin", 1, 220500), ("piano", 7, 50000), ("violin", 12, 44100),
    ("noise", 60, 36608), ("chirp", 0, 66150), ("piano", 21, 123457),
]


def reference_get_cqt():
    """The reference's own ``get_CQT`` when its tree is importable with real librosa, else the call it wraps."""
    import librosa

    ref_root = os.environ.get("AST_REFERENCE_ROOT", "/root/reference")
    if os.path.exists(os.path.join(ref_root, "utilityFunctions.py")):
        try:
            sys.path.insert(0, ref_root)
            uf = importlib.import_module("utilityFunctions")
            return lambda y: uf.get_CQT(y).numpy(), "utilityFunctions.get_CQT (unmodified reference)"
        except Exception as e:  # e.g. matplotlib / torchaudio missing: fall through to the wrapped call
            print(f"reference module not importable ({e}); calling librosa.cqt directly")
        finally:
            sys.path.remove(ref_root)

    def call(y):
        # utilityFunctions.py:52-58: cqt -> stack real / imag -> (2, T, 84)
        c = librosa.cqt(np.asarray(y, dtype=np.float32), sr=22050, n_bins=84, hop_length=256)
        return np.stack([np.real(c).T, np.imag(c).T]).astype(np.float32)

    return call, "librosa.cqt(y, sr=22050, n_bins=84, hop_length=256) as at utilityFunctions.py:52"


def main() -> int:
    try:
        import librosa
        import soxr
    except ImportError as e:
        print(f"cannot pin the CQT here: {e}.  Install librosa and soxr (see the module docstring) and re-run.")
        return 2
    synth = importlib.import_module("audio_style_transfer_b200.synth")
    get_cqt, how = reference_get_cqt()
    out = {"librosa_version": np.array(librosa.__version__), "soxr_version": np.array(soxr.__version__),
           "numpy_version": np.array(np.__version__), "how": np.array(how),
           "cases": np.array([f"{k}:{i}:{n}" for k, i, n in CASES])}
    for k, (kind, cid, n) in enumerate(CASES):
        y = synth.clip(kind, cid, n)
        v = np.asarray(get_cqt(y), dtype=np.float32)
        assert v.shape == (2, 1 + n // 256, 84), v.shape
        out[f"cqt_{k}"] = v
        print(f"{kind}:{cid}:{n} -> {v.shape}, max |V| {np.abs(v).max():.4f}")
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} (librosa {librosa.__version__}, soxr {soxr.__version__})")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
