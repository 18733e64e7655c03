"""Generate ``tests/golden/load_audio.npz`` by running the UNMODIFIED reference ``load_audio``.
TEST INFRASTRUCTURE ONLY.  Run in the build container (``/root/reference`` and ``torchaudio`` present)::

    python -m oracle.make_golden_audio

``torchaudio.load`` needs ``torchcodec`` (absent here), so it is patched to serve in-memory clips keyed by a fake
path; everything after the decode - pad / cut, ``torchaudio.functional.resample``, stereo mean - runs exactly as
``utilityFunctions.py:105-122`` ships it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402


def clip(seed, channels, n, sr):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = np.stack([0.3 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t + rng.uniform(0, 6.28)) * np.exp(-2.0 * t)
                  + 0.05 * rng.standard_normal(n) for c in range(channels)])
    return x.astype(np.float32)


CASES = {
    # name: (channels, samples, orig_sr, cut_time_seconds)
    "stereo_44k_cut": (2, 9000, 44100, 0.15),     # longer than the cut: truncated to 6615 samples, 2:1, stereo mean
    "stereo_44k_pad": (2, 5000, 44100, 0.15),     # shorter than the cut: zero-padded
    "mono_44k": (1, 7001, 44100, 0.15),
    "mono_48k": (1, 6000, 48000, 0.1),            # 320:147 polyphase filter bank
    "stereo_16k_up": (2, 2000, 16000, 0.1),       # upsampling 16 k -> 22.05 k (320:441)
    "mono_22k": (1, 3000, 22050, 0.1),            # no resample, pad / cut only
}


def main():
    import torchaudio

    uf, _dl = ref_loader.load()
    store = {}
    out = {}
    for i, (name, (ch, n, sr, cut)) in enumerate(CASES.items()):
        x = clip(100 + i, ch, n, sr)
        store[name] = (torch.from_numpy(x.copy()), sr)
        out[f"{name}.in"] = x
        out[f"{name}.meta"] = np.array([sr, cut], dtype=np.float64)
    real_load = getattr(torchaudio, "load", None)
    torchaudio.load = lambda path: store[path]
    try:
        for name, (_ch, _n, _sr, cut) in CASES.items():
            y, sr_out = uf.load_audio(name, sample_rate=22050, cut_time_seconds=cut)
            assert sr_out == 22050
            out[f"{name}.out"] = y.numpy().astype(np.float32)
    finally:
        if real_load is not None:
            torchaudio.load = real_load
    path = os.path.join(ROOT, "tests", "golden", "load_audio.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith(".out")})


if __name__ == "__main__":
    main()
