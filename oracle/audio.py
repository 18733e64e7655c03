"""CPU restatement of ``load_audio``'s arithmetic (``utilityFunctions.py:105-122``).  TEST INFRASTRUCTURE ONLY.

``torchaudio.functional.resample`` (defaults: ``sinc_interp_hann``, ``lowpass_filter_width=6``,
``rolloff=0.99``) restated from ``torchaudio/functional/functional.py`` (2.11):
``_get_sinc_resample_kernel`` (float64 arithmetic, float32 taps) and ``_apply_sinc_resample_kernel``
(zero padding ``(width, width + orig)``, ``conv1d`` with stride ``orig``, crop to ``ceil(new * L / orig)``).
Pinned by ``tests/golden/load_audio.npz``, which ``oracle/make_golden_audio.py`` generates by running the
UNMODIFIED reference ``load_audio`` (with ``torchaudio.load`` patched to serve in-memory clips).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def resample_geometry(orig_sr: int, new_sr: int):
    g = math.gcd(int(orig_sr), int(new_sr))
    orig, new = int(orig_sr) // g, int(new_sr) // g
    base = min(orig, new) * 0.99
    width = math.ceil(6 * orig / base)
    return orig, new, width


def resample_taps(orig_sr: int, new_sr: int) -> np.ndarray:
    """``(new, 2 * width + orig)`` float32, functional.py ``_get_sinc_resample_kernel`` with ``dtype=None``."""
    orig, new, width = resample_geometry(orig_sr, new_sr)
    base = min(orig, new) * 0.99
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = t * base
    t = np.clip(t, -6.0, 6.0)
    window = np.cos(t * math.pi / 6.0 / 2.0) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0, 1.0, np.sin(t) / t)
    return (sinc * window * (base / orig)).astype(np.float32)


def resample(x: np.ndarray, orig_sr: int, new_sr: int) -> np.ndarray:
    """``(..., L)`` -> ``(..., ceil(new * L / orig))``; float64 accumulation of the float32 taps."""
    x = np.asarray(x)
    if orig_sr == new_sr:
        return x.astype(np.float32)
    orig, new, width = resample_geometry(orig_sr, new_sr)
    taps = resample_taps(orig_sr, new_sr).astype(np.float64)
    lead = x.shape[:-1]
    flat = x.reshape(-1, x.shape[-1]).astype(np.float64)
    length = flat.shape[1]
    padded = np.pad(flat, ((0, 0), (width, width + orig)))
    n_steps = (padded.shape[1] - taps.shape[1]) // orig + 1
    # frames[n, k] = padded[n * orig + k]
    idx = np.arange(n_steps)[:, None] * orig + np.arange(taps.shape[1])[None, :]
    out = np.einsum("bnk,ik->bni", padded[:, idx], taps).reshape(flat.shape[0], -1)
    target = -(-new * length // orig)
    return out[:, :target].reshape(lead + (target,)).astype(np.float32)


def load_audio_from_array(waveform: np.ndarray, orig_sample_rate: int, sample_rate: int = 22050,
                          cut_time_seconds: float = 10):
    """Everything of ``load_audio`` after ``torchaudio.load``: ``(C, L)`` float32 -> ``(1 or C, L')``."""
    waveform = np.asarray(waveform, dtype=np.float32)
    cut = int(cut_time_seconds * orig_sample_rate)
    if waveform.shape[-1] < cut:
        waveform = np.concatenate([waveform, np.zeros((waveform.shape[0], cut - waveform.shape[-1]), np.float32)], axis=-1)
    waveform = waveform[:, :cut]
    if orig_sample_rate != sample_rate:
        waveform = resample(waveform, orig_sample_rate, sample_rate)
    if waveform.shape[0] == 2:
        waveform = waveform.astype(np.float32).mean(axis=0, keepdims=True, dtype=np.float32)
    return waveform.astype(np.float32), sample_rate
