"""Synthetic clips: there is no dataset on the GPU box, so every test and
bench input is generated from a seed.  Each clip depends only on ``(kind, clip_id)`` so any
clip of a sharded job can be regenerated anywhere (counter-based Philox stream, seed
``1000 + clip_id``).

piano-like : 6 notes, MIDI uniform in 36..96, sum_{h=1..8} h^-1.5 sin(2 pi h f t) exp(-(t - t0)/tau)
violin-like: 3 sustained notes, 5.5 Hz +-0.5 % vibrato, 12 harmonics h^-1
both + white noise sigma = 1e-3, RMS-normalised to 0.07 (the dataset's target level,
``Preprocessing_Dataset/unifies_violin_datasets.py:22``).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 22050
CLIP_SAMPLES = 220500  # 10 s, utilityFunctions.py:105
TARGET_RMS = 0.07


def _rng(clip_id: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=1000 + int(clip_id)))


def _finish(y: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    y = y + rng.standard_normal(y.shape[0]) * 1e-3
    rms = np.sqrt(np.mean(y * y))
    return (y * (TARGET_RMS / max(rms, 1e-12))).astype(np.float32)


def piano_clip(clip_id: int, n_samples: int = CLIP_SAMPLES, sr: int = SAMPLE_RATE) -> np.ndarray:
    rng = _rng(clip_id)
    t = np.arange(n_samples, dtype=np.float64) / sr
    y = np.zeros(n_samples)
    for _ in range(6):
        f = 440.0 * 2.0 ** ((rng.integers(36, 97) - 69) / 12.0)
        t0 = rng.uniform(0.0, 0.9 * n_samples / sr)
        tau = rng.uniform(0.3, 1.5)
        env = np.where(t >= t0, np.exp(-(t - t0) / tau), 0.0)
        for h in range(1, 9):
            if h * f < 0.5 * sr:
                y += h**-1.5 * np.sin(2 * np.pi * h * f * (t - t0)) * env
    return _finish(y, rng)


def violin_clip(clip_id: int, n_samples: int = CLIP_SAMPLES, sr: int = SAMPLE_RATE) -> np.ndarray:
    rng = _rng(clip_id)
    t = np.arange(n_samples, dtype=np.float64) / sr
    y = np.zeros(n_samples)
    for _ in range(3):
        f = 440.0 * 2.0 ** ((rng.integers(55, 97) - 69) / 12.0)
        ph = rng.uniform(0, 2 * np.pi)
        vib = 0.005 * f / 5.5 * np.sin(2 * np.pi * 5.5 * t + ph)  # phase deviation of a +-0.5 % vibrato
        for h in range(1, 13):
            if h * f < 0.5 * sr:
                y += (1.0 / h) * np.sin(2 * np.pi * h * f * t + h * vib)
    return _finish(y, rng)


def noise_clip(clip_id: int, n_samples: int = CLIP_SAMPLES) -> np.ndarray:
    """``randn * 0.1`` stress clip."""
    return (_rng(clip_id).standard_normal(n_samples) * 0.1).astype(np.float32)


def chirp_clip(n_samples: int = CLIP_SAMPLES, sr: int = SAMPLE_RATE, f0: float = 30.0, f1: float = 10000.0) -> np.ndarray:
    """Full-scale +-1 exponential chirp 30 Hz -> 10 kHz."""
    t = np.arange(n_samples, dtype=np.float64) / sr
    dur = n_samples / sr
    k = (f1 / f0) ** (1.0 / dur)
    phase = 2 * np.pi * f0 * (k**t - 1.0) / np.log(k)
    return np.sin(phase).astype(np.float32)


def clip(kind: str, clip_id: int, n_samples: int = CLIP_SAMPLES) -> np.ndarray:
    if kind == "piano":
        return piano_clip(clip_id, n_samples)
    if kind == "violin":
        return violin_clip(clip_id, n_samples)
    if kind == "noise":
        return noise_clip(clip_id, n_samples)
    if kind == "chirp":
        return chirp_clip(n_samples)
    raise ValueError(f"unknown clip kind {kind!r}")


def batch(n_clips: int, n_samples: int = CLIP_SAMPLES, first_id: int = 0) -> np.ndarray:
    """``(n_clips, n_samples)`` float32: even ids piano-like, odd ids violin-like."""
    out = np.empty((n_clips, n_samples), dtype=np.float32)
    for i in range(n_clips):
        cid = first_id + i
        out[i] = piano_clip(cid, n_samples) if cid % 2 == 0 else violin_clip(cid, n_samples)
    return out
