"""CPU restatement of ``mse_spectrogram`` (``evaluation_reconstruction.py:105-118``) and
``instrumentation_similarity`` (``evaluation_style_transfer.py:111-119``).  TEST INFRASTRUCTURE ONLY.

``librosa.stft(y, n_fft=1024, hop_length=256)`` with librosa >= 0.10 defaults: ``win_length = n_fft``,
``window="hann"`` (``scipy.signal.get_window(..., fftbins=True)``: the periodic Hann, identical to
``torch.hann_window``), ``center=True``, ``pad_mode="constant"`` (zeros), complex64 output of shape
``(1 + n_fft / 2, 1 + len(y) // hop)``.  librosa is not installed here: the transform is pinned
instead on ``torch.stft(center=True, pad_mode="constant")`` in ``tests/test_metrics.py``, which computes the
same thing.
"""
from __future__ import annotations

import numpy as np

from .spectral import HOP, N_FFT, hann_periodic


def librosa_stft_mag(y, n_fft: int = N_FFT, hop_length: int = HOP) -> np.ndarray:
    """``np.abs(librosa.stft(y, n_fft, hop_length))`` -> ``(1 + n_fft / 2, T)`` float32."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    n_frames = 1 + len(y) // hop_length
    idx = np.arange(n_frames)[:, None] * hop_length + np.arange(n_fft)[None, :]
    spec = np.fft.rfft(ypad[idx] * hann_periodic(n_fft)[None, :], axis=1).astype(np.complex64)
    return np.abs(spec).T.astype(np.float32)


def mse_spectrogram(original_audio, generated_audio, sr: int = 22050) -> float:
    spec_orig = librosa_stft_mag(original_audio)
    spec_gen = librosa_stft_mag(generated_audio)
    min_time = min(spec_orig.shape[1], spec_gen.shape[1])
    return float(np.mean((spec_orig[:, :min_time] - spec_gen[:, :min_time]) ** 2))


def instrumentation_similarity(audio1, audio2, sr: int = 22050) -> float:
    """``evaluation_style_transfer.py:111-119``: ``librosa.stft`` with its defaults (``n_fft=2048``,
    ``hop_length = win_length // 4 = 512``), magnitudes summed over time per bin (a float32 ``np.sum``), Pearson
    correlation of the two 1025-bin profiles through ``scipy.stats.pearsonr`` like the reference, NaN -> 0.0."""
    import warnings

    from scipy.stats import pearsonr

    s1 = librosa_stft_mag(audio1, n_fft=2048, hop_length=512)
    s2 = librosa_stft_mag(audio2, n_fft=2048, hop_length=512)
    energy1 = np.sum(s1, axis=1)
    energy2 = np.sum(s2, axis=1)
    min_len = min(len(energy1), len(energy2))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # ConstantInputWarning for silence
        corr, _ = pearsonr(energy1[:min_len], energy2[:min_len])
    return float(corr) if not np.isnan(corr) else 0.0
