"""Drop-in mirror of the spectrogram-domain evaluation helpers on the far side of the iSTFT.

``mse_spectrogram`` keeps the signature of ``evaluation_reconstruction.py:105-118`` /
``evaluation_style_transfer.py:111-119`` (NumPy arrays or tensors in, Python float out, ``inf`` on failure) and
runs both STFTs and the reduction on the device; ``instrumentation_similarity`` keeps
``evaluation_style_transfer.py:111-119`` (two 2048-point STFTs, per-bin energy profiles and their Pearson
correlation in one pass on the device); ``reconstruct_audio_from_sections`` keeps
``evaluation_reconstruction.py:161-189`` (section 0 only, ``np.zeros(22050)`` on any exception).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .frontend import F_STFT, default_frontend, _ptr, _stream_ptr
from .utilityFunctions import _cuda_device

N_FFT = 1024        # evaluation_reconstruction.py / evaluation_style_transfer.py module constants
HOP_LENGTH = 256


def _device_signal(x, device):
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    return x.reshape(-1).to(device=device, dtype=torch.float32).contiguous()


def mse_spectrogram_device(original_audio, generated_audio) -> torch.Tensor:
    """The metric as a 0-d float64 tensor on the device (no host synchronisation)."""
    dev = _cuda_device(original_audio if isinstance(original_audio, torch.Tensor) else None)
    fe = default_frontend(dev)
    a, b = _device_signal(original_audio, fe.device), _device_signal(generated_audio, fe.device)
    nbytes = fe.lib.ast_mse_workspace_bytes(fe._plan, a.numel(), b.numel())
    ws = fe._workspace(nbytes)
    out = torch.empty(1, dtype=torch.float64, device=fe.device)
    with torch.cuda.device(fe.device):
        _lib.check(fe.lib.ast_mse_spectrogram(fe._plan, _ptr(a), a.numel(), _ptr(b), b.numel(), _ptr(ws), nbytes, _ptr(out),
                                              _stream_ptr(fe.device)))
    return out[0]


def mse_spectrogram(original_audio, generated_audio, sr=22050):
    """Calculate MSE between spectrograms (``evaluation_reconstruction.py:105-118``)."""
    try:
        return float(mse_spectrogram_device(original_audio, generated_audio).item())
    except Exception as e:  # the reference prints and returns inf
        print(f"Error in mse_spectrogram: {e}")
        return float("inf")


def instrumentation_similarity_device(audio1, audio2) -> torch.Tensor:
    """The correlation as a 0-d float64 tensor on the device (no host synchronisation)."""
    dev = _cuda_device(audio1 if isinstance(audio1, torch.Tensor) else None)
    fe = default_frontend(dev)
    a, b = _device_signal(audio1, fe.device), _device_signal(audio2, fe.device)
    nbytes = fe.lib.ast_instrumentation_similarity_workspace_bytes(fe._plan, a.numel(), b.numel())
    ws = fe._workspace(nbytes)
    out = torch.empty(1, dtype=torch.float64, device=fe.device)
    with torch.cuda.device(fe.device):
        _lib.check(fe.lib.ast_instrumentation_similarity(fe._plan, _ptr(a), a.numel(), _ptr(b), b.numel(), _ptr(ws), nbytes,
                                                         _ptr(out), _stream_ptr(fe.device)))
    return out[0]


def instrumentation_similarity(audio1, audio2, sr=22050):
    """``evaluation_style_transfer.instrumentation_similarity`` (``:111-119``): Pearson correlation of the per-bin
    energy profiles of ``|librosa.stft(audio)|`` (n_fft 2048, hop 512); 0.0 where the reference's is NaN.  Like the
    reference it does not catch exceptions (an empty signal raises)."""
    return float(instrumentation_similarity_device(audio1, audio2).item())


def reconstruct_audio_from_sections(stft_sections, batch_idx=None, sample_idx=None):
    """``evaluation_reconstruction.reconstruct_audio_from_sections`` (``:161-189``), same three positional arguments
    as its call sites (``:345-350``): ``stft_sections (1, S, 2, 287, 513)`` -> the inverse STFT of section 0 as a NumPy
    array; on any exception the message (with the batch / sample indices, as the reference prints them) and
    ``np.zeros(22050)``."""
    try:
        first = stft_sections[0, 0]
        if first.shape[-1] != F_STFT:
            raise RuntimeError(f"expected {F_STFT} frequency bins, got {first.shape[-1]}")
        fe = default_frontend(_cuda_device(first))
        audio = fe.istft(first.unsqueeze(0), layout="flat")[0]
        return audio.cpu().numpy()
    except Exception as e:
        print(f"⚠️ Error in audio reconstruction (batch {batch_idx}, sample {sample_idx}): {e}")
        return np.zeros(22050)
