"""The reference's CPU path, as close as it can be run on a box without ``/root/reference`` and without
librosa.  TEST INFRASTRUCTURE / ``bench.py`` CPU ARM ONLY - never on the product path.

* STFT, iSTFT, normalise, concat, section cut / merge are restated with the SAME torch calls the
  reference makes (``torch.stft`` / ``torch.istft`` with a CPU Hann window, eager elementwise ops, the
  Python section loops) - line for line what ``utilityFunctions.py:12-37, :62-82, :240-283`` and
  ``dataloader.py:9-18`` execute, so the timing is the reference's timing for those functions.
* ``get_CQT`` runs the restated ``librosa.cqt`` of ``oracle/cqt.py`` (NumPy / SciPy, polyphase
  decimator) and is labelled "port" in every number that includes it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import cqt as _cqt

WINDOW_SIZE = 287
OVERLAP_FRAMES = 96


def get_STFT(waveform, n_fft=1024, hop_length=256):
    if waveform.ndim == 1:
        waveform = waveform.unsqueeze(0)
    window = torch.hann_window(n_fft)
    stft = torch.stft(waveform, n_fft=n_fft, hop_length=hop_length, window=window, return_complex=True).squeeze(0)
    return torch.stack([torch.real(stft), torch.imag(stft)], dim=-1).permute(2, 1, 0)


def get_CQT(waveform):
    if isinstance(waveform, torch.Tensor):
        waveform = waveform.cpu().numpy()
    return torch.from_numpy(_cqt.get_CQT(waveform.squeeze())).float()


def inverse_STFT(stft_tensor, n_fft=1024, hop_length=256):
    stft_tensor = stft_tensor.permute(0, 2, 1)
    spec = torch.complex(stft_tensor[0], stft_tensor[1]).unsqueeze(0)
    window = torch.hann_window(n_fft)
    return torch.istft(spec, n_fft=n_fft, hop_length=hop_length, window=window, return_complex=False).squeeze(0)


def normalize(x, mean, std, eps=1e-8):
    if mean.ndim == 2:
        mean = mean.unsqueeze(1)
        std = std.unsqueeze(1)
    return (x - mean) / (std + eps)


def get_overlap_windows(spectrogram, window_size=WINDOW_SIZE, overlap_frames=OVERLAP_FRAMES):
    channels, n_time, n_freq = spectrogram.shape
    step = window_size - overlap_frames
    sections = []
    for start in range(0, n_time, step):
        end = min(start + window_size, n_time)
        if end - start < window_size * 0.5:
            break
        section = spectrogram[:, start:end, :]
        pad = window_size - (end - start)
        if pad > 0:
            section = torch.cat([section, torch.zeros((channels, pad, n_freq))], dim=1)
        sections.append(section)
        if end == n_time:
            break
    return torch.stack(sections, dim=0)


def sections2spectrogram(sections, original_size, overlap=OVERLAP_FRAMES):
    n_sections, _, wind, n_freq = sections.shape
    hop = wind - overlap
    n_time = hop * (n_sections - 1) + wind
    full = torch.zeros((2, n_time, n_freq))
    count = torch.zeros((1, n_time, 1))
    for i in range(n_sections):
        full[:, i * hop : i * hop + wind, :] += sections[i]
        count[:, i * hop : i * hop + wind, :] += 1.0
    return (full / count.clamp(min=1.0))[:, :original_size, :]


def features_clip(wave_1xL: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """``DualInstrumentDataset.__getitem__`` for one clip (``dataloader.py:100-112``) -> ``(S, 2, 287, 597)``."""
    stft = normalize(get_STFT(wave_1xL), mean[:, :513], std[:, :513])
    cq = normalize(get_CQT(wave_1xL), mean[:, 513:], std[:, 513:])
    return get_overlap_windows(torch.cat((stft, cq), dim=2))


def features_batch(wave: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """Per-clip loop + collate copy, as the reference's DataLoader does (``dataloader.py:123-147``)."""
    items = [features_clip(wave[i : i + 1], mean, std) for i in range(wave.shape[0])]
    out = torch.empty((len(items),) + tuple(items[0].shape), dtype=items[0].dtype)
    for i, it in enumerate(items):
        out[i] = it
    return out


def istft_batch(sections: torch.Tensor, overlap: int = OVERLAP_FRAMES, original_size: int = 862) -> torch.Tensor:
    """The inference notebook's reconstruction loop (``style_transfer_inference_test.ipynb`` cell 4:36-42)."""
    outs = [inverse_STFT(sections2spectrogram(sections[i], original_size, overlap)) for i in range(sections.shape[0])]
    return torch.stack(outs)
