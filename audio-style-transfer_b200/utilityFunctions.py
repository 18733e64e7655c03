"""Drop-in mirror of the reference's ``utilityFunctions.py`` hot-path functions.

Same names, defaults, argument meaning, output shapes / dtypes and error behaviour as the
reference (file:line cited per function); the arithmetic runs in the sm_100a kernels behind
``libast_frontend.so``.  Results are returned on the device of the input (CPU in -> CPU out,
like the reference; CUDA in -> CUDA out with no host round trip).  There is no CPU fallback:
without a CUDA device every function raises.

Not mirrored (out of scope): ``inverse_CQT`` (dead code in the reference),
``plot_stft`` / ``plot_cqt`` (matplotlib).  ``load_audio`` decodes the file on the host and runs everything
after the decode (pad / cut, resample, stereo mean - SURVEY.md §8f-1) on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .frontend import F_CQT, F_STFT, HOP, N_FFT, SAMPLE_RATE, default_frontend

WINDOW_SIZE = 287          # utilityFunctions.py:8
OVERLAP_PERCENTAGE = 0.3   # utilityFunctions.py:9
OVERLAP_FRAMES = 96        # utilityFunctions.py:10


def _home(t) -> torch.device:
    return t.device if isinstance(t, torch.Tensor) else torch.device("cpu")


def _cuda_device(t) -> torch.device:
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("audio-style-transfer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _check_geometry(n_fft, hop_length):
    if (n_fft, hop_length) != (N_FFT, HOP):
        raise NotImplementedError(
            f"only n_fft={N_FFT}, hop_length={HOP} (the values every reference call site uses) are built; "
            f"got n_fft={n_fft}, hop_length={hop_length}")


def _mono(waveform: torch.Tensor) -> torch.Tensor:
    if waveform.ndim == 1:
        waveform = waveform.unsqueeze(0)  # utilityFunctions.py:21-22
    if waveform.ndim != 2 or waveform.shape[0] != 1:
        # the reference squeezes dim 0 after torch.stft, which only yields (freq, time) for one channel
        raise ValueError(f"expected a mono waveform of shape (1, samples) or (samples,), got {tuple(waveform.shape)}")
    return waveform


def get_STFT(waveform, n_fft=1024, hop_length=256):
    """``utilityFunctions.get_STFT`` (``utilityFunctions.py:12-37``).

    Input: audio of shape (1, samples) or (samples,).  Output: ``(2, T, 513)`` float32,
    channel 0 real / channel 1 imaginary, ``T = 1 + samples // 256``."""
    _check_geometry(n_fft, hop_length)
    waveform = _mono(waveform)
    if waveform.shape[1] <= n_fft // 2:
        raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, "
                           f"but got: padding ({n_fft // 2}, {n_fft // 2}) at dimension 2 of input {list(waveform.shape)}")
    fe = default_frontend(_cuda_device(waveform))
    return fe.stft(waveform)[0].to(_home(waveform))


def get_CQT(waveform, sample_rate=22050, n_bins=84, hop_length=256):
    """``utilityFunctions.get_CQT`` (``utilityFunctions.py:39-60``): ``librosa.cqt(y, sr=22050,
    n_bins=84, hop_length=256)`` real / imaginary stacked -> ``(2, T, 84)`` float32.
    Accepts a tensor or an ndarray (``:47-50``)."""
    if (sample_rate, n_bins, hop_length) != (SAMPLE_RATE, F_CQT, HOP):
        raise NotImplementedError(
            f"only sample_rate={SAMPLE_RATE}, n_bins={F_CQT}, hop_length={HOP} are built; "
            f"got {sample_rate}, {n_bins}, {hop_length}")
    home = _home(waveform)
    if not isinstance(waveform, torch.Tensor):
        waveform = torch.from_numpy(np.ascontiguousarray(np.asarray(waveform, dtype=np.float32)))
    waveform = waveform.squeeze()  # utilityFunctions.py:50
    if waveform.ndim != 1:
        raise ValueError(f"expected a mono waveform, got shape {tuple(waveform.shape)} after squeeze")
    if waveform.shape[0] <= N_FFT // 2:
        raise RuntimeError("waveform too short for the front-end (needs more than 512 samples)")
    fe = default_frontend(_cuda_device(waveform))
    return fe.cqt(waveform.unsqueeze(0))[0].to(home)


def inverse_STFT(stft_tensor, n_fft=1024, hop_length=256):
    """``utilityFunctions.inverse_STFT`` (``utilityFunctions.py:62-82``): ``(2, T, 513)`` ->
    ``(256 * (T - 1),)`` float32 (``torch.istft`` with the periodic Hann window, centre-trimmed)."""
    _check_geometry(n_fft, hop_length)
    if stft_tensor.ndim != 3 or stft_tensor.shape[0] != 2:
        raise ValueError(f"expected (2, time, freq), got {tuple(stft_tensor.shape)}")
    if stft_tensor.shape[2] != F_STFT:
        raise RuntimeError(f"istft expects {F_STFT} frequency bins for n_fft={n_fft}, got {stft_tensor.shape[2]}")
    fe = default_frontend(_cuda_device(stft_tensor))
    return fe.istft(stft_tensor.unsqueeze(0), layout="flat")[0].to(_home(stft_tensor))


def get_overlap_windows(spectrogram, window_size=WINDOW_SIZE, overlap_frames=OVERLAP_FRAMES):
    """``utilityFunctions.get_overlap_windows`` (``utilityFunctions.py:240-263``):
    ``(2, time, freq)`` -> ``(n_sections, 2, window_size, freq)``, last section zero-padded."""
    fe = default_frontend(_cuda_device(spectrogram))
    return fe.overlap_windows(spectrogram, window_size, overlap_frames).to(_home(spectrogram))


def sections2spectrogram(sections, original_size, overlap=OVERLAP_FRAMES):
    """``utilityFunctions.sections2spectrogram`` (``utilityFunctions.py:265-283``)."""
    fe = default_frontend(_cuda_device(sections))
    return fe.sections_merge(sections, original_size, overlap).to(_home(sections))


def concat_stft_cqt(stft, cqt):
    """``utilityFunctions.concat_stft_cqt`` (``utilityFunctions.py:285-299``), same ``ValueError``s."""
    if stft.ndim != 3 or cqt.ndim != 3:
        raise ValueError(f"Both tensors must be 3D, got {stft.ndim}D e {cqt.ndim}D.")
    if stft.shape[0] != cqt.shape[0] or stft.shape[1] != cqt.shape[1]:
        raise ValueError(f"Channel/Time mismatch: stft {stft.shape[:2]} vs cqt {cqt.shape[:2]}")
    fe = default_frontend(_cuda_device(stft))
    return fe.concat(stft, cqt).to(_home(stft))


def _decode_file(file_path):
    """Host-side decode: ``torchaudio.load`` where it works (it needs ``torchcodec``), else PCM / float WAV
    through ``scipy.io.wavfile``.  Returns ``(channels, samples)`` float32 in [-1, 1] and the file's rate."""
    try:
        import torchaudio

        waveform, sr = torchaudio.load(file_path)
        return waveform.to(torch.float32), int(sr)
    except (ImportError, RuntimeError, OSError, AttributeError):
        pass
    from scipy.io import wavfile

    sr, data = wavfile.read(file_path)
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.T)), int(sr)


def load_audio_tensor(waveform, orig_sample_rate, sample_rate=22050, cut_time_seconds=10):
    """Everything ``load_audio`` does after ``torchaudio.load`` (``utilityFunctions.py:109-120``), on the device:
    ``(C, L)`` at ``orig_sample_rate`` -> ``(1, L')`` at ``sample_rate`` for mono / stereo input."""
    if waveform.ndim != 2:
        raise ValueError(f"expected (channels, samples), got {tuple(waveform.shape)}")
    if waveform.shape[0] not in (1, 2):
        raise NotImplementedError("only mono and stereo files are handled on the device (the reference averages "
                                  "exactly two channels, utilityFunctions.py:119)")
    fe = default_frontend(_cuda_device(waveform))
    out = fe.load_audio(waveform, orig_sample_rate, sample_rate, cut_time_seconds)
    return out.to(_home(waveform)), sample_rate


def load_audio(file_path, sample_rate=22050, cut_time_seconds=10):
    """``utilityFunctions.load_audio`` (``utilityFunctions.py:105-122``): returns ``(waveform (1, L'), sample_rate)``."""
    waveform, orig_sample_rate = _decode_file(file_path)
    return load_audio_tensor(waveform, orig_sample_rate, sample_rate, cut_time_seconds)
