"""CPU restatement of the torch half of the spectral front-end.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it restates.  The arithmetic is plain
NumPy in float64 (the reference runs torch CPU in float32 - complex64 FFT); the
restatement is pinned against outputs of the unmodified reference captured by
``oracle/make_golden.py`` (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import numpy as np

from . import cqt as _cqt

WINDOW_SIZE = 287  # utilityFunctions.py:8
OVERLAP_FRAMES = 96  # utilityFunctions.py:10
N_FFT = 1024
HOP = 256
F_STFT = N_FFT // 2 + 1
F_CQT = 84


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """``torch.hann_window(n)`` (periodic=True), ``utilityFunctions.py:24`` / ``:76``."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


# --------------------------------------------------------------------------- a1
def get_STFT(waveform, n_fft: int = N_FFT, hop_length: int = HOP, dtype=np.float32) -> np.ndarray:
    """``utilityFunctions.get_STFT`` (``utilityFunctions.py:12-37``): ``torch.stft(center=True,
    pad_mode="reflect", window=hann, onesided, normalized=False)`` -> ``(2, T, n_fft // 2 + 1)``,
    ``T = 1 + L // hop``."""
    y = np.asarray(waveform, dtype=np.float64).reshape(-1)
    pad = n_fft // 2
    if len(y) <= pad:
        raise RuntimeError("reflect padding needs more than n_fft // 2 samples")  # torch.stft raises too
    ypad = np.pad(y, (pad, pad), mode="reflect")
    n_frames = 1 + len(y) // hop_length
    idx = np.arange(n_frames)[:, None] * hop_length + np.arange(n_fft)[None, :]
    frames = ypad[idx] * hann_periodic(n_fft)[None, :]
    spec = np.fft.rfft(frames, axis=1)  # (T, F)
    out = np.stack([spec.real, spec.imag], axis=0)  # (2, T, F)
    out[1, :, 0] = 0.0
    if n_fft % 2 == 0:
        out[1, :, -1] = 0.0
    return out.astype(dtype)


# --------------------------------------------------------------------------- a2
def get_CQT(waveform, sample_rate=22050, n_bins=84, hop_length=256) -> np.ndarray:
    """``utilityFunctions.get_CQT`` (``utilityFunctions.py:39-60``) - see ``oracle/cqt.py``."""
    return _cqt.get_CQT(waveform, sample_rate, n_bins, hop_length)


# --------------------------------------------------------------------------- a3
def normalize(x, mean, std, eps: float = 1e-8) -> np.ndarray:
    """``dataloader.normalize`` (``dataloader.py:9-13``), float32 arithmetic like the reference."""
    x = np.asarray(x, dtype=np.float32)
    mean = np.asarray(mean, dtype=np.float32)
    std = np.asarray(std, dtype=np.float32)
    if mean.ndim == 2:
        mean = mean[:, None, :]
        std = std[:, None, :]
    return (x - mean) / (std + np.float32(eps))


# --------------------------------------------------------------------------- a4
def concat_stft_cqt(stft, cqt) -> np.ndarray:
    """``utilityFunctions.concat_stft_cqt`` (``utilityFunctions.py:285-299``)."""
    stft = np.asarray(stft)
    cqt = np.asarray(cqt)
    if stft.ndim != 3 or cqt.ndim != 3:
        raise ValueError(f"Both tensors must be 3D, got {stft.ndim}D e {cqt.ndim}D.")
    if stft.shape[0] != cqt.shape[0] or stft.shape[1] != cqt.shape[1]:
        raise ValueError(f"Channel/Time mismatch: stft {stft.shape[:2]} vs cqt {cqt.shape[:2]}")
    return np.concatenate([stft, cqt], axis=2)


# --------------------------------------------------------------------------- a5
def section_starts(n_time: int, window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES):
    """Start frame of every section ``get_overlap_windows`` emits (``utilityFunctions.py:249-261``)."""
    step = window_size - overlap_frames
    starts = []
    for start in range(0, n_time, step):
        end = min(start + window_size, n_time)
        if end - start < window_size * 0.5:
            break
        starts.append(start)
        if end == n_time:
            break
    return starts


def n_sections(n_time: int, window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES) -> int:
    return len(section_starts(n_time, window_size, overlap_frames))


def get_overlap_windows(spectrogram, window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES) -> np.ndarray:
    """``utilityFunctions.get_overlap_windows`` (``utilityFunctions.py:240-263``):
    ``(2, T, F)`` -> ``(S, 2, window_size, F)``, last section zero-padded."""
    spectrogram = np.asarray(spectrogram)
    channels, n_time, n_freq = spectrogram.shape
    starts = section_starts(n_time, window_size, overlap_frames)
    if not starts:
        raise RuntimeError("stack expects a non-empty TensorList")  # torch.stack([]) in the reference
    out = np.zeros((len(starts), channels, window_size, n_freq), dtype=spectrogram.dtype)
    for i, s in enumerate(starts):
        e = min(s + window_size, n_time)
        out[i, :, : e - s, :] = spectrogram[:, s:e, :]
    return out


# --------------------------------------------------------------------------- a6
def custom_collate_fn(batch):
    """``dataloader.custom_collate_fn`` (``dataloader.py:123-147``): the first ``B // 2`` items'
    piano sections then the same items' violin sections; labels ``[0] * B/2 + [1] * B/2``."""
    batch_size = len(batch)
    half = batch_size // 2
    piano = [batch[i]["piano"] for i in range(half)]
    violin = [batch[i]["violin"] for i in range(half)]
    shape = piano[0].shape
    out = np.empty((batch_size,) + tuple(shape), dtype=piano[0].dtype)
    for i in range(half):
        out[i] = piano[i]
        out[i + half] = violin[i]
    labels = np.concatenate([np.zeros(half, dtype=np.int64), np.ones(half, dtype=np.int64)])
    return out, labels


# --------------------------------------------------------------------------- a7
def sections2spectrogram(sections, original_size: int, overlap: int = OVERLAP_FRAMES) -> np.ndarray:
    """``utilityFunctions.sections2spectrogram`` (``utilityFunctions.py:265-283``):
    overlap-average then crop to ``original_size`` frames.  float32 accumulate like the reference."""
    sections = np.asarray(sections, dtype=np.float32)
    n_sec, _, wind, n_freq = sections.shape
    hop = wind - overlap
    n_time = hop * (n_sec - 1) + wind
    full = np.zeros((2, n_time, n_freq), dtype=np.float32)
    count = np.zeros((1, n_time, 1), dtype=np.float32)
    for i in range(n_sec):
        s = i * hop
        full[:, s : s + wind, :] += sections[i]
        count[:, s : s + wind, :] += 1.0
    full = full / np.maximum(count, 1.0)
    return full[:, :original_size, :]


# --------------------------------------------------------------------------- a8
def inverse_STFT(stft_tensor, n_fft: int = N_FFT, hop_length: int = HOP, dtype=np.float32) -> np.ndarray:
    """``utilityFunctions.inverse_STFT`` (``utilityFunctions.py:62-82``): ``torch.istft(center=True,
    hann, onesided, length=None)``.  irfft (1/N, imaginary parts of DC / Nyquist ignored) -> x window ->
    overlap-add -> / sum w^2 -> trim ``n_fft // 2`` each side -> ``hop * (T - 1)`` samples."""
    spec = np.asarray(stft_tensor, dtype=np.float64)
    n_frames = spec.shape[1]
    X = spec[0] + 1j * spec[1]  # (T, F)
    w = hann_periodic(n_fft)
    frames = np.fft.irfft(X, n=n_fft, axis=1) * w[None, :]
    total = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(total)
    env = np.zeros(total)
    for t in range(n_frames):
        y[t * hop_length : t * hop_length + n_fft] += frames[t]
        env[t * hop_length : t * hop_length + n_fft] += w * w
    start = n_fft // 2
    end = total - n_fft // 2
    y, env = y[start:end], env[start:end]
    if np.any(np.abs(env) < 1e-11):
        raise RuntimeError("window overlap add min: NOLA violated")  # torch.istft raises the same way
    return (y / env).astype(dtype)


# --------------------------------------------------------------------------- a9
def reconstruct_audio_from_sections(sections, n_fft: int = N_FFT, hop_length: int = HOP) -> np.ndarray:
    """``evaluation_reconstruction.reconstruct_audio_from_sections`` (``:161-189``): iSTFT of
    section 0 only of a ``(1, S, 2, T, F)`` batch; any failure returns one second of silence."""
    try:
        sections = np.asarray(sections)
        if sections.ndim == 5:
            sections = sections[0]
        return inverse_STFT(sections[0], n_fft, hop_length)
    except Exception:
        return np.zeros(22050, dtype=np.float32)


# --------------------------------------------------------------------------- a10
def clip_features(waveform) -> np.ndarray:
    """``merged = cat(get_STFT(audio), get_CQT(audio))`` (``compute_separated_stats.py:22-24``)."""
    return concat_stft_cqt(get_STFT(waveform), get_CQT(waveform))


def compute_stats(clips, features=None):
    """``compute_stats`` (``compute_separated_stats.py:16-43``; ``compute_unified_stats.py:25-50``):
    mean over clips of the per-clip per-bin mean over T, and sqrt of the mean over clips of the
    per-clip UNBIASED variance over T.  Returns ``(mean (2, 597), std (2, 597))`` float32.

    The reference accumulates in float32; this restatement accumulates in float64 (the CUDA path
    does too) and rounds once at the end.
    """
    sum_all = None
    sum_sq_all = None
    count = 0
    for i, clip in enumerate(clips):
        merged = np.asarray(features[i] if features is not None else clip_features(clip), dtype=np.float64)
        clip_mean = merged.mean(axis=1)
        clip_var = merged.var(axis=1, ddof=1)
        if sum_all is None:
            sum_all, sum_sq_all = clip_mean.copy(), clip_var.copy()
        else:
            sum_all += clip_mean
            sum_sq_all += clip_var
        count += 1
    mean = sum_all / count
    std = np.sqrt(sum_sq_all / count)
    return mean.astype(np.float32), std.astype(np.float32)


def split_stats(mean: np.ndarray, std: np.ndarray):
    """The four arrays of the ``train_set_stats/*.npz`` contract (``compute_separated_stats.py:46-62``,
    ``dataloader.py:48-59``)."""
    return dict(
        stft_mean=mean[:, :F_STFT], stft_std=std[:, :F_STFT], cqt_mean=mean[:, F_STFT:], cqt_std=std[:, F_STFT:]
    )


# --------------------------------------------------------------------------- whole path
def features_sections(waveform, mean=None, std=None, eps: float = 1e-8,
                      window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES) -> np.ndarray:
    """``DualInstrumentDataset.__getitem__`` for one clip (``dataloader.py:94-112``):
    STFT + CQT -> normalise each -> concat -> sections ``(S, 2, 287, 597)``.
    ``mean`` / ``std`` are ``(2, 597)`` (STFT stats then CQT stats) or ``None`` (no normalisation,
    as in ``evaluation_style_transfer.process_audio``, ``:135-139``)."""
    stft = get_STFT(waveform)
    cq = get_CQT(waveform)
    if mean is not None:
        stft = normalize(stft, mean[:, :F_STFT], std[:, :F_STFT], eps)
        cq = normalize(cq, mean[:, F_STFT:], std[:, F_STFT:], eps)
    return get_overlap_windows(concat_stft_cqt(stft, cq), window_size, overlap_frames)
