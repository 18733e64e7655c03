"""Drop-in mirror of the reference's ``dataloader.py`` hot-path pieces, plus the batched GPU
collate that replaces the per-item CPU pipeline.

``normalize`` / ``concat_stft_cqt`` / ``custom_collate_fn`` keep the reference signatures
(``dataloader.py:9-18``, ``:123-147``).  ``SpectralBatcher`` is the B200-first path: it takes
raw waveforms of a balanced piano / violin batch and produces the ``(B, S, 2, 287, 597)`` float32
tensor and the ``(B,)`` int64 labels the encoders consume in ONE fused C-ABI call, directly in
device memory (no CPU -> CUDA -> CPU -> CUDA ping-pong, no discarded half batch).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .frontend import F_CQT, F_STFT, F_TOTAL, SAMPLE_RATE, FrontEnd, default_frontend
from .utilityFunctions import _cuda_device, _decode_file, _home

DEFAULT_STATS_DIR = "train_set_stats"  # dataloader.py:43-44, :68


def normalize(x, mean, std, eps=1e-8):
    """``dataloader.normalize`` (``dataloader.py:9-13``): ``(x - mean[:, None, :]) / (std[:, None, :] + eps)``."""
    fe = default_frontend(_cuda_device(x))
    if mean.ndim == 2:
        return fe.normalize(x, mean, std, eps).to(_home(x))
    # already broadcast-shaped statistics (mean.ndim != 2): the reference applies them as they are
    m = mean.reshape(mean.shape[0], mean.shape[-1])
    s = std.reshape(std.shape[0], std.shape[-1])
    return fe.normalize(x, m, s, eps).to(_home(x))


def concat_stft_cqt(stft, cqt):
    """``dataloader.concat_stft_cqt`` (``dataloader.py:15-18``): moves both to the compute device and
    concatenates on the frequency axis; the result stays on the device (as in the reference when CUDA
    is available)."""
    fe = default_frontend(_cuda_device(stft))
    return fe.concat(stft, cqt)


def custom_collate_fn(batch):
    """``dataloader.custom_collate_fn`` (``dataloader.py:123-147``): first ``B // 2`` items' piano
    sections, then the same items' violin sections; labels ``[0] * B/2 + [1] * B/2`` (int64).
    Pure data movement (``torch`` copies), kept for callers that still build per-item dicts."""
    batch_size = len(batch)
    half_batch = batch_size // 2
    piano_sections = [batch[i]["piano"] for i in range(half_batch)]
    violin_sections = [batch[i]["violin"] for i in range(half_batch)]
    sample_shape = piano_sections[0].shape
    result = torch.empty((batch_size,) + tuple(sample_shape), dtype=piano_sections[0].dtype,
                         device=piano_sections[0].device)
    for i in range(half_batch):
        result[i] = piano_sections[i]
        result[i + half_batch] = violin_sections[i]
    labels = torch.cat([torch.zeros(half_batch, dtype=torch.long), torch.ones(half_batch, dtype=torch.long)])
    return result, labels


# --------------------------------------------------------------------------------------------
# stats artefacts (a11): train_set_stats/*.npz with keys stft_mean, stft_std, cqt_mean, cqt_std
# --------------------------------------------------------------------------------------------
def load_stats_npz(path: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """One ``.npz`` of the reference's contract (``dataloader.py:48-59``) -> ``(mean, std)`` each
    ``(2, 597)`` float32 (STFT columns then CQT columns)."""
    z = np.load(path)
    for key, shape in (("stft_mean", (2, F_STFT)), ("stft_std", (2, F_STFT)), ("cqt_mean", (2, F_CQT)), ("cqt_std", (2, F_CQT))):
        if key not in z.files or z[key].shape != shape:
            raise ValueError(f"{path}: expected key {key} of shape {shape}")
    mean = np.concatenate([z["stft_mean"], z["cqt_mean"]], axis=1).astype(np.float32)
    std = np.concatenate([z["stft_std"], z["cqt_std"]], axis=1).astype(np.float32)
    return torch.from_numpy(mean), torch.from_numpy(std)


def save_stats_npz(path: str, mean, std) -> None:
    """Writes the four arrays exactly as ``compute_separated_stats.py:52-62`` does."""
    mean = np.asarray(mean, dtype=np.float32)
    std = np.asarray(std, dtype=np.float32)
    np.savez(path, stft_mean=mean[:, :F_STFT], stft_std=std[:, :F_STFT], cqt_mean=mean[:, F_STFT:], cqt_std=std[:, F_STFT:])


def dummy_stats() -> Tuple[torch.Tensor, torch.Tensor]:
    """``_create_dummy_separate_stats`` (``dataloader.py:80-89``): zeros / ones."""
    return torch.zeros(2, F_TOTAL), torch.ones(2, F_TOTAL)


class SpectralBatcher:
    """Balanced piano / violin batch -> encoder input, on the GPU.

    Mirrors ``DualInstrumentDataset`` statistics handling (``dataloader.py:34-89``):
    ``use_separate_stats=True`` loads ``stats_stft_cqt_piano.npz`` / ``stats_stft_cqt_violin.npz``
    from ``stats_dir``; otherwise the unified file ``stats_path`` (default
    ``train_set_stats/stats_unified_stft_cqt.npz``); missing files fall back to zeros / ones with
    the reference's warning."""

    def __init__(self, stats_path: Optional[str] = None, use_separate_stats: bool = True,
                 stats_dir: str = DEFAULT_STATS_DIR, device=None, frontend: Optional[FrontEnd] = None):
        self.fe = frontend if frontend is not None else default_frontend(device)
        dev = self.fe.device
        if use_separate_stats:
            pp = os.path.join(stats_dir, "stats_stft_cqt_piano.npz")
            vp = os.path.join(stats_dir, "stats_stft_cqt_violin.npz")
            if os.path.exists(pp) and os.path.exists(vp):
                (pm, ps), (vm, vs) = load_stats_npz(pp), load_stats_npz(vp)
            else:
                print("⚠️ Warning: Separate stats files not found. Using dummy normalization.")
                print(f"  Expected: {pp}, {vp}")
                (pm, ps), (vm, vs) = dummy_stats(), dummy_stats()
        else:
            if stats_path is None:
                stats_path = os.path.join(stats_dir, "stats_unified_stft_cqt.npz")
            if os.path.exists(stats_path):
                pm, ps = load_stats_npz(stats_path)
            else:
                print(f"⚠️ Warning: Combined stats file {stats_path} not found. Using dummy normalization.")
                pm, ps = dummy_stats()
            vm, vs = pm, ps
        self.mean = torch.stack([pm, vm]).to(dev)  # (2 instruments, 2, 597)
        self.std = torch.stack([ps, vs]).to(dev)

    def __call__(self, piano_waves: torch.Tensor, violin_waves: torch.Tensor,
                 lengths: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``piano_waves`` / ``violin_waves``: ``(B/2, L)`` each.  Returns ``((B, S, 2, 287, 597) float32
        on the device, (B,) int64 labels)`` laid out as ``custom_collate_fn`` does (piano rows first)."""
        half = piano_waves.shape[0]
        if violin_waves.shape != piano_waves.shape:
            raise ValueError("piano and violin batches must have the same shape")
        dev = self.fe.device
        wave = torch.cat([piano_waves.to(dev), violin_waves.to(dev)], dim=0)
        labels = torch.cat([torch.zeros(half, dtype=torch.long), torch.ones(half, dtype=torch.long)])
        idx = labels.to(dev)
        feats, _ = self.fe.features(wave, lengths=lengths, mean=self.mean[idx], std=self.std[idx], layout="sections")
        return feats, labels


def collate_waveforms(piano_waves: Sequence[torch.Tensor], violin_waves: Sequence[torch.Tensor],
                      batcher: SpectralBatcher) -> Tuple[torch.Tensor, torch.Tensor]:
    """Convenience: lists of equal-length mono clips -> one fused batch (see ``SpectralBatcher``)."""
    p = torch.stack([w.reshape(-1) for w in piano_waves])
    v = torch.stack([w.reshape(-1) for w in violin_waves])
    return batcher(p, v)


# --------------------------------------------------------------------------------------------
# DualInstrumentDataset / get_dataloader (dataloader.py:20-121, :149-172) on the device
# --------------------------------------------------------------------------------------------
def _list_audio(directory: str):
    # dataloader.py:22-31: sorted, ".mp3" or ".wav"
    return sorted(os.path.join(directory, f) for f in os.listdir(directory) if f.endswith(".mp3") or f.endswith(".wav"))


def load_clips(paths: Sequence[str], fe: FrontEnd, sample_rate: int = SAMPLE_RATE, cut_time_seconds: float = 10,
               skip_errors: bool = False):
    """``load_audio`` for a list of files -> ``(len(paths), L')`` float32 on the device.  Files are decoded on the
    host and grouped by (channels, rate) so that each group is ONE device call (pad / cut + resample + mono mix).

    ``skip_errors=True`` is the statistics scripts' behaviour (``compute_separated_stats.py:21-38`` catches per-file
    errors, prints them and leaves the file out of the count): undecodable files are reported and dropped, and the
    return value is ``(clips or None, kept_paths)``."""
    if skip_errors:
        decoded, kept = [], []
        for p in paths:
            try:
                decoded.append(_decode_file(p))
                kept.append(p)
            except Exception as e:  # the reference prints and carries on with the next file
                print(f"Errore con {p}: {e}")
        if not kept:
            return None, []
        return load_clips_decoded(decoded, fe, sample_rate, cut_time_seconds), kept
    return load_clips_decoded([_decode_file(p) for p in paths], fe, sample_rate, cut_time_seconds)


def load_clips_decoded(decoded, fe: FrontEnd, sample_rate: int = SAMPLE_RATE, cut_time_seconds: float = 10) -> torch.Tensor:
    """The device part of :func:`load_clips` for already decoded ``(waveform (C, n), rate)`` pairs."""
    paths = decoded
    out = None
    groups = {}
    for i, (w, sr) in enumerate(decoded):
        groups.setdefault((int(w.shape[0]), int(sr)), []).append(i)
    for (channels, sr), idx in groups.items():
        cut = int(cut_time_seconds * sr)
        n_max = min(max(int(decoded[i][0].shape[1]) for i in idx), cut)
        batch = torch.zeros((len(idx), channels, max(n_max, 1)), dtype=torch.float32)
        lengths = torch.zeros(len(idx), dtype=torch.int32)
        for k, i in enumerate(idx):
            n = min(int(decoded[i][0].shape[1]), cut)
            batch[k, :, :n] = decoded[i][0][:, :n]
            lengths[k] = n
        y = fe.load_audio(batch, sr, sample_rate, cut_time_seconds, lengths=lengths)
        if out is None:
            out = torch.empty((len(paths), y.shape[1]), dtype=torch.float32, device=fe.device)
        if y.shape[1] != out.shape[1]:
            raise ValueError("files of different rates must resample to the same clip length")
        out[torch.tensor(idx, device=fe.device)] = y
    return out


class DualInstrumentDataset:
    """``dataloader.DualInstrumentDataset`` (``dataloader.py:20-121``) with the same constructor, file listing and
    statistics handling; ``__getitem__`` returns the reference's dict, computed by the device path."""

    def __init__(self, piano_dir, violin_dir, stats_path=None, use_separate_stats=True, stats_dir: str = DEFAULT_STATS_DIR,
                 device=None):
        self.piano_files = _list_audio(piano_dir)
        self.violin_files = _list_audio(violin_dir)
        self.length = min(len(self.piano_files), len(self.violin_files))
        self.use_separate_stats = use_separate_stats
        self.batcher = SpectralBatcher(stats_path, use_separate_stats, stats_dir, device)

    def __len__(self):
        return self.length

    def waves(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        fe = self.batcher.fe
        both = load_clips([self.piano_files[i] for i in indices] + [self.violin_files[i] for i in indices], fe)
        return both[: len(indices)], both[len(indices):]

    def __getitem__(self, idx):
        p, v = self.waves([idx])
        feats, _ = self.batcher(p, v)
        return {"piano": feats[0], "violin": feats[1], "piano_label": 0, "violin_label": 1}


class GpuDataLoader:
    """What ``get_dataloader`` returns: an iterable of ``((B, S, 2, 287, 597) float32, (B,) int64)`` batches.

    Batch composition follows the reference exactly: ``DataLoader(batch_size=B, drop_last=True)`` hands B items
    to ``custom_collate_fn``, which keeps the first B/2 of them (``dataloader.py:133-136``) - so batch k holds the
    piano and the violin clips of items ``order[k B : k B + B/2]``.  Only those B/2 items are decoded and
    transformed here (the reference also transforms, then discards, the other half)."""

    def __init__(self, dataset: DualInstrumentDataset, batch_size: int, shuffle: bool, generator: Optional[torch.Generator] = None):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), bool(shuffle), generator

    def __len__(self):
        return len(self.dataset) // self.batch_size if self.batch_size > 0 else 0

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            # What DataLoader(shuffle=True).__iter__ does with the global RNG (dataloader.py:172): one draw for the
            # iterator's base seed, then RandomSampler seeds a fresh generator from a second draw - so the shuffle
            # order under torch.manual_seed(s) equals the reference DataLoader's (tests/test_cabi_host.py)
            gen = self.generator
            if gen is None:
                torch.empty((), dtype=torch.int64).random_()
                gen = torch.Generator()
                gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            order = torch.randperm(n, generator=gen).tolist()
        else:
            order = list(range(n))
        half = self.batch_size // 2
        for k in range(len(self)):
            items = order[k * self.batch_size: k * self.batch_size + half]
            p, v = self.dataset.waves(items)
            yield self.dataset.batcher(p, v)


def get_dataloader(piano_dir, violin_dir, batch_size=8, shuffle=True, stats_path=None, use_separate_stats=True):
    """``dataloader.get_dataloader`` (``dataloader.py:149-172``), same arguments and batch layout; batches are
    produced in device memory."""
    if batch_size % 2 != 0:
        print(f"Warning: batch_size={batch_size} is odd. Rounding down to {batch_size-1} for balanced batches.")
        batch_size = batch_size - 1
    dataset = DualInstrumentDataset(piano_dir, violin_dir, stats_path, use_separate_stats)
    return GpuDataLoader(dataset, batch_size, shuffle)
