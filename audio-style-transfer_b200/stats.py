"""Dataset statistics (replaces ``Preprocessing_Dataset/compute_separated_stats.py`` and
``compute_unified_stats.py``).

The reference's statistics are *not* the pooled mean / std: per clip it takes the per-bin mean over
time and the UNBIASED per-bin variance over time of the raw ``(2, T, 597)`` features, then
``mean = sum(clip_mean) / N`` and ``std = sqrt(sum(clip_var) / N)``
(``compute_separated_stats.py:27-42``).  Both sums are plain sums over clips, so clips shard across
ranks with no communication and ONE all-reduce(sum) of ``(G, 2, 2, 597) + (G,)`` float64 values
combines the partial moments (NCCL over NVLink on GPUs; ``gloo`` in the CPU tests).  The unified
statistics are the sum of the per-instrument groups (``compute_unified_stats.py`` walks both
directories).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .dataloader import save_stats_npz
from .frontend import F_TOTAL, FrontEnd

GROUP_NAMES = ("piano", "violin")


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous block partition of ``n_items`` clips over ``world_size`` ranks (first ranks get the
    remainder), so the union over ranks is exactly ``range(n_items)`` with no overlap."""
    base, rem = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def allreduce_accumulators(acc: torch.Tensor, counts: torch.Tensor, group=None) -> None:
    """In-place sum of the partial moments over all ranks: the only collective of the whole path."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([acc.reshape(-1), counts.reshape(-1)])  # one message (~19 KB per group)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    acc.copy_(flat[: acc.numel()].reshape(acc.shape))
    counts.copy_(flat[acc.numel():].reshape(counts.shape))


def finalize(acc: torch.Tensor, count) -> Tuple[np.ndarray, np.ndarray]:
    """``acc (2, 2, 597)`` float64 (sum of clip means, sum of clip variances), ``count`` clips ->
    ``(mean, std)`` float32 ``(2, 597)`` via the C-ABI's ``ast_stats_finalize``."""
    lib = _lib.load()
    a = np.ascontiguousarray(acc.detach().cpu().numpy().astype(np.float64).reshape(2, 2, F_TOTAL))
    mean = np.zeros((2, F_TOTAL), dtype=np.float32)
    std = np.zeros((2, F_TOTAL), dtype=np.float32)
    _lib.check(lib.ast_stats_finalize(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(count),
                                      mean.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                      std.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
    return mean, std


def finalize_all(acc: torch.Tensor, counts: torch.Tensor, names: Sequence[str] = GROUP_NAMES) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """Per-group statistics plus ``"unified"`` (all groups pooled as the reference's unified script does)."""
    out = {}
    for g, name in enumerate(names[: acc.shape[0]]):
        if float(counts[g]) > 0:
            out[name] = finalize(acc[g], counts[g])
    if float(counts.sum()) > 0:
        out["unified"] = finalize(acc.sum(0), counts.sum())
    return out


def compute_stats(frontend: FrontEnd, batches: Iterable, n_groups: int = 2, group=None):
    """Streams ``(wave (B, L), group_ids (B,) | None[, lengths (B,)])`` batches of THIS rank's shard through
    the fused stats path, all-reduces once, and returns ``(acc, counts)`` (identical on every rank)."""
    acc, counts = frontend.new_stats_accumulator(n_groups)
    for item in batches:
        wave, gids = item[0], item[1]
        lengths = item[2] if len(item) > 2 else None
        frontend.stats_accumulate(wave, acc, counts, lengths=lengths, group_ids=gids)
    allreduce_accumulators(acc, counts, group)
    return acc, counts


def write_reference_npz(out_dir: str, results: Dict[str, Tuple[np.ndarray, np.ndarray]]) -> Dict[str, str]:
    """Writes ``stats_stft_cqt_piano.npz`` / ``stats_stft_cqt_violin.npz`` / ``stats_unified_stft_cqt.npz``
    with the keys and shapes ``DualInstrumentDataset`` loads (``dataloader.py:43-59``, ``:68``)."""
    import os

    names = {"piano": "stats_stft_cqt_piano.npz", "violin": "stats_stft_cqt_violin.npz", "unified": "stats_unified_stft_cqt.npz"}
    paths = {}
    os.makedirs(out_dir, exist_ok=True)
    for key, (mean, std) in results.items():
        paths[key] = os.path.join(out_dir, names.get(key, f"stats_{key}.npz"))
        save_stats_npz(paths[key], mean, std)
    return paths


def main(argv=None) -> int:
    """``compute_separated_stats.py`` / ``compute_unified_stats.py`` as one command::

        python -m audio_style_transfer_b200.stats --piano-dir D1 --violin-dir D2 --out-dir train_set_stats [--batch 32]

    Every rank of a ``torchrun`` launch takes a contiguous shard of each file list (one all-reduce at the end);
    rank 0 writes ``stats_stft_cqt_piano.npz``, ``stats_stft_cqt_violin.npz`` and ``stats_unified_stft_cqt.npz``
    with the keys ``DualInstrumentDataset`` loads."""
    import argparse
    import os

    import torch.distributed as dist

    from .dataloader import _list_audio, load_clips

    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--piano-dir", required=True)
    ap.add_argument("--violin-dir", required=True)
    ap.add_argument("--out-dir", default="train_set_stats")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fe = FrontEnd(torch.device("cuda", local))
    lists = [_list_audio(args.piano_dir), _list_audio(args.violin_dir)]

    def batches():
        for g, files in enumerate(lists):
            mine = [files[i] for i in shard_range(len(files), rank, world)]
            for k in range(0, len(mine), args.batch):
                # per-file errors are printed and the file is left out of the count (compute_separated_stats.py:21-38)
                wave, kept = load_clips(mine[k: k + args.batch], fe, skip_errors=True)
                if wave is None:
                    continue
                yield wave, torch.full((wave.shape[0],), g, dtype=torch.int32, device=fe.device)

    # every rank must reach the all-reduce, whatever happened to its shard: a rank that raised before the collective
    # would leave the others hanging in NCCL.  A failed rank contributes zeros and the error is re-raised afterwards.
    error = None
    acc, counts = fe.new_stats_accumulator(2)
    try:
        for wave, gids in batches():
            fe.stats_accumulate(wave, acc, counts, group_ids=gids)
    except Exception as e:  # noqa: BLE001 - reported below, after the collective
        error = e
        acc.zero_()
        counts.zero_()
    failed = torch.tensor([1.0 if error is not None else 0.0], dtype=torch.float64, device=fe.device)
    allreduce_accumulators(acc, counts)
    if world > 1:
        dist.all_reduce(failed)
    if float(failed[0]) > 0:
        if world > 1:
            dist.destroy_process_group()
        raise RuntimeError(f"statistics pass failed on {int(failed[0])} rank(s)") from error
    if rank == 0:
        paths = write_reference_npz(args.out_dir, finalize_all(acc, counts))
        for key, path in paths.items():
            print(f"{key}: {path}")
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
