"""The N > 1 host logic on CPU (gloo, world_size 2): contiguous sharding of clips over ranks and the
single all-reduce of partial per-bin moments.  The moments themselves are synthetic
here (the feature kernels need a GPU); what is checked is that sharded + all-reduced == serial."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_style_transfer_b200 import stats
    rng = np.random.default_rng(123)  # same stream on every rank: "clip i" is reproducible anywhere
    clip_mean = rng.standard_normal((n_clips, 2, 597))
    clip_var = rng.random((n_clips, 2, 597))
    kinds = np.arange(n_clips) % 2
    acc = torch.zeros(2, 2, 2, 597, dtype=torch.float64)
    counts = torch.zeros(2, dtype=torch.float64)
    for i in stats.shard_range(n_clips, rank, world):
        acc[kinds[i], 0] += torch.from_numpy(clip_mean[i])
        acc[kinds[i], 1] += torch.from_numpy(clip_var[i])
        counts[kinds[i]] += 1
    stats.allreduce_accumulators(acc, counts)
    res = stats.finalize_all(acc, counts)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), acc=acc.numpy(), counts=counts.numpy(),
             **{f"{k}_mean": v[0] for k, v in res.items()}, **{f"{k}_std": v[1] for k, v in res.items()})
    if rank == 0:
        stats.write_reference_npz(os.path.join(out_dir, "npz"), res)
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from audio_style_transfer_b200 import stats
    for n in (0, 1, 7, 8, 100000, 12501):
        for world in (1, 2, 3, 4, 8):
            got = [i for r in range(world) for i in stats.shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(stats.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert len(stats.shard_range(100000, 3, 8)) == 12500


def test_two_rank_allreduce_equals_serial(tmp_path):
    n_clips = 37
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_clips, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in r0.files:
        assert np.array_equal(r0[k], r1[k]), k  # every rank ends with identical statistics
    rng = np.random.default_rng(123)
    clip_mean = rng.standard_normal((n_clips, 2, 597))
    clip_var = rng.random((n_clips, 2, 597))
    kinds = np.arange(n_clips) % 2
    assert r0["counts"].tolist() == [19.0, 18.0]
    for g, name in enumerate(("piano", "violin")):
        m = clip_mean[kinds == g].mean(0)
        s = np.sqrt(clip_var[kinds == g].mean(0))
        assert np.allclose(r0[f"{name}_mean"], m, rtol=1e-6, atol=1e-7)
        assert np.allclose(r0[f"{name}_std"], s, rtol=1e-6)
    assert np.allclose(r0["unified_mean"], clip_mean.mean(0), rtol=1e-6, atol=1e-7)
    assert np.allclose(r0["unified_std"], np.sqrt(clip_var.mean(0)), rtol=1e-6)
    # the written files follow the reference's npz contract (dataloader.py:48-59)
    from audio_style_transfer_b200 import dataloader as dl
    for fname in ("stats_stft_cqt_piano.npz", "stats_stft_cqt_violin.npz", "stats_unified_stft_cqt.npz"):
        z = np.load(tmp_path / "npz" / fname)
        assert sorted(z.files) == ["cqt_mean", "cqt_std", "stft_mean", "stft_std"]
        assert z["stft_mean"].shape == (2, 513) and z["cqt_std"].shape == (2, 84) and z["stft_std"].dtype == np.float32
        mean, std = dl.load_stats_npz(str(tmp_path / "npz" / fname))
        assert tuple(mean.shape) == (2, 597)
