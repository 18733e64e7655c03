"""N > 1 on real GPUs (skipped with fewer than 2 devices): clips sharded over ranks, one NCCL all-reduce of the
partial per-bin moments, result identical on every rank and equal to the single-GPU pass (config 4 at test size)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CLIPS, N_SAMPLES = 12, 40000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _clips():
    synth = importlib.import_module("audio_style_transfer_b200.synth")
    waves = np.stack([synth.clip("piano" if i % 2 == 0 else "violin", 100 + i, N_SAMPLES) for i in range(N_CLIPS)])
    kinds = np.arange(N_CLIPS) % 2
    return waves, kinds


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from audio_style_transfer_b200 import stats
    from audio_style_transfer_b200.frontend import FrontEnd
    fe = FrontEnd(f"cuda:{rank}")
    waves, kinds = _clips()
    mine = list(stats.shard_range(N_CLIPS, rank, world))
    batches = []
    for i in range(0, len(mine), 4):  # several batches per rank: accumulation across calls
        idx = mine[i : i + 4]
        batches.append((torch.from_numpy(waves[idx]).cuda(), torch.from_numpy(kinds[idx].astype(np.int32)).cuda()))
    acc, counts = stats.compute_stats(fe, batches, n_groups=2)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), acc=acc.cpu().numpy(), counts=counts.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_stats_with_nccl_allreduce(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["acc"], r1["acc"]) and np.array_equal(r0["counts"], r1["counts"])
    assert r0["counts"].tolist() == [6.0, 6.0]
    from audio_style_transfer_b200.frontend import FrontEnd
    fe = FrontEnd("cuda:0")
    waves, kinds = _clips()
    acc, counts = fe.new_stats_accumulator(2)
    fe.stats_accumulate(torch.from_numpy(waves).cuda(), acc, counts, group_ids=torch.from_numpy(kinds.astype(np.int32)))
    torch.cuda.synchronize()
    # same per-clip moments, different summation order across ranks: float64 sums agree to ~1e-15 relative
    assert np.allclose(r0["acc"], acc.cpu().numpy(), rtol=1e-12, atol=1e-18)
