"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, its
host-only helpers agree with the oracle, and compute entry points fail loudly without a GPU."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

from oracle import cqt as oc
from oracle import spectral as osp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from audio_style_transfer_b200 import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from audio_style_transfer_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ast_frontend.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ast_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.ast_version()


def test_geometry_helpers_match_reference_goldens(lib, golden):
    g = golden("sections.npz")
    got = np.array([lib.ast_num_sections(T, 287, 96) for T in range(144, 2000)], dtype=np.int32)
    assert np.array_equal(got, g["section_counts_144_2000"])
    got86 = np.array([lib.ast_num_sections(T, 287, 86) for T in range(144, 1200)], dtype=np.int32)
    assert np.array_equal(got86, g["section_counts86_144_1200"])
    for T in list(range(0, 144)):
        assert lib.ast_num_sections(T, 287, 96) == 0
    for W, ov in ((64, 16), (10, 5), (287, 0), (100, 50)):
        for T in range(1, 400):
            assert lib.ast_num_sections(T, W, ov) == osp.n_sections(T, W, ov), (W, ov, T)
    assert lib.ast_num_frames(220500) == 862 and lib.ast_num_frames(40000) == 157
    assert lib.ast_istft_length(862) == 220416 and lib.ast_istft_length(860) == 219904 and lib.ast_istft_length(1) == 0


def test_plan_constants_match_oracle(lib):
    taps = np.zeros(400, dtype=np.float64)
    n = ctypes.c_int32(0)
    assert lib.ast_host_decimator_taps(taps.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 400, ctypes.byref(n)) == 0
    assert n.value == 385
    ref = oc.decimator_taps()
    assert np.abs(taps[:385] - ref).max() < 1e-15
    kern = np.zeros((12, 256, 2), dtype=np.float64)
    assert lib.ast_host_cqt_kernel(kern.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    K = oc.time_domain_kernel()
    assert np.abs(kern[..., 0] + 1j * kern[..., 1] - K).max() < 1e-12 * np.abs(K).max()
    lengths = np.zeros(84, dtype=np.float64)
    assert lib.ast_host_cqt_lengths(lengths.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    freqs = oc.cqt_frequencies()
    ref_len, _ = oc.wavelet_lengths(freqs, oc.SR, oc.relative_bandwidth(freqs))
    assert np.abs(lengths - ref_len).max() < 1e-9


def test_stats_finalize_is_mean_of_means_and_sqrt_mean_var(lib):
    rng = np.random.default_rng(0)
    acc = rng.random((2, 2, 597))
    mean = np.zeros((2, 597), dtype=np.float32)
    std = np.zeros((2, 597), dtype=np.float32)
    rc = lib.ast_stats_finalize(acc.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 7.0,
                                mean.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                std.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    assert rc == 0
    assert np.array_equal(mean, (acc[0] / 7.0).astype(np.float32))
    assert np.array_equal(std, np.sqrt(acc[1] / 7.0).astype(np.float32))
    assert lib.ast_stats_finalize(acc.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 0.0,
                                  mean.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                  std.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) < 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audio_style_transfer_b200 import utilityFunctions as uf
    from audio_style_transfer_b200.frontend import FrontEnd
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FrontEnd()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uf.get_STFT(torch.zeros(1, 4096))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uf.inverse_STFT(torch.zeros(2, 10, 513))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "audio-style-transfer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_hostmem_cpulist_parser_and_scoped_binding():
    # host-side placement helper of the host-buffer API: parser + "never raises, always restores"
    import os

    hostmem = importlib.import_module("audio_style_transfer_b200.hostmem")
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    with hostmem.device_local_affinity(0) as info:      # no CUDA device here: must report, not raise
        assert info["bound"] in (True, False) and "why" in info
    assert os.sched_getaffinity(0) == before


def test_loader_shuffle_order_equals_torch_dataloader():
    """GpuDataLoader draws its permutation exactly as DataLoader(shuffle=True) + RandomSampler do under the same
    torch.manual_seed (the reference's get_dataloader, dataloader.py:172): no file is touched here, only the order."""
    import torch
    from torch.utils.data import DataLoader

    dl = importlib.import_module("audio_style_transfer_b200.dataloader")

    class _Waves:
        batcher = staticmethod(lambda p, v: (p, v))

        def __len__(self):
            return 12

        def waves(self, items):
            return list(items), list(items)

    torch.manual_seed(11)
    want = [int(x) for b in DataLoader(list(range(12)), batch_size=4, shuffle=True, drop_last=True) for x in b[:2]]
    torch.manual_seed(11)
    got = [i for p, _ in dl.GpuDataLoader(_Waves(), 4, True) for i in p]
    assert got == want
