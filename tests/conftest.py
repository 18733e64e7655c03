import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def piano_stats():
    z = np.load(os.path.join(GOLDEN, "train_set_stats", "stats_stft_cqt_piano.npz"))
    mean = np.concatenate([z["stft_mean"], z["cqt_mean"]], axis=1).astype(np.float32)
    std = np.concatenate([z["stft_std"], z["cqt_std"]], axis=1).astype(np.float32)
    return mean, std


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.sqrt(((a - b) ** 2).sum() / max((b * b).sum(), 1e-300))
