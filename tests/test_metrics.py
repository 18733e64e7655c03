"""mse_spectrogram (evaluation_reconstruction.py:105-118): oracle vs torch.stft with constant padding
(CPU; librosa itself is not installable here), CUDA path vs the oracle (GPU)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import metrics as om


def signals(seed, n):
    synth = importlib.import_module("audio_style_transfer_b200.synth")
    a = synth.piano_clip(seed, n)
    rng = np.random.default_rng(seed)
    b = (a * rng.uniform(0.8, 1.2) + 0.01 * rng.standard_normal(n)).astype(np.float32)
    return a, b


@pytest.mark.parametrize("n", [300, 22050, 50001])
def test_oracle_stft_magnitude_equals_torch_constant_pad(n):
    a, _ = signals(1, n)
    ref = torch.stft(torch.from_numpy(a), 1024, 256, window=torch.hann_window(1024), center=True, pad_mode="constant",
                     return_complex=True).abs().numpy()
    got = om.librosa_stft_mag(a)
    assert got.shape == ref.shape == (513, 1 + n // 256)
    assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1.0)


def test_oracle_mse_properties():
    a, b = signals(2, 30000)
    assert om.mse_spectrogram(a, a) == 0.0
    m = om.mse_spectrogram(a, b)
    assert m > 0 and abs(om.mse_spectrogram(b, a) - m) <= 1e-12
    # min_time cropping (:111-113): a longer second signal only contributes its first frames
    b_long = np.concatenate([b, np.ones(4096, np.float32)])
    sa, sb = om.librosa_stft_mag(a), om.librosa_stft_mag(b_long)
    assert sb.shape[1] == sa.shape[1] + 16
    assert om.mse_spectrogram(a, b_long) == float(np.mean((sa - sb[:, : sa.shape[1]]) ** 2))


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb", [(300, 300), (22050, 22050), (219904, 220500), (50001, 40000)])
def test_gpu_mse_spectrogram_matches_oracle(na, nb):
    ev = importlib.import_module("audio_style_transfer_b200.evaluation")
    a, _ = signals(3, na)
    _, b = signals(3, nb)
    ref = om.mse_spectrogram(a, b)
    got = ev.mse_spectrogram(a, b)
    assert isinstance(got, float) and abs(got - ref) <= 1e-5 * ref + 1e-12
    assert ev.mse_spectrogram(a, a) == 0.0
    # tensors on the device are accepted, and the call is deterministic
    got2 = ev.mse_spectrogram(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert got2 == got


@pytest.mark.gpu
def test_gpu_reconstruct_audio_from_sections_mirror():
    ev = importlib.import_module("audio_style_transfer_b200.evaluation")
    from oracle import spectral as osp

    rng = np.random.default_rng(0)
    sec = rng.standard_normal((1, 4, 2, 287, 513)).astype(np.float32)
    y = ev.reconstruct_audio_from_sections(torch.from_numpy(sec))
    ref = osp.reconstruct_audio_from_sections(sec)
    assert y.shape == ref.shape == (73216,) and np.abs(y - ref).max() <= 2e-5
    bad = ev.reconstruct_audio_from_sections(torch.zeros(1, 4, 2, 287, 100))   # wrong bin count -> the reference's fallback
    assert bad.shape == (22050,) and not bad.any()
