"""mse_spectrogram (evaluation_reconstruction.py:105-118) and instrumentation_similarity
(evaluation_style_transfer.py:111-119), SURVEY.md 8f-4: oracle vs torch.stft with constant padding (CPU; librosa
itself is not installable here), CUDA path vs the oracle (GPU)."""
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
    y = ev.reconstruct_audio_from_sections(torch.from_numpy(sec), 0, 3)   # three positional arguments, as evaluation_reconstruction.py:345-350 calls it
    assert np.array_equal(y, ev.reconstruct_audio_from_sections(torch.from_numpy(sec)))
    ref = osp.reconstruct_audio_from_sections(sec)
    assert y.shape == ref.shape == (73216,) and np.abs(y - ref).max() <= 2e-5
    bad = ev.reconstruct_audio_from_sections(torch.zeros(1, 4, 2, 287, 100))   # wrong bin count -> the reference's fallback
    assert bad.shape == (22050,) and not bad.any()


@pytest.mark.parametrize("n", [700, 22050, 50001])
def test_oracle_default_stft_magnitude_equals_torch_constant_pad(n):
    # librosa.stft defaults (instrumentation_similarity): n_fft 2048, hop 512
    a, _ = signals(4, n)
    ref = torch.stft(torch.from_numpy(a), 2048, 512, window=torch.hann_window(2048), center=True, pad_mode="constant",
                     return_complex=True).abs().numpy()
    got = om.librosa_stft_mag(a, n_fft=2048, hop_length=512)
    assert got.shape == ref.shape == (1025, 1 + n // 512)
    assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1.0)


def test_oracle_instrumentation_similarity_properties():
    a, b = signals(5, 40000)
    assert abs(om.instrumentation_similarity(a, a) - 1.0) <= 1e-6
    r = om.instrumentation_similarity(a, b)
    assert -1.0 <= r <= 1.0 and abs(om.instrumentation_similarity(b, a) - r) <= 1e-6
    # a gain does not change the correlation; silence has a constant profile -> NaN -> 0.0 (:119)
    assert abs(om.instrumentation_similarity(a, 0.25 * b) - r) <= 1e-5
    assert om.instrumentation_similarity(a, np.zeros(40000, np.float32)) == 0.0
    # a sine against noise: the noise profile is flat, the sine's is a single peak
    t = np.arange(40000) / 22050.0
    tone = (0.3 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    noise = np.random.default_rng(0).standard_normal(40000).astype(np.float32)
    assert abs(om.instrumentation_similarity(tone, noise)) < 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb", [(300, 900), (700, 700), (22050, 22050), (219904, 220500), (50001, 40000), (661500, 220500)])
def test_gpu_instrumentation_similarity_matches_oracle(na, nb):
    ev = importlib.import_module("audio_style_transfer_b200.evaluation")
    a, _ = signals(6, max(na, 64))
    _, b = signals(7, max(nb, 64))
    a, b = a[:na], b[:nb]
    ref = om.instrumentation_similarity(a, b)
    got = ev.instrumentation_similarity(a, b)
    assert isinstance(got, float) and abs(got - ref) <= 2e-5
    assert abs(ev.instrumentation_similarity(a, a) - 1.0) <= 1e-6
    got2 = ev.instrumentation_similarity(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert got2 == got                     # deterministic; device tensors accepted


@pytest.mark.gpu
def test_gpu_instrumentation_similarity_silence_and_errors():
    ev = importlib.import_module("audio_style_transfer_b200.evaluation")
    a, _ = signals(8, 30000)
    assert ev.instrumentation_similarity(a, np.zeros(30000, np.float32)) == 0.0
    assert ev.instrumentation_similarity(np.zeros(100, np.float32), np.zeros(100, np.float32)) == 0.0
    with pytest.raises(RuntimeError):
        ev.instrumentation_similarity(a, np.zeros(0, np.float32))
