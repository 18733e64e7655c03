"""load_audio (utilityFunctions.py:105-122): oracle vs the reference's own outputs (CPU), host tap
tables of the C library vs the oracle (CPU, no compute), and the CUDA path vs both (GPU)."""
import ctypes
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import audio as oa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["stereo_44k_cut", "stereo_44k_pad", "mono_44k", "mono_48k", "stereo_16k_up", "mono_22k"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "load_audio.npz"))


def tol(name):
    # 2:1 has 28 taps; the 320:147 and 320:441 filter banks have ~350 taps, summed in float32 by torch's conv1d
    return 5e-7 if "44k" in name or "22k" in name else 2e-5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_load_audio(golden, name):
    sr, cut = golden[f"{name}.meta"]
    y, sr_out = oa.load_audio_from_array(golden[f"{name}.in"], int(sr), 22050, float(cut))
    ref = golden[f"{name}.out"]
    assert sr_out == 22050 and y.shape == ref.shape and y.dtype == np.float32
    assert np.abs(y - ref).max() <= tol(name)


def test_resample_geometry_and_taps_match_torchaudio_formula():
    assert oa.resample_geometry(44100, 22050) == (2, 1, 13)
    assert oa.resample_geometry(48000, 22050) == (320, 147, 14)
    taps = oa.resample_taps(44100, 22050)
    assert taps.shape == (1, 28) and taps.dtype == np.float32
    assert abs(float(taps.sum()) - 1.0) < 2e-3          # unit DC gain up to the window's ripple
    assert np.allclose(taps[0, :27], taps[0, :27][::-1], atol=1e-7) and taps[0, 27] == 0  # symmetric about x[2 m]


@pytest.mark.parametrize("rates", [(44100, 22050), (48000, 22050), (16000, 22050), (22050, 22050)])
def test_library_taps_equal_oracle(rates):
    lib = importlib.import_module("audio_style_transfer_b200._lib").load()
    o, n, w = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    assert lib.ast_resample_geometry(rates[0], rates[1], ctypes.byref(o), ctypes.byref(n), ctypes.byref(w)) == 0
    assert (o.value, n.value, w.value) == oa.resample_geometry(*rates)
    count = n.value * (2 * w.value + o.value)
    buf = (ctypes.c_float * count)()
    assert lib.ast_host_resample_taps(rates[0], rates[1], buf, count) == 0
    got = np.frombuffer(buf, dtype=np.float32).reshape(n.value, -1)
    assert np.abs(got - oa.resample_taps(*rates)).max() <= 1e-7
    for n_in in (0, 1, 6615, 441000):
        assert lib.ast_resample_length(n_in, rates[0], rates[1]) == -(-n_in * n.value // o.value)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def fe():
    frontend = importlib.import_module("audio_style_transfer_b200.frontend")
    return frontend.FrontEnd("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_load_audio_matches_reference_golden(fe, golden, name):
    uf = importlib.import_module("audio_style_transfer_b200.utilityFunctions")
    sr, cut = golden[f"{name}.meta"]
    x = torch.from_numpy(golden[f"{name}.in"])
    y, sr_out = uf.load_audio_tensor(x, int(sr), 22050, float(cut))      # CPU in -> CPU out, like the reference
    ref = golden[f"{name}.out"]
    assert sr_out == 22050 and tuple(y.shape) == ref.shape and y.dtype == torch.float32 and y.device.type == "cpu"
    assert np.abs(y.numpy() - ref).max() <= tol(name)


@pytest.mark.gpu
def test_gpu_load_audio_full_size_batch_and_ragged(fe):
    # BASELINE-size clips: 10 s stereo at 44.1 kHz -> 220 500 mono samples; ragged lengths read zeros past the end
    rng = np.random.default_rng(3)
    B, L = 3, 441000
    x = (0.1 * rng.standard_normal((B, 2, L))).astype(np.float32)
    lengths = np.array([441000, 300001, 17], dtype=np.int32)
    y = fe.load_audio(torch.from_numpy(x), 44100, 22050, 10, lengths=torch.from_numpy(lengths)).cpu().numpy()
    assert y.shape == (B, 220500)
    for b in range(B):
        xb = x[b].copy()
        xb[:, lengths[b]:] = 0
        ref, _ = oa.load_audio_from_array(xb, 44100, 22050, 10)
        assert np.abs(y[b] - ref[0]).max() <= 5e-7
    # linearity (size-independent property) and the exact zero response of an empty clip
    z = fe.load_audio(torch.from_numpy(2.0 * x), 44100, 22050, 10, lengths=torch.from_numpy(lengths)).cpu().numpy()
    assert np.abs(z - 2.0 * y).max() <= 1e-6
    e = fe.load_audio(torch.zeros(1, 2, 1000), 44100, 22050, 10).cpu().numpy()
    assert e.shape == (1, 220500) and not e.any()


@pytest.mark.gpu
def test_gpu_load_audio_wav_file_roundtrip(tmp_path):
    from scipy.io import wavfile

    uf = importlib.import_module("audio_style_transfer_b200.utilityFunctions")
    rng = np.random.default_rng(5)
    pcm = (rng.uniform(-0.5, 0.5, (30000, 2)) * 32767).astype(np.int16)
    path = str(tmp_path / "clip.wav")
    wavfile.write(path, 44100, pcm)
    y, sr = uf.load_audio(path, sample_rate=22050, cut_time_seconds=1)
    assert sr == 22050 and tuple(y.shape) == (1, 22050)
    ref, _ = oa.load_audio_from_array((pcm.T.astype(np.float32) / 32768.0), 44100, 22050, 1)
    assert np.abs(y.numpy() - ref).max() <= 5e-7
    with pytest.raises(NotImplementedError):
        uf.load_audio_tensor(torch.zeros(3, 100), 44100)
