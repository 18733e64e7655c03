"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden, made by
oracle/make_golden.py).  Tolerances: the reference computes in float32 (complex64 FFT); the
oracle in float64, so agreement is limited by the reference's own rounding:
1e-5 x max|X| on STFT bins, 2e-6 abs on iSTFT samples of a 0.07-RMS signal."""
import hashlib
import importlib

import numpy as np
import pytest

from conftest import rel_l2, rel_max
from oracle import spectral as osp

synth = importlib.import_module("audio_style_transfer_b200.synth")


def sums(x):
    x = np.asarray(x, dtype=np.float64)
    return np.array([x.sum(), np.abs(x).sum(), (x * x).sum(), x.size])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_stft_matches_reference(golden, tag):
    g = golden("stft_istft.npz")
    s = osp.get_STFT(g[f"wave_{tag}"])
    assert s.shape == (2, 157, 513) and s.dtype == np.float32
    ref = g[f"stft_{tag}_frames"]
    assert rel_max(s[:, g["frame_idx"], :], ref) < 1e-5
    assert rel_l2(s[:, g["frame_idx"], :], ref) < 1e-6
    got, want = sums(s), g[f"stft_{tag}_sums"]
    assert got[3] == want[3]
    assert abs(got[1] - want[1]) < 1e-5 * want[1]
    assert abs(got[2] - want[2]) < 1e-5 * want[2]
    assert np.all(s[1, :, 0] == 0) and np.all(s[1, :, 512] == 0)


def test_stft_tiny_and_1d_input(golden):
    g = golden("stft_istft.npz")
    s = osp.get_STFT(g["wave_tiny"])
    assert s.shape == g["stft_tiny"].shape == (2, 7, 513)
    assert rel_max(s, g["stft_tiny"]) < 1e-5


@pytest.mark.parametrize("tag", ["a", "b"])
def test_istft_matches_reference(golden, tag):
    g = golden("stft_istft.npz")
    s = osp.get_STFT(g[f"wave_{tag}"], dtype=np.float64)
    y = osp.inverse_STFT(s)
    assert y.shape == (256 * 156,)
    assert np.abs(y[:3072] - g[f"istft_{tag}_head"]).max() < 2e-6
    assert np.abs(y[20000:22048] - g[f"istft_{tag}_mid"]).max() < 2e-6
    assert np.abs(y[-3072:] - g[f"istft_{tag}_tail"]).max() < 2e-6
    want = g[f"istft_{tag}_sums"]
    assert abs(sums(y)[2] - want[2]) < 1e-5 * want[2]


def test_istft_of_arbitrary_spectrogram_ignores_imag_dc_nyquist(golden):
    g = golden("stft_istft.npz")
    y = osp.inverse_STFT(g["spec_rand"])
    assert y.shape == g["istft_rand"].shape == (256 * 11,)
    assert rel_max(y, g["istft_rand"]) < 1e-5
    spec2 = g["spec_rand"].copy()
    spec2[1, :, 0] = 123.0
    spec2[1, :, 512] = -7.0
    assert np.array_equal(osp.inverse_STFT(spec2), y)
    assert rel_max(osp.inverse_STFT(g["stft_tiny"]), g["istft_tiny"]) < 1e-5


def test_full_clip_checksums_and_roundtrip(golden):
    g = golden("stft_istft.npz")
    wave = synth.piano_clip(0)
    sha = np.frombuffer(hashlib.sha256(wave.tobytes()).digest(), dtype=np.uint8)
    if not np.array_equal(sha, g["wave_full_sha256"]):
        pytest.skip("numpy Philox/normal stream differs from the one the golden was made with")
    s = osp.get_STFT(wave)
    assert s.shape == (2, 862, 513)
    assert rel_max(s[:, g["stft_full_idx"], :], g["stft_full_frames"]) < 1e-5
    want = g["stft_full_sums"]
    assert abs(sums(s)[2] - want[2]) < 1e-5 * want[2]
    y = osp.inverse_STFT(osp.get_STFT(wave), dtype=np.float64)  # float32 spectrogram, like the reference
    assert y.shape == (220416,)
    w64 = wave[:220416].astype(np.float64)
    snr = 10 * np.log10(np.sum(w64 ** 2) / max(np.sum((w64 - y) ** 2), 1e-300))
    assert snr > 120.0 and float(g["roundtrip_full_snr_db"]) > 120.0


def test_overlap_windows_match_reference(golden):
    g = golden("sections.npz")
    spec = g["spec"]
    assert np.array_equal(osp.get_overlap_windows(spec), g["windows_default"])
    assert np.array_equal(osp.get_overlap_windows(spec, 287, 86), g["windows_86"])
    assert np.array_equal(osp.get_overlap_windows(spec, 64, 16), g["windows_small"])
    counts = [osp.n_sections(T) for T in range(144, 2000)]
    assert np.array_equal(np.array(counts), g["section_counts_144_2000"])
    # SURVEY §8a a5 closed form for the default geometry
    assert all(c == 1 + (T - 144) // 191 for c, T in zip(counts, range(144, 2000)))
    counts86 = [osp.n_sections(T, 287, 86) for T in range(144, 1200)]
    assert np.array_equal(np.array(counts86), g["section_counts86_144_1200"])
    with pytest.raises(RuntimeError):
        osp.get_overlap_windows(spec[:, :143])


def test_sections2spectrogram_matches_reference(golden):
    g = golden("sections.npz")
    sec = g["sections"]
    for key, (size, ov) in {"merged_96_862": (862, 96), "merged_96_700": (700, 96), "merged_86_890": (890, 86)}.items():
        got = osp.sections2spectrogram(sec, size, ov)
        assert got.shape == g[key].shape
        assert np.abs(got - g[key]).max() < 1e-6
    assert np.array_equal(osp.sections2spectrogram(g["sections_single"], 287), g["merged_single"])
    # cut -> merge is the identity on the covered frames
    spec = g["spec"]
    merged = osp.sections2spectrogram(osp.get_overlap_windows(spec), 500)
    n_cov = 191 * 1 + 287
    assert np.abs(merged[:, :n_cov] - spec[:, :n_cov]).max() < 1e-6


def test_normalize_concat_collate_match_reference(golden, piano_stats):
    g = golden("normalize_collate.npz")
    mean, std = piano_stats
    got = osp.normalize(g["x"], mean[:, :513], std[:, :513])
    ref = g["x_norm_piano"]
    # identical float32 formula -> bit-exact, including the two std == 0 columns (x * 1e8)
    assert np.array_equal(got, ref)
    assert np.array_equal(osp.normalize(g["xq"], mean[:, 513:], std[:, 513:]), g["xq_norm_piano"])
    assert np.array_equal(osp.concat_stft_cqt(g["x"], g["xq"]), g["concat"])
    with pytest.raises(ValueError):
        osp.concat_stft_cqt(g["x"][0], g["xq"])
    with pytest.raises(ValueError):
        osp.concat_stft_cqt(g["x"], g["xq"][:, :5])
    items = [{"piano": g[f"item{i}_piano"], "violin": g[f"item{i}_violin"]} for i in range(4)]
    batch, labels = osp.custom_collate_fn(items)
    assert np.array_equal(batch, g["collate_batch"]) and np.array_equal(labels, g["collate_labels"])
    assert labels.dtype == np.int64


def test_compute_stats_matches_reference(golden):
    g = golden("stats.npz")
    mean, std = osp.compute_stats(list(g["waves"]))
    assert mean.shape == std.shape == (2, 597)
    ref_mean, ref_std = g["mean"], g["std"]
    # mean abs err <= 1e-6 + 1e-4 |mean|, std rel err <= 1e-5 (float32 reference sums)
    assert np.all(np.abs(mean - ref_mean) <= 1e-6 + 1e-4 * np.abs(ref_mean))
    nz = ref_std > 0
    assert np.all(np.abs(std[nz] - ref_std[nz]) <= 2e-5 * ref_std[nz])
    assert np.all(std[~nz] == 0)
    parts = osp.split_stats(mean, std)
    assert parts["stft_mean"].shape == (2, 513) and parts["cqt_std"].shape == (2, 84)
