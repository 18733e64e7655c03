"""Parity of the CUDA path (through the C-ABI) against the CPU oracle and the golden vectors captured
from the unmodified reference.  Tolerances (stated per test):

* STFT bins        max-abs <= 1e-5 x max|X| per clip, rel-L2 <= 1e-6, imag DC / Nyquist exactly 0
* CQT bins         max-abs <= 1e-5 x max|V| vs the restated oracle (same decimator taps)
* normalised bins  the same raw tolerance scaled by 1 / (std + eps) per column (std has exact zeros)
* iSTFT            max-abs <= 2e-6 on 0.07-RMS audio, round-trip SNR >= 120 dB
* stats            mean abs err <= 1e-6 + 1e-4 |mean|, std rel err <= 1e-5 (STD_RTOL)
"""
import importlib

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max
from oracle import cqt as oc
from oracle import spectral as osp

pytestmark = pytest.mark.gpu

synth = importlib.import_module("audio_style_transfer_b200.synth")


@pytest.fixture(scope="module")
def fe():
    from audio_style_transfer_b200.frontend import FrontEnd
    return FrontEnd("cuda:0")


@pytest.fixture(scope="module")
def uf():
    return importlib.import_module("audio_style_transfer_b200.utilityFunctions")


@pytest.fixture(scope="module")
def dl():
    return importlib.import_module("audio_style_transfer_b200.dataloader")


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------------------- STFT (a1)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_stft_matches_reference_golden(fe, golden, tag):
    g = golden("stft_istft.npz")
    out = fe.stft(cuda(g[f"wave_{tag}"])[None])[0].cpu().numpy()
    assert out.shape == (2, 157, 513) and out.dtype == np.float32
    ref = g[f"stft_{tag}_frames"]
    got = out[:, g["frame_idx"], :]
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    assert rel_l2(got, ref) <= 1e-6
    want = g[f"stft_{tag}_sums"]
    o64 = out.astype(np.float64)
    assert abs((o64 * o64).sum() - want[2]) <= 1e-5 * want[2]
    assert np.all(out[1, :, 0] == 0.0) and np.all(out[1, :, 512] == 0.0)
    full = osp.get_STFT(g[f"wave_{tag}"], dtype=np.float64)
    assert np.abs(out - full).max() <= 1e-5 * np.abs(full).max()


def test_stft_tiny_clip_and_dropin_signature(uf, golden):
    g = golden("stft_istft.npz")
    w = torch.from_numpy(g["wave_tiny"])
    out = uf.get_STFT(w)  # 1-D CPU input -> CPU output, like the reference (utilityFunctions.py:21-22)
    assert out.device.type == "cpu" and tuple(out.shape) == (2, 7, 513)
    assert rel_max(out.numpy(), g["stft_tiny"]) <= 1e-5
    out2 = uf.get_STFT(w.cuda().unsqueeze(0))
    assert out2.is_cuda and torch.equal(out2.cpu(), out)
    with pytest.raises(RuntimeError):
        uf.get_STFT(torch.zeros(1, 512))  # torch.stft: reflect pad must be smaller than the input
    with pytest.raises(NotImplementedError):
        uf.get_STFT(w, n_fft=2048)


def test_stft_known_answers(fe):
    n = 40000
    t = np.arange(n) / 22050.0
    # pure tone at bin centre k: |X[k]| = A * sum(w) / 2 = 256 A
    k, A = 100, 0.3
    tone = (A * np.sin(2 * np.pi * (k * 22050.0 / 1024) * t)).astype(np.float32)
    out = fe.stft(cuda(tone)[None])[0].cpu().numpy()
    mag = np.hypot(out[0], out[1])
    assert abs(mag[80, k] - 256 * A) < 1e-3 * 256 * A and mag[80].argmax() == k
    # zero input -> exact zeros
    assert torch.count_nonzero(fe.stft(torch.zeros(2, n, device="cuda"))) == 0
    # linearity
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    sa, sb, sab = (fe.stft(cuda(x)[None])[0].cpu().numpy().astype(np.float64) for x in (a, b, 2 * a - 3 * b))
    assert np.abs(sab - (2 * sa - 3 * sb)).max() <= 2e-5 * np.abs(sab).max()


# ------------------------------------------------------------------------------- CQT (a2)
@pytest.mark.parametrize("kind,n", [("piano", 40000), ("violin", 40000), ("noise", 36608), ("chirp", 60001)])
def test_cqt_matches_oracle(fe, kind, n):
    w = synth.clip(kind, 5, n)
    V = oc.cqt(w)  # (84, T) complex128
    out = fe.cqt(cuda(w)[None])[0].cpu().numpy()
    T = 1 + n // 256
    assert out.shape == (2, T, 84)
    got = out[0].T + 1j * out[1].T
    scale = np.abs(V).max()
    assert np.abs(got - V).max() <= 1e-5 * scale, np.abs(got - V).max() / scale
    # tensor-core path: the accumulate rounds toward zero: ~3e-6; the FMA path gives ~6e-7
    assert rel_l2(out, np.stack([V.real.T, V.imag.T])) <= 5e-6


def test_cqt_known_answers_and_dropin(fe, uf):
    freqs = oc.cqt_frequencies()
    lengths, _ = oc.wavelet_lengths(freqs, oc.SR, oc.relative_bandwidth(freqs))
    n = 110250
    t = np.arange(n) / 22050.0
    for k in (3, 45, 80):
        w = (0.3 * np.sin(2 * np.pi * freqs[k] * t)).astype(np.float32)
        out = fe.cqt(cuda(w)[None])[0].cpu().numpy()
        mag = np.hypot(out[0], out[1])
        mid = mag.shape[0] // 2
        assert mag[mid].argmax() == k
        assert abs(mag[mid, k] - 0.3 * np.sqrt(lengths[k]) / 2) < 2e-3 * mag[mid, k]
    assert torch.count_nonzero(fe.cqt(torch.zeros(1, 220500, device="cuda"))) == 0
    # reference signature: tensor or ndarray in, (2, T, 84) float32 out; shape pinned by test_correctness.ipynb cell 3
    w = synth.piano_clip(0)
    a = uf.get_CQT(torch.from_numpy(w).unsqueeze(0))
    b = uf.get_CQT(w[None, :])
    assert tuple(a.shape) == (2, 862, 84) and a.dtype == torch.float32 and a.device.type == "cpu"
    assert torch.equal(a, b)
    with pytest.raises(NotImplementedError):
        uf.get_CQT(w, n_bins=96)


# ------------------------------------------------------------------------------- features (a1-a6)
def _check_normalised(got, raw_ref, mean, std, tol_raw, eps=1e-8):
    """got: normalised CUDA output (2, T, F); raw_ref: un-normalised oracle (float64)."""
    want = (raw_ref - mean[:, None, :].astype(np.float64)) / (std[:, None, :].astype(np.float64) + eps)
    bound = tol_raw / (std[:, None, :].astype(np.float64) + eps) + 2e-6 * np.abs(want)
    err = np.abs(got.astype(np.float64) - want)
    assert np.all(err <= bound), float((err / bound).max())


def test_config1_full_clip_sections_normalised(fe, piano_stats):
    """BASELINE config 1: one 10 s clip -> (4, 2, 287, 597) normalised with the shipped piano stats."""
    mean, std = piano_stats
    w = synth.piano_clip(0)
    out, counts = fe.features(cuda(w)[None], mean=cuda(mean), std=cuda(std), layout="sections")
    assert tuple(out.shape) == (1, 4, 2, 287, 597) and out.is_contiguous() and int(counts[0]) == 4
    out = out[0].cpu().numpy()
    raw = np.concatenate([osp.get_STFT(w, dtype=np.float64),
                          np.stack([oc.cqt(w).real.T, oc.cqt(w).imag.T])], axis=2)  # (2, 862, 597)
    sx, sv = np.abs(raw[..., :513]).max(), np.abs(raw[..., 513:]).max()
    for s in range(4):
        seg = raw[:, s * 191 : s * 191 + 287]
        _check_normalised(out[s][..., :513], seg[..., :513], mean[:, :513], std[:, :513], 1e-5 * sx)
        _check_normalised(out[s][..., 513:], seg[..., 513:], mean[:, 513:], std[:, 513:], 1e-5 * sv)
    # the two std == 0 columns (imag DC / Nyquist) must be exactly (0 - mean) / 1e-8 = 0
    assert np.all(out[:, 1, :, 0] == 0.0) and np.all(out[:, 1, :, 512] == 0.0)
    # same thing through the oracle's float32 pipeline (dataloader.py:100-112)
    ref32 = osp.features_sections(w, mean, std)
    nz = std > 1e-3
    assert rel_max(out[:, :, :, nz[0]][:, 0], ref32[:, :, :, nz[0]][:, 0]) < 1e-4


def test_features_flat_raw_equals_separate_calls(fe):
    w = cuda(synth.batch(3, 40000))
    flat, counts = fe.features(w, layout="flat")
    assert tuple(flat.shape) == (3, 2, 157, 597) and counts.tolist() == [157] * 3
    assert torch.equal(flat[..., :513], fe.stft(w))
    assert torch.equal(flat[..., 513:], fe.cqt(w))
    # sections of raw features == get_overlap_windows of the flat tensor (evaluation_style_transfer.py:135-139)
    sec, cnt = fe.features(w, layout="sections")
    assert cnt.tolist() == [1, 1, 1]
    ref = osp.get_overlap_windows(flat[0].cpu().numpy())
    assert np.array_equal(sec[0].cpu().numpy(), ref)


def test_config3_variable_length_batch(fe, piano_stats):
    """BASELINE config 3: ragged padded batch; each clip must equal the run on the unpadded clip."""
    mean, std = piano_stats
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(44100, 120001, (6,), generator=g).tolist() + [36608, 120000]
    lmax = 120000
    wave = np.zeros((len(lengths), lmax), dtype=np.float32)
    for i, L in enumerate(lengths):
        wave[i, :L] = synth.clip("piano" if i % 2 == 0 else "violin", 30 + i, L)
        wave[i, L:] = 7.0  # garbage in the padding must never be read
    out, counts = fe.features(cuda(wave), lengths=torch.tensor(lengths), mean=cuda(mean), std=cuda(std))
    s_max = osp.n_sections(1 + lmax // 256)
    assert tuple(out.shape) == (len(lengths), s_max, 2, 287, 597)
    for i, L in enumerate(lengths):
        single, c1 = fe.features(cuda(wave[i, :L])[None], mean=cuda(mean), std=cuda(std))
        s_i = osp.n_sections(1 + L // 256)
        assert int(counts[i]) == s_i == int(c1[0]) == single.shape[1]
        assert torch.equal(out[i, :s_i], single[0]), i
        assert torch.count_nonzero(out[i, s_i:]) == 0
        # rows past the clip's last frame inside the last section are zeros (padding after normalisation)
        T_i = 1 + L // 256
        last_rows = T_i - (s_i - 1) * 191
        if last_rows < 287:
            assert torch.count_nonzero(out[i, s_i - 1, :, last_rows:]) == 0
    flat, frames = fe.features(cuda(wave), lengths=torch.tensor(lengths), layout="flat")
    for i, L in enumerate(lengths):
        T_i = 1 + L // 256
        assert int(frames[i]) == T_i and torch.count_nonzero(flat[i, :, T_i:]) == 0
        ref = osp.get_STFT(wave[i, :L], dtype=np.float64)
        assert np.abs(flat[i, :, :T_i, :513].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("batch,n", [(1, 36608), (3, 70001), (150, 44100), (301, 40000)])
def test_cqt_batch_shapes_exercise_the_decimator_tile_scheduler(fe, batch, n):
    """The six decimator stages run as ONE launch whose tiles wait on each other through completion counters:
    cover fewer tiles than CTAs, several tiles per CTA and more clips than SMs.  Every clip of the batch must equal
    the same clip run alone (bit for bit: the arithmetic per tile does not depend on the schedule), and one of them
    the oracle."""
    base = np.stack([synth.clip("piano" if i % 2 == 0 else "violin", 200 + i, n) for i in range(min(batch, 5))])
    gains = np.random.default_rng(batch).uniform(0.5, 1.5, batch).astype(np.float32)
    wave = base[np.arange(batch) % len(base)] * gains[:, None]
    out = fe.cqt(cuda(wave))
    torch.cuda.synchronize()
    assert tuple(out.shape) == (batch, 2, 1 + n // 256, 84)
    for i in sorted({0, batch // 2, batch - 1}):
        single = fe.cqt(cuda(wave[i])[None])
        assert torch.equal(out[i], single[0]), i
    V = oc.cqt(wave[batch - 1].astype(np.float64))
    ref = np.stack([V.real.T, V.imag.T])
    assert np.abs(out[batch - 1].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("batch", [1, 2, 9])
def test_back_to_back_calls_share_a_workspace_safely(fe, piano_stats, batch):
    """The feature call is a chain of programmatic dependent launches (prologue -> decimator -> CQT projection -> STFT)
    whose next call starts behind "the previous kernel" only; with few clips the STFT finishes long before the CQT
    projection, so this is the case where a missing ordering would let call k + 1 rewrite the statistics table / zero the
    completion counters under call k.  40 un-synchronised calls on alternating inputs (and an iSTFT straight after a
    feature call) must equal the same calls made one at a time."""
    mean, std = piano_stats
    mean_d, std_d = torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    n = 60000
    xs = [cuda(np.stack([synth.clip("piano" if (i + j) % 2 else "violin", 300 + 7 * j + i, n) for i in range(batch)]))
          for j in range(2)]
    refs = []
    for x in xs:
        sec, counts = fe.features(x, mean=mean_d, std=std_d, layout="sections")
        torch.cuda.synchronize()
        refs.append((sec.clone(), counts.clone()))
    outs = []
    for k in range(40):
        sec, counts = fe.features(xs[k % 2], mean=mean_d if k % 3 else mean_d * 1.0, std=std_d, layout="sections")
        if k >= 38:
            outs.append((k % 2, sec, counts))
        else:
            y = fe.istft(sec, layout="sections", overlap=96)      # reads the STFT columns of the call just enqueued
    torch.cuda.synchronize()
    for which, sec, counts in outs:
        assert torch.equal(sec, refs[which][0]) and torch.equal(counts, refs[which][1])
    y_ref = fe.istft(refs[1][0], layout="sections", overlap=96)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)


def test_per_clip_statistics_and_batcher(fe, dl, piano_stats):
    import os
    from conftest import GOLDEN
    batcher = dl.SpectralBatcher(stats_dir=os.path.join(GOLDEN, "train_set_stats"), frontend=fe)
    p = cuda(np.stack([synth.piano_clip(40, 40000), synth.piano_clip(41, 40000)]))
    v = cuda(np.stack([synth.violin_clip(42, 40000), synth.violin_clip(43, 40000)]))
    batch, labels = batcher(p, v)
    assert tuple(batch.shape) == (4, 1, 2, 287, 597) and labels.tolist() == [0, 0, 1, 1] and labels.dtype == torch.int64
    pm, ps = dl.load_stats_npz(os.path.join(GOLDEN, "train_set_stats", "stats_stft_cqt_piano.npz"))
    vm, vs = dl.load_stats_npz(os.path.join(GOLDEN, "train_set_stats", "stats_stft_cqt_violin.npz"))
    assert torch.equal(batch[:2], fe.features(p, mean=pm, std=ps)[0])
    assert torch.equal(batch[2:], fe.features(v, mean=vm, std=vs)[0])
    unified = dl.SpectralBatcher(use_separate_stats=False, stats_dir=os.path.join(GOLDEN, "train_set_stats"), frontend=fe)
    um, us = dl.load_stats_npz(os.path.join(GOLDEN, "train_set_stats", "stats_unified_stft_cqt.npz"))
    assert torch.equal(unified(p, v)[0][2:], fe.features(v, mean=um, std=us)[0])


# ------------------------------------------------------------------------------- iSTFT (a7-a9)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_istft_matches_reference_golden(fe, uf, golden, tag):
    g = golden("stft_istft.npz")
    spec = torch.from_numpy(osp.get_STFT(g[f"wave_{tag}"]))
    y = uf.inverse_STFT(spec)
    assert y.device.type == "cpu" and tuple(y.shape) == (256 * 156,)
    y = y.numpy()
    assert np.abs(y[:3072] - g[f"istft_{tag}_head"]).max() <= 2e-6
    assert np.abs(y[20000:22048] - g[f"istft_{tag}_mid"]).max() <= 2e-6
    assert np.abs(y[-3072:] - g[f"istft_{tag}_tail"]).max() <= 2e-6
    want = g[f"istft_{tag}_sums"]
    assert abs((y.astype(np.float64) ** 2).sum() - want[2]) <= 1e-5 * want[2]


def test_istft_arbitrary_spectrogram(fe, uf, golden):
    g = golden("stft_istft.npz")
    y = uf.inverse_STFT(torch.from_numpy(g["spec_rand"]).cuda())
    assert y.is_cuda and tuple(y.shape) == (256 * 11,)
    assert rel_max(y.cpu().numpy(), g["istft_rand"]) <= 1e-5
    spec2 = g["spec_rand"].copy()
    spec2[1, :, 0] = 123.0
    spec2[1, :, 512] = -7.0  # torch.istft ignores the imaginary part of DC / Nyquist
    assert torch.equal(uf.inverse_STFT(torch.from_numpy(spec2).cuda()), y)
    assert rel_max(uf.inverse_STFT(torch.from_numpy(g["stft_tiny"])).numpy(), g["istft_tiny"]) <= 1e-5
    assert tuple(uf.inverse_STFT(torch.zeros(2, 1, 513)).shape) == (0,)
    # a 597-wide feature tensor is accepted by the batched call; only [:513] is read (test_correctness.ipynb cell 11)
    wide = torch.cat([torch.from_numpy(g["spec_rand"]), torch.randn(2, 12, 84)], dim=2).cuda()
    assert torch.equal(fe.istft(wide[None])[0], y)


@pytest.mark.parametrize("overlap,original", [(96, 862), (96, 700), (86, 890), (96, 0)])
def test_istft_from_sections_equals_merge_then_istft(fe, overlap, original):
    """sections2spectrogram + inverse_STFT fused (style_transfer_inference_test.ipynb cell 4:36-42)."""
    g = torch.Generator().manual_seed(3)
    sec = torch.randn(2, 4, 2, 287, 513, generator=g)
    got = fe.istft(sec.cuda(), layout="sections", overlap=overlap, original_size=original).cpu().numpy()
    for b in range(2):
        merged = osp.sections2spectrogram(sec[b].numpy(), original if original > 0 else 10**9, overlap)
        want = osp.inverse_STFT(merged, dtype=np.float64)
        assert got[b].shape == want.shape
        assert np.abs(got[b] - want).max() <= 1e-5 * np.abs(want).max()


def test_reconstruct_audio_from_sections_variant(fe):
    """evaluation_reconstruction.py:161-189: iSTFT of section 0 only -> 73 216 samples."""
    g = torch.Generator().manual_seed(4)
    sec = torch.randn(1, 4, 2, 287, 513, generator=g)
    want = osp.reconstruct_audio_from_sections(sec.numpy())
    got = fe.istft(sec[:, 0].cuda(), layout="flat")[0].cpu().numpy()
    assert got.shape == want.shape == (73216,)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_roundtrip_snr_full_clip(fe, golden):
    w = synth.piano_clip(0)
    spec = fe.stft(cuda(w)[None])
    y = fe.istft(spec)[0].cpu().numpy().astype(np.float64)
    assert y.shape == (220416,)
    w64 = w[:220416].astype(np.float64)
    snr = 10 * np.log10((w64**2).sum() / max(((w64 - y) ** 2).sum(), 1e-300))
    assert snr >= 120.0, snr
    # features -> [:513] -> merge -> iSTFT (config 1's loop): covers frames 0..859 -> 219 904 samples
    sec, _ = fe.features(cuda(w)[None])
    y2 = fe.istft(sec, layout="sections", overlap=96, original_size=862)[0].cpu().numpy().astype(np.float64)
    assert y2.shape == (219904,)
    # the last 3 hops see a truncated overlap-add (frames 860/861 dropped by the section cut): compare the interior
    n = 219904 - 1024
    snr2 = 10 * np.log10((w64[:n] ** 2).sum() / max(((w64[:n] - y2[:n]) ** 2).sum(), 1e-300))
    assert snr2 >= 120.0, snr2


# ------------------------------------------------------------------------------- small operators
def test_small_operators_match_reference_golden(fe, uf, dl, golden, piano_stats):
    g = golden("sections.npz")
    spec = torch.from_numpy(g["spec"])
    assert np.array_equal(uf.get_overlap_windows(spec).numpy(), g["windows_default"])
    assert np.array_equal(uf.get_overlap_windows(spec.cuda(), 287, 86).cpu().numpy(), g["windows_86"])
    assert np.array_equal(uf.get_overlap_windows(spec, 64, 16).numpy(), g["windows_small"])
    with pytest.raises(RuntimeError):
        uf.get_overlap_windows(spec[:, :143])
    sec = torch.from_numpy(g["sections"])
    for key, (size, ov) in {"merged_96_862": (862, 96), "merged_96_700": (700, 96), "merged_86_890": (890, 86)}.items():
        got = uf.sections2spectrogram(sec, size, ov).numpy()
        assert got.shape == g[key].shape and np.array_equal(got, g[key]), key
    assert np.array_equal(uf.sections2spectrogram(torch.from_numpy(g["sections_single"]), 287).numpy(), g["merged_single"])
    n = golden("normalize_collate.npz")
    mean, std = piano_stats
    got = dl.normalize(torch.from_numpy(n["x"]), torch.from_numpy(mean[:, :513]), torch.from_numpy(std[:, :513]))
    assert got.device.type == "cpu" and np.array_equal(got.numpy(), n["x_norm_piano"])  # same float32 formula: bit-exact
    got = dl.normalize(torch.from_numpy(n["xq"]).cuda(), torch.from_numpy(mean[:, 513:]), torch.from_numpy(std[:, 513:]))
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), n["xq_norm_piano"])
    cat = uf.concat_stft_cqt(torch.from_numpy(n["x"]), torch.from_numpy(n["xq"]))
    assert np.array_equal(cat.numpy(), n["concat"])
    assert dl.concat_stft_cqt(torch.from_numpy(n["x"]), torch.from_numpy(n["xq"])).is_cuda  # dataloader.py:15-18
    with pytest.raises(ValueError):
        uf.concat_stft_cqt(torch.zeros(2, 5), torch.zeros(2, 5, 3))
    with pytest.raises(ValueError):
        uf.concat_stft_cqt(torch.zeros(2, 5, 3), torch.zeros(2, 6, 3))
    items = [{"piano": torch.from_numpy(n[f"item{i}_piano"]), "violin": torch.from_numpy(n[f"item{i}_violin"])} for i in range(4)]
    batch, labels = dl.custom_collate_fn(items)
    assert np.array_equal(batch.numpy(), n["collate_batch"]) and np.array_equal(labels.numpy(), n["collate_labels"])


# ------------------------------------------------------------------------------- stats (a10)
# std relative error <= 1e-5 (the golden is the reference's own float32 compute_stats)
STD_RTOL = 1e-5


def test_stats_match_reference_golden(fe, golden):
    g = golden("stats.npz")
    acc, counts = fe.new_stats_accumulator(1)
    fe.stats_accumulate(cuda(g["waves"][:3]), acc, counts)
    fe.stats_accumulate(cuda(g["waves"][3:]), acc, counts)  # accumulation across calls
    torch.cuda.synchronize()
    assert counts.tolist() == [5.0]
    stats = importlib.import_module("audio_style_transfer_b200.stats")
    mean, std = stats.finalize(acc[0], counts[0])
    ref_mean, ref_std = g["mean"], g["std"]
    assert np.all(np.abs(mean - ref_mean) <= 1e-6 + 1e-4 * np.abs(ref_mean))
    nz = ref_std > 0
    rel = np.abs(std[nz] - ref_std[nz]) / ref_std[nz]
    print(f"stats vs the reference's compute_stats: std rel err max {rel.max():.2e} (STFT columns "
          f"{(np.abs(std[:, :513] - ref_std[:, :513])[nz[:, :513]] / ref_std[:, :513][nz[:, :513]]).max():.2e}), "
          f"mean abs err max {np.abs(mean - ref_mean).max():.2e}")
    assert np.all(rel <= STD_RTOL)
    assert np.all(std[~nz] == 0)
    # per-instrument groups + unified = sum of the groups (compute_unified_stats.py walks both trees)
    acc2, counts2 = fe.new_stats_accumulator(2)
    fe.stats_accumulate(cuda(g["waves"]), acc2, counts2, group_ids=torch.from_numpy(g["clip_kinds"]))
    assert counts2.tolist() == [3.0, 2.0]
    assert torch.allclose(acc2.sum(0), acc[0], rtol=1e-12, atol=1e-12)
    o_mean, o_std = osp.compute_stats(list(g["waves"][:3]))
    m0, s0 = stats.finalize(acc2[0], counts2[0])
    assert np.all(np.abs(m0 - o_mean) <= 1e-6 + 1e-4 * np.abs(o_mean))
    nz = o_std > 0
    assert np.all(np.abs(s0[nz] - o_std[nz]) <= STD_RTOL * o_std[nz])


def test_stats_ragged_and_constant(fe):
    lengths = [40000, 52000, 36608]
    wave = np.zeros((3, 52000), dtype=np.float32)
    for i, L in enumerate(lengths):
        wave[i, :L] = synth.noise_clip(60 + i, L)
    acc, counts = fe.new_stats_accumulator(1)
    fe.stats_accumulate(cuda(wave), acc, counts, lengths=torch.tensor(lengths))
    acc1, counts1 = fe.new_stats_accumulator(1)
    for i, L in enumerate(lengths):
        fe.stats_accumulate(cuda(wave[i, :L])[None], acc1, counts1)
    assert torch.allclose(acc, acc1, rtol=1e-9, atol=1e-12) and counts.tolist() == counts1.tolist() == [3.0]
    # a constant spectrogram has zero variance
    feats = torch.full((2, 2, 50, 597), 1.5, device="cuda")
    acc, counts = fe.new_stats_accumulator(1)
    fe.stats_accumulate_features(feats, acc, counts)
    assert torch.all(acc[0, 0] == 3.0) and torch.all(acc[0, 1] == 0.0) and counts.tolist() == [2.0]
    # accumulators reach the kernels as raw pointers: anything but the float64 device buffers is refused up front
    with pytest.raises(ValueError):
        fe.stats_accumulate(cuda(wave), acc.cpu(), counts)
    with pytest.raises(ValueError):
        fe.stats_accumulate(cuda(wave), acc.float(), counts)
    with pytest.raises(ValueError):
        fe.stats_accumulate_features(feats, acc, counts.cpu())


# ------------------------------------------------------------------------------- full-size properties
def test_full_batch_properties(fe, piano_stats):
    """BASELINE config 2 size (64 clips x 10 s): size-independent properties instead of the slow oracle."""
    mean, std = piano_stats
    rng = np.random.default_rng(5)
    base = synth.batch(4)
    wave = torch.from_numpy(base).cuda().repeat(16, 1)  # 64 clips
    gains = torch.from_numpy(rng.uniform(0.5, 2.0, 64).astype(np.float32)).cuda()
    wave = wave * gains[:, None]
    raw, _ = fe.features(wave, layout="sections")
    assert tuple(raw.shape) == (64, 4, 2, 287, 597)
    # homogeneity: features(g x) = g features(x), clip by clip
    for i in (5, 38, 63):
        ref = raw[i % 4] * (gains[i] / gains[i % 4])
        assert (raw[i] - ref).abs().max() <= 2e-5 * ref.abs().max()
    # batch invariance: clip i of the batch == the same clip alone (bit-exact: same kernels, same order)
    alone, _ = fe.features(wave[17:18], layout="sections")
    assert torch.equal(alone[0], raw[17])
    # overlapping rows of consecutive sections are identical copies
    assert torch.equal(raw[:, 0, :, 191:], raw[:, 1, :, :96])
    # normalised = (raw - mean) / (std + eps) column-wise
    nrm, _ = fe.features(wave, mean=cuda(mean), std=cuda(std), layout="sections")
    m = cuda(mean)[None, None, :, None, :]
    s = cuda(std)[None, None, :, None, :]
    want = (raw - m) / (s + 1e-8)
    ok = (nrm - want).abs() <= 2e-6 * want.abs() + 1e-30
    assert bool(ok.all())
    # round trip through the decoder-shaped tensor
    y = fe.istft(raw[..., :513].contiguous(), layout="sections", overlap=96, original_size=862)
    assert tuple(y.shape) == (64, 219904)
    n = 219904 - 1024
    err = (y[:, :n] - wave[:, :n]).double().pow(2).sum(1)
    sig = wave[:, :n].double().pow(2).sum(1)
    assert float((10 * torch.log10(sig / err)).min()) >= 120.0


def test_full_batch_against_oracle(fe, piano_stats):
    """BASELINE configs[1] at full size AGAINST THE ORACLE: 64 x 10 s clips built from 4 distinct clips (2 piano-like,
    2 violin-like) at 64 different gains.  STFT and CQT are linear, so the fp64 oracle of clip i is gain_i x the oracle
    of its base clip; every one of the 64 x (4, 2, 287, 597) outputs is compared, raw and normalised."""
    mean, std = piano_stats
    base = synth.batch(4, first_id=40)
    gains = np.random.default_rng(9).uniform(0.6, 1.6, 64).astype(np.float32)
    wave = np.tile(base, (16, 1)) * gains[:, None]
    raw_ref = []
    for i in range(4):
        V = oc.cqt(base[i])
        raw_ref.append(np.concatenate([osp.get_STFT(base[i], dtype=np.float64), np.stack([V.real.T, V.imag.T])], axis=2))
    raw, counts = fe.features(cuda(wave), layout="sections")
    nrm, _ = fe.features(cuda(wave), mean=cuda(mean), std=cuda(std), layout="sections")
    assert counts.tolist() == [4] * 64
    raw, nrm = raw.cpu().numpy(), nrm.cpu().numpy()
    worst_s = worst_c = 0.0
    for i in range(64):
        # the clip really fed to the kernels is float32(base * gain): its oracle differs from gain x oracle(base) by
        # the rounding of the product, <= 2^-24 relative per sample - far below the 1e-5 bound
        ref = raw_ref[i % 4] * float(gains[i])
        sx, sv = np.abs(ref[..., :513]).max(), np.abs(ref[..., 513:]).max()
        for s in range(4):
            seg = ref[:, s * 191 : s * 191 + 287]
            worst_s = max(worst_s, np.abs(raw[i, s][..., :513] - seg[..., :513]).max() / sx)
            worst_c = max(worst_c, np.abs(raw[i, s][..., 513:] - seg[..., 513:]).max() / sv)
            if i % 8 == 3:  # the normalised tensor of every 8th clip, all sections
                _check_normalised(nrm[i, s][..., :513], seg[..., :513], mean[:, :513], std[:, :513], 1e-5 * sx)
                _check_normalised(nrm[i, s][..., 513:], seg[..., 513:], mean[:, 513:], std[:, 513:], 1e-5 * sv)
    print(f"full batch vs oracle: STFT {worst_s:.2e}, CQT {worst_c:.2e} of max")
    assert worst_s <= 1e-5 and worst_c <= 1e-5, (worst_s, worst_c)


def test_stft_tapered_tail_deals_every_frame_once(fe, piano_stats, monkeypatch):
    """The STFT kernel deals the last clips of a large batch as half / quarter rows of the grid (shorter CTAs at the end:
    stft.cu, launch_stft).  Which CTA transforms a frame pair must not change a bit of it: the default taper, none, and
    odd / clamped settings give identical tensors (flat and sections layouts, ragged lengths included)."""
    mean, std = (cuda(a) for a in piano_stats)
    small = cuda(synth.batch(40, 52000))               # 2 pairs per warp here: half rows only
    lengths = torch.tensor([52000 - 997 * (i % 9) for i in range(40)], dtype=torch.int32)
    big = fe.synth_clips(64)                           # configs[1]: 4 pairs per warp, three waves of CTAs, quarter rows too
    monkeypatch.setenv("AST_STFT_TAPER", "0")
    ref_flat = fe.stft(small, lengths=lengths).clone()
    ref_feat = fe.features(small, mean=mean, std=std, layout="flat")[0][..., :513].clone()
    ref_big = fe.features(big, mean=mean, std=std, layout="sections")[0][..., :513].clone()
    for setting in (None, "3,2", "5", "0,7", "64,64", "1,0", "2,61"):
        if setting is None:
            monkeypatch.delenv("AST_STFT_TAPER")
        else:
            monkeypatch.setenv("AST_STFT_TAPER", setting)
        assert torch.equal(fe.stft(small, lengths=lengths), ref_flat), setting
        assert torch.equal(fe.features(small, mean=mean, std=std, layout="flat")[0][..., :513], ref_feat), setting
        assert torch.equal(fe.features(big, mean=mean, std=std, layout="sections")[0][..., :513], ref_big), setting


def test_rows_that_are_not_16_byte_aligned_take_the_register_path(fe):
    """A batch whose row stride is not a multiple of 4 floats cannot be described by a tensor map: the CQT projection
    falls back to its register-staged producers (cqt_tc_kernel<0>) and must give the same numbers as the TMA path gives
    on an aligned copy of the same clips."""
    base = synth.batch(3, 40002)                      # stride 40002: 8-byte aligned rows only
    x = cuda(base)
    assert x.stride(0) % 4 == 2
    got, _ = fe.features(x, layout="flat")
    padded = torch.zeros((3, 40004), device="cuda")
    padded[:, :40002] = x
    ref, _ = fe.features(padded[:, :40002], layout="flat")   # stride 40004: TMA path
    assert padded[:, :40002].stride(0) % 4 == 0
    assert torch.equal(got[..., :513], ref[..., :513])
    assert (got[..., 513:] - ref[..., 513:]).abs().max() <= 1e-5 * ref[..., 513:].abs().max()
    V = oc.cqt(base[1])
    want = np.stack([V.real.T, V.imag.T])
    assert rel_max(got[1, :, :, 513:].cpu().numpy(), want) <= 1e-5


def test_features_host_pipeline_equals_device_call(fe, piano_stats):
    """The host-buffer API (chunked H2D / kernels / D2H over several streams) returns exactly what the
    device-resident call returns, for chunk sizes that do and do not divide the batch."""
    mean, std = piano_stats
    wave = torch.from_numpy(synth.batch(7, 40000))
    ref, _ = fe.features(wave.cuda(), mean=cuda(mean), std=cuda(std))
    for chunk, pinned in ((3, True), (16, False), (2, True)):
        host_in = wave.pin_memory() if pinned else wave
        host_out = torch.empty(tuple(ref.shape), dtype=torch.float32)
        host_out = host_out.pin_memory() if pinned else host_out
        got = fe.features_host(host_in, host_out, mean=cuda(mean), std=cuda(std), chunk=chunk)
        assert got is host_out and torch.equal(host_out, ref.cpu()), chunk
    with pytest.raises(ValueError):
        fe.features_host(wave, torch.empty(1, 2, 3))


def test_features_host_full_size_chunks_repeatedly(fe, piano_stats):
    """The persistent tensor-core grids of a feature call wait on counters written by their own CTAs, so two feature
    calls must never be in flight at once (include/ast_frontend.h).  features_host therefore issues every kernel on ONE
    stream and only the copies on side streams: 50 passes over 64 full-length clips in chunks of 16 (each chunk's grids
    are 148 CTAs wide, as in the bench) over 3 rotating buffers must finish and reproduce the device-resident call."""
    mean, std = piano_stats
    wave = torch.from_numpy(np.tile(synth.batch(4, first_id=8), (16, 1)))
    ref, _ = fe.features(wave.cuda(), mean=cuda(mean), std=cuda(std))
    ref = ref.cpu()
    host_in = wave.pin_memory()
    host_out = torch.empty(tuple(ref.shape), dtype=torch.float32).pin_memory()
    for it in range(50):
        if it % 10 == 0:
            host_out.zero_()
        fe.features_host(host_in, host_out, mean=cuda(mean), std=cuda(std), chunk=16, n_streams=3)
        if it % 10 == 0 or it == 49:
            assert torch.equal(host_out, ref), it


@pytest.mark.parametrize("env", [{"AST_DECIMATOR": "fma"}, {"AST_CQT": "fma"}, {"AST_DECIMATOR": "fma", "AST_CQT": "fma"},
                                 {"AST_OVERLAP": "0"}, {"AST_CQT_TMA": "0"}, {"AST_DECIMATOR": "tf32"}])
def test_diagnostic_kernel_variants_agree(fe, tmp_path, env):
    """The FMA-pipe twins of the two tensor-core kernels (AST_DECIMATOR=fma, AST_CQT=fma), the TF32-split decimator
    (AST_DECIMATOR=tf32; the default splits into FP16 pairs) and the serial launch order (AST_OVERLAP=0) are diagnostics
    of the same library, selected when a plan is created: run them in a fresh process
    and compare with the default path (and so, transitively, with the oracle)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    wave = np.stack([synth.clip("piano", 300, 50000), synth.clip("violin", 301, 50000)])
    np.save(tmp_path / "wave.npy", wave)
    code = (
        "import importlib, sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "fe = importlib.import_module('audio_style_transfer_b200.frontend').FrontEnd('cuda:0')\n"
        f"x = torch.from_numpy(np.load({str(tmp_path / 'wave.npy')!r})).cuda()\n"
        "f, n = fe.features(x, layout='flat')\n"
        f"np.save({str(tmp_path / 'out.npy')!r}, f.cpu().numpy())\n")
    subprocess.run([sys.executable, "-c", code], check=True, env={**os.environ, **env}, timeout=300)
    got = np.load(tmp_path / "out.npy")
    ref, _ = fe.features(cuda(wave), layout="flat")
    ref = ref.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got[..., :513], ref[..., :513])                       # the STFT does not depend on the switches
    assert np.abs(got[..., 513:] - ref[..., 513:]).max() <= 1e-5 * np.abs(ref[..., 513:]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("gain", [2.0 ** -30, 1e-4, 37.0, 3.0e4, 2.0 ** 40])
def test_cqt_is_homogeneous_over_the_float_range(fe, gain):
    """The decimator stages its operands as FP16 pairs scaled PER TILE by a power of two taken from the tile's own largest
    magnitude, so the result must not depend on the input's scale: CQT(g x) = g CQT(x) for gains far outside FP16's range
    (a fixed scale would overflow at |x| > 65504 / 2^k and flush small signals to zero).  The clip has a loud and a
    40 dB quieter half, so one clip holds tiles of very different scales too.  librosa's CQT (utilityFunctions.py:52) is
    linear; the tolerance is the CQT parity tolerance."""
    wave = synth.clip("violin", 77, 60000).copy()
    wave[30000:] *= 0.01
    base = fe.cqt(cuda(wave[None]))[0].cpu().numpy().astype(np.float64)
    scaled = (wave.astype(np.float64) * gain).astype(np.float32)
    got = fe.cqt(cuda(scaled[None]))[0].cpu().numpy().astype(np.float64) / gain
    assert np.isfinite(got).all()
    assert np.abs(got - base).max() <= 1e-5 * np.abs(base).max()
    # the quiet half on its own terms: relative to ITS largest value (frames well inside it)
    q0 = 30000 // 256 + 40   # (the lowest octave's filters span 64 frames)
    assert np.abs(got[:, q0:] - base[:, q0:]).max() <= 2e-5 * np.abs(base[:, q0:]).max()


@pytest.mark.gpu
def test_cqt_does_not_depend_on_the_batch_a_clip_travels_in(fe):
    """The decimator deals its tiles differently with the batch size (chain stages on a few CTAs up to 64 clips, round
    robin beyond; a CTA's first 64 tiles come from a decoded table in shared memory, later ones are decoded on the
    fly: 320 full-length clips give a CTA ~ 67) and the CQT projection's CTAs draw their tiles from a queue in the
    feature call.  None of this may change a single bit of a clip's result: every tile's arithmetic, including its FP16
    staging scale, depends on the tile alone.  (custom_collate_fn, dataloader.py:123-147, batches clips independently.)"""
    base = np.stack([synth.clip("piano" if i % 2 == 0 else "violin", 500 + i, 220500) for i in range(4)])
    gains = np.random.default_rng(3).uniform(0.5, 1.5, 320).astype(np.float32)
    wave = base[np.arange(320) % 4] * gains[:, None]
    big = fe.cqt(cuda(wave)).cpu().numpy()
    for n in (1, 3, 64, 65):
        small = fe.cqt(cuda(wave[:n])).cpu().numpy()
        assert np.array_equal(small, big[:n]), n
    # the fused feature call (queue-fed projection, STFT in between) against the CQT-only call
    feats, _ = fe.features(cuda(wave[:64]), layout="flat")
    assert np.array_equal(feats[..., 513:].cpu().numpy(), big[:64])
