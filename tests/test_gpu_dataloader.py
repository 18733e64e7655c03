"""get_dataloader / DualInstrumentDataset / the stats command on real files: WAV files in
two directories -> batches laid out as dataloader.py:123-172 does, compared with the CPU oracle run file by file."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import audio as oa
from oracle import spectral as osp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATS_DIR = os.path.join(ROOT, "tests", "golden", "train_set_stats")


def write_dataset(tmp_path, n_files=4):
    from scipy.io import wavfile

    synth = importlib.import_module("audio_style_transfer_b200.synth")
    dirs = {}
    clips = {}
    for g, name in enumerate(("piano", "violin")):
        d = tmp_path / name
        d.mkdir()
        dirs[name] = str(d)
        for i in range(n_files):
            n = 44100 * 3 + 1000 * i                      # 3 s stereo at 44.1 kHz: padded to 10 s by load_audio
            mono = synth.piano_clip(50 + 10 * g + i, n) if g == 0 else synth.violin_clip(50 + 10 * g + i, n)
            stereo = np.stack([mono, 0.5 * mono[::-1]], axis=1)
            pcm = np.clip(stereo * 32767.0 * 4.0, -32767, 32767).astype(np.int16)
            path = str(d / f"clip_{i:02d}.wav")
            wavfile.write(path, 44100, pcm)
            clips[path] = pcm.T.astype(np.float32) / 32768.0
        (d / "notes.txt").write_text("ignored: not .wav / .mp3")
    return dirs, clips


def oracle_wave(clip_cxl):
    y, _ = oa.load_audio_from_array(clip_cxl, 44100, 22050, 10)
    return y[0]


@pytest.mark.gpu
def test_get_dataloader_batches_match_reference_layout(tmp_path, monkeypatch):
    dl = importlib.import_module("audio_style_transfer_b200.dataloader")
    dirs, clips = write_dataset(tmp_path)
    monkeypatch.setattr(dl, "DEFAULT_STATS_DIR", STATS_DIR)
    # SpectralBatcher's default argument was bound at definition time: build the dataset explicitly as well
    ds = dl.DualInstrumentDataset(dirs["piano"], dirs["violin"], stats_dir=STATS_DIR)
    loader = dl.GpuDataLoader(ds, batch_size=4, shuffle=False)
    assert len(ds) == 4 and len(loader) == 1
    assert [os.path.basename(p) for p in ds.piano_files] == [f"clip_{i:02d}.wav" for i in range(4)]
    batches = list(loader)
    assert len(batches) == 1
    x, labels = batches[0]
    assert tuple(x.shape) == (4, 4, 2, 287, 597) and x.dtype == torch.float32 and x.is_cuda
    assert labels.tolist() == [0, 0, 1, 1] and labels.dtype == torch.int64
    # rows 0, 1 = piano items 0, 1; rows 2, 3 = violin items 0, 1 (custom_collate_fn keeps the first B/2 items)
    pm, ps = dl.load_stats_npz(os.path.join(STATS_DIR, "stats_stft_cqt_piano.npz"))
    vm, vs = dl.load_stats_npz(os.path.join(STATS_DIR, "stats_stft_cqt_violin.npz"))
    xs = x.cpu().numpy()
    for row, (files, m, s) in zip((0, 2), ((ds.piano_files, pm, ps), (ds.violin_files, vm, vs))):
        wave = oracle_wave(clips[files[0]])
        ref = osp.features_sections(wave, m.numpy(), s.numpy())
        raw = osp.get_overlap_windows(osp.clip_features(wave))
        scale = 1.0 / (s.numpy() + 1e-8)                                     # per-column error amplification
        err = np.abs(xs[row] - ref) / scale[None, :, None, :]
        assert err[..., :513].max() <= 1e-5 * np.abs(raw[..., :513]).max()
        assert err[..., 513:].max() <= 1e-5 * np.abs(raw[..., 513:]).max()
    # __getitem__ returns the reference's dict, equal to the batch rows
    item = ds[1]
    assert set(item) == {"piano", "violin", "piano_label", "violin_label"} and item["violin_label"] == 1
    assert torch.equal(item["piano"], x[1]) and torch.equal(item["violin"], x[3])
    # odd batch sizes are rounded down with the reference's warning
    monkeypatch.chdir(tmp_path)
    os.symlink(STATS_DIR, tmp_path / "train_set_stats")
    loader2 = dl.get_dataloader(dirs["piano"], dirs["violin"], batch_size=3, shuffle=False)
    assert loader2.batch_size == 2 and len(loader2) == 2
    x2, l2 = next(iter(loader2))
    assert tuple(x2.shape) == (2, 4, 2, 287, 597) and l2.tolist() == [0, 1]
    assert torch.equal(x2[0], x[0]) and torch.equal(x2[1], x[2])


@pytest.mark.gpu
def test_stats_command_writes_reference_npz(tmp_path):
    stats = importlib.import_module("audio_style_transfer_b200.stats")
    dl = importlib.import_module("audio_style_transfer_b200.dataloader")
    dirs, clips = write_dataset(tmp_path, n_files=2)
    out_dir = str(tmp_path / "out")
    assert stats.main(["--piano-dir", dirs["piano"], "--violin-dir", dirs["violin"], "--out-dir", out_dir, "--batch", "2"]) == 0
    files = sorted(os.listdir(out_dir))
    assert files == ["stats_stft_cqt_piano.npz", "stats_stft_cqt_violin.npz", "stats_unified_stft_cqt.npz"]
    z = np.load(os.path.join(out_dir, "stats_stft_cqt_piano.npz"))
    assert sorted(z.files) == ["cqt_mean", "cqt_std", "stft_mean", "stft_std"]
    assert z["stft_mean"].shape == (2, 513) and z["cqt_std"].shape == (2, 84) and z["stft_std"].dtype == np.float32
    piano = sorted(p for p in clips if os.sep + "piano" + os.sep in p)
    mean, std = osp.compute_stats([oracle_wave(clips[p]) for p in piano])
    got_mean = np.concatenate([z["stft_mean"], z["cqt_mean"]], axis=1)
    got_std = np.concatenate([z["stft_std"], z["cqt_std"]], axis=1)
    assert np.abs(got_mean - mean).max() <= 1e-6 + 1e-4 * np.abs(mean).max()
    nz = std > 1e-6
    assert (np.abs(got_std - std)[nz] / std[nz]).max() <= 2e-5
    # the files load through the reference's contract
    m, s = dl.load_stats_npz(os.path.join(out_dir, "stats_unified_stft_cqt.npz"))
    assert tuple(m.shape) == (2, 597) and tuple(s.shape) == (2, 597)


@pytest.mark.gpu
def test_synth_clips_are_a_pure_function_of_the_clip_id():
    """ast_synth_clips (the configs[3] workload generator): any chunking / any rank regenerates the same clip bit for
    bit, piano-like below violin_from_id and violin-like from there on, RMS near the dataset's 0.07."""
    fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
    whole = fe.synth_clips(12, first_clip_id=100, violin_from_id=106, n_samples=50000)
    assert tuple(whole.shape) == (12, 50000) and whole.dtype == torch.float32
    parts = torch.cat([fe.synth_clips(5, first_clip_id=100, violin_from_id=106, n_samples=50000),
                       fe.synth_clips(7, first_clip_id=105, violin_from_id=106, n_samples=50000)])
    assert torch.equal(whole, parts)
    rms = whole.double().pow(2).mean(1).sqrt()
    assert float(rms.min()) > 0.02 and float(rms.max()) < 0.2 and bool(torch.isfinite(whole).all())
    assert not torch.equal(whole[0], whole[1])
    # a prefix of a clip is the same clip (sample i depends on (id, i) only, up to the onset scaling by the length)
    buf = torch.zeros((3, 50008), device="cuda")
    out = fe.synth_clips(3, first_clip_id=100, violin_from_id=106, n_samples=50000, out=buf)
    assert torch.equal(out, whole[:3]) and float(buf[:, 50000:].abs().max()) == 0.0
