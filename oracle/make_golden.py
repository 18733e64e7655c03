"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (``/root/reference`` must exist)::

    python -m oracle.make_golden

Everything written here comes out of the reference's own functions
(``utilityFunctions.py`` / ``dataloader.py`` imported through ``oracle/ref_loader.py``,
``compute_stats`` lifted verbatim out of ``Preprocessing_Dataset/compute_separated_stats.py``
with ``ast`` because the script has module-level side effects).  ``get_CQT`` cannot run
(librosa absent) - no CQT golden exists; see ``oracle/cqt.py``.
"""
from __future__ import annotations

import ast
import hashlib
import importlib
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import cqt as oracle_cqt  # noqa: E402

synth = importlib.import_module("audio_style_transfer_b200.synth")


def checksums(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return np.array([x.sum(), np.abs(x).sum(), (x * x).sum(), x.size], dtype=np.float64)


def lift_compute_stats(uf):
    """Compile ``compute_stats`` exactly as written in compute_separated_stats.py:16-43."""
    path = os.path.join(ref_loader.REFERENCE_ROOT, "Preprocessing_Dataset", "compute_separated_stats.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_stats"][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {
        "torch": torch,
        "tqdm": lambda it, **kw: it,
        "load_audio": lambda clip: (clip, 22050),  # "files" are already waveforms
        "get_STFT": uf.get_STFT,
        # librosa is absent: the CQT half comes from the restatement (documented gap)
        "get_CQT": lambda audio: torch.from_numpy(oracle_cqt.get_CQT(audio.numpy())),
        "concat_stft_cqt": lambda a, b: torch.cat((a, b), dim=2),
        "print": lambda *a, **k: None,
    }
    exec(compile(mod, path, "exec"), ns)
    return ns["compute_stats"]


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(1)
    uf, dl = ref_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)

    # ------------------------------------------------------------------ stats artefacts (a11)
    dst = os.path.join(GOLDEN, "train_set_stats")
    os.makedirs(dst, exist_ok=True)
    for name in ("stats_stft_cqt_piano.npz", "stats_stft_cqt_violin.npz", "stats_unified_stft_cqt.npz"):
        shutil.copyfile(os.path.join(ref_loader.REFERENCE_ROOT, "train_set_stats", name), os.path.join(dst, name))
    piano = np.load(os.path.join(dst, "stats_stft_cqt_piano.npz"))

    # ------------------------------------------------------------------ STFT / iSTFT (a1, a8)
    wave_a = synth.piano_clip(1, 40000)
    wave_b = synth.violin_clip(2, 40000)
    wave_tiny = synth.noise_clip(3, 1536)
    out = {"wave_a": wave_a, "wave_b": wave_b, "wave_tiny": wave_tiny}
    frame_idx = np.array([0, 1, 2, 3, 77, 78, 79, 153, 154, 155, 156])
    for tag, w in (("a", wave_a), ("b", wave_b)):
        s = uf.get_STFT(torch.from_numpy(w).unsqueeze(0))
        assert tuple(s.shape) == (2, 157, 513)
        s_np = s.contiguous().numpy()
        out[f"stft_{tag}_frames"] = s_np[:, frame_idx, :]
        out[f"stft_{tag}_sums"] = checksums(s_np)
        y = uf.inverse_STFT(s).numpy()
        assert y.shape == (256 * 156,)
        out[f"istft_{tag}_head"] = y[:3072]
        out[f"istft_{tag}_mid"] = y[20000:22048]
        out[f"istft_{tag}_tail"] = y[-3072:]
        out[f"istft_{tag}_sums"] = checksums(y)
    out["frame_idx"] = frame_idx
    s_tiny = uf.get_STFT(torch.from_numpy(wave_tiny))  # 1-D input path, utilityFunctions.py:21-22
    out["stft_tiny"] = s_tiny.contiguous().numpy()
    out["istft_tiny"] = uf.inverse_STFT(s_tiny).numpy()
    # iSTFT of a non-STFT-consistent spectrogram (decoder output is arbitrary), incl. non-zero imag DC/Nyquist
    g = torch.Generator().manual_seed(11)
    spec_rand = torch.randn(2, 12, 513, generator=g)
    out["spec_rand"] = spec_rand.numpy()
    out["istft_rand"] = uf.inverse_STFT(spec_rand).numpy()

    # full-length clip: waveform regenerated from the seed (hash stored), checksums + a few frames
    wave_full = synth.piano_clip(0)
    s_full = uf.get_STFT(torch.from_numpy(wave_full).unsqueeze(0)).contiguous().numpy()
    assert s_full.shape == (2, 862, 513)
    full_idx = np.array([0, 1, 2, 430, 431, 859, 860, 861])
    out["wave_full_sha256"] = np.frombuffer(hashlib.sha256(wave_full.tobytes()).digest(), dtype=np.uint8)
    out["stft_full_frames"] = s_full[:, full_idx, :]
    out["stft_full_idx"] = full_idx
    out["stft_full_sums"] = checksums(s_full)
    y_full = uf.inverse_STFT(torch.from_numpy(s_full)).numpy()
    assert y_full.shape == (220416,)
    out["istft_full_sums"] = checksums(y_full)
    out["roundtrip_full_snr_db"] = np.array(
        10 * np.log10(np.sum(wave_full[:220416].astype(np.float64) ** 2)
                      / np.sum((wave_full[:220416].astype(np.float64) - y_full) ** 2)))
    np.savez_compressed(os.path.join(GOLDEN, "stft_istft.npz"), **out)

    # ------------------------------------------------------------------ sections (a5, a7)
    out = {}
    g = torch.Generator().manual_seed(5)
    spec = torch.randn(2, 500, 7, generator=g)
    out["spec"] = spec.numpy()
    out["windows_default"] = uf.get_overlap_windows(spec).numpy()
    out["windows_86"] = uf.get_overlap_windows(spec, window_size=287, overlap_frames=86).numpy()
    out["windows_small"] = uf.get_overlap_windows(spec, window_size=64, overlap_frames=16).numpy()
    counts = []
    probe = torch.zeros(2, 2000, 1)
    for T in range(144, 2000):
        counts.append(uf.get_overlap_windows(probe[:, :T]).shape[0])
    out["section_counts_144_2000"] = np.array(counts, dtype=np.int32)
    counts86 = [uf.get_overlap_windows(probe[:, :T], 287, 86).shape[0] for T in range(144, 1200)]
    out["section_counts86_144_1200"] = np.array(counts86, dtype=np.int32)
    sec = torch.randn(4, 2, 287, 5, generator=g)
    out["sections"] = sec.numpy()
    out["merged_96_862"] = uf.sections2spectrogram(sec, 862).numpy()
    out["merged_96_700"] = uf.sections2spectrogram(sec, 700).numpy()
    out["merged_86_890"] = uf.sections2spectrogram(sec, 890, overlap=86).numpy()
    sec1 = torch.randn(1, 2, 287, 5, generator=g)
    out["sections_single"] = sec1.numpy()
    out["merged_single"] = uf.sections2spectrogram(sec1, 287).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "sections.npz"), **out)

    # ------------------------------------------------------------------ normalise / concat / collate (a3, a4, a6)
    out = {}
    x = torch.randn(2, 20, 513, generator=g) * 3
    mean = torch.tensor(piano["stft_mean"]).float()
    std = torch.tensor(piano["stft_std"]).float()
    out["x"] = x.numpy()
    out["x_norm_piano"] = dl.normalize(x, mean, std).numpy()
    xq = torch.randn(2, 20, 84, generator=g)
    out["xq"] = xq.numpy()
    out["xq_norm_piano"] = dl.normalize(xq, torch.tensor(piano["cqt_mean"]), torch.tensor(piano["cqt_std"])).numpy()
    out["concat"] = uf.concat_stft_cqt(x, xq).numpy()
    items = []
    for i in range(4):
        items.append({"piano": torch.randn(2, 2, 5, 3, generator=g), "violin": torch.randn(2, 2, 5, 3, generator=g),
                      "piano_label": 0, "violin_label": 1})
    for i, it in enumerate(items):
        out[f"item{i}_piano"] = it["piano"].numpy()
        out[f"item{i}_violin"] = it["violin"].numpy()
    batch, labels = dl.custom_collate_fn(items)
    out["collate_batch"] = batch.numpy()
    out["collate_labels"] = labels.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "normalize_collate.npz"), **out)

    # ------------------------------------------------------------------ dataset statistics (a10)
    compute_stats = lift_compute_stats(uf)
    clips = [torch.from_numpy(synth.piano_clip(10 + i, 40000)).unsqueeze(0) for i in range(3)]
    clips += [torch.from_numpy(synth.violin_clip(20 + i, 40000)).unsqueeze(0) for i in range(2)]
    mean, std = compute_stats(clips, "golden")
    np.savez_compressed(
        os.path.join(GOLDEN, "stats.npz"),
        clip_kinds=np.array([0, 0, 0, 1, 1], dtype=np.int32),
        clip_ids=np.array([10, 11, 12, 20, 21], dtype=np.int32),
        n_samples=np.array(40000),
        waves=np.stack([c.numpy()[0] for c in clips]),
        mean=mean.numpy(),
        std=std.numpy(),
    )
    for f in sorted(os.listdir(GOLDEN)):
        p = os.path.join(GOLDEN, f)
        if os.path.isfile(p):
            print(f"{f:32s} {os.path.getsize(p):9d} B")


if __name__ == "__main__":
    main()
