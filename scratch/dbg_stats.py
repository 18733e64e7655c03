import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("audio_style_transfer_b200.synth")
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
lengths = [40000, 52000, 36608]
wave = np.zeros((3, 52000), dtype=np.float32)
for i, L in enumerate(lengths):
    wave[i, :L] = synth.noise_clip(60 + i, L)
for i, L in enumerate(lengths):
    acc, counts = fe.new_stats_accumulator(1)
    w = torch.from_numpy(wave).cuda()
    fe.stats_accumulate(w[i:i+1], acc, counts, lengths=torch.tensor([L]))
    acc1, counts1 = fe.new_stats_accumulator(1)
    fe.stats_accumulate(torch.from_numpy(wave[i, :L]).cuda()[None], acc1, counts1)
    d = (acc - acc1).abs() / acc1.abs().clamp(min=1e-30)
    print(i, L, "stft mean", float(d[0, 0, :, :513].max()), "stft var", float(d[0, 1, :, :513].max()),
          "cqt mean", float(d[0, 0, :, 513:].max()), "cqt var", float(d[0, 1, :, 513:].max()))
    # against the two-pass float64 path on stored features
    feats, frames = fe.features(torch.from_numpy(wave[i, :L]).cuda()[None], layout="flat")
    acc2, counts2 = fe.new_stats_accumulator(1)
    fe.stats_accumulate_features(feats, acc2, counts2)
    d2 = (acc1 - acc2).abs() / acc2.abs().clamp(min=1e-30)
    print("   fused vs two-pass f64: mean", float(d2[0, 0].max()), "var", float(d2[0, 1].max()), "var stft", float(d2[0,1,:,:513].max()), "var cqt", float(d2[0,1,:,513:].max()))
