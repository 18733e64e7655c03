"""configs[2] alone: the ragged 64-clip batch of bench.py (lengths U{44100..220500}, seed 7), ms per feature call."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
dl = importlib.import_module("audio_style_transfer_b200.dataloader")
mean, std = dl.load_stats_npz(bench.STATS_NPZ); mean, std = mean.cuda(), std.cuda()
wave = torch.from_numpy(bench.make_clips()).cuda()
g7 = torch.Generator().manual_seed(7)
L = torch.randint(44100, 220501, (64,), generator=g7, dtype=torch.int64).to(torch.int32).cuda()
out = torch.empty((64, 4, 2, 287, 597), device="cuda")
for _ in range(3): fe.features(wave, lengths=L, mean=mean, std=std, layout="sections", out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): fe.features(wave, lengths=L, mean=mean, std=std, layout="sections", out=out)
e1.record(); torch.cuda.synchronize()
print("ragged ms per step", e0.elapsed_time(e1) / 30, "valid audio-s/s", float(L.sum()) / 22050 / (e0.elapsed_time(e1) / 30e3))
