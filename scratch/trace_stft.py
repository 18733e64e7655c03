"""Cycles per phase of a frame pair in stft_kernel (diagnostic build: scratch/build_variant.sh trace -DAST_STFT_TRACE,
copied over the package library).  usage: python scratch/trace_stft.py"""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
dl = importlib.import_module("audio_style_transfer_b200.dataloader")
mean, std = dl.load_stats_npz(bench.STATS_NPZ)
wave = torch.from_numpy(bench.make_clips()).cuda()
out = torch.empty((64, 4, 2, 287, 597), device="cuda")
os.environ.setdefault("AST_OVERLAP", "1")
fe.features(wave, mean=mean.cuda(), std=std.cuda(), layout="sections", out=out)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 8))()
fe.lib.ast_debug_stft_trace(buf)
a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
names = ["loop/tail", "loads+window", "fft32 #1", "twiddle+exchange", "fft32 #2", "separate+normalise+stage", "fence+flush", "-"]
pairs = 12  # iterations per warp at the bench size
tot = a[:, :7].sum(1).mean()
print("mean cycles per pair per warp:", tot / pairs)
for i, n in enumerate(names[:7]):
    print(f"  {n:28s} {a[:, i].mean() / pairs:9.0f}  {100 * a[:, i].mean() / tot:5.1f}%")

buf2 = (ctypes.c_ulonglong * (4096 * 3))()
fe.lib.ast_debug_stft_cta_times(buf2)
t = np.array(buf2[:], dtype=np.uint64).reshape(4096, 3)[:576].astype(np.int64)
t0 = t[:, 0].min()
start, end, sm = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2]
print(f"CTAs: start us min/median/max {start.min():.1f}/{np.median(start):.1f}/{start.max():.1f}; end {end.min():.1f}/{np.median(end):.1f}/{end.max():.1f}; "
      f"duration {np.min(end - start):.1f}/{np.median(end - start):.1f}/{np.max(end - start):.1f}")
per_sm = np.bincount(sm.astype(int), minlength=148)
print("CTAs per SM: min/max", per_sm.min(), per_sm.max(), "SMs used", (per_sm > 0).sum())
order = np.argsort(start)
print("start times of CTAs 0,100,200,...:", [round(float(start[order[i]]), 1) for i in range(0, 576, 64)])
dur = end - start
bx, by = np.arange(576) % 9, np.arange(576) // 9
print("duration by blockIdx.x (tile in clip):", [round(float(dur[bx == i].mean()), 1) for i in range(9)])
print("duration by clip (first 16):", [round(float(dur[by == i].mean()), 1) for i in range(16)])
sm_mean = np.array([dur[sm == i].mean() for i in range(148)])
print("duration by SM: min/median/max", round(float(sm_mean.min()), 1), round(float(np.median(sm_mean)), 1), round(float(sm_mean.max()), 1))
print("SMs sorted by mean duration (id:us):", [(int(i), round(float(sm_mean[i]), 1)) for i in np.argsort(sm_mean)[::12]])
print("mean duration on SMs with 3 CTAs vs 4:", round(float(sm_mean[per_sm == 3].mean()), 1), round(float(sm_mean[per_sm == 4].mean()), 1))
# second call (warm)
fe.features(wave, mean=mean.cuda(), std=std.cuda(), layout="sections", out=out)
torch.cuda.synchronize()
fe.lib.ast_debug_stft_cta_times(buf2)
t = np.array(buf2[:], dtype=np.uint64).reshape(4096, 3)[:576].astype(np.int64)
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
dur = end - start
print(f"warm call: end min/median/max {end.min():.1f}/{np.median(end):.1f}/{end.max():.1f}")
print("warm duration by blockIdx.x:", [round(float(dur[bx == i].mean()), 1) for i in range(9)])
