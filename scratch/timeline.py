"""Gantt chart of one feature call (diagnostic build: scratch/build_variant.sh timeline -DAST_TIMELINE, library copied over the
package library).  Per kernel: CTA start / end distribution; per 10 us bin: SMs holding a decimator CTA, a CQT CTA, and
the number of resident STFT CTAs.  usage (on the GPU box): cp scratch/variants/lib_timeline.so audio-style-transfer_b200/libast_frontend.so; python scratch/timeline.py"""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
dl = importlib.import_module("audio_style_transfer_b200.dataloader")
mean, std = dl.load_stats_npz(bench.STATS_NPZ)
mean, std = mean.cuda(), std.cuda()
wave = torch.from_numpy(bench.make_clips()).cuda()
out = torch.empty((64, 4, 2, 287, 597), device="cuda")
for _ in range(3):
    fe.features(wave, mean=mean, std=std, layout="sections", out=out)
torch.cuda.synchronize()

def fetch(name, n):
    buf = (ctypes.c_ulonglong * (8192 * 4))()
    getattr(fe.lib, "ast_debug_timeline_" + name)(buf)
    return np.array(buf[:], dtype=np.uint64).reshape(8192, 4)[:n].astype(np.int64)

n_stft = int(os.environ.get("N_STFT", "1728"))
T = {"dec": fetch("dec", 148), "cqt": fetch("cqt", 148), "stft": fetch("stft", n_stft)}
T = {k: v[v[:, 0] > 0] for k, v in T.items()}
t0 = min(v[:, 0].min() for v in T.values())
for k, v in T.items():
    s, e = (v[:, 0] - t0) / 1e3, (v[:, 1] - t0) / 1e3
    q = lambda a: "/".join(f"{x:.1f}" for x in np.percentile(a, [0, 10, 50, 90, 100]))
    print(f"{k:5s} CTAs {len(v):5d}  start us (min/p10/med/p90/max) {q(s)}   end {q(e)}   duration {q(e - s)}")
end = max(v[:, 1].max() for v in T.values())
print("whole call (first CTA start -> last CTA end): %.1f us" % ((end - t0) / 1e3))
print("  t us | SMs with decimator CTA | SMs with CQT CTA | resident STFT CTAs (of 592) | SMs with nothing")
for t in np.arange(0, (end - t0) / 1e3, 10.0):
    ts = t0 + t * 1e3 + 5e3
    on = {k: v[(v[:, 0] <= ts) & (v[:, 1] > ts)] for k, v in T.items()}
    busy = set(on["dec"][:, 2]) | set(on["cqt"][:, 2]) | set(on["stft"][:, 2])
    print(f"  {t:5.0f} | {len(on['dec']):4d} | {len(on['cqt']):4d} | {len(on['stft']):4d} | {148 - len(busy):4d}")
# the decimator's tail CTAs and the CQT CTAs that follow them on the same SM
d, c = T["dec"], T["cqt"]
late = d[np.argsort(d[:, 1])][-8:]
print("last decimator CTAs end at", [round((x - t0) / 1e3, 1) for x in late[:, 1]])
sm_dec_end = {int(r[2]): r[1] for r in d}
gap = [(r[0] - sm_dec_end.get(int(r[2]), r[0])) / 1e3 for r in c]
print("CQT CTA start minus the end of the decimator CTA on its SM: min/med/max %.1f/%.1f/%.1f us" % (min(gap), np.median(gap), max(gap)))
print("decimator: TMEM release after the end stamp: med %.2f us" % np.median((d[:, 3] - d[:, 1]) / 1e3))
print("CQT: TMEM release after the end stamp: med/max %.2f/%.2f us" % (np.median((c[:, 3] - c[:, 1]) / 1e3), np.max((c[:, 3] - c[:, 1]) / 1e3)))
# turnaround on an SM: start of a STFT CTA minus the latest earlier end of any CTA on the same SM (4 slots per SM: take CTAs
# that start after the first wave)
s_ = T["stft"]
ends = {}
for k, v in T.items():
    for r in v:
        ends.setdefault(int(r[2]), []).append(r[1])
first_wave = np.sort(s_[:, 0])[591]
gaps = []
for r in s_[s_[:, 0] > first_wave + 2000]:
    e = np.array(ends[int(r[2])])
    e = e[e <= r[0]]
    if len(e): gaps.append((r[0] - e.max()) / 1e3)
print("STFT CTA start minus the latest earlier CTA end on its SM (after the first wave): p10/med/p90 %.2f/%.2f/%.2f us" % tuple(np.percentile(gaps, [10, 50, 90])))
sm_cqt_end = {int(r[2]): r[1] for r in c}
g2 = [(r[0] - sm_cqt_end[int(r[2])]) / 1e3 for r in s_[np.argsort(s_[:, 0])][:592] if int(r[2]) in sm_cqt_end]
print("first-wave STFT CTA start minus the CQT CTA's end stamp on its SM: p10/med/p90 %.1f/%.1f/%.1f us" % tuple(np.percentile(g2, [10, 50, 90])))
