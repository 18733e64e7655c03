"""One process, many settings: feature step (64 clips x 10 s) against AST_STFT_ITERS x AST_STFT_TAPER (both read per launch)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
dl = importlib.import_module("audio_style_transfer_b200.dataloader")
mean, std = dl.load_stats_npz(bench.STATS_NPZ)
mean, std = mean.cuda(), std.cuda()
wave = torch.from_numpy(bench.make_clips()).cuda()
out = torch.empty((64, 4, 2, 287, 597), device="cuda")
step = lambda: fe.features(wave, mean=mean, std=std, layout="sections", out=out)
def timed(n=100):
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
settings = [(None, None), ("8", None), ("8", "8,4"), ("8", "12,6"), ("8", "16,8"), ("8", "8,8"), ("6", None), ("6", "8"), ("6", "12"),
            ("4", "3,2"), ("4", "4,3"), ("4", "5,2"), (None, None)]
for it, tp in settings:
    for k, v in (("AST_STFT_ITERS", it), ("AST_STFT_TAPER", tp)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    print(json.dumps({"iters": it, "taper": tp, "ms": round(timed(), 5)}), flush=True)
