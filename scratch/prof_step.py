"""The bench's feature step and iSTFT leg alone (64 clips x 10 s), for ncu captures and quick A/B timing:

    python scratch/prof_step.py [--steps K] [--legs features,istft,stats]

Prints the CUDA-event time per step of each leg (back-to-back launches, as bench.py times them)."""
import argparse, importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--legs", default="features,istft")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--profile", action="store_true", help="per-kernel events (serialised launches)")
a = ap.parse_args()
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
dl = importlib.import_module("audio_style_transfer_b200.dataloader")
lib = importlib.import_module("audio_style_transfer_b200._lib")
mean, std = dl.load_stats_npz(bench.STATS_NPZ)
mean, std = mean.cuda(), std.cuda()
wave = torch.from_numpy(bench.make_clips()).cuda()
out = torch.empty((64, 4, 2, 287, 597), device="cuda")
legs = {"features": lambda: fe.features(wave, mean=mean, std=std, layout="sections", out=out)}
fe.features(wave, mean=mean, std=std, layout="sections", out=out)
spec = out[..., :513].contiguous()
legs["istft"] = lambda: fe.istft(spec, layout="sections", overlap=96, original_size=862)
acc, counts = fe.new_stats_accumulator(2)
gid = (torch.arange(64, device="cuda") % 2).to(torch.int32)
legs["stats"] = lambda: fe.stats_accumulate(wave, acc, counts, group_ids=gid)
res = {}
for name in a.legs.split(","):
    fn = legs[name]
    for _ in range(a.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / a.steps
if a.profile:
    lib.profile_enable(True)
    for name in a.legs.split(","):
        for _ in range(a.steps):
            legs[name]()
    torch.cuda.synchronize()
    res["kernels"] = {k: t / n for k, (t, n) in lib.profile_collect().items()}
    lib.profile_enable(False)
print(json.dumps(res))
