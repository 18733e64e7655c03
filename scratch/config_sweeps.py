"""BASELINE.json configs[3] and configs[4] at full size on one GPU (numbers for DESIGN.md, not bench lines):
 - dataset statistics over 100 000 synthetic clips (50 000 'piano' + 50 000 'violin'), streamed in chunks of 256;
 - iSTFT reconstruction sweep, decoder-shaped input (B, 4, 2, 287, 513) for B = 1 ... 4096."""
import importlib, json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
stats = importlib.import_module("audio_style_transfer_b200.stats")
out = {}

def timed(fn, k):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

# ---- configs[3]: 100k clips, chunks generated on the device from a counter-based seed (clip id -> seed)
n_clips, chunk = 100000, 256
acc, counts = fe.new_stats_accumulator(2)
g = torch.Generator(device="cuda")
x = torch.empty(chunk, 220500, device="cuda")
gid = torch.empty(chunk, dtype=torch.int32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
done = 0
while done < n_clips:
    n = min(chunk, n_clips - done)
    g.manual_seed(1000 + done)
    x[:n].normal_(0.0, 0.07, generator=g)
    gid[:n] = (torch.arange(done, done + n, device="cuda") >= n_clips // 2).to(torch.int32)
    fe.stats_accumulate(x[:n], acc, counts, group_ids=gid[:n])
    done += n
torch.cuda.synchronize(); dt = time.perf_counter() - t0
res = stats.finalize_all(acc, counts)
out["stats_100k"] = {"clips": n_clips, "seconds": dt, "clips_per_s": n_clips / dt, "audio_s_per_s": n_clips * 10.0 / dt,
                     "counts": [float(c) for c in counts.cpu()], "groups": sorted(res),
                     "note": "includes the on-device normal_() generation of every chunk (56 MB per 256 clips)"}
print(json.dumps(out["stats_100k"]))

# ---- configs[4]: iSTFT sweep
sweep = []
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    spec = torch.randn(B, 4, 2, 287, 513, device="cuda")
    k = 50 if B <= 256 else 5
    ms = timed(lambda: fe.istft(spec, layout="sections", overlap=96, original_size=862), k)
    sweep.append({"B": B, "ms": ms, "audio_s_per_s": B * 219904 / 22050 / (ms * 1e-3),
                  "gbs": B * 5591008 / (ms * 1e-3) / 1e9})
    print(json.dumps(sweep[-1]))
    del spec
out["istft_sweep"] = sweep
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_sweeps.json"), "w"), indent=1)
