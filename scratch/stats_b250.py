"""ast_stats_accumulate at the bench's call size (250 clips x 10 s): CUDA-event time per call."""
import importlib, os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
B = int(os.environ.get("B", "250"))
wave = (torch.randn(B, 220500, device="cuda") * 0.07)
acc, counts = fe.new_stats_accumulator(2)
gid = (torch.arange(B, device="cuda") % 2).to(torch.int32)
for _ in range(3): fe.stats_accumulate(wave, acc, counts, group_ids=gid)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fe.stats_accumulate(wave, acc, counts, group_ids=gid)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"B": B, "ms_per_call": ms, "audio_s_per_s": B * 10 / (ms * 1e-3)}))
