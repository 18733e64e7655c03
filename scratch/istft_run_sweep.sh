#!/bin/bash
# diagnostic: iSTFT time at B = 64 vs the run length (segments per CTA)
for r in 0 14 16 18 20 24 28 38 44 58; do
  AST_ISTFT_RUN=$r python - <<'PY'
import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
spec = torch.randn(64, 4, 2, 287, 513, device="cuda")
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("run", os.environ["AST_ISTFT_RUN"], "%.4f ms" % timed(lambda: fe.istft(spec, layout="sections", overlap=96, original_size=862)))
PY
done
