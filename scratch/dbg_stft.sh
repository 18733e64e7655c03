#!/bin/bash
# diagnostic: STFT kernel time with its global stores predicated off (AST_STFT_DEBUG=1), sections and flat layouts
for d in 0 1; do
  AST_STFT_DEBUG=$d python - <<'PY'
import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
x = torch.randn(64, 220500, device="cuda") * 0.07
def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("AST_STFT_DEBUG=%s  stft flat: %.4f ms" % (os.environ.get("AST_STFT_DEBUG"), timed(lambda: fe.stft(x))))
PY
done
