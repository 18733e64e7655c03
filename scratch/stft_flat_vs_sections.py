"""Diagnostic: one STFT launch in each layout (for ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
x = torch.randn(64, 220500, device="cuda") * 0.07
for _ in range(2):
    fe.stft(x)                                   # flat (B, 2, 862, 513)
    fe.features(x, layout="sections")            # sections (B, 4, 2, 287, 597): stft + decimator + cqt
torch.cuda.synchronize()
