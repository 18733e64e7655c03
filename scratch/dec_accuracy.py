"""Scratch: per-octave accuracy of the decimation cascade (reads the octave buffers out of the workspace)."""
import sys, os, importlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cqt as oc
frontend = importlib.import_module("audio_style_transfer_b200.frontend")
synth = importlib.import_module("audio_style_transfer_b200.synth")

def pad_len(n): return (n + 8 + 7) & ~7
def octave_len(L, o): return (L + (1 << o) - 1) >> o
def octave_offset(L, o): return sum(pad_len(octave_len(L, i)) for i in range(1, o))

fe = frontend.FrontEnd("cuda:0")
L = 220500
w = synth.noise_clip(1, L)
x = torch.from_numpy(w).cuda()[None]
out = fe.cqt(x)
torch.cuda.synchronize()
ws = fe._ws
table_bytes = (8 * 2 * 597 * 1 + 255) // 256 * 256
f = ws[table_bytes:].view(torch.float32)
sigs = oc.octave_signals(w.astype(np.float64))
for o in range(1, 7):
    n = octave_len(L, o)
    got = f[octave_offset(L, o): octave_offset(L, o) + n].cpu().numpy().astype(np.float64)
    ref = sigs[o]
    err = got - ref
    scale = np.sqrt((ref ** 2).mean())
    # signed gain error: least-squares gain of got vs ref
    gain = (got * ref).sum() / (ref * ref).sum() - 1.0
    print(f"octave {o}: rms err/rms {np.sqrt((err**2).mean())/scale:.3e}  max err/rms {np.abs(err).max()/scale:.3e}  gain-1 {gain:+.3e}")
V = oc.cqt(w)
g = out[0].cpu().numpy(); got = g[0].T + 1j * g[1].T
print("cqt rel max err", np.abs(got - V).max() / np.abs(V).max())
