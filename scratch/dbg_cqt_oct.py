"""Per-octave error of the CUDA CQT against the oracle (which octave / frame range is wrong)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cqt as oc
synth = importlib.import_module("audio_style_transfer_b200.synth")
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
w = synth.piano_clip(3, n)
got = fe.cqt(torch.from_numpy(w).cuda()[None])[0].cpu().numpy()      # (2, T, 84)
V = oc.cqt(w)
ref = np.stack([V.real.T, V.imag.T])
mx = np.abs(ref).max()
for o in range(7):
    c0 = 84 - 12 * (o + 1)
    e = np.abs(got[:, :, c0:c0 + 12] - ref[:, :, c0:c0 + 12]).max(axis=(0, 2)) / mx   # per frame
    bad = np.nonzero(e > 1e-5)[0]
    print(f"octave {o}: max err {e.max():.2e}  bad frames {len(bad)}/{len(e)}" + (f" first {bad[:6]} last {bad[-3:]}" if len(bad) else ""),
          "| got/ref energy ratio", float(np.abs(got[:, :, c0:c0+12]).sum() / np.abs(ref[:, :, c0:c0+12]).sum()))
