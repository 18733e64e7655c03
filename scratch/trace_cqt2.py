"""CTA 0's whole block sequence from the -DAST_TRACE build: per block the splitter's wait for the boxes, split time and
the MMA warp's phases, labelled with the block's octave."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
lib_mod = importlib.import_module("audio_style_transfer_b200._lib")
lib_mod.LIB_PATH = os.path.join(ROOT, "scratch", "libast_trace.so")
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
x = (torch.randn(64, 220500, device="cuda") * 0.07)
os.environ.setdefault("AST_OVERLAP", "1")
for _ in range(3):
    fe.features(x, layout="sections")
torch.cuda.synchronize()
lib = lib_mod.load()
buf = np.zeros((3, 512, 6), dtype=np.int64)
assert lib.ast_debug_cqt_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))) == 0
octs = []
for tile in range(0, 3136, 148):
    o = tile // 448
    octs += [o] * (8 if o == 0 else 4 if o == 1 else 2 if o == 2 else 1)
P, M = buf[0], buf[1]
t0 = P[0, 0]
print("item oct | splitter: top->landed  split  fence | mma: top->full_ok  ->issued | period")
prev = None
for i, o in enumerate(octs):
    p, m = P[i] - t0, M[i] - t0
    print(i, o, "|", p[2] - p[0], p[3] - p[2], p[4] - p[3], "|", m[1] - m[0], m[2] - m[1], "|", (p[0] - prev) if prev is not None else 0)
    prev = p[0]
print("total", (M[len(octs) - 1, 2] - t0))
