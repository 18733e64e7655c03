"""Diagnostic: run the CQT tensor-core kernel from a -DAST_TRACE build and print CTA 0's pipeline timeline."""
import ctypes, importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("audio_style_transfer_b200.build")
out = os.path.join(ROOT, "scratch", "libast_trace.so")
if "--build" in sys.argv or not os.path.exists(out):
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-DAST_TRACE", "-o", out] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
    subprocess.check_call(cmd)
    if "--build" in sys.argv:
        sys.exit(0)
import torch
lib_mod = importlib.import_module("audio_style_transfer_b200._lib")
lib_mod.LIB_PATH = out
fe_mod = importlib.import_module("audio_style_transfer_b200.frontend")
fe = fe_mod.FrontEnd("cuda:0")
x = (torch.randn(64, 220500, device="cuda") * 0.07)
for _ in range(3):
    f, _n = fe.features(x, layout="sections")
torch.cuda.synchronize()
lib = lib_mod.load()
buf = np.zeros((3, 512, 6), dtype=np.int64)
rc = lib.ast_debug_cqt_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
assert rc == 0, rc
t0 = buf[0, 0, 0]
P, M, E = buf[0] - t0, buf[1] - t0, buf[2] - t0
n_items = int((buf[0, :, 0] > 0).sum())
print("items", n_items)
print("producer warp0: item | top  loads_issued  empty_ok  stored  arrived | MMA: wait_start full_ok issued")
for i in range(min(n_items, 100)):
    print(i, P[i, :5].tolist(), M[i, :3].tolist())
print("epilogue warp8: tile | wait_start acc_full drained stored")
for i in range(int((buf[2, :, 0] > 0).sum())):
    print(i, E[i, :4].tolist())
d = np.diff(P[:n_items, 0])
print("mean item period", d.mean(), "median", np.median(d))
print("producer: issue loads", (P[:n_items,1]-P[:n_items,0]).mean(), "wait empty", (P[:n_items,2]-P[:n_items,1]).mean(),
      "split+store (incl load wait)", (P[:n_items,3]-P[:n_items,2]).mean(), "fence+arrive", (P[:n_items,4]-P[:n_items,3]).mean())
print("mma: wait full", (M[:n_items,1]-M[:n_items,0]).mean(), "issue", (M[:n_items,2]-M[:n_items,1]).mean())
# TMA path (splitters): stamps are 0 top, 2 waits done (box landed + operand stage free), 3 split + stored, 4 arrived, 1 advanced
print("splitter (TMA path): wait raw/empty", (P[:n_items:2,2]-P[:n_items:2,0]).mean(), "split+store", (P[:n_items:2,3]-P[:n_items:2,2]).mean(),
      "fence+arrive", (P[:n_items:2,4]-P[:n_items:2,3]).mean(), "advance", (P[:n_items:2,1]-P[:n_items:2,4]).mean(),
      "| period of group 0 (2 blocks)", np.diff(P[:n_items:2,0]).mean())
nt = int((buf[2, :, 0] > 0).sum())
Ev = E[:nt][E[:nt, 0] > 0]
print("epilogue group 0: wait acc", (Ev[:,1]-Ev[:,0]).mean(), "drain TMEM", (Ev[:,2]-Ev[:,1]).mean(), "scale+transpose+store", (Ev[:,3]-Ev[:,2]).mean(),
      "| tile period (2 tiles)", np.diff(Ev[:,0]).mean())
print("epilogue detail: scale+transpose", (Ev[:,4]-Ev[:,2]).mean(), "load_ctx(next)", (Ev[:,5]-Ev[:,4]).mean(), "store loop", (Ev[:,3]-Ev[:,5]).mean())
print("total cycles CTA 0:", max(P[:n_items].max(), M[:n_items].max(), E[:nt].max()))
