"""Diagnostic: run the decimator kernel from a -DAST_TRACE build and print CTA 0's pipeline timeline per tile."""
import ctypes, importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("audio_style_transfer_b200.build")
out = os.environ.get("TRACE_LIB", os.path.join(ROOT, "scratch", "libast_trace.so"))
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
    f = fe.cqt(x)
torch.cuda.synchronize()
lib = lib_mod.load()
buf = np.zeros((4, 64, 8), dtype=np.int64)
assert lib.ast_debug_dec_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))) == 0
t0 = buf[0, 0, 0]
n_tiles = int((buf[2, :, 0] > 0).sum())
print("tiles of CTA 0:", n_tiles, "(cycles since the first producer stamp)")
print("tile | G0: top h_empty_ok s0_ok s0_arr s1_ok s1_arr deps_ok loads_issued | G1: same (s2, s3)")
for i in range(n_tiles):
    print(i, (buf[0, i] - t0).tolist(), (buf[1, i] - t0).tolist())
print("tile | MMA: top acc_empty_ok full0 full1 full2 full3 issued_all | EPI: top acc_full drained stored")
for i in range(n_tiles):
    print(i, (buf[2, i, :7] - t0).tolist(), (buf[3, i, :4] - t0).tolist())
