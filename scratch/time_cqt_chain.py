"""Diagnostic: device time of the decimator + CQT chain alone (fe.cqt), of the STFT alone, and of the fused features call."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
out_c = torch.empty(64, 2, 862, 84, device="cuda")
print("cqt chain (memset + decimator + cqt_tc), flat 84-wide rows: %.4f ms" % timed(lambda: fe.cqt(x)))
print("stft alone, flat 513-wide rows: %.4f ms" % timed(lambda: fe.stft(x)))
print("features, sections: %.4f ms" % timed(lambda: fe.features(x, layout="sections")))
