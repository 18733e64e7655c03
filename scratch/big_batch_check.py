"""Sanity at a large batch: 1024 clips x 10 s through the fused feature path; a few clips must equal their single-clip runs."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty(1024, 220500, device="cuda").normal_(0, 0.07, generator=g)
out, n = fe.features(x, layout="sections")
torch.cuda.synchronize()
print(tuple(out.shape), int(n.min()), int(n.max()))
for i in (0, 511, 1023):
    single, _ = fe.features(x[i:i + 1], layout="sections")
    print(i, "equal:", bool(torch.equal(out[i], single[0])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fe.features(x, layout="sections", out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("1024 clips: %.3f ms  %.2f M audio-s/s" % (ms, 1024 * 10 / ms / 1e3))
