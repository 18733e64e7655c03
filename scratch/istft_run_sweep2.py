"""One process: iSTFT leg (64 x (4,2,287,513) sections -> audio) against AST_ISTFT_RUN (segments per CTA, read per launch)."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fe = importlib.import_module("audio_style_transfer_b200.frontend").FrontEnd("cuda:0")
spec = torch.randn(64, 4, 2, 287, 513, device="cuda")
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
step = lambda: fe.istft(spec, layout="sections", overlap=96, original_size=862)
def timed(n=60):
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for r in (None, 20, 24, 28, 32, 36, 40, 44, 52, 58, 64, 72, 96, None):
    if r is None: os.environ.pop("AST_ISTFT_RUN", None)
    else: os.environ["AST_ISTFT_RUN"] = str(r)
    y = step()
    if ref is None: ref = y.clone()
    print(json.dumps({"run": r, "ms": round(timed(), 5), "bit_identical_to_default": bool(torch.equal(y, ref))}), flush=True)
