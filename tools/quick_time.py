"""Scratch timing of the eval forward at the north-star shape (not the bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stemgnn_b200 import synthetic as tp
from models.base_model import Model

B, N, W, H = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 358, 12, 3))]
dev = torch.device("cuda:0")
m = Model(N, 2, W, 5, horizon=H)
m.load_state_dict(tp.synthetic_params(N, W, H, 5, seed=0))
m = m.to(dev).eval()
x, _ = tp.synthetic_batch(B, N, W, H)
x = x.to(dev)
with torch.no_grad():
    for _ in range(5):
        m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    e0.record()
    for _ in range(iters):
        m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
print(f"forward B={B} N={N}: {ms:.3f} ms  -> {B / ms * 1e3:.0f} windows/s")
