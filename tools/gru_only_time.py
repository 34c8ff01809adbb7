"""Scratch: time the GRU recurrence kernel alone through the profiling hook."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stemgnn_b200 import synthetic as tp
from models.base_model import Model
from stemgnn_b200 import _lib
B, N, W, H = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 358, 12, 3))]
dev = torch.device("cuda:0")
m = Model(N, 2, W, 5, horizon=H); m.load_state_dict(tp.synthetic_params(N, W, H, 5, seed=0)); m = m.to(dev).eval()
x = tp.synthetic_batch(B, N, W, H)[0].to(dev)
lib = _lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); e1.record(); torch.cuda.synchronize()
with torch.no_grad():
    for _ in range(3): m(x)
    tot = 0.0
    for _ in range(10):
        lib.stemgnn_profile_gru(e0.cuda_event, e1.cuda_event); m(x); lib.stemgnn_profile_gru(None, None)
        torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
print(f"B={B} N={N}: gru kernel {tot/10*1000:.1f} us = {tot/10*1e6/N*1.965:.0f} cycles/step")
