"""Scratch timing of one training step at the north-star shape (not the bench; see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stemgnn_b200 import synthetic as tp
from models.base_model import Model

B, N, W, H = 32, 358, 12, 3
dev = torch.device("cuda:0")
m = Model(N, 2, W, 5, horizon=H)
m.load_state_dict(tp.synthetic_params(N, W, H, 5, seed=0))
m = m.to(dev).train()
x, y = tp.synthetic_batch(B, N, W, H)
x, y = x.to(dev), y.to(dev)
opt = torch.optim.RMSprop(m.parameters(), lr=1e-4, eps=1e-8)
def step():
    m.zero_grad()
    f, _ = m(x)
    loss = torch.nn.functional.mse_loss(f, y)
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
print(f"train step: {e0.elapsed_time(e1) / iters:.3f} ms")
