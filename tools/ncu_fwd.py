"""Eager eval forwards at cfg2 for an ncu capture (tools/gpu_shot_ncu.sh): python tools/ncu_fwd.py [n_forwards]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                             # noqa: E402
from stemgnn_b200 import synthetic as tp                 # noqa: E402
from models.base_model import Model                      # noqa: E402

B, N, W, H = 32, 358, 12, 3
dev = torch.device("cuda:0")
m = Model(N, 2, W, 5, horizon=H)
m.load_state_dict(tp.synthetic_params(N, W, H, 5, seed=0))
m = m.to(dev).eval()
m.use_cuda_graph = False
x, _ = tp.synthetic_batch(B, N, W, H)
x = x.to(dev)
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        m(x)
torch.cuda.synchronize()
print("done")
