"""Manual check (not collected by pytest; run on the GPU box): large-shape sanity (cfg4 global batch, cfg5 N=2048 shard) — forward + backward finite and
consistent with the host port where the port is affordable."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_port as tp
from models.base_model import Model

def run(B, N, W, H, check):
    dev = torch.device("cuda:0")
    p = tp.synthetic_params(N, W, H, 5, seed=1)
    m = Model(N, 2, W, 5, horizon=H); m.load_state_dict(p); m = m.to(dev).eval()
    x, y = tp.synthetic_batch(B, N, W, H, seed=2)
    torch.cuda.synchronize(); t0 = time.time()
    f, a = m(x.to(dev))
    loss = torch.nn.functional.mse_loss(f, y.to(dev)); loss.backward()
    torch.cuda.synchronize(); dt = time.time() - t0
    ok = bool(torch.isfinite(f).all()) and all(torch.isfinite(q.grad).all() for q in m.parameters() if q.grad is not None)
    msg = f"B={B} N={N} W={W} H={H}: fwd+bwd {dt*1e3:.1f} ms, finite={ok}, loss={float(loss):.5f}"
    if check:
        with torch.no_grad():
            f_ref, _ = tp.model_forward(x, p)
        msg += f", max|f-f_ref|={float((f.detach().cpu()-f_ref).abs().max()):.2e}"
    print(msg, flush=True)

run(128, 325, 12, 12, True)      # cfg4 global batch on one GPU
run(256, 140, 12, 3, True)       # large batch
run(4, 2048, 12, 3, True)        # cfg5 node count (generic per-step GRU path)
