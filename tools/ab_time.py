"""A/B timing of the eval forward and the fused train step at cfg2 in ONE process per setting (the library reads its
STEMGNN_* switches once).  `python tools/ab_time.py [prof]` — with `prof`, also prints a per-kernel table (torch profiler,
CUPTI activity records: no replay, unlike ncu) of eager forwards and train steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                             # noqa: E402
from stemgnn_b200 import synthetic as tp                 # noqa: E402
from stemgnn_b200.trainer import FusedTrainer            # noqa: E402
from models.base_model import Model                      # noqa: E402

B, N, W, H = 32, 358, 12, 3
dev = torch.device("cuda:0")
m = Model(N, 2, W, 5, horizon=H)
m.load_state_dict(tp.synthetic_params(N, W, H, 5, seed=0))
m = m.to(dev).eval()
x, y = tp.synthetic_batch(B, N, W, H)
x, y = x.to(dev), y.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("STEMGNN_")) or "default"


def timed(fn, iters):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


with torch.no_grad():
    for _ in range(6):
        m(x)
    med, best = timed(lambda: m(x), 60)
print(f"[{tag}] eval forward: median {med:.4f} ms  best {best:.4f} ms  ({B / med * 1e3:.0f} windows/s)", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "prof":
    from torch.profiler import ProfilerActivity, profile
    m.use_cuda_graph = False
    with torch.no_grad():
        m(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            for _ in range(5):
                m(x)
            torch.cuda.synchronize()
    print("---- eager eval forward x5: kernels by total time ----")
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    m.use_cuda_graph = True

m.train()
tr = FusedTrainer(m, lr=1e-4, warmup_eager=2)
for _ in range(5):
    tr.step(x, y)
torch.cuda.synchronize()
med, best = timed(lambda: tr.step(x, y), 30)
print(f"[{tag}] fused train step: median {med:.4f} ms  best {best:.4f} ms", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "prof":
    tr2 = FusedTrainer(m, lr=1e-4, use_graph=False)
    tr2.step(x, y)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(3):
            tr2.step(x, y)
        torch.cuda.synchronize()
    print("---- eager train step x3: kernels by total time ----")
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
