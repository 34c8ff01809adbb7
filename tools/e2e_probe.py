import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stemgnn_b200 import synthetic as tp
from models.base_model import Model
dev = torch.device("cuda", 0)
m = Model(358, 2, 12, 5, horizon=3); m.load_state_dict(tp.synthetic_params(358, 12, 3, 5, seed=0)); m = m.to(dev).eval()
x_host = tp.synthetic_batch(32, 358, 12, 3)[0].pin_memory()
out_host = torch.empty(32, 3, 358).pin_memory()
x_dev = x_host.to(dev)
def wall(fn, n=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    for g in (True, False):
        m.use_cuda_graph = g
        for _ in range(4): m(x_dev)
        print("graph", g, "resident x: %.3f ms/step wall" % wall(lambda: m(x_dev)))
        def e2e():
            xd = x_host.to(dev, non_blocking=True); f, _ = m(xd); out_host.copy_(f, non_blocking=True)
        for _ in range(3): e2e()
        print("graph", g, "e2e: %.3f ms/step wall" % wall(e2e))
        def e2e2():
            x_dev.copy_(x_host, non_blocking=True); f, _ = m(x_dev); out_host.copy_(f, non_blocking=True)
        print("graph", g, "e2e (static device x): %.3f ms/step wall" % wall(e2e2))
    print("has graph", m._rt.get("cuda_graph") is not None, "use", m.use_cuda_graph)
