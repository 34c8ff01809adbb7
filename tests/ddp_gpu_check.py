"""torchrun check (not collected by pytest): 2+ ranks, one batch each; the gradients produced by the
CUDA backward + its single NCCL all-reduce must equal the mean of the per-rank gradients.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/ddp_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port as tp          # noqa: E402  (seeded weights / inputs only)
from models.base_model import Model           # noqa: E402
from stemgnn_b200 import ddp                  # noqa: E402


def main():
    rank, world, local = ddp.init_from_env()
    dev = torch.device("cuda", local)
    N, W, H = 70, 12, 3
    torch.manual_seed(5 + rank)                                   # different init per rank on purpose
    m = Model(N, 2, W, 5, horizon=H).to(dev)
    ddp.attach(m)                                                 # broadcasts rank 0's parameters
    sig = torch.cat([p.detach().reshape(-1)[:4] for p in m.parameters()])
    gathered = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(gathered, sig)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "parameters differ after broadcast"

    x, y = tp.synthetic_batch(8, N, W, H, seed=50 + rank)
    mask = (torch.rand(8, N, N, generator=torch.Generator().manual_seed(rank)) >= 0.5)
    m.train()
    f, _ = m(x.to(dev), dropout_mask=mask)
    torch.nn.functional.mse_loss(f, y.to(dev)).backward()
    avg = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m.parameters()])

    m._ddp["enabled"] = False                                     # same step WITHOUT the collective
    m.zero_grad()
    f, _ = m(x.to(dev), dropout_mask=mask)
    torch.nn.functional.mse_loss(f, y.to(dev)).backward()
    local_g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m.parameters()])
    dist.all_reduce(local_g)
    local_g /= world
    err = float((avg - local_g).abs().max() / local_g.abs().max())
    assert err < 1e-5, f"all-reduced gradient mismatch: {err}"
    if rank == 0:
        print(f"ddp_gpu_check ok: world={world}, grad elems={avg.numel()}, rel err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
