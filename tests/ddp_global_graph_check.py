"""torchrun check (driven by tests/test_ddp_global_graph_gpu.py; also usable by hand): with `ddp.attach(model,
global_graph=True)` a world-size-N data-parallel step equals the single-process step on the concatenated batch — forecast of
every shard, the attention matrix and every parameter gradient (reference semantics: `torch.mean(attention, dim=0)` over the
WHOLE batch, base_model.py:140).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/ddp_global_graph_check.py
    DDP_ONE_GPU=1 ... (all ranks on cuda:0, gloo backend: exercises the same hooks on a single-GPU box)
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port as tp          # noqa: E402  (seeded weights / inputs only)
from models.base_model import Model           # noqa: E402
from stemgnn_b200 import ddp                  # noqa: E402


def flat_grads(m):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m.parameters()])


def main():
    one_gpu = os.environ.get("DDP_ONE_GPU") == "1"
    if one_gpu:
        os.environ["LOCAL_RANK"] = "0"
    rank, world, local = ddp.init_from_env(backend="gloo" if one_gpu else None)
    dev = torch.device("cuda", 0 if one_gpu else local)
    torch.cuda.set_device(dev)
    N, W, H, Bs = 53, 12, 3, 6
    Bt = Bs * world
    params = tp.synthetic_params(N, W, H, 5, seed=21, scale_mode="trained")
    x, y = tp.synthetic_batch(Bt, N, W, H, seed=77)
    mask = (torch.rand(Bt, N, N, generator=torch.Generator().manual_seed(3)) >= 0.5)
    sl = slice(rank * Bs, (rank + 1) * Bs)

    m = Model(N, 2, W, 5, horizon=H)
    m.load_state_dict(params)
    m = m.to(dev).train()
    ddp.attach(m, global_graph=True)
    f, att = m(x[sl].to(dev), dropout_mask=mask[sl])
    torch.nn.functional.mse_loss(f, y[sl].to(dev)).backward()
    g_ddp = flat_grads(m)

    # reference of the check: ONE process, the whole batch
    s = Model(N, 2, W, 5, horizon=H)
    s.load_state_dict(params)
    s = s.to(dev).train()
    fs, att_s = s(x.to(dev), dropout_mask=mask)
    torch.nn.functional.mse_loss(fs, y.to(dev)).backward()
    g_single = flat_grads(s)

    e_att = float((att - att_s).abs().max() / att_s.abs().max())
    e_f = float((f - fs[sl]).abs().max() / fs.abs().max())
    e_g = float((g_ddp - g_single).abs().max() / g_single.abs().max())
    # per-shard graphs (the default) must NOT reproduce the global step: the check has teeth
    m2 = Model(N, 2, W, 5, horizon=H)
    m2.load_state_dict(params)
    m2 = m2.to(dev).train()
    ddp.attach(m2, global_graph=False)
    f2, att2 = m2(x[sl].to(dev), dropout_mask=mask[sl])
    e_att_local = float((att2 - att_s).abs().max() / att_s.abs().max())
    assert e_att < 1e-5, f"attention mismatch {e_att}"
    assert e_f < 1e-4, f"forecast mismatch {e_f}"
    assert e_g < 1e-3, f"gradient mismatch {e_g}"
    assert e_att_local > 10 * max(e_att, 1e-7), f"shard-local graph unexpectedly equals the global one ({e_att_local})"

    # the captured-graph trainer runs the same exchange inside its step graph
    from stemgnn_b200.trainer import FusedTrainer
    t = Model(N, 2, W, 5, horizon=H)
    t.load_state_dict(params)
    t = t.to(dev).train()
    t.dropout_rate = 0.0
    ddp.attach(t, global_graph=True)
    # (gloo stages CUDA tensors through the host: not capturable, so the one-GPU variant runs the step body eagerly)
    tr = FusedTrainer(t, lr=1e-3, warmup_eager=1, use_graph=not one_gpu)
    u = Model(N, 2, W, 5, horizon=H)
    u.load_state_dict(params)
    u = u.to(dev).train()
    u.dropout_rate = 0.0
    tu = FusedTrainer(u, lr=1e-3, warmup_eager=1, use_graph=not one_gpu)
    for _ in range(3):                     # eager step, capture, replay
        tr.step(x[sl].to(dev), y[sl].to(dev))
        tu.step(x.to(dev), y.to(dev))
    torch.cuda.synchronize()
    pd = torch.cat([p.detach().reshape(-1) for p in t.parameters()])
    pu = torch.cat([p.detach().reshape(-1) for p in u.parameters()])
    e_p = float((pd - pu).abs().max() / pu.abs().max())
    assert e_p < 1e-4, f"parameters after 3 fused steps differ: {e_p}"
    if rank == 0:
        print(f"ddp_global_graph_check ok: world={world} backend={dist.get_backend()} attention {e_att:.1e} "
              f"forecast {e_f:.1e} grads {e_g:.1e} (shard-local graph differs by {e_att_local:.1e}); "
              f"fused trainer params after 3 steps {e_p:.1e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
