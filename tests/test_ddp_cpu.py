"""world_size-2 gloo tests (CPU) of the data-parallel host logic: shard bookkeeping, parameter
broadcast, the single flat-gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from stemgnn_b200 import ddp
    r, w, _ = ddp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                      # replicas start DIFFERENT on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    unused = torch.nn.Parameter(torch.ones(3))          # a parameter that never gets a gradient
    net.register_parameter("unused", unused)
    ddp.broadcast_parameters(net)
    ref = [p.detach().clone() for p in net.parameters()]
    # every rank: gradient of its own shard
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(16, 6, generator=g), torch.randn(16, 2, generator=g)
    idx = ddp.shard_indices(16, rank, world, epoch=3, shuffle=True)
    loss = torch.nn.functional.mse_loss(net(X[idx]), Y[idx])
    loss.backward()
    ddp.average_gradients(net)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])
    buf = torch.arange(4, dtype=torch.float32) + rank
    ddp.allreduce_mean_(buf)
    # numpy payloads are pickled by value: torch tensors would travel as shared-memory file descriptors, which
    # vanish if this worker exits before the parent has read the queue
    q.put((rank, [t.numpy().copy() for t in ref], flat.detach().numpy().copy(), idx, buf.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out = [(r, [torch.from_numpy(t) for t in ps], torch.from_numpy(g), i, torch.from_numpy(b))
           for r, ps, g, i, b in out]
    (_, p0, g0, i0, b0), (_, p1, g1, i1, b1) = out
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)                               # broadcast made the replicas identical
    assert torch.allclose(g0, g1)                               # one all-reduce: identical mean gradients
    assert sorted(i0 + i1) == list(range(16))                   # shards partition the epoch
    assert torch.equal(b0, torch.tensor([0.5, 1.5, 2.5, 3.5])) and torch.equal(b0, b1)
    # the mean gradient equals the full-batch gradient of the union (equal shard sizes, MSE mean)
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(16, 6, generator=g), torch.randn(16, 2, generator=g)
    torch.nn.functional.mse_loss(net(X), Y).backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    # (the parameter registered directly on the Sequential comes first in .parameters())
    assert float(g0[:3].abs().max()) == 0.0                      # the unused parameter stays zero
    assert torch.allclose(g0[3:], full, atol=1e-6)


def test_shard_indices_padding_and_determinism():
    from stemgnn_b200 import ddp
    a = [ddp.shard_indices(10, r, 4, epoch=1) for r in range(4)]
    assert all(len(s) == 3 for s in a)                           # padded by wrap-around to 12
    assert set(sum(a, [])) == set(range(10))
    assert ddp.shard_indices(10, 2, 4, epoch=1) == a[2]
    assert ddp.shard_indices(10, 2, 4, epoch=2) != a[2]
    b = [ddp.shard_indices(10, r, 4, shuffle=False, drop_last=True) for r in range(4)]
    assert sorted(sum(b, [])) == list(range(8))
    assert ddp.world_size() == 1
    t = torch.ones(3)
    assert ddp.allreduce_mean_(t) is t                           # single process: no-op


def _graph_allreduce_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from stemgnn_b200 import ddp, runtime
    ddp.init_from_env(backend="gloo")
    # a stand-in "workspace": the library hands the hook raw pointers INTO it; the hook must average exactly that range
    ws = torch.zeros(4096, dtype=torch.uint8)
    view = ws[256:256 + 4 * 10].view(torch.float32)
    view.copy_(torch.arange(10, dtype=torch.float32) * (rank + 1))
    sentinel = ws[:256].clone(), ws[256 + 40:].clone()
    ga = runtime.GraphAllreduce(ws, None)
    ga.fn(ws.data_ptr() + 256, 10, None, None)           # what stemgnn_model_forward does through the function pointer
    ga.check()
    ok_bounds = torch.equal(ws[:256], sentinel[0]) and torch.equal(ws[256 + 40:], sentinel[1])
    # a pointer outside the workspace is refused, the error is parked (ctypes would swallow it) and re-raised by check()
    ga.fn(ws.data_ptr() + 4096, 4, None, None)
    refused = False
    try:
        ga.check()
    except RuntimeError:
        refused = True
    q.put((rank, view.numpy().copy(), ok_bounds, refused, ga.calls))
    dist.barrier()
    dist.destroy_process_group()


def test_graph_allreduce_hook_maps_pointers_into_the_workspace():
    """Host side of `stemgnn_fwd_opts_t.graph_allreduce` (global-batch graph semantics for data-parallel runs)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_graph_allreduce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(10, dtype=torch.float32) * 1.5).numpy()      # mean of 1x and 2x
    for _, got, ok_bounds, refused, calls in out:
        assert (got == want).all() and ok_bounds and refused and calls == 1
