"""Data-parallel training of the hot path: one process per GPU (torchrun), replicated parameters,
batches sharded across ranks, ONE all-reduce of the flat fp32 gradient per step.

The reference has no distributed code; this is the only collective the build adds.  The CUDA backward
(`stemgnn_model_backward`) writes every parameter gradient into one flat buffer, so the collective is a
single NCCL all-reduce of P(N) floats (5.8 MB at N=358) issued on the backward stream right after the
last backward kernel — see `runtime.StemGNNFunction.backward`.  Parameters that never receive a
gradient (`stock_block.1.backcast_short_cut.*`, base_model.py:70-74) are zeros in the buffer on every
rank, so no `find_unused_parameters` machinery is needed.

Graph semantics: by default each replica builds the latent graph from ITS shard of the batch
(`torch.mean(attention, dim=0)`, base_model.py:140) — standard DDP semantics, not identical to one
process with the global batch.  `attach(model, global_graph=True)` restores the global-batch semantics
with a second, small exchange: the batch-mean attention (N x N) and its degree vector are all-reduced in
the forward before the Laplacian is formed, and the gradient with respect to that attention is
all-reduced in the backward (`stemgnn_fwd_opts_t.graph_allreduce`, `runtime.GraphAllreduce`).  With
equal shards an N-rank step then equals the single-process step on the concatenated batch (up to fp32
summation order) — `tests/ddp_global_graph_check.py`.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_mean_(flat, group=None):
    """In-place mean over ranks of one flat gradient buffer (the single collective of a step)."""
    ws = world_size(group)
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(ws)
    return flat


def broadcast_parameters(module, src=0, group=None):
    """Replicas start from rank `src`'s parameters and buffers."""
    if world_size(group) > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=src, group=group)     # in-place on the Parameter: bumps ._version
        if hasattr(module, "invalidate_runtime"):
            module.invalidate_runtime()                      # folded / pre-split weight caches are stale now


def attach(model, group=None, global_graph=False):
    """Marks a stemgnn_b200 Model for data-parallel training: its backward all-reduces the flat
    gradient buffer before handing gradients to autograd.  global_graph=True additionally exchanges the
    batch-mean attention so that every replica uses the graph of the GLOBAL batch (module docstring)."""
    ws = world_size(group)
    model._ddp = {"group": group, "enabled": ws > 1, "rank": dist.get_rank(group) if ws > 1 else 0,
                  "global_graph": bool(global_graph) and ws > 1}
    broadcast_parameters(model, 0, group)
    return model


def average_gradients(module, group=None):
    """Generic fallback for modules whose gradients are separate tensors: flatten, one all-reduce,
    scatter back (parameters without a gradient contribute zeros, so every rank agrees on the layout)."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params or world_size(group) == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    allreduce_mean_(flat, group)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n


def shard_indices(n_items, rank, world, epoch=0, shuffle=True, seed=0, drop_last=False):
    """Indices of this rank's shard of an epoch: a seeded permutation common to all ranks, padded by
    wrap-around to a multiple of `world` (or truncated with drop_last), then strided by rank."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        order = torch.randperm(n_items, generator=g).tolist()
    else:
        order = list(range(n_items))
    if drop_last:
        order = order[:n_items - n_items % world]
    elif len(order) % world:
        order += order[:world - len(order) % world]
    return order[rank::world]


def train_ddp(train_data, valid_data, args, result_file):
    """Data-parallel counterpart of `models.handler.train` (same arguments / return value; run under
    torchrun with one process per GPU).  `args.batch_size` is the PER-GPU batch.  Rank 0 validates and
    checkpoints; every rank returns the same (metrics, normalize_statistic)."""
    import json
    import time

    import numpy as np
    import torch.nn as nn
    import torch.utils.data as torch_data

    from data_loader.forecast_dataloader import ForecastDataset
    from models import handler
    from models.base_model import Model

    rank, world, local_rank = init_from_env()
    device = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device(args.device)
    node_cnt = train_data.shape[1]
    torch.manual_seed(0)
    model = Model(node_cnt, 2, args.window_size, args.multi_layer, horizon=args.horizon).to(device)
    attach(model)
    stat = handler._norm_statistic(train_data, args.norm_method)
    if rank == 0 and stat is not None:
        os.makedirs(result_file, exist_ok=True)
        with open(os.path.join(result_file, 'norm_stat.json'), 'w') as f:
            json.dump(stat, f)
    if args.optimizer == 'RMSProp':
        optim = torch.optim.RMSprop(params=model.parameters(), lr=args.lr, eps=1e-08)
    else:
        optim = torch.optim.Adam(params=model.parameters(), lr=args.lr, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.ExponentialLR(optimizer=optim, gamma=args.decay_rate)
    kw = dict(window_size=args.window_size, horizon=args.horizon, normalize_method=args.norm_method,
              norm_statistic=stat)
    train_set = ForecastDataset(train_data, **kw)
    valid_loader = torch_data.DataLoader(ForecastDataset(valid_data, **kw), batch_size=args.batch_size,
                                         shuffle=False, num_workers=0)
    criterion = nn.MSELoss(reduction='mean').to(device)
    best, stale, metrics = np.inf, 0, {}
    for epoch in range(args.epoch):
        t0 = time.time()
        model.train()
        idx = shard_indices(len(train_set), rank, world, epoch=epoch, shuffle=True)
        loader = torch_data.DataLoader(torch_data.Subset(train_set, idx), batch_size=args.batch_size,
                                       shuffle=False, drop_last=False, num_workers=0)
        total, cnt = torch.zeros((), device=device), 0
        for inputs, target in loader:
            inputs, target = inputs.to(device), target.to(device)
            model.zero_grad()
            forecast, _ = model(inputs)
            loss = criterion(forecast, target)
            loss.backward()                       # flat-gradient all-reduce happens inside
            optim.step()
            total += loss.detach()                # no per-step host sync (reference: float(loss))
            cnt += 1
        if rank == 0:
            print('| end of epoch {:3d} | time: {:5.2f}s | train_total_loss {:5.4f}'.format(
                epoch, time.time() - t0, float(total) / max(cnt, 1)))
            handler.save_model(model, result_file, epoch)
        if (epoch + 1) % args.exponential_decay_step == 0:
            sched.step()
        if (epoch + 1) % args.validate_freq == 0:
            flag = torch.zeros(2, device=device)
            if rank == 0:
                metrics = handler.validate(model, valid_loader, device, args.norm_method, stat, node_cnt,
                                           args.window_size, args.horizon, result_file=result_file)
                if best > metrics['mae']:
                    best, stale = metrics['mae'], 0
                    handler.save_model(model, result_file)
                else:
                    stale += 1
                flag[0], flag[1] = stale, 1.0
            if world > 1:
                dist.broadcast(flag, src=0)
            stale = int(flag[0])
        if args.early_stop and stale >= args.early_stop_step:
            break
    if world > 1:
        obj = [metrics]
        dist.broadcast_object_list(obj, src=0)
        metrics = obj[0]
    return metrics, stat
