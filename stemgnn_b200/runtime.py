"""Host-side glue between torch tensors and the C ABI (include/stemgnn_b200.h).

torch is used here for device memory (caching allocator), streams and autograd bookkeeping only;
all arithmetic of the hot path happens inside libstemgnn_b200.so.  There is no CPU path: tensors
that are not on a CUDA device are rejected with a RuntimeError.
"""
import ctypes
from ctypes import byref

import torch

from . import _lib
from ._lib import BlockPtrs, Dims, FwdOpts, ModelPtrs

GEMM_AUTO, GEMM_FP32, GEMM_TC, GEMM_BF16 = 0, 1, 2, 3

# order in which parameters are handed to autograd (names = reference state_dict keys)
_BLOCK_FIELDS = [("weight", "weight"), ("forecast_w", "forecast.weight"), ("forecast_b", "forecast.bias"),
                 ("forecast_result_w", "forecast_result.weight"),
                 ("forecast_result_b", "forecast_result.bias"),
                 ("backcast_w", "backcast.weight"), ("backcast_b", "backcast.bias"),
                 ("shortcut_w", "backcast_short_cut.weight"), ("shortcut_b", "backcast_short_cut.bias")]
_TOP_FIELDS = [("weight_key", "weight_key"), ("weight_query", "weight_query"),
               ("gru_w_ih", "GRU.weight_ih_l0"), ("gru_w_hh", "GRU.weight_hh_l0"),
               ("gru_b_ih", "GRU.bias_ih_l0"), ("gru_b_hh", "GRU.bias_hh_l0"),
               ("fc0_w", "fc.0.weight"), ("fc0_b", "fc.0.bias"), ("fc2_w", "fc.2.weight"),
               ("fc2_b", "fc.2.bias")]


def param_slots(stack_cnt=2):
    """[(state_dict key, setter)] where setter(struct, ptr) stores a pointer into a ModelPtrs."""
    slots = []
    for field, key in _TOP_FIELDS:
        slots.append((key, (lambda s, p, f=field: setattr(s, f, p))))
    for i in range(stack_cnt):
        for field, key in _BLOCK_FIELDS:
            if i != 0 and field in ("backcast_w", "backcast_b"):
                continue
            slots.append((f"stock_block.{i}.{key}", (lambda s, p, i=i, f=field: setattr(s.block[i], f, p))))
        for g in range(6):
            for side in ("left", "right"):
                for kind, suffix in (("w", "weight"), ("b", "bias")):
                    fld = f"glu_{side}_{kind}"
                    slots.append((f"stock_block.{i}.GLUs.{g}.linear_{side}.{suffix}",
                                  (lambda s, p, i=i, f=fld, g=g: getattr(s.block[i], f).__setitem__(g, p))))
    return slots


_SLOTS = param_slots()
PARAM_KEYS = [k for k, _ in _SLOTS]


def _check_cuda_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"stemgnn_b200: `{name}` is on {t.device}; the B200 path needs CUDA tensors "
                           "(there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"stemgnn_b200: `{name}` must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"stemgnn_b200: `{name}` must be contiguous")


def build_ptrs(tensors_by_key):
    """ModelPtrs filled from {state_dict key: tensor or None}."""
    st = ModelPtrs()
    for key, setter in _SLOTS:
        t = tensors_by_key.get(key)
        if t is None:
            setter(st, None)
        else:
            _check_cuda_f32(t, key)
            setter(st, t.data_ptr())
    return st


def build_block_ptrs(tensors_by_key, prefix):
    full = {k: None for k in PARAM_KEYS}
    for k, v in tensors_by_key.items():
        full[k] = v
    st = build_ptrs(full)
    idx = int(prefix.split(".")[1])
    out = BlockPtrs()
    ctypes.memmove(byref(out), byref(st.block[idx]), ctypes.sizeof(BlockPtrs))
    return out


def workspace_bytes(dims, training):
    return int(_lib.load().stemgnn_workspace_bytes(byref(dims), 1 if training else 0))


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def alloc_workspace(dims, training, device):
    n = workspace_bytes(dims, training)
    # torch's caching allocator returns >= 512-byte aligned blocks
    return torch.empty(n, dtype=torch.uint8, device=device)


def make_opts(alpha, p, training, seed=0, offset=0, mask=None, gemm_mode=GEMM_AUTO, reuse_folded=False, graph_mode=0):
    o = FwdOpts()
    o.leaky_alpha = float(alpha)
    o.dropout_p = float(p)
    o.training = int(bool(training))
    o.dropout_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    o.dropout_offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    o.dropout_mask = mask.data_ptr() if mask is not None else None
    o.gemm_mode = int(gemm_mode)
    o.reuse_folded = int(bool(reuse_folded))
    o.graph_mode = int(graph_mode)
    o.dropout_offset_dev = None
    return o


class GraphAllreduce:
    """`stemgnn_fwd_opts_t.graph_allreduce` for data-parallel runs with GLOBAL-batch graph semantics: the library calls back
    (on the host, while it issues its launches) with a device buffer inside `workspace`; the buffer is replaced by its mean over
    the ranks with one `torch.distributed.all_reduce` enqueued on the current stream.  ctypes swallows exceptions raised in
    callbacks, so they are parked here and re-raised by `check()` after the C call returned."""

    def __init__(self, workspace, group=None):
        self.ws, self.group, self.error, self.calls = workspace, group, None, 0
        self.fn = _lib.ALLREDUCE_FN(self._cb)       # keep the trampoline alive as long as this object

    def _cb(self, ptr, n, _user, _stream):
        try:
            from . import ddp
            off = int(ptr) - self.ws.data_ptr()
            if off < 0 or off + 4 * n > self.ws.numel():
                raise RuntimeError("graph_allreduce: buffer outside the workspace")
            ddp.allreduce_mean_(self.ws[off:off + 4 * n].view(torch.float32), self.group)
            self.calls += 1
        except BaseException as exc:                 # noqa: BLE001 - re-raised in check()
            self.error = exc

    def install(self, opts):
        opts.graph_allreduce = self.fn
        opts.graph_allreduce_user = None
        return self

    def check(self):
        if self.error is not None:
            err, self.error = self.error, None
            raise RuntimeError("stemgnn_b200: graph all-reduce failed") from err


def model_forward_raw(dims, ptrs, opts, x, workspace, want_mul_L=False):
    """One call of stemgnn_model_forward on the current stream.  Returns (forecast, attention, mul_L)."""
    lib = _lib.load()
    _check_cuda_f32(x, "x")
    dev = x.device
    h_shape = (dims.B, dims.H, dims.N)
    forecast = torch.empty(h_shape, dtype=torch.float32, device=dev)
    attention = torch.empty((dims.N, dims.N), dtype=torch.float32, device=dev)
    mul_L = torch.empty((4, dims.N, dims.N), dtype=torch.float32, device=dev) if want_mul_L else None
    rc = lib.stemgnn_model_forward(byref(dims), byref(ptrs), byref(opts), x.data_ptr(),
                                   forecast.data_ptr(), attention.data_ptr(),
                                   mul_L.data_ptr() if want_mul_L else None,
                                   workspace.data_ptr(), workspace.numel(), _stream_ptr(dev))
    _lib.check(rc, "stemgnn_model_forward")
    return forecast, attention, mul_L


class StemGNNFunction(torch.autograd.Function):
    """autograd seam: forward/backward are single C-ABI calls (handler.py:161-164)."""

    @staticmethod
    def forward(ctx, x, cfg, *params):
        dims, alpha, p_drop, use_dropout, seed, offset, mask, gemm_mode = cfg[:8]
        tensors = dict(zip(PARAM_KEYS, params))
        ptrs = build_ptrs(tensors)
        opts = make_opts(alpha, p_drop if use_dropout else 0.0, True, seed, offset,
                         mask if use_dropout else None, gemm_mode)
        ws = alloc_workspace(dims, True, x.device)
        ddp_state = cfg[8] if len(cfg) > 8 else None
        ctx.graph_ar = None
        if ddp_state is not None and ddp_state.get("global_graph"):
            ctx.graph_ar = GraphAllreduce(ws, ddp_state.get("group")).install(opts)
        forecast, attention, _ = model_forward_raw(dims, ptrs, opts, x, ws)
        if ctx.graph_ar is not None:
            ctx.graph_ar.check()
        ctx.save_for_backward(x, *params)
        ctx.cfg = cfg
        ctx.ws = ws
        ctx.opts = opts
        ctx.mask = mask
        ctx.set_materialize_grads(False)
        return forecast, attention

    @staticmethod
    def backward(ctx, d_forecast, d_attention):
        lib = _lib.load()
        x, *params = ctx.saved_tensors
        dims = ctx.cfg[0]
        dev = x.device
        tensors = dict(zip(PARAM_KEYS, params))
        ptrs = build_ptrs(tensors)
        sizes = [t.numel() for t in params]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        grads, off = [], 0
        for t, n in zip(params, sizes):
            grads.append(flat[off:off + n].view_as(t))
            off += n
        gptrs = build_ptrs(dict(zip(PARAM_KEYS, grads)))
        need_dx = ctx.needs_input_grad[0]
        d_x = torch.empty_like(x) if need_dx else None
        if d_forecast is None:
            d_forecast = torch.zeros((dims.B, dims.H, dims.N), dtype=torch.float32, device=dev)
        d_forecast = d_forecast.contiguous()
        d_att = d_attention.contiguous() if d_attention is not None else None
        rc = lib.stemgnn_model_backward(byref(dims), byref(ptrs), byref(ctx.opts), x.data_ptr(),
                                        d_forecast.data_ptr(),
                                        d_att.data_ptr() if d_att is not None else None,
                                        byref(gptrs), d_x.data_ptr() if need_dx else None,
                                        ctx.ws.data_ptr(), ctx.ws.numel(), _stream_ptr(dev))
        _lib.check(rc, "stemgnn_model_backward")
        if ctx.graph_ar is not None:
            ctx.graph_ar.check()
            ctx.graph_ar = None
        ctx.ws = None
        ddp_group = ctx.cfg[8] if len(ctx.cfg) > 8 else None
        if ddp_group is not None:                 # the ONE collective of a training step
            from . import ddp
            ddp.allreduce_mean_(flat, ddp_group.get("group"))
        # stock_block.1.backcast_short_cut.* never reaches the output (base_model.py:70-74): the
        # reference leaves its .grad as None
        out = [None if k.startswith("stock_block.1.backcast_short_cut") else g
               for k, g in zip(PARAM_KEYS, grads)]
        return (d_x, None) + tuple(out)
