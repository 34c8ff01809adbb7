"""Captured-graph training step (reference: models/handler.py:126-130, 156-166).

The reference's step is `zero_grad -> forward -> MSELoss -> backward -> optimizer.step -> float(loss)`: ~210 kernel
launches driven from Python plus one host synchronisation per step.  `FusedTrainer` keeps the same arithmetic but
  * flattens parameters, gradients and optimiser state into three contiguous fp32 buffers (the model's Parameters
    become views of the flat buffer, so `state_dict()` / checkpoints are unaffected);
  * runs the whole step as ONE CUDA graph replay: memset(grad) -> stemgnn_model_forward (training, Philox dropout with a
    device-side offset counter) -> stemgnn_mse_loss_grad -> stemgnn_model_backward -> [NCCL all-reduce of the flat
    gradient] -> stemgnn_optimizer_step (one fused RMSprop / Adam launch) -> stemgnn_counters_tick;
  * accumulates the loss in a device scalar that is read once per epoch (`pop_loss()`), not once per step;
  * keeps learning rate and step count in device memory so the schedule changes without re-capturing.
There is no CPU path: the model must live on a CUDA device.
"""
import ctypes
from ctypes import byref

import torch

from . import _lib, runtime

RMSPROP, ADAM = 0, 1


class FusedTrainer:
    def __init__(self, model, optimizer="RMSProp", lr=1e-4, alpha=0.99, betas=(0.9, 0.999), eps=1e-8,
                 use_graph=True, warmup_eager=2, seed=None):
        params = model._ordered_params()
        if not params[0].is_cuda:
            raise RuntimeError("FusedTrainer needs the model on a CUDA device (no CPU fallback)")
        self.model, self.dev = model, params[0].device
        self.kind = RMSPROP if optimizer == "RMSProp" else ADAM
        self.h0, self.h1 = (alpha, 0.0) if self.kind == RMSPROP else (betas[0], betas[1])
        self.eps = eps
        self.use_graph, self.warmup_eager = use_graph, warmup_eager
        self.lib = _lib.load()
        # ---- flat buffers: parameters become views (same values, same state_dict) ----------------------------------
        sizes = [p.numel() for p in params]
        self.n = sum(sizes)
        self.flat_p = torch.empty(self.n, dtype=torch.float32, device=self.dev)
        self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=self.dev)
        self.s1 = torch.zeros(self.n, dtype=torch.float32, device=self.dev)
        self.s2 = torch.zeros(self.n if self.kind == ADAM else 1, dtype=torch.float32, device=self.dev)
        off, gviews = 0, []
        with torch.no_grad():
            for p, n in zip(params, sizes):
                self.flat_p[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
                gviews.append(self.flat_g[off:off + n].view_as(p))
                off += n
        model.invalidate_runtime()
        self._grad_views = dict(zip(runtime.PARAM_KEYS, gviews))
        self.gptrs = runtime.build_ptrs(self._grad_views)
        self.lr_dev = torch.tensor([lr], dtype=torch.float32, device=self.dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.drop_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.lr = lr
        seed = int(torch.initial_seed()) if seed is None else int(seed)
        rank = (getattr(model, "_ddp", None) or {}).get("rank", 0)
        self.seed = ((seed ^ 0x5DEECE66D) + rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        self._slots = {}            # batch size -> dict(x, y, forecast, attention, d_forecast, ws, dims, opts, graph, calls)
        self.steps_done = 0

    # ---- public API ------------------------------------------------------------------------------------------------
    def set_lr(self, lr):
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    def pop_loss(self):
        """Sum of the per-step mean losses since the last call (ONE host sync; handler.py:166 syncs every step)."""
        v = float(self.loss_dev.item())
        self.loss_dev.zero_()
        return v

    def grads(self):
        """{state_dict key: gradient view} of the last step (views of the flat gradient buffer)."""
        return self._grad_views

    def state_dict(self):
        return {"kind": self.kind, "lr": self.lr, "step": int(self.step_dev.item()), "dropout_ctr": int(self.drop_dev.item()),
                "state1": self.s1.detach().cpu(), "state2": self.s2.detach().cpu(), "seed": self.seed}

    def load_state_dict(self, sd):
        if int(sd["kind"]) != self.kind:
            raise RuntimeError("optimizer kind mismatch")
        self.set_lr(sd["lr"])
        self.step_dev.fill_(int(sd["step"]))
        self.drop_dev.fill_(int(sd["dropout_ctr"]))
        self.s1.copy_(sd["state1"])
        self.s2.copy_(sd["state2"])
        self.seed = int(sd["seed"])
        for slot in self._slots.values():      # the seed is baked into captured graphs
            slot["graph"], slot["calls"] = None, 0

    def step(self, x, y):
        """One training step on the batch (x (B,W,N), y (B,H,N)); asynchronous — nothing is copied back."""
        m = self.model
        B = int(x.shape[0])
        slot = self._slots.get(B)
        if slot is None:
            slot = self._slots[B] = self._make_slot(B)
        slot["x"].copy_(x, non_blocking=True)
        slot["y"].copy_(y, non_blocking=True)
        if not self.use_graph or slot["calls"] < self.warmup_eager:
            self._body(slot)
        else:
            if slot["graph"] is None:
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body(slot)             # capture does not execute: replay below runs this step
                slot["graph"] = g
            slot["graph"].replay()
        slot["calls"] += 1
        self.steps_done += 1
        m._dropout_calls += 1

    # ---- internals -------------------------------------------------------------------------------------------------
    def _make_slot(self, B):
        m = self.model
        dims = m._dims(B)
        dev = self.dev
        use_dropout = m.dropout_rate > 0
        opts = runtime.make_opts(m.alpha, m.dropout_rate if use_dropout else 0.0, True, seed=self.seed, offset=0,
                                 gemm_mode=m.gemm_mode)
        opts.dropout_offset_dev = self.drop_dev.data_ptr()
        ws = runtime.alloc_workspace(dims, True, dev)
        ddp_state = getattr(m, "_ddp", None) or {}
        graph_ar = None
        if ddp_state.get("enabled") and ddp_state.get("global_graph"):    # global-batch graph: 2 small all-reduces (ddp.py)
            graph_ar = runtime.GraphAllreduce(ws, ddp_state.get("group")).install(opts)
        return {"dims": dims, "opts": opts, "graph_ar": graph_ar,
                "x": torch.empty(B, m.time_step, m.unit, dtype=torch.float32, device=dev),
                "y": torch.empty(B, m.horizon, m.unit, dtype=torch.float32, device=dev),
                "forecast": torch.empty(B, m.horizon, m.unit, dtype=torch.float32, device=dev),
                "attention": torch.empty(m.unit, m.unit, dtype=torch.float32, device=dev),
                "d_forecast": torch.empty(B, m.horizon, m.unit, dtype=torch.float32, device=dev),
                "ws": ws, "graph": None, "calls": 0,
                "drop_inc": (B * m.unit * m.unit + 3) // 4 + 1}

    def _body(self, s):
        lib, m = self.lib, self.model
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        ptrs = m._runtime()["ptrs"]
        self.flat_g.zero_()
        rc = lib.stemgnn_model_forward(byref(s["dims"]), byref(ptrs), byref(s["opts"]), s["x"].data_ptr(),
                                       s["forecast"].data_ptr(), s["attention"].data_ptr(), None,
                                       s["ws"].data_ptr(), s["ws"].numel(), st)
        _lib.check(rc, "stemgnn_model_forward")
        rc = lib.stemgnn_mse_loss_grad(s["forecast"].data_ptr(), s["y"].data_ptr(), s["forecast"].numel(),
                                       s["d_forecast"].data_ptr(), self.loss_dev.data_ptr(), st)
        _lib.check(rc, "stemgnn_mse_loss_grad")
        rc = lib.stemgnn_model_backward(byref(s["dims"]), byref(ptrs), byref(s["opts"]), s["x"].data_ptr(),
                                        s["d_forecast"].data_ptr(), None, byref(self.gptrs), None,
                                        s["ws"].data_ptr(), s["ws"].numel(), st)
        _lib.check(rc, "stemgnn_model_backward")
        if s["graph_ar"] is not None:
            s["graph_ar"].check()
        ddp_state = getattr(m, "_ddp", None)
        if ddp_state and ddp_state.get("enabled"):          # the ONE collective of a training step
            from . import ddp
            ddp.allreduce_mean_(self.flat_g, ddp_state.get("group"))
        rc = lib.stemgnn_optimizer_step(self.kind, self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.s1.data_ptr(),
                                        self.s2.data_ptr(), self.n, self.lr_dev.data_ptr(), self.h0, self.h1, self.eps,
                                        self.step_dev.data_ptr(), st)
        _lib.check(rc, "stemgnn_optimizer_step")
        rc = lib.stemgnn_counters_tick(self.step_dev.data_ptr(), self.drop_dev.data_ptr(), s["drop_inc"], st)
        _lib.check(rc, "stemgnn_counters_tick")
