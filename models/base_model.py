"""Drop-in `models.base_model` for microsoft/StemGNN whose forward/backward run on hand-written
sm_100a CUDA kernels (stemgnn_b200) instead of ATen.

Same public surface as the reference file (models/base_model.py @ dc7dea68): classes `GLU`,
`StockBlockLayer`, `Model`; the `Model(units, stack_cnt, time_step, multi_layer, horizon=1,
dropout_rate=0.5, leaky_rate=0.2, device='cpu')` constructor; `forward(x:(B,W,N)) ->
(forecast (B,H,N) | (B,1,N), attention (N,N))`; identical `state_dict()` keys/shapes and the same
parameter initialisers, so reference checkpoints load and `models.handler` / the reference
`main.py` work unchanged.  The modules below are parameter CONTAINERS: all arithmetic happens in
libstemgnn_b200.so through one C-ABI call per forward and one per backward.  There is no CPU
implementation here — calling the model with CPU tensors raises.
"""
import ctypes

import torch
import torch.nn as nn

from stemgnn_b200 import _lib, runtime


class GLU(nn.Module):
    """Parameter container for the reference GLU (base_model.py:6-13): two Linear layers whose
    product `left * sigmoid(right)` is evaluated inside the fused GEMM epilogue."""

    def __init__(self, input_channel, output_channel):
        super().__init__()
        self.linear_left = nn.Linear(input_channel, output_channel)
        self.linear_right = nn.Linear(input_channel, output_channel)

    def forward(self, x):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        runtime._check_cuda_f32(x2, "x")
        out = torch.empty(x2.shape[0], self.linear_left.out_features, device=x.device)
        lib = _lib.load()
        rc = lib.stemgnn_glu_gemm(x2.shape[0], out.shape[1], x2.shape[1], x2.data_ptr(), x2.shape[1],
                                  self.linear_left.weight.data_ptr(), self.linear_left.bias.data_ptr(),
                                  self.linear_right.weight.data_ptr(), self.linear_right.bias.data_ptr(),
                                  out.data_ptr(), out.shape[1], 0, runtime._stream_ptr(x.device))
        _lib.check(rc, "stemgnn_glu_gemm")
        return out.reshape(*lead, out.shape[1])


class StockBlockLayer(nn.Module):
    """Parameters of one spectral block (reference base_model.py:16-44); `forward` and
    `spe_seq_cell` keep the reference signatures for stage-level use (inference only)."""

    def __init__(self, time_step, unit, multi_layer, stack_cnt=0):
        super().__init__()
        self.time_step, self.unit, self.stack_cnt, self.multi = time_step, unit, stack_cnt, multi_layer
        T = time_step * multi_layer
        self.weight = nn.Parameter(torch.empty(1, 3 + 1, 1, T, T))
        nn.init.xavier_normal_(self.weight)
        self.forecast = nn.Linear(T, T)
        self.forecast_result = nn.Linear(T, time_step)
        if stack_cnt == 0:
            self.backcast = nn.Linear(T, time_step)
        self.backcast_short_cut = nn.Linear(time_step, time_step)
        self.output_channel = 4 * multi_layer
        d = time_step * self.output_channel
        # stage-level calls run the same tensor-core chain as the fused Model.forward path (fp16 hi/lo split operands
        # keep fp32 parity at the stage boundary); GEMM_FP32 forces the FFMA2 kernels
        self.gemm_mode = runtime.GEMM_AUTO
        self.GLUs = nn.ModuleList()
        for fan_in in (4 * time_step, d, d):
            self.GLUs.append(GLU(fan_in, d))    # real chain  (even index)
            self.GLUs.append(GLU(fan_in, d))    # imag chain  (odd index)

    def __setstate__(self, state):
        # a block unpickled from a REFERENCE-made whole-module checkpoint has no gemm_mode
        self.__dict__.update(state)
        self.__dict__.setdefault("gemm_mode", runtime.GEMM_AUTO)

    # -- helpers ---------------------------------------------------------------------------------
    def _block_ptrs(self):
        prefix = f"stock_block.{self.stack_cnt}"
        tensors = {f"{prefix}.{k}": v for k, v in self.named_parameters()}
        return runtime.build_block_ptrs(tensors, prefix)

    def _dims(self, B, N):
        return _lib.Dims(B, N, self.time_step, 1, self.multi)

    @torch.no_grad()
    def spe_seq_cell(self, input):
        """(B,4,1,N,W) or (B,4,N,W) -> (B,4,N,T): rfft -> 3 GLU layers on re/im -> irfft."""
        if input.dim() == 5:
            input = input.reshape(input.shape[0], -1, input.shape[3], input.shape[4])
        g = input.contiguous().float()
        runtime._check_cuda_f32(g, "input")
        B, _, N, W = g.shape
        dims = self._dims(B, N)
        ws = runtime.alloc_workspace(dims, False, g.device)
        out = torch.empty(B, 4, N, W * self.multi, device=g.device)
        bp = self._block_ptrs()
        rc = _lib.load().stemgnn_spe_seq_cell_forward(
            ctypes.byref(dims), ctypes.byref(bp), self.gemm_mode, g.data_ptr(), out.data_ptr(),
            ws.data_ptr(), ws.numel(), runtime._stream_ptr(g.device))
        _lib.check(rc, "stemgnn_spe_seq_cell_forward")
        return out

    @torch.no_grad()
    def forward(self, x, mul_L):
        """x: (B,1,N,W), mul_L: (4,N,N) -> (forecast (B,N,W), backcast (B,1,N,W) | None)."""
        xb = x.reshape(x.shape[0], x.shape[-2], x.shape[-1]).contiguous().float()
        mul_L = mul_L.contiguous().float()
        runtime._check_cuda_f32(xb, "x")
        B, N, W = xb.shape
        dims = self._dims(B, N)
        ws = runtime.alloc_workspace(dims, False, xb.device)
        forecast = torch.empty(B, N, W, device=xb.device)
        backcast = torch.empty(B, N, W, device=xb.device) if self.stack_cnt == 0 else None
        bp = self._block_ptrs()
        rc = _lib.load().stemgnn_block_forward(
            ctypes.byref(dims), ctypes.byref(bp), self.stack_cnt, self.gemm_mode, xb.data_ptr(),
            mul_L.data_ptr(), forecast.data_ptr(), backcast.data_ptr() if backcast is not None else None,
            ws.data_ptr(), ws.numel(), runtime._stream_ptr(xb.device))
        _lib.check(rc, "stemgnn_block_forward")
        return forecast, (backcast.unsqueeze(1) if backcast is not None else None)


class Model(nn.Module):
    def __init__(self, units, stack_cnt, time_step, multi_layer, horizon=1, dropout_rate=0.5,
                 leaky_rate=0.2, device='cpu'):
        super().__init__()
        if stack_cnt != 2:
            # the reference forward hard-codes result[0] + result[1] (base_model.py:174)
            raise ValueError("stemgnn_b200 supports stack_cnt == 2 (as does the reference forward)")
        self.unit = units
        self.stack_cnt = stack_cnt
        self.alpha = leaky_rate
        self.time_step = time_step
        self.horizon = horizon
        self.multi_layer = multi_layer
        self.dropout_rate = dropout_rate
        self.weight_key = nn.Parameter(torch.zeros(size=(units, 1)))
        nn.init.xavier_uniform_(self.weight_key.data, gain=1.414)
        self.weight_query = nn.Parameter(torch.zeros(size=(units, 1)))
        nn.init.xavier_uniform_(self.weight_query.data, gain=1.414)
        self.GRU = nn.GRU(time_step, units)        # parameter container; aten::gru is never called
        self.stock_block = nn.ModuleList(
            [StockBlockLayer(time_step, units, multi_layer, stack_cnt=i) for i in range(stack_cnt)])
        self.fc = nn.Sequential(nn.Linear(int(time_step), int(time_step)), nn.LeakyReLU(),
                                nn.Linear(int(time_step), horizon))
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.gemm_mode = runtime.GEMM_AUTO
        self.graph_mode = "poly"    # "eig": opt-in fused Laplacian + Jacobi eigendecomposition path (eval only)
        self.use_cuda_graph = True  # eval forwards with frozen weights replay ONE captured CUDA graph (26 launches -> 1)
        self._dropout_calls = 0
        self._philox_ctr = 0       # next free Philox block (dropout masks of successive steps never overlap)
        self._rt = None            # runtime cache (pointer struct, workspaces): never pickled
        self.to(device)

    # -- runtime cache ------------------------------------------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_rt"] = None
        state.pop("_ddp", None)          # process groups are not picklable (whole-module checkpoints)
        return state

    def __setstate__(self, state):
        # also accepts the __dict__ of a module pickled by the REFERENCE class (handler.py:24 saves whole modules):
        # fill what the reference object does not carry
        self.__dict__.update(state)
        d = self.__dict__
        d.setdefault("_rt", None)
        d.setdefault("gemm_mode", runtime.GEMM_AUTO)
        d.setdefault("graph_mode", "poly")
        d.setdefault("use_cuda_graph", True)
        d.setdefault("_dropout_calls", 0)
        d.setdefault("_philox_ctr", 0)
        if "dropout_rate" not in d:
            drop = d.get("_modules", {}).get("dropout")
            d["dropout_rate"] = float(getattr(drop, "p", 0.5))
        self._rt = None

    def _apply(self, fn, *a, **kw):
        self._rt = None
        return super()._apply(fn, *a, **kw)

    def _ordered_params(self):
        named = dict(self.named_parameters())
        return [named[k] for k in runtime.PARAM_KEYS]

    def invalidate_runtime(self):
        """Drops the cached pointer struct, workspaces and folded-weight state (call after writing parameters
        through `.data`, which does not bump `._version`)."""
        self._rt = None

    def _runtime(self):
        # raw device pointers are cached: re-validate them on every call so that a replaced Parameter
        # (`load_state_dict(assign=True)`, `p.data = ...`, `module.weight = nn.Parameter(...)`) can never leave
        # stale pointers behind
        params = self._ordered_params()
        sig = tuple((id(p), p.data_ptr()) for p in params)
        if self._rt is None or self._rt["ptr_sig"] != sig:
            self._rt = {"params": params, "ptrs": runtime.build_ptrs(dict(zip(runtime.PARAM_KEYS, params))),
                        "ptr_sig": sig, "ws": {}}
        return self._rt

    def _dims(self, B):
        return _lib.Dims(int(B), self.unit, self.time_step, self.horizon, self.multi_layer)

    # -- reference-compatible stage methods (inference only) -------------------------------------------
    @torch.no_grad()
    def latent_correlation_layer(self, x):
        """x (B,W,N) -> (mul_L (4,N,N), attention (N,N))   [reference base_model.py:136-149]"""
        _, attention, mul_L = self._forward_eval(x, want_mul_L=True)
        return mul_L, attention

    @torch.no_grad()
    def cheb_polynomial(self, laplacian):
        """[0, L, 2LL, 2L(2LL) - L] on the library's fp32 GEMM   [reference base_model.py:121-134]"""
        lap = laplacian.contiguous().float()
        runtime._check_cuda_f32(lap, "laplacian")
        N = lap.shape[0]
        out = torch.zeros(4, N, N, device=lap.device)
        out[1] = lap
        lib, st = _lib.load(), runtime._stream_ptr(lap.device)
        _lib.check(lib.stemgnn_sgemm(N, N, N, 2.0, lap.data_ptr(), N, 0, lap.data_ptr(), N, 0, 0.0,
                                     out[2].data_ptr(), N, st), "stemgnn_sgemm")
        out[3] = lap
        _lib.check(lib.stemgnn_sgemm(N, N, N, 2.0, lap.data_ptr(), N, 0, out[2].data_ptr(), N, 0, -1.0,
                                     out[3].data_ptr(), N, st), "stemgnn_sgemm")
        return out

    # -- forward ----------------------------------------------------------------------------------------
    def _forward_eval(self, x, want_mul_L=False):
        rt = self._runtime()
        dims = self._dims(x.shape[0])
        key = (x.shape[0], x.device)
        ws = rt["ws"].get(key)
        # the DFT-folded weights inside the workspace stay valid while no parameter was written to
        versions = tuple(p._version for p in rt["params"])
        graph_mode = 1 if self.graph_mode == "eig" else 0
        reuse = ws is not None and rt.get("folded_for") == (key, versions, self.gemm_mode)
        if ws is None:
            rt["ws"].clear()
            ws = rt["ws"][key] = runtime.alloc_workspace(dims, False, x.device)
        sig = (key, versions, self.gemm_mode, graph_mode)
        if reuse and self.use_cuda_graph and not want_mul_L:
            # frozen weights: the whole forward (prep, GRU recurrence, graph, two spectral blocks, head) is replayed as one
            # CUDA graph; inputs / outputs live in static buffers and are copied in / out (stream-ordered, no host sync)
            cg = rt.get("cuda_graph")
            if cg is None or cg["sig"] != sig:
                cg = None
                try:
                    xs = torch.empty_like(x)
                    xs.copy_(x)
                    opts = runtime.make_opts(self.alpha, 0.0, False, gemm_mode=self.gemm_mode, reuse_folded=True,
                                             graph_mode=graph_mode)
                    torch.cuda.synchronize(x.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        f_s, a_s, _ = runtime.model_forward_raw(dims, rt["ptrs"], opts, xs, ws, False)
                    rt["captures"] = rt.get("captures", 0) + 1
                    cg = rt["cuda_graph"] = {"sig": sig, "graph": g, "x": xs, "forecast": f_s, "attention": a_s, "opts": opts}
                except Exception:                       # capture unsupported here: stay on the eager launch path
                    self.use_cuda_graph = False
                    rt.pop("cuda_graph", None)
            if cg is not None:
                cg["x"].copy_(x)
                cg["graph"].replay()
                return cg["forecast"].clone(), cg["attention"].clone(), None
        rt["folded_for"] = None
        rt.pop("cuda_graph", None) if not reuse else None
        opts = runtime.make_opts(self.alpha, 0.0, False, gemm_mode=self.gemm_mode, reuse_folded=reuse,
                                 graph_mode=graph_mode)
        out = runtime.model_forward_raw(dims, rt["ptrs"], opts, x, ws, want_mul_L)
        rt["folded_for"] = (key, versions, self.gemm_mode)       # only after a successful call
        return out

    def inference_session(self, batch_size):
        """Frozen-weight inference from / to host buffers: one H2D, one CUDA-graph replay, one D2H per call
        (`stemgnn_b200/session.py`; the reference's `inference` loop, handler.py:34-40, without per-call host work)."""
        from stemgnn_b200.session import InferenceSession
        return InferenceSession(self, batch_size)

    def forward(self, x, dropout_mask=None):
        """x: (B, W, N) float32 CUDA tensor.  `dropout_mask` (optional, tests): explicit {0,1}
        keep-mask (B,N,N) replacing the Philox mask in training mode."""
        if x.dim() != 3 or x.shape[1] != self.time_step or x.shape[2] != self.unit:
            raise RuntimeError(f"expected x of shape (B,{self.time_step},{self.unit}), got {tuple(x.shape)}")
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        runtime._check_cuda_f32(x, "x")
        rt = self._runtime()
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in rt["params"]))
        use_dropout = self.training and (self.dropout_rate > 0 or dropout_mask is not None)
        if not needs_grad and not use_dropout:
            forecast, attention, _ = self._forward_eval(x)
        else:
            seed, offset = 0, 0
            if use_dropout and dropout_mask is None:
                seed = int(torch.initial_seed()) ^ 0x5DEECE66D
                rank = (getattr(self, "_ddp", None) or {}).get("rank", 0)
                seed = (seed + rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF    # replicas draw different masks
                self._dropout_calls += 1
                # running Philox block counter: batches of different size (drop_last=False) never overlap
                offset = self._philox_ctr = getattr(self, "_philox_ctr", 0) + 1
                self._philox_ctr += (x.shape[0] * self.unit * self.unit + 3) // 4
            mask = None
            if dropout_mask is not None:
                mask = dropout_mask.to(device=x.device, dtype=torch.uint8).contiguous()
            ddp_state = getattr(self, "_ddp", None)
            cfg = (self._dims(x.shape[0]), self.alpha, self.dropout_rate, use_dropout, seed, offset, mask,
                   self.gemm_mode, ddp_state if ddp_state and ddp_state.get("enabled") else None)
            forecast, attention = runtime.StemGNNFunction.apply(x, cfg, *rt["params"])
        if self.horizon == 1:
            return forecast.reshape(x.shape[0], 1, self.unit), attention
        return forecast, attention
