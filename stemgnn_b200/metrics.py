"""Device-side validation metrics: the counterpart of the reference's
`de_normalized` + `utils.math_utils.evaluate` pair used by `handler.validate` (handler.py:74-82), computed by
`stemgnn_eval_metrics` on tensors that never leave HBM.  One D2H copy of (6, N) float64 per validation."""
import ctypes

import numpy as np
import torch

from . import _lib

_METHOD = {None: 0, "": 0, "z_score": 1, "min_max": 2}


def device_evaluate(forecast_norm, target_norm, normalize_method, statistic):
    """forecast_norm (count,H,N) float64 CUDA, target_norm (count,H,N) float32 CUDA  ->
    (score, score_by_node, score_norm) with score = (MAPE, MAE, RMSE) exactly as `evaluate(target, forecast)`,
    `evaluate(..., by_node=True)` and `evaluate(target_norm, forecast_norm)` of the reference return them."""
    lib = _lib.load()
    if not forecast_norm.is_cuda:
        raise RuntimeError("device_evaluate needs CUDA tensors (no CPU fallback)")
    f = forecast_norm.to(torch.float64).contiguous()
    t = target_norm.to(torch.float32).contiguous()
    count, H, N = f.shape
    dev = f.device
    method = _METHOD[normalize_method] if statistic else 0
    scale = shift = None
    if method == 1:
        std = np.asarray([1 if s == 0 else s for s in statistic["std"]], dtype=np.float64)
        scale, shift = std, np.asarray(statistic["mean"], dtype=np.float64)
    elif method == 2:
        lo = np.asarray(statistic["min"], dtype=np.float64)
        scale, shift = np.asarray(statistic["max"], dtype=np.float64) - lo + 1e-8, lo
    sc = torch.from_numpy(scale).to(dev) if method else None
    sh = torch.from_numpy(shift).to(dev) if method else None
    rows = count * H
    chunks = int(max(1, min(256, (rows + 63) // 64)))
    partial = torch.empty(chunks * 6 * N, dtype=torch.float64, device=dev)
    sums = torch.empty(6, N, dtype=torch.float64, device=dev)
    rc = lib.stemgnn_eval_metrics(f.data_ptr(), t.data_ptr(), count, H, N, method,
                                  sc.data_ptr() if method else None, sh.data_ptr() if method else None,
                                  partial.data_ptr(), chunks, sums.data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "stemgnn_eval_metrics")
    s = sums.cpu().numpy()                         # the ONE device->host copy of a validation
    per_node = float(rows)
    total = per_node * N

    def triple(k):
        return (s[k].sum() / total, s[k + 1].sum() / total, np.sqrt(s[k + 2].sum() / total))
    by_node = (s[0] / per_node, s[1] / per_node, np.sqrt(s[2] / per_node))
    return triple(0), by_node, triple(3)
