"""Forecast error metrics with the semantics of the reference `utils/math_utils.py`
(microsoft/StemGNN): MAPE is |err|/|truth| + 1e-5 clipped at 5 (reference :32-34), MAE and RMSE
are plain means; `evaluate` reduces over everything, per step, per node, or per (step,node).
Inputs are arrays shaped [count, time_step, node]; ground truth comes first."""
import numpy as np

_MAPE_CLIP = 5.0


def MAPE(v, v_, axis=None):
    ratio = (np.abs(v_ - v) / np.abs(v) + 1e-5).astype(np.float64)
    return np.mean(np.minimum(ratio, _MAPE_CLIP), axis)


def masked_MAPE(v, v_, axis=None):
    """MAPE that ignores positions whose ground truth is exactly zero (unused by the handler)."""
    zero = (v == 0)
    ratio = np.abs(v_ - v) / np.abs(v)
    if not np.any(zero):
        return np.mean(ratio, axis).astype(np.float64)
    out = np.ma.masked_array(ratio, mask=zero).mean(axis=axis)
    return out.filled(np.nan) if isinstance(out, np.ma.MaskedArray) else out


def RMSE(v, v_, axis=None):
    return np.sqrt(np.mean(np.square(v_ - v), axis)).astype(np.float64)


def MAE(v, v_, axis=None):
    return np.mean(np.abs(v_ - v), axis).astype(np.float64)


def evaluate(y, y_hat, by_step=False, by_node=False):
    """-> (MAPE, MAE, RMSE) of prediction `y_hat` against truth `y`."""
    if by_step and by_node:
        axis = 0
    elif by_step:
        axis = (0, 2)
    elif by_node:
        axis = (0, 1)
    else:
        axis = None
    return MAPE(y, y_hat, axis), MAE(y, y_hat, axis), RMSE(y, y_hat, axis)
