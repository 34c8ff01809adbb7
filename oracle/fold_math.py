"""TEST INFRASTRUCTURE — numpy statement of the two weight folds the CUDA path relies on
(stemgnn_b200/csrc/spectral.cu: fold_in_kernel, irfft_table_kernel + the RI / woutT GEMMs), so that the
algebra is pinned on the CPU independently of any kernel:

  fold_in :  Linear(W0) o two-sided rfft_W          ==  Linear(W0') on the time samples
  fold_out:  {forecast, backcast} o sum_k (.)@weight[k] o irfft_T  ==  one (T+W) x 8T map on [real3 | imag3]

`tests/test_fold_math.py` checks both against the direct (unfolded) oracle with hypothesis-generated shapes.
"""
import numpy as np


def twiddle_in(W, chain):
    """tw[f, t]: real part (chain 0) / imaginary part (chain 1) of exp(-2 pi i f t / W)."""
    f, t = np.meshgrid(np.arange(W), np.arange(W), indexing="ij")
    ang = 2.0 * np.pi * ((f * t) % W) / W
    return np.cos(ang) if chain == 0 else -np.sin(ang)


def fold_in(w, W, chain, kfirst=1, nk=3):
    """w: (d, 4W) acting on spectra laid out [k*W + f]  ->  (d, nk*W) acting on samples [k'*W + t]."""
    d = w.shape[0]
    tw = twiddle_in(W, chain)                                   # (f, t)
    out = np.zeros((d, nk * W), dtype=np.float64)
    for kp in range(nk):
        blk = w[:, (kp + kfirst) * W:(kp + kfirst + 1) * W].astype(np.float64)   # (d, f)
        out[:, kp * W:(kp + 1) * W] = blk @ tw
    return out


def irfft_table(T):
    """ic[chain, f, t]: contribution of Re (0) / Im (1) of bin f to sample t of irfft(n=T) — bins 0..T/2 only,
    Im of DC and (even T) Nyquist ignored, norm 1/T."""
    ic = np.zeros((2, T, T), dtype=np.float64)
    half = T // 2
    t = np.arange(T)
    for f in range(half + 1):
        edge = f == 0 or (T % 2 == 0 and f == half)
        ang = 2.0 * np.pi * ((f * t) % T) / T
        ic[0, f] = (1.0 if edge else 2.0) * np.cos(ang) / T
        ic[1, f] = 0.0 if edge else -2.0 * np.sin(ang) / T
    return ic


def fold_out(weight, forecast_w, backcast_w):
    """weight: (4, T, T); forecast_w: (T, T); backcast_w: (W, T) or None  ->  woutT (T [+ W], 8T) such that
    pre = [real3 | imag3] @ woutT.T  with real3/imag3 laid out [k*T + f]."""
    K, T, _ = weight.shape
    ic = irfft_table(T)
    ri = np.zeros((2, K * T, T), dtype=np.float64)               # RI[c][k*T+f][u] = sum_t ic[c][f][t] weight[k][t][u]
    for c in range(2):
        for k in range(K):
            ri[c, k * T:(k + 1) * T] = ic[c] @ weight[k].astype(np.float64)
    ri = ri.reshape(2 * K * T, T)
    heads = forecast_w if backcast_w is None else np.concatenate([forecast_w, backcast_w], axis=0)
    return heads.astype(np.float64) @ ri.T
