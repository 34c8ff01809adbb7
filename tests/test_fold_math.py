"""The DFT folds used by the CUDA path (oracle/fold_math.py restates them) are algebraically identical to the
reference's rfft -> Linear and irfft -> weight -> Linear pipelines (oracle/stemgnn_oracle.py)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import fold_math as fm


@settings(max_examples=25, deadline=None)
@given(W=st.integers(2, 16), d=st.integers(1, 12), rows=st.integers(1, 6), seed=st.integers(0, 10_000))
def test_fold_in_equals_rfft_then_linear(W, d, rows, seed):
    rng = np.random.default_rng(seed)
    g = rng.normal(size=(rows, 4, W))
    g[:, 0] = 0.0                                           # Chebyshev term 0 is zeros (base_model.py:129)
    w = rng.normal(size=(2, d, 4 * W))
    spec = np.fft.fft(g, axis=-1)
    for chain, part in ((0, spec.real), (1, spec.imag)):
        direct = part.reshape(rows, 4 * W) @ w[chain].T                       # Linear on [k*W + f]
        folded = g[:, 1:].reshape(rows, 3 * W) @ fm.fold_in(w[chain], W, chain).T
        np.testing.assert_allclose(folded, direct, rtol=1e-9, atol=1e-9)
        full = g.reshape(rows, 4 * W) @ fm.fold_in(w[chain], W, chain, kfirst=0, nk=4).T   # stage-level API
        np.testing.assert_allclose(full, direct, rtol=1e-9, atol=1e-9)


@settings(max_examples=25, deadline=None)
@given(T=st.integers(2, 20), W=st.integers(1, 6), rows=st.integers(1, 5), seed=st.integers(0, 10_000),
       with_backcast=st.booleans())
def test_fold_out_equals_irfft_weight_heads(T, W, rows, seed, with_backcast):
    rng = np.random.default_rng(seed)
    real3 = rng.normal(size=(rows, 4, T))
    imag3 = rng.normal(size=(rows, 4, T))
    weight = rng.normal(size=(4, T, T))
    fw = rng.normal(size=(T, T))
    bw = rng.normal(size=(W, T)) if with_backcast else None
    # reference order: irfft over bins 0..T/2 (Im of DC / Nyquist ignored), sum_k y_k @ weight[k], then Linears
    y = np.fft.irfft((real3 + 1j * imag3)[..., : T // 2 + 1], n=T, axis=-1)
    igfted = np.einsum("rkt,ktu->ru", y, weight)
    direct = igfted @ (fw if bw is None else np.concatenate([fw, bw], 0)).T
    a = np.concatenate([real3.reshape(rows, 4 * T), imag3.reshape(rows, 4 * T)], axis=1)
    folded = a @ fm.fold_out(weight, fw, bw).T
    np.testing.assert_allclose(folded, direct, rtol=1e-8, atol=1e-8)


def test_dead_bins_are_exact_zero_rows():
    T = 60
    wt = fm.fold_out(np.random.default_rng(0).normal(size=(4, T, T)), np.eye(T), None)     # (T, 8T)
    cols = wt.reshape(T, 2, 4, T)                                                            # [chain][k][f]
    assert np.all(cols[:, 0, :, T // 2 + 1:] == 0.0)          # real bins above Nyquist never contribute
    assert np.all(cols[:, 1, :, 0] == 0.0) and np.all(cols[:, 1, :, T // 2:] == 0.0)   # Im(DC), Im(>= Nyquist)
    live = int((np.abs(cols[:, 0]).sum(axis=(0, 1)) > 0).sum()), int((np.abs(cols[:, 1]).sum(axis=(0, 1)) > 0).sum())
    assert live == (31, 29)                                    # 124 + 116 live columns of 480
