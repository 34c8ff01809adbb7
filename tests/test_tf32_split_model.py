"""The arithmetic model behind csrc/spec_tc.cu, on the CPU: (1) the single-pass errors MEASURED on the B200 (recorded below from
tests/test_spec_tc_gpu.py, same seeded inputs) identify operand truncation, not rounding; (2) the 3xTF32 split of that model is
an fp32-level product for any operand magnitude."""
import numpy as np
import pytest
import torch

from oracle import tf32_model as tm

# (M, N, K) -> max|C_gpu - C_fp64| / max|C_fp64| of ONE kind::tf32 pass, measured on a B200 (profiles/README.md)
MEASURED_SINGLE_PASS = {(128, 16, 8): 6.42e-4, (77, 32, 37): 9.35e-4, (300, 48, 358): 7.19e-4, (513, 256, 100): 7.92e-4}


def _inputs(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)          # the generator of tests/test_spec_tc_gpu.py::_tc_gemm
    ld = (K + 3) // 4 * 4
    A = torch.randn(M, ld, generator=g)
    B = torch.randn(N, ld, generator=g)
    return A[:, :K].numpy().copy(), B[:, :K].numpy().copy()


def _rel(c, ref):
    return float(np.abs(c - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("shape", sorted(MEASURED_SINGLE_PASS))
def test_measured_single_pass_error_is_the_truncation_model(shape):
    a, b = _inputs(*shape)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    e_trunc = _rel(tm.single_pass(a, b, tm.tf32_truncate), ref)
    e_round = _rel(tm.single_pass(a, b, tm.tf32_round_nearest), ref)
    meas = MEASURED_SINGLE_PASS[shape]
    assert abs(e_trunc - meas) <= 0.01 * meas, (e_trunc, meas)       # agrees to the printed 3 digits
    assert abs(e_round - meas) >= 0.3 * meas, (e_round, meas)        # a rounding tensor core would have been ~2x better


@pytest.mark.parametrize("scale", [1.0, 3e-4, 1e4])
def test_split_model_is_fp32_level_at_any_magnitude(scale):
    rng = np.random.default_rng(7)
    a = rng.standard_normal((96, 358)).astype(np.float32)
    b = (rng.standard_normal((48, 358)) * scale).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    e_split = _rel(tm.split_3xtf32(a, b), ref)
    e_fp32 = _rel((a @ b.T).astype(np.float64), ref)
    e_single = _rel(tm.single_pass(a, b), ref)
    assert e_split < 1e-6 and e_split < 3 * max(e_fp32, 2e-7)
    assert e_single > 100 * e_split


def test_hi_lo_are_exact_and_lo_fits_tf32_range():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(10000) * 10.0 ** rng.uniform(-20, 20, 10000)).astype(np.float32)
    hi = tm.tf32_truncate(x)
    lo = x - hi
    assert np.array_equal(hi + lo, x)                                   # exact decomposition in fp32
    nz = hi != 0
    assert np.all(np.abs(lo[nz]) < np.abs(hi[nz]) * 2.0 ** -10 * 1.0001)  # lo carries the 13 dropped bits
