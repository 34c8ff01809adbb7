"""tcgen05 GEMMs with 3xTF32 split operands (csrc/spec_tc.cu): unit parity against fp64 references through the
C ABI test hooks, and the stage taps of the spectral block (graph Fourier transform, fused output map + heads) against the
reference goldens in the default mode."""
import numpy as np
import pytest
import torch

from oracle import torch_port as tp
from tests.helpers import assert_close, build_model, cases, golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tc_gemm(M, N, K, split, seed, lda=None, ldb=None, scale_b=1.0):
    from stemgnn_b200 import _lib, runtime
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    lda, ldb = lda or (K + 3) // 4 * 4, ldb or (K + 3) // 4 * 4
    A = torch.randn(M, lda, generator=g)
    B = torch.randn(N, ldb, generator=g) * scale_b
    A[:, K:] = float("nan")          # the pitch padding must never be read as data (logical width K, TMA zero fill)
    B[:, K:] = float("nan")
    Ad, Bd = A.to(DEV), B.to(DEV)
    C = torch.full((M, N), float("nan"), device=DEV)
    rc = lib.stemgnn_tc_gemm(M, N, K, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, C.data_ptr(), N, split,
                             runtime._stream_ptr(torch.device(DEV)))
    _lib.check(rc, "tc_gemm")
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ B[:, :K].double().t()
    return C.cpu().double(), ref


@pytest.mark.parametrize("M,N,K", [(128, 16, 8), (77, 32, 37), (300, 48, 358), (1074, 32, 358), (1000, 240, 480),
                                   (513, 256, 100), (11456, 80, 480), (480, 240, 1433)])
def test_tc3_gemm_split_is_fp32_level(M, N, K):
    out, ref = _tc_gemm(M, N, K, 1, seed=M + N + K)
    assert torch.isfinite(out).all()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item() / scale
    out1, _ = _tc_gemm(M, N, K, 0, seed=M + N + K)
    err1 = (out1 - ref).abs().max().item() / scale
    print(f"M={M} N={N} K={K}: max|err|/max|ref|  3xTF32 split {err:.2e}   single TF32 pass {err1:.2e}")
    # fp32-level: the residual is the tensor core's truncation of the running sum at every accumulating MMA, linear in K
    # (measured 4e-7 at K = 8 ... 1.2e-5 at K = 1433 with a single accumulator; csrc/spec_tc.cu keeps the cross terms apart)
    assert err < 3e-6 + 8e-9 * K, "3xTF32 split operands must give an fp32-level product"
    assert err1 < 5e-3 and err < 0.05 * err1


def test_tc3_gemm_wide_dynamic_range():
    """hi/lo are split per element (exponent-independent): tiny and large magnitudes keep the same relative accuracy."""
    out, ref = _tc_gemm(256, 64, 358, 1, seed=5, scale_b=3e-4)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-6, err


@pytest.mark.parametrize("B,N,W", [(32, 358, 12), (8, 70, 12), (5, 37, 8), (4, 325, 12), (3, 24, 12), (2, 140, 12)])
def test_gft_tc_vs_fp64(B, N, W):
    from stemgnn_b200 import _lib, runtime
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * 1000 + N)
    mul_L = torch.randn(4, N, N, generator=g) / N ** 0.5
    x = torch.randn(B, W, N, generator=g)
    Np = (N + 3) // 4 * 4
    scratch = torch.full((4 * N * Np + B * W * Np,), float("nan"), device=DEV)
    G = torch.full((B * N, 3 * W), float("nan"), device=DEV)
    Ld, xd = mul_L.to(DEV), x.to(DEV)           # keep the device copies alive across the call
    rc = lib.stemgnn_gft_forward(Ld.data_ptr(), xd.data_ptr(), G.data_ptr(), B, N, W,
                                 scratch.data_ptr(), runtime._stream_ptr(torch.device(DEV)))
    _lib.check(rc, "gft_forward")
    torch.cuda.synchronize()
    # gfted[b,k,n,t] = sum_m mul_L[k][n][m] x[b][t][m]   (base_model.py:63 with x as (B,1,N,W))
    ref = torch.einsum("knm,btm->bnkt", mul_L[1:].double(), x.double()).reshape(B * N, 3 * W)
    out = G.cpu().double()
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    print(f"B={B} N={N} W={W}: gfted max|err|/max|ref| = {err:.2e}")
    assert err < 6e-6


@pytest.mark.parametrize("name", ["tiny_taps", "odd_h1_taps", "multi2_w8"])
def test_block_stage_taps_default_mode_vs_reference_golden(name):
    """StockBlockLayer through the stage API in the DEFAULT mode: the tcgen05 graph Fourier transform, the kind::f16 GLU
    chain and the fused output-map + heads kernel against the reference's block outputs at the north_star tolerance."""
    from stemgnn_b200 import runtime
    c = cases("forward")[name]
    g = golden(name)
    if "block0.forecast" not in g.files:
        pytest.skip("no stage taps in this golden")
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    mul_L = torch.from_numpy(g["mul_L"]).to(DEV)
    X = x.permute(0, 2, 1).contiguous().unsqueeze(1).to(DEV)
    blk = m.stock_block[0]
    blk.gemm_mode = runtime.GEMM_AUTO
    with torch.no_grad():
        forecast, backcast = blk(X, mul_L)
    assert_close(forecast, g["block0.forecast"], msg=name + " block0.forecast")
    assert_close(backcast, g["block0.backcast"], msg=name + " block0.backcast")
    e = np.abs(forecast.cpu().numpy() - g["block0.forecast"]).max()
    print(f"{name}: block0.forecast max|err| {e:.2e}")
