"""Fused Laplacian + Jacobi eigendecomposition kernel (opt-in `graph_mode="eig"`).
The reference computes no eigendecomposition, so the oracle is `torch.linalg.eigh` in float64 on the
reference's own Laplacian, and — eigenvectors of (near-)degenerate eigenvalues being non-unique — only INVARIANTS are
compared: eigenvalues, ||U^T U - I||, ||L U - U Lambda||, and U p(Lambda) U^T against the polynomial stack."""
import numpy as np
import pytest
import torch

from oracle import stemgnn_oracle as so, torch_port as tp
from tests.helpers import assert_close, build_model, case_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _attention(N, seed, sharp):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(N, N, generator=g) * sharp
    return torch.softmax(logits, dim=1)            # rows sum to 1, not symmetric — like base_model.py:140


def _laplacian64(a):
    a = a.double()
    deg = a.sum(1)
    asym = 0.5 * (a + a.t())
    dh = 1.0 / (deg.sqrt() + 1e-7)
    return dh[:, None] * (torch.diag(deg) - asym) * dh[None, :]


@pytest.mark.parametrize("N,sharp", [(2, 1.0), (6, 1.0), (53, 2.0), (140, 2.0), (228, 1.0), (357, 2.0), (358, 0.05), (512, 2.0)])
def test_laplacian_eig_invariants(N, sharp):
    from stemgnn_b200.eig import laplacian_eig
    a = _attention(N, N, sharp)
    lam, U, info = laplacian_eig(a.to(DEV))
    assert lam.shape == (N,) and U.shape == (N, N) and info["sweeps"] >= 1
    L = _laplacian64(a)
    ref = torch.linalg.eigvalsh(L)
    lam64, U64 = lam.double().cpu(), U.double().cpu()
    # fp32 Jacobi: ~10 sweeps x (n-1) rotations per column, each with ~1e-7 rounding -> a few 1e-5 at n = 512
    assert (lam64 - ref).abs().max().item() < 4e-5, (lam64 - ref).abs().max().item()
    assert (U64.t() @ U64 - torch.eye(N, dtype=torch.float64)).abs().max().item() < 5e-5
    assert (L @ U64 - U64 * lam64[None, :]).abs().max().item() < 5e-5
    # U p(Lambda) U^T reproduces the polynomial stack [0, L, 2L^2, 4L^3 - L] (base_model.py:121-134)
    stack = so.cheb_polynomial(L.float().numpy())
    for k, pk in ((1, lam64), (2, 2 * lam64 ** 2), (3, 4 * lam64 ** 3 - lam64)):
        rec = (U64 * pk[None, :]) @ U64.t()
        assert (rec - torch.from_numpy(stack[k]).double()).abs().max().item() < 2e-4, k


@pytest.mark.parametrize("N", [140, 358])
def test_model_eig_mode_matches_poly_mode(N):
    """Model.graph_mode = 'eig' (Jacobi path) against the default polynomial path and the reference port."""
    c = dict(B=8, N=N, W=12, H=3, multi=5, pseed=7, mode="trained")
    p = case_params(c)
    m = build_model(c, DEV, p).eval()
    x, _ = tp.synthetic_batch(8, N, 12, 3, seed=3)
    with torch.no_grad():
        f_ref, _ = tp.model_forward(x, p)
        f_poly, _ = m(x.to(DEV))
        mul_poly, _ = m.latent_correlation_layer(x.to(DEV))
        m.graph_mode = "eig"
        f_eig, _ = m(x.to(DEV))
        mul_eig, _ = m.latent_correlation_layer(x.to(DEV))
    assert_close(mul_eig, mul_poly, rtol=1e-3, atol=2e-4, msg="U p(Lambda) U^T vs polynomial stack")
    assert_close(f_eig, f_ref, msg="forecast in eig mode")
    assert not torch.equal(f_eig, f_poly)          # really a different path
