"""Opt-in eigendecomposition of the latent graph Laplacian (north_star; reference hooks `get_laplacian` /
`graph_fft`, models/base_model.py:106-119,164-165 — dead code in the reference, so this is not on the default forward).
Thin host wrapper over `stemgnn_laplacian_eig_forward` (fused Laplacian build + cluster-resident Jacobi sweeps)."""
import ctypes

import torch

from . import _lib


def laplacian_eig(attention_raw, degree=None, max_sweeps=30, tol=1e-6):
    """attention_raw (N,N) CUDA float32: batch-mean softmax attention BEFORE symmetrisation (base_model.py:140);
    degree (N): its row sums (computed here when omitted).  Returns (eigenvalues (N,) ascending, eigenvectors (N,N)
    with column j <-> eigenvalue j, info dict) of L = D^(diag(deg) - (A + A^T)/2)D^ (base_model.py:141-147)."""
    lib = _lib.load()
    a = attention_raw.contiguous().float()
    if not a.is_cuda:
        raise RuntimeError("laplacian_eig needs CUDA tensors (no CPU fallback)")
    N = a.shape[0]
    deg = (a.sum(dim=1) if degree is None else degree).contiguous().float()
    n = (N + 1) & ~1
    lam = torch.empty(n, dtype=torch.float32, device=a.device)
    U = torch.empty(n, n, dtype=torch.float32, device=a.device)
    info = torch.zeros(4, dtype=torch.int32, device=a.device)
    rc = lib.stemgnn_laplacian_eig_forward(a.data_ptr(), deg.data_ptr(), N, lam.data_ptr(), U.data_ptr(), info.data_ptr(),
                                           int(max_sweeps), float(tol),
                                           ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    _lib.check(rc, "stemgnn_laplacian_eig_forward")
    if n != N:                      # drop the padding pair: the unit vector e_{n-1}
        keep = U[n - 1].abs() < 0.5
        lam, U = lam[keep], U[:N][:, keep]
    order = torch.argsort(lam)
    inf = info.cpu()
    return lam[order], U[:, order].contiguous(), {"sweeps": int(inf[0]), "converged": bool(inf[1])}
