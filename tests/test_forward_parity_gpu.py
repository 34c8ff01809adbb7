"""GPU parity of the CUDA forward path (through the C ABI) against the oracle and against the golden
vectors minted from the unmodified reference.  Tolerance: rtol 1e-3 / atol 1e-4 fp32 (north_star)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import stemgnn_oracle as so, torch_port as tp
from tests.helpers import RTOL, ATOL, assert_close, build_model, case_params, cases, golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

FWD_CASES = ["tiny_taps", "odd_h1_taps", "multi2_w8", "cfg1_shape", "cfg1_trained", "cfg2_shape"]


def _lib():
    from stemgnn_b200 import _lib as L
    return L, L.load()


# ---------------------------------------------------------------------------------------------
# fp32 GEMM building block
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (128, 64, 16), (130, 67, 19), (358, 358, 358),
                                   (1074, 384, 358), (257, 72, 480), (33, 240, 36)])
@pytest.mark.parametrize("a_km,b_nk", [(0, 1), (0, 0), (1, 1), (1, 0)])
def test_sgemm_all_layouts(M, N, K, a_km, b_nk):
    from stemgnn_b200 import runtime
    L, lib = _lib()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(K, N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = (2.0 * A.double() @ Bm.double() - 0.5 * C0.double()).float()
    Ad = (A.t().contiguous() if a_km else A.contiguous()).to(DEV)
    Bd = (Bm.t().contiguous() if b_nk else Bm.contiguous()).to(DEV)
    Cd = C0.clone().to(DEV)
    rc = lib.stemgnn_sgemm(M, N, K, 2.0, Ad.data_ptr(), Ad.shape[1], a_km, Bd.data_ptr(), Bd.shape[1],
                           b_nk, -0.5, Cd.data_ptr(), N, runtime._stream_ptr(torch.device(DEV)))
    L.check(rc, "sgemm")
    assert_close(Cd, ref, rtol=1e-4, atol=1e-4 * max(1.0, K ** 0.5 / 4), msg="sgemm")


def test_glu_gemm_fp32():
    from models.base_model import GLU
    torch.manual_seed(3)
    glu = GLU(36, 240).to(DEV)
    x = torch.randn(3, 77, 36, device=DEV)
    out = glu(x)
    xd = x.double().cpu()
    l = xd @ glu.linear_left.weight.double().cpu().t() + glu.linear_left.bias.double().cpu()
    r = xd @ glu.linear_right.weight.double().cpu().t() + glu.linear_right.bias.double().cpu()
    assert_close(out, (l * torch.sigmoid(r)).float(), rtol=1e-4, atol=1e-5, msg="glu")


# ---------------------------------------------------------------------------------------------
# GRU + key/query  (cluster path and generic path)
# ---------------------------------------------------------------------------------------------
def _gru_call(c, path):
    from stemgnn_b200 import runtime
    L, lib = _lib()
    p = case_params(c, DEV)
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    xd = x.to(DEV)
    dims = L.Dims(c["B"], c["N"], c["W"], c["H"], c["multi"])
    ptrs = runtime.build_ptrs({k: p.get(k) for k in runtime.PARAM_KEYS})
    ws = runtime.alloc_workspace(dims, False, xd.device)
    key = torch.empty(c["B"], c["N"], device=DEV)
    query = torch.empty(c["B"], c["N"], device=DEV)
    out = torch.empty(c["N"], c["B"], c["N"], device=DEV)
    rc = lib.stemgnn_gru_keyquery_forward(ctypes.byref(dims), ctypes.byref(ptrs), xd.data_ptr(),
                                          key.data_ptr(), query.data_ptr(), out.data_ptr(), path,
                                          ws.data_ptr(), ws.numel(), runtime._stream_ptr(xd.device))
    L.check(rc, "gru")
    torch.cuda.synchronize()
    return x, p, key, query, out


@pytest.mark.parametrize("path", [3, 2, 1], ids=["tensorcore", "cluster", "generic"])
@pytest.mark.parametrize("name", ["tiny_taps", "odd_h1_taps", "multi2_w8", "cfg1_trained", "cfg2_shape"])
def test_gru_keyquery_vs_oracle(name, path):
    c = cases("forward")[name]
    x, p, key, query, out = _gru_call(c, path)
    pc = {k: v.cpu() for k, v in p.items()}
    with torch.no_grad():   # (S,B,H) aten::gru on the host in float64 (independent of the host BLAS's fp32 code path)
        ref = tp._gru(x.permute(2, 0, 1).contiguous().double(), {k: v.double() for k, v in pc.items()}).float()
    ref_key = torch.einsum("sbh,s->bh", ref.double(), pc["weight_key"][:, 0].double()).float()
    ref_query = torch.einsum("sbh,s->bh", ref.double(), pc["weight_query"][:, 0].double()).float()
    assert_close(out, ref, rtol=1e-4, atol=2e-6, msg="gru_out")
    assert_close(key, ref_key, rtol=1e-4, atol=1e-5, msg="key")
    assert_close(query, ref_query, rtol=1e-4, atol=1e-5, msg="query")
    if "gru_out" in golden(name).files:
        assert_close(out, golden(name)["gru_out"], rtol=1e-4, atol=2e-6, msg="gru_out vs reference golden")


@pytest.mark.parametrize("path", [3, 2], ids=["tensorcore", "cluster"])
def test_gru_ragged_batch_and_padding(path):
    """B not a multiple of the sequences-per-cluster group; N not a multiple of the unit slice."""
    c = dict(B=7, N=53, W=12, H=3, multi=5, pseed=77, mode="trained")
    x, p, key, query, out = _gru_call(c, path)
    pc = {k: v.cpu() for k, v in p.items()}
    with torch.no_grad():
        ref = tp._gru(x.permute(2, 0, 1).contiguous().double(), {k: v.double() for k, v in pc.items()}).float()
    assert_close(out, ref, rtol=1e-4, atol=2e-6, msg="gru_out ragged")


@pytest.mark.parametrize("B,N", [(32, 358), (64, 228), (33, 325)])
def test_gru_tensorcore_repeatability(B, N):
    """The tcgen05 recurrence exchanges h through DSMEM with mbarrier-signalled st.async stores: 40 back-to-back calls
    (with cache-perturbing work in between) must all reproduce the host GRU — a lost or early signal shows up here."""
    c = dict(B=B, N=N, W=12, H=3, multi=5, pseed=N, mode="trained")
    from stemgnn_b200 import runtime
    L, lib = _lib()
    p = case_params(c, DEV)
    x, _ = tp.synthetic_batch(B, N, 12, 3, seed=7)
    xd = x.to(DEV)
    dims = L.Dims(B, N, 12, 3, 5)
    ptrs = runtime.build_ptrs({k: p.get(k) for k in runtime.PARAM_KEYS})
    ws = runtime.alloc_workspace(dims, False, xd.device)
    key = torch.empty(B, N, device=DEV); query = torch.empty(B, N, device=DEV); out = torch.empty(N, B, N, device=DEV)
    pc = {k: v.cpu().double() for k, v in p.items()}
    with torch.no_grad():
        ref = tp._gru(x.permute(2, 0, 1).contiguous().double(), pc).float().to(DEV)
    junk = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)
    worst = 0.0
    for i in range(40):
        if i % 3 == 1:
            junk.zero_()
        rc = lib.stemgnn_gru_keyquery_forward(ctypes.byref(dims), ctypes.byref(ptrs), xd.data_ptr(), key.data_ptr(),
                                              query.data_ptr(), out.data_ptr(), 3, ws.data_ptr(), ws.numel(),
                                              runtime._stream_ptr(xd.device))
        L.check(rc, "gru")
        worst = max(worst, float((out - ref).abs().max()))
    assert worst < 3e-6, worst


# ---------------------------------------------------------------------------------------------
# attention -> Laplacian -> Chebyshev stack
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N", [(4, 24), (3, 37), (32, 140), (32, 358)])
@pytest.mark.parametrize("drop", [False, True])
def test_graph_forward_vs_oracle(B, N, drop):
    from stemgnn_b200 import runtime
    L, lib = _lib()
    g = torch.Generator().manual_seed(B * 1000 + N)
    key = torch.randn(B, N, generator=g) * 0.7
    query = torch.randn(B, N, generator=g) * 0.7
    mask = (torch.rand(B, N, N, generator=g) >= 0.5) if drop else None
    # oracle (numpy fp32): rebuild the attention from key/query exactly as base_model.py:156-161
    data = key[:, :, None].numpy() + query[:, None, :].numpy()
    data = np.where(data >= 0, data, 0.2 * data)
    e = np.exp(data - data.max(axis=2, keepdims=True))
    att = e / e.sum(axis=2, keepdims=True)
    if drop:
        att = att * mask.numpy().astype(np.float32) / 0.5
    lap, attention, _ = so.laplacian_from_attention(att.astype(np.float32))
    mul_ref = so.cheb_polynomial(lap)
    dims = L.Dims(B, N, 12, 3, 5)
    ws = runtime.alloc_workspace(dims, False, torch.device(DEV))
    mask_d = mask.to(torch.uint8).to(DEV) if drop else None      # must outlive the call
    opts = runtime.make_opts(0.2, 0.5 if drop else 0.0, drop, mask=mask_d)
    kd, qd = key.to(DEV), query.to(DEV)
    a_out = torch.empty(N, N, device=DEV)
    m_out = torch.empty(4, N, N, device=DEV)
    rc = lib.stemgnn_graph_forward(ctypes.byref(dims), ctypes.byref(opts), kd.data_ptr(), qd.data_ptr(),
                                   a_out.data_ptr(), m_out.data_ptr(), ws.data_ptr(), ws.numel(),
                                   runtime._stream_ptr(torch.device(DEV)))
    L.check(rc, "graph")
    assert_close(a_out, attention, rtol=1e-4, atol=1e-7, msg="attention")
    assert_close(m_out, mul_ref, rtol=1e-3, atol=2e-6, msg="mul_L")
    assert float((a_out - a_out.t()).abs().max()) == 0.0          # exactly symmetric
    assert float(m_out[0].abs().max()) == 0.0                       # first Chebyshev term is zeros


def test_philox_dropout_statistics():
    """Train-mode mask from (seed, offset): keep-rate ~ 1-p, and E[attention] is preserved."""
    from stemgnn_b200 import runtime
    L, lib = _lib()
    B, N = 16, 96
    g = torch.Generator().manual_seed(5)
    kd, qd = torch.randn(B, N, generator=g).to(DEV), torch.randn(B, N, generator=g).to(DEV)
    dims = L.Dims(B, N, 12, 3, 5)
    ws = runtime.alloc_workspace(dims, False, torch.device(DEV))
    outs = []
    for seed, p in ((1, 0.0), (1, 0.5), (2, 0.5), (1, 0.5)):
        opts = runtime.make_opts(0.2, p, p > 0, seed=seed, offset=3)
        a_out = torch.empty(N, N, device=DEV)
        m_out = torch.empty(4, N, N, device=DEV)
        L.check(lib.stemgnn_graph_forward(ctypes.byref(dims), ctypes.byref(opts), kd.data_ptr(),
                                          qd.data_ptr(), a_out.data_ptr(), m_out.data_ptr(),
                                          ws.data_ptr(), ws.numel(),
                                          runtime._stream_ptr(torch.device(DEV))), "graph")
        outs.append(a_out.cpu())
    ev, d1, d2, d1b = outs
    assert torch.equal(d1, d1b)                        # same (seed, offset) -> same mask
    assert not torch.equal(d1, d2)                     # different seed -> different mask
    assert abs(float(d1.sum() / ev.sum()) - 1.0) < 0.02   # inverted dropout preserves the mean


# ---------------------------------------------------------------------------------------------
# spectral block (stage-level methods of the drop-in classes)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny_taps", "odd_h1_taps", "multi2_w8"])
def test_stock_block_and_spe_seq_cell_vs_reference_golden(name):
    c = cases("forward")[name]
    g = golden(name)
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    mul_L = torch.from_numpy(g["mul_L"]).to(DEV)
    X = x.permute(0, 2, 1).contiguous().unsqueeze(1).to(DEV)              # (B,1,N,W)
    for i in range(2):
        gfted = torch.matmul(mul_L.unsqueeze(1).cpu(), X.unsqueeze(1).cpu()).to(DEV)   # (B,4,1,N,W)
        iff = m.stock_block[i].spe_seq_cell(gfted)
        assert_close(iff, g[f"block{i}.iffted"], msg=f"block{i}.iffted")
        fc, X2 = m.stock_block[i](X, mul_L)
        assert_close(fc, g[f"block{i}.forecast"], msg=f"block{i}.forecast")
        if i == 0:
            assert_close(X2, g["block0.backcast"], msg="block0.backcast")
            X = X2
        else:
            assert X2 is None


def test_latent_correlation_layer_vs_reference_golden():
    for name in ["tiny_taps", "odd_h1_taps", "multi2_w8"]:
        c = cases("forward")[name]
        g = golden(name)
        m = build_model(c, DEV).eval()
        x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
        mul_L, attention = m.latent_correlation_layer(x.to(DEV))
        assert_close(mul_L, g["mul_L"], rtol=1e-3, atol=1e-5, msg=name + " mul_L")
        assert_close(attention, g["attention"], rtol=1e-3, atol=1e-6, msg=name + " attention")
        ch = m.cheb_polynomial(mul_L[1])
        assert_close(ch, g["mul_L"], rtol=1e-3, atol=1e-5, msg=name + " cheb_polynomial")


# ---------------------------------------------------------------------------------------------
# full forward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", FWD_CASES)
def test_model_forward_vs_reference_golden(name):
    c = cases("forward")[name]
    g = golden(name)
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    with torch.no_grad():
        forecast, attention = m(x.to(DEV))
    assert forecast.shape == g["forecast"].shape
    assert_close(forecast, g["forecast"], msg=name + " forecast")
    if "attention" in g.files:
        assert_close(attention, g["attention"], msg=name + " attention")
    else:
        assert_close(attention[::7], g["attention_rows7"], msg=name + " attention")
    mae = float(np.mean(np.abs(forecast.cpu().numpy() - g["forecast"])))   # utils.math_utils.MAE
    assert mae < 2e-5, f"MAE vs reference {mae}"


@pytest.mark.parametrize("B,N,H", [(64, 228, 3), (32, 325, 12), (1, 140, 3), (33, 512, 3)])
def test_model_forward_vs_oracle_other_configs(B, N, H):
    """cfg3 / cfg4-shard shapes, batch 1, and the largest N of the cluster GRU path — checked against
    the torch port of the reference run on the host CPU."""
    c = dict(B=B, N=N, W=12, H=H, multi=5, pseed=100 + N, mode="trained")
    p = case_params(c)
    m = build_model(c, DEV, p).eval()
    x, _ = tp.synthetic_batch(B, N, 12, H, seed=99)
    with torch.no_grad():
        f_ref, a_ref = tp.model_forward(x, p)
        forecast, attention = m(x.to(DEV))
    assert_close(forecast, f_ref, msg="forecast")
    assert_close(attention, a_ref, msg="attention")


def test_model_forward_large_n_generic_gru_path():
    """N above the cluster-GRU limit (512) takes the generic per-step path."""
    c = dict(B=4, N=600, W=12, H=3, multi=5, pseed=5, mode="init")
    p = case_params(c)
    m = build_model(c, DEV, p).eval()
    x, _ = tp.synthetic_batch(4, 600, 12, 3, seed=98)
    with torch.no_grad():
        f_ref, a_ref = tp.model_forward(x, p)
        forecast, attention = m(x.to(DEV))
    assert_close(forecast, f_ref, msg="forecast")
    assert_close(attention, a_ref, msg="attention")


def test_forward_properties_full_size():
    """Size-independent invariants at the north-star shape (B,N,W)=(32,358,12)."""
    c = cases("forward")["cfg2_shape"]
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(32, 358, 12, 3, seed=7)
    xd = x.to(DEV)
    with torch.no_grad():
        f1, a1 = m(xd)
        f2, a2 = m(xd)
        mul_L, att = m.latent_correlation_layer(xd)
    assert torch.equal(f1, f2) and torch.equal(a1, a2)                 # deterministic
    assert torch.isfinite(f1).all()
    assert float((a1 - a1.t()).abs().max()) == 0.0                      # symmetric
    assert abs(float(a1.sum()) - 358.0) < 1e-2                          # softmax rows sum to 1
    L1 = mul_L[1].double()
    assert float((mul_L[2].double() - 2 * L1 @ L1).abs().max()) < 1e-5  # Chebyshev recurrences
    assert float((mul_L[3].double() - (2 * L1 @ mul_L[2].double() - L1)).abs().max()) < 1e-5
    # permuting the batch permutes the forecast and leaves the (batch-mean) graph unchanged
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        f3, a3 = m(xd[perm.to(DEV)].contiguous())
    assert_close(f3, f1[perm.to(DEV)], rtol=1e-4, atol=1e-5, msg="batch permutation")
    assert_close(a3, a1, rtol=1e-4, atol=1e-7, msg="graph invariance")


def test_cpu_tensor_is_rejected_loudly():
    c = cases("forward")["tiny_taps"]
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)


def test_folded_weight_cache_is_invalidated_by_parameter_updates():
    """Eval forwards reuse the DFT-folded weights in the workspace until a parameter changes."""
    c = cases("forward")["tiny_taps"]
    m = build_model(c, DEV).eval()
    x, _ = tp.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    xd = x.to(DEV)
    with torch.no_grad():
        f1, _ = m(xd)
        f2, _ = m(xd)                                   # cached folds
        assert torch.equal(f1, f2)
        m.stock_block[0].GLUs[0].linear_left.weight.mul_(1.5)      # touches a folded weight
        m.stock_block[1].weight.add_(0.01)                          # touches the output fold
        f3, _ = m(xd)
    p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        f_ref, _ = tp.model_forward(x, p)
    assert not torch.equal(f1, f3)
    assert_close(f3, f_ref, msg="forward after in-place weight update")


@pytest.mark.parametrize("B,N,W,H,multi", [(6, 25, 28, 28, 5), (1, 5, 12, 1, 5), (2, 17, 4, 2, 3), (5, 130, 12, 3, 1)])
def test_unusual_shapes_forward_and_backward_vs_port(B, N, W, H, multi):
    """COVID-19 config of the reference README (W=28, H=28: d=560 exceeds the tensor-core tile, fp32 path),
    single-window batches, tiny graphs, multi_layer=1 — forward and every gradient against the port."""
    c = dict(B=B, N=N, W=W, H=H, multi=multi, pseed=900 + N, mode="trained")
    p = case_params(c)
    m = build_model(c, DEV, p).eval()
    x, y = tp.synthetic_batch(B, N, W, H, seed=5)
    xd = x.to(DEV)
    f, a = m(xd)
    torch.nn.functional.mse_loss(f, y.to(DEV) if H > 1 else y.to(DEV)).backward()
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f_ref, a_ref = tp.model_forward(x, pr)
    torch.nn.functional.mse_loss(f_ref, y).backward()
    assert_close(f, f_ref, msg="forecast")
    assert_close(a, a_ref, msg="attention")
    named = dict(m.named_parameters())
    for k, v in pr.items():
        if v.grad is None:
            continue
        g, r = named[k].grad.cpu(), v.grad
        scale = max(float(r.abs().max()), 1e-12)
        assert float((g - r).abs().max()) <= 1e-2 * scale + 1e-9, f"{k}: grad mismatch"
