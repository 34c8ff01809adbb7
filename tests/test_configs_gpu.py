"""BASELINE.json configs on the B200 (VERDICT r1 "N3"): every named configuration runs through the CUDA path and
is compared with the reference — the UNMODIFIED reference itself when its staged copy travelled to the box
(git-ignored baseline/_ref/, oracle/fetch_reference.py), else the torch port of it.

  cfg1  ECG_data.csv N=140 W=12 H=3 batch=32: forward on real ECG rows vs the live reference, and one full epoch of
        `python main.py` (reference file, unchanged) on cuda:0 through run_main.py;
  cfg3  (64,228,12,3) bf16 tensor-core mode with its own, looser, stated tolerance;
  cfg4  global batch (128,325,12,12): forward + backward;
  cfg5  N=2048 (>= 8 windows): forward + backward.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shim, torch_port as tp
from oracle.cpu_reference import CpuReference
from tests.helpers import assert_close, build_model, case_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ecg_csv():
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        p = os.path.join(root, "dataset", "ECG_data.csv")
        if os.path.isfile(p):
            return p
    return None


def test_cfg1_ecg_forward_vs_live_reference():
    """configs[0]: z-scored ECG rows, reference default-seed weights, CUDA forward vs the unmodified reference."""
    csv = _ecg_csv()
    if csv is None or not ref_shim.reference_available():
        pytest.skip("staged reference (baseline/_ref) not present on this box")
    data = np.loadtxt(csv, delimiter=",")
    data = (data - data.mean(0)) / data.std(0)
    hi = [12 + 131 * i for i in range(32)]
    x = torch.from_numpy(np.stack([data[h - 12:h] for h in hi]).astype(np.float32))
    ref_model = ref_shim.build_reference_model(140, 12, 5, 3, seed=0).eval()      # main.py:52 seed
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    from models.base_model import Model
    m = Model(140, 2, 12, 5, horizon=3)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        f_ref, a_ref = ref_model(x)
        f, a = m(x.to(DEV))
    assert_close(f, f_ref, msg="cfg1 forecast vs live reference")
    assert_close(a, a_ref, msg="cfg1 attention vs live reference")
    mae = float((f.cpu() - f_ref).abs().mean())
    assert mae < 2e-5, mae


def test_cfg1_reference_main_py_one_epoch_on_cuda(tmp_path):
    """`python main.py --epoch 1 --device cuda:0` (the reference's own file) against the drop-in packages."""
    if _ecg_csv() is None or not ref_shim.reference_available():
        pytest.skip("staged reference (baseline/_ref) not present on this box")
    env = dict(os.environ, STEMGNN_WORKDIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_main.py"), "--epoch", "1", "--device", "cuda:0"],
                       capture_output=True, text=True, timeout=900, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "Performance on test set" in out, out[-3000:]
    # normalised validation MAE after one epoch: the reference reaches RAW MAE ~0.40 / 0.35; train-mode
    # dropout makes this statistical — require the same ballpark
    raw = [l for l in out.splitlines() if "RAW" in l and "MAE" in l]
    assert raw, out[-3000:]
    maes = [float(l.split("MAE")[1].split(";")[0].replace(":", "").split()[0]) for l in raw]
    assert all(0.05 < v < 0.9 for v in maes), maes
    assert os.path.exists(os.path.join(str(tmp_path), "output", "ECG_data", "train", "_stemgnn.pt"))


def _fwd_bwd_vs_cpu(B, N, H, seed, rtol_g, use_reference):
    c = dict(B=B, N=N, W=12, H=H, multi=5, pseed=seed, mode="init")
    p = case_params(c)
    x, y = tp.synthetic_batch(B, N, 12, H, seed=seed + 1)
    m = build_model(c, DEV, p)
    m.eval()                                           # eval: no dropout, gradients still flow
    f, a = m(x.to(DEV))
    loss = torch.nn.functional.mse_loss(f, y.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f_ref, a_ref = tp.model_forward(x, pr)
    torch.nn.functional.mse_loss(f_ref, y).backward()
    if use_reference and ref_shim.reference_available():      # the port itself is re-checked against the live reference
        ref = CpuReference(N, 12, H, 5, p)
        f_live, _ = ref.forward(x)
        assert_close(f_ref.detach(), f_live, rtol=1e-4, atol=1e-5, msg="port vs live reference")
    assert_close(f.detach(), f_ref.detach(), msg="forecast")
    assert_close(a.detach(), a_ref.detach(), msg="attention")
    named = dict(m.named_parameters())
    worst = 0.0
    for k, g_ref in ((k, v.grad) for k, v in pr.items()):
        g = named[k].grad
        if g_ref is None or float(g_ref.abs().max()) == 0.0:
            continue
        assert g is not None, k
        rel = float((g.cpu() - g_ref).abs().max() / g_ref.abs().max())
        worst = max(worst, rel)
        assert rel < rtol_g, f"{k}: gradient rel err {rel:.3e}"
    return worst


def test_cfg4_global_batch_forward_backward():
    """configs[3] at its global batch on one GPU: (128,325,12,12)."""
    _fwd_bwd_vs_cpu(128, 325, 12, 41, 1e-2, True)    # default mode: weight / input gradient GEMMs on TF32 tensor cores


def test_cfg5_n2048_forward_backward():
    """configs[4] node count: (8,2048,12,3) — the GRU runs beyond the single-cluster envelope."""
    _fwd_bwd_vs_cpu(8, 2048, 3, 43, 1e-2, False)


def test_cfg3_bf16_mode():
    """configs[2] "bf16 tensor-core GFT": (64,228,12,3) with Model.gemm_mode = GEMM_BF16.
    Stated tolerance for this mode: rtol 1e-2 / atol 2e-3 on the forecast (bf16 operands carry 8 mantissa bits;
    SURVEY §7.3 measured 7.7e-2 stage-level error for a bf16 GFT, which the heads attenuate), attention exact-path."""
    from stemgnn_b200 import runtime
    if not hasattr(runtime, "GEMM_BF16"):
        pytest.skip("bf16 mode not built")
    c = dict(B=64, N=228, W=12, H=3, multi=5, pseed=328, mode="trained")
    p = case_params(c)
    m = build_model(c, DEV, p).eval()
    m.gemm_mode = runtime.GEMM_BF16
    x, _ = tp.synthetic_batch(64, 228, 12, 3, seed=99)
    with torch.no_grad():
        f_ref, a_ref = tp.model_forward(x, p)
        f, a = m(x.to(DEV))
    assert_close(a, a_ref, msg="attention (GRU/attention stay in split-precision fp32)")
    assert_close(f, f_ref, rtol=1e-2, atol=2e-3, msg="bf16 forecast")
