"""TEST INFRASTRUCTURE — mints tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference `models.base_model.Model` is imported through oracle/ref_shim.py, loaded with the
seeded synthetic weights of oracle/torch_port.synthetic_params (so the weights can be regenerated
on the GPU box without the reference) and run on the seeded inputs of synthetic_batch.  Every
array written here is an output of reference code; nothing of the reference's source is stored.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, torch_port  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (B, N, W, H, multi, param_seed, scale_mode, what)
CASES = {
    "tiny_taps":      dict(B=4, N=24, W=12, H=3, multi=5, pseed=11, mode="init", taps=True),
    "odd_h1_taps":    dict(B=3, N=37, W=12, H=1, multi=5, pseed=12, mode="trained", taps=True),
    "multi2_w8":      dict(B=5, N=19, W=8, H=2, multi=2, pseed=13, mode="trained", taps=True),
    "cfg1_shape":     dict(B=32, N=140, W=12, H=3, multi=5, pseed=14, mode="init", taps=False),
    "cfg1_trained":   dict(B=29, N=140, W=12, H=3, multi=5, pseed=15, mode="trained", taps=False),
    "cfg2_shape":     dict(B=32, N=358, W=12, H=3, multi=5, pseed=16, mode="init", taps=False),
}
GRAD_CASES = {
    "grad_multi2":    dict(B=4, N=21, W=12, H=3, multi=2, pseed=21, mode="trained", p_drop=None),
    "grad_tiny":      dict(B=4, N=24, W=12, H=3, multi=5, pseed=22, mode="trained", p_drop=None),
    "grad_dropmask":  dict(B=4, N=24, W=12, H=3, multi=2, pseed=23, mode="trained", p_drop=0.5),
}


def _ref_model(c):
    cls = ref_shim.load_reference_model_class()
    m = cls(c["N"], 2, c["W"], c["multi"], horizon=c["H"])
    p = torch_port.synthetic_params(c["N"], c["W"], c["H"], c["multi"], seed=c["pseed"],
                                    scale_mode=c["mode"])
    missing = m.load_state_dict(p, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m, p


class _MaskDropout(torch.nn.Module):
    """stands in for model.dropout (base_model.py:103,161) with an explicit keep-mask."""
    def __init__(self, mask, p):
        super().__init__()
        self.mask, self.p = mask, p

    def forward(self, x):
        return x * self.mask / (1.0 - self.p)


def forward_case(name, c):
    m, _ = _ref_model(c)
    m.eval()
    x, _ = torch_port.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=1234)
    out = {}
    with torch.no_grad():
        forecast, attention = m(x)
        out["forecast"] = forecast.numpy()
        if c["N"] <= 140:
            out["attention"] = attention.numpy()
        else:
            out["attention_rows7"] = attention[::7].numpy()
        if c["taps"]:
            gru_out, _ = m.GRU(x.permute(2, 0, 1).contiguous())
            out["gru_out"] = gru_out.numpy()
            mul_L, _ = m.latent_correlation_layer(x)
            out["mul_L"] = mul_L.numpy()
            X = x.unsqueeze(1).permute(0, 1, 3, 2).contiguous()
            for i in range(2):
                blk = m.stock_block[i]
                gfted = torch.matmul(mul_L.unsqueeze(1), X.unsqueeze(1))
                out[f"block{i}.iffted"] = blk.spe_seq_cell(gfted).numpy()
                fc, X = blk(X, mul_L)
                out[f"block{i}.forecast"] = fc.numpy()
                if X is not None:
                    out[f"block{i}.backcast"] = X.numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def grad_case(name, c):
    m, p = _ref_model(c)
    x, y = torch_port.synthetic_batch(c["B"], c["N"], c["W"], c["H"], seed=4321)
    out = {}
    if c["p_drop"] is None:
        m.eval()
    else:
        g = torch.Generator().manual_seed(99)
        mask = (torch.rand(c["B"], c["N"], c["N"], generator=g) >= c["p_drop"]).float()
        m.dropout = _MaskDropout(mask, c["p_drop"])
        m.train()
    x.requires_grad_(True)
    forecast, attention = m(x)
    loss = torch.nn.functional.mse_loss(forecast, y)        # handler.py:140,162
    loss.backward()
    out["forecast"] = forecast.detach().numpy()
    out["attention"] = attention.detach().numpy()
    out["loss"] = np.float32(loss.item())
    out["grad.x"] = x.grad.numpy()
    for k, v in m.named_parameters():
        if v.grad is None:
            out["nograd." + k] = np.zeros(1, np.float32)
            continue
        gnp = v.grad.numpy()
        if gnp.size <= 20000:
            out["grad." + k] = gnp
        else:                                   # large GLU weights: strided sample + moments
            flat = gnp.reshape(-1)
            out["gradsample." + k] = flat[::53].copy()
            out["gradsum." + k] = np.array([flat.astype(np.float64).sum(),
                                            np.abs(flat).astype(np.float64).sum()])
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "loss", out["loss"], len(out), "arrays")


if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for n, c in CASES.items():
        forward_case(n, c)
    for n, c in GRAD_CASES.items():
        grad_case(n, c)
    import json
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump({"forward": CASES, "grad": GRAD_CASES,
                   "generator": "oracle/make_golden.py", "torch": torch.__version__,
                   "reference": "microsoft/StemGNN @ dc7dea68 via oracle/ref_shim.py"}, f, indent=1)
