"""Seeded synthetic parameters and inputs for StemGNN-shaped problems (no reference needed).

Used by bench.py, the tests and the oracle: weights follow the reference's init distributions
(base_model.py:23-31, :88-101; nn.Linear / nn.GRU defaults = U(-1/sqrt(fan), 1/sqrt(fan))), inputs follow
SURVEY.md §8(d).  Pure data generation — no model arithmetic lives here.
"""
import torch


def param_shapes(N, W, H, multi=5, stack_cnt=2):
    """Reference state_dict keys -> shapes (checked vs the reference in
    tests/test_oracle_golden.py)."""
    T, d = multi * W, 4 * multi * W
    s = {"weight_key": (N, 1), "weight_query": (N, 1),
         "GRU.weight_ih_l0": (3 * N, W), "GRU.weight_hh_l0": (3 * N, N),
         "GRU.bias_ih_l0": (3 * N,), "GRU.bias_hh_l0": (3 * N,)}
    for i in range(stack_cnt):
        q = f"stock_block.{i}"
        s[q + ".weight"] = (1, 4, 1, T, T)
        s[q + ".forecast.weight"], s[q + ".forecast.bias"] = (T, T), (T,)
        s[q + ".forecast_result.weight"], s[q + ".forecast_result.bias"] = (W, T), (W,)
        if i == 0:
            s[q + ".backcast.weight"], s[q + ".backcast.bias"] = (W, T), (W,)
        s[q + ".backcast_short_cut.weight"], s[q + ".backcast_short_cut.bias"] = (W, W), (W,)
        for g in range(6):
            fan_in = 4 * W if g < 2 else d
            for side in ("left", "right"):
                s[f"{q}.GLUs.{g}.linear_{side}.weight"] = (d, fan_in)
                s[f"{q}.GLUs.{g}.linear_{side}.bias"] = (d,)
    s["fc.0.weight"], s["fc.0.bias"] = (W, W), (W,)
    s["fc.2.weight"], s["fc.2.bias"] = (H, W), (H,)
    return s


def synthetic_params(N, W, H, multi=5, seed=0, dtype=torch.float32, scale_mode="init"):
    """Seeded stand-in weights with the reference's init *distributions* (base_model.py:23-31,
    :88-101; nn.Linear / nn.GRU defaults = U(-1/sqrt(fan), 1/sqrt(fan))).  Used where the real
    reference cannot be instantiated (GPU box).  `scale_mode="trained"` widens key/query and
    GRU weights so the attention is far from uniform (stress case for parity tests)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for k, shp in param_shapes(N, W, H, multi).items():
        if k in ("weight_key", "weight_query"):
            bound = 1.414 * (6.0 / (N + 1)) ** 0.5
            if scale_mode == "trained":
                bound *= 4.0
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif k.startswith("GRU."):
            bound = 1.0 / N ** 0.5
            if scale_mode == "trained":
                bound *= 3.0
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif k.endswith(".weight") and len(shp) == 5:
            T = shp[-1]
            # xavier_normal_ on (1,4,1,T,T): fan_in = 4*T*T, fan_out = T*T (base_model.py:26)
            t = torch.randn(shp, generator=g, dtype=torch.float64) * (2.0 / (4 * T * T + T * T)) ** 0.5
        else:
            fan_in = shp[-1] if len(shp) == 2 else None
            if fan_in is None:       # bias: bound by the matching weight's fan_in
                wkey = k[:-4] + "weight"
                fan_in = param_shapes(N, W, H, multi)[wkey][-1]
            bound = 1.0 / fan_in ** 0.5
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound
        p[k] = t.to(dtype)
    return p


def synthetic_batch(B, N, W, H, seed=1234, dtype=torch.float32):
    """g=Generator().manual_seed(1234); x=randn(B,W,N); y=randn(B,H,N)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, W, N, generator=g, dtype=torch.float32).to(dtype)
    y = torch.randn(B, H, N, generator=g, dtype=torch.float32).to(dtype)
    return x, y
