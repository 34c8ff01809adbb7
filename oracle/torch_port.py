"""TEST INFRASTRUCTURE — torch (CPU) restatement of the reference hot path.  NOT product code.

Why a second restatement next to `stemgnn_oracle.py` (numpy):
  * the reference IS PyTorch-on-CPU code; its arithmetic lives in ATen (aten::gru, addmm,
    bmm, _fft_r2c/_fft_c2r, softmax).  This port issues the same ATen ops in the same order
    with the modern `torch.fft` API, so timing it on the GPU box's host cores is the honest
    "reference CPU path" (`bench.py --impl reference`, `cpu_baseline.kind == "port"`);
    the unmodified reference cannot travel to the GPU box (/root/reference is absent there).
  * it is differentiable, which gives the oracle for the backward kernels (autograd through
    the reference-ordered ops), pinned by the gradient goldens in tests/golden/.

Each function cites the reference lines it follows (models/base_model.py @ dc7dea68).
Parameters: dict of tensors keyed exactly like the reference `state_dict()`.
"""
import torch
import torch.nn.functional as F


def _gru(x_seq, p):
    """base_model.py:92,137 — aten::gru, 1 layer, batch_first=False, h0=0."""
    hidden = p["GRU.weight_hh_l0"].shape[1]
    h0 = x_seq.new_zeros(1, x_seq.shape[1], hidden)
    flat = [p["GRU.weight_ih_l0"], p["GRU.weight_hh_l0"], p["GRU.bias_ih_l0"], p["GRU.bias_hh_l0"]]
    out, _ = torch._VF.gru(x_seq, h0, flat, True, 1, 0.0, False, False, False)
    return out


def self_graph_attention(inp_bsh, p, alpha, dropout_mask, dropout_p):
    """base_model.py:151-162 (same materialisation order: repeat/view/add, leakyrelu, softmax)."""
    inp = inp_bsh.permute(0, 2, 1).contiguous()
    bat, N, _ = inp.shape
    key = torch.matmul(inp, p["weight_key"])
    query = torch.matmul(inp, p["weight_query"])
    data = key.repeat(1, 1, N).view(bat, N * N, 1) + query.repeat(1, N, 1)
    data = data.squeeze(2).view(bat, N, -1)
    data = F.leaky_relu(data, alpha)
    att = F.softmax(data, dim=2)
    if dropout_mask is not None:
        att = att * dropout_mask / (1.0 - dropout_p)
    return att


def cheb_polynomial(lap):
    """base_model.py:121-134."""
    N = lap.size(0)
    lap = lap.unsqueeze(0)
    first = torch.zeros([1, N, N], device=lap.device, dtype=lap.dtype)
    second = lap
    third = (2 * torch.matmul(lap, second)) - first
    forth = 2 * torch.matmul(lap, third) - second
    return torch.cat([first, second, third, forth], dim=0)


def latent_correlation_layer(x, p, alpha=0.2, dropout_mask=None, dropout_p=0.5):
    """base_model.py:136-149 (including the dense diag matmuls at :144-147)."""
    inp = _gru(x.permute(2, 0, 1).contiguous(), p)
    inp = inp.permute(1, 0, 2).contiguous()
    attention = self_graph_attention(inp, p, alpha, dropout_mask, dropout_p)
    attention = torch.mean(attention, dim=0)
    degree = torch.sum(attention, dim=1)
    attention = 0.5 * (attention + attention.T)
    degree_l = torch.diag(degree)
    d_hat = torch.diag(1 / (torch.sqrt(degree) + 1e-7))
    lap = torch.matmul(d_hat, torch.matmul(degree_l - attention, d_hat))
    return cheb_polynomial(lap), attention


def _glu(x, p, prefix):
    """base_model.py:12-13."""
    return torch.mul(F.linear(x, p[prefix + ".linear_left.weight"], p[prefix + ".linear_left.bias"]),
                     torch.sigmoid(F.linear(x, p[prefix + ".linear_right.weight"],
                                            p[prefix + ".linear_right.bias"])))


def spe_seq_cell(inp, p, prefix):
    """base_model.py:46-59 with torch.fft standing in for the removed torch.rfft/irfft."""
    B, k, c, N, W = inp.shape
    inp = inp.view(B, -1, N, W)
    ffted = torch.view_as_real(torch.fft.fft(inp, dim=-1))
    real = ffted[..., 0].permute(0, 2, 1, 3).contiguous().reshape(B, N, -1)
    img = ffted[..., 1].permute(0, 2, 1, 3).contiguous().reshape(B, N, -1)
    for i in range(3):
        real = _glu(real, p, f"{prefix}.GLUs.{2 * i}")
        img = _glu(img, p, f"{prefix}.GLUs.{2 * i + 1}")
    real = real.reshape(B, N, 4, -1).permute(0, 2, 1, 3).contiguous()
    img = img.reshape(B, N, 4, -1).permute(0, 2, 1, 3).contiguous()
    spec = torch.cat([real.unsqueeze(-1), img.unsqueeze(-1)], dim=-1)
    return torch.fft.irfft(torch.view_as_complex(spec), n=spec.shape[-2], dim=-1)


def stock_block_forward(x, mul_L, p, prefix, stack_idx):
    """base_model.py:61-75.  x: (B,1,N,W)."""
    mul_L = mul_L.unsqueeze(1)
    x = x.unsqueeze(1)
    gfted = torch.matmul(mul_L, x)
    gconv_input = spe_seq_cell(gfted, p, prefix).unsqueeze(2)
    igfted = torch.sum(torch.matmul(gconv_input, p[prefix + ".weight"]), dim=1)
    fsrc = torch.sigmoid(F.linear(igfted, p[prefix + ".forecast.weight"],
                                  p[prefix + ".forecast.bias"]).squeeze(1))
    forecast = F.linear(fsrc, p[prefix + ".forecast_result.weight"],
                        p[prefix + ".forecast_result.bias"])
    if stack_idx == 0:
        short = F.linear(x, p[prefix + ".backcast_short_cut.weight"],
                         p[prefix + ".backcast_short_cut.bias"]).squeeze(1)
        back = torch.sigmoid(F.linear(igfted, p[prefix + ".backcast.weight"],
                                      p[prefix + ".backcast.bias"]) - short)
    else:
        back = None
    return forecast, back


def model_forward(x, p, stack_cnt=2, alpha=0.2, dropout_mask=None, dropout_p=0.5):
    """base_model.py:167-179.  x: (B,W,N) -> (forecast, attention)."""
    mul_L, attention = latent_correlation_layer(x, p, alpha, dropout_mask, dropout_p)
    X = x.unsqueeze(1).permute(0, 1, 3, 2).contiguous()
    result = []
    for i in range(stack_cnt):
        fc, X = stock_block_forward(X, mul_L, p, f"stock_block.{i}", i)
        result.append(fc)
    f = result[0] + result[1]
    f = F.linear(F.leaky_relu(F.linear(f, p["fc.0.weight"], p["fc.0.bias"]), 0.01),
                 p["fc.2.weight"], p["fc.2.bias"])
    if f.size(-1) == 1:
        return f.unsqueeze(1).squeeze(-1), attention
    return f.permute(0, 2, 1).contiguous(), attention


# --------------------------------------------------------------------------------------
# deterministic parameters / inputs (shared with bench.py; they live in the product tree because they are
# plain data generators, not part of the oracle's arithmetic)
# --------------------------------------------------------------------------------------
from stemgnn_b200.synthetic import param_shapes, synthetic_batch, synthetic_params  # noqa: E402,F401
