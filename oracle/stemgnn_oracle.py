"""TEST INFRASTRUCTURE — CPU oracle for the StemGNN hot path.  NOT product code.

A plain-numpy restatement of the reference forward
(`/root/reference/models/base_model.py`, microsoft/StemGNN @ dc7dea68).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
anything under `oracle/`; the product (`stemgnn_b200/`, `models/`) never does.

Parity pin: the reference ships no tests or golden vectors for this path, so
the pin is the reference ITSELF, imported unmodified in the build container through
`oracle/ref_shim.py`; `oracle/make_golden.py` wrote `tests/golden/*.npz` from it and
`tests/test_oracle_golden.py` checks this restatement against every tensor in them.
Residual risk (stated in DESIGN.md): torch 1.7.1's `irfft(onesided=False)` is emulated by
`torch.fft.irfft` on the modern stack (ref_shim.py, patch 2).

All functions take/return numpy arrays.  `dtype=np.float32` follows the reference's fp32
arithmetic; pass `np.float64` for a high-precision truth when judging which of two fp32
results is closer.

Parameters are passed as a dict with exactly the reference's `state_dict()` keys
(`weight_key`, `GRU.weight_hh_l0`, `stock_block.0.GLUs.3.linear_left.weight`, ...).
"""
import numpy as np

K_ORDER = 4  # base_model.py:23 "3 + 1" Chebyshev terms


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _leaky_relu(x, slope):
    return np.where(x >= 0, x, x * slope)


def _linear(x, w, b):
    """nn.Linear: y = x @ w.T + b."""
    return x @ w.T + b


# --------------------------------------------------------------------------------------
# latent correlation layer
# --------------------------------------------------------------------------------------
def gru_forward(x_seq, w_ih, w_hh, b_ih, b_hh):
    """nn.GRU(time_step, unit) as called at base_model.py:92,137 — batch_first=False,
    one layer, h0 = 0, gate order [r, z, n].

    x_seq: (S, B, W) with S = node index (sequence axis), returns (S, B, Hd), Hd = units.
      r = sig(W_ir x + b_ir + W_hr h + b_hr);  z likewise
      n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1 - z) * n + z * h
    """
    S, B, _ = x_seq.shape
    Hd = w_hh.shape[1]
    dt = x_seq.dtype
    h = np.zeros((B, Hd), dtype=dt)
    out = np.empty((S, B, Hd), dtype=dt)
    for s in range(S):
        gi = x_seq[s] @ w_ih.T + b_ih
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, :Hd] + gh[:, :Hd])
        z = _sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        n = np.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
        h = ((1.0 - z) * n + z * h).astype(dt)
        out[s] = h
    return out


def self_graph_attention(gru_out_bsh, weight_key, weight_query, alpha=0.2, dropout_mask=None,
                         dropout_p=0.5):
    """base_model.py:151-162.  `gru_out_bsh`: (B, S, Hd) (already permuted as at :138).
    After the permute at :152 the graph node i is GRU hidden unit i and key/query contract the
    SEQUENCE axis:  key[b,i] = sum_s h_s[b,i] wk[s].
    dropout_mask: None (eval) or a {0,1} array (B,N,N); kept entries are scaled by 1/(1-p)."""
    inp = np.transpose(gru_out_bsh, (0, 2, 1))            # (B, Hd, S)
    key = inp @ weight_key                                 # (B, N, 1)
    query = inp @ weight_query                             # (B, N, 1)
    data = key + np.transpose(query, (0, 2, 1))            # data[b,i,j] = key[b,i] + query[b,j]
    data = _leaky_relu(data, alpha)
    data = data - data.max(axis=2, keepdims=True)
    e = np.exp(data)
    att = e / e.sum(axis=2, keepdims=True)
    if dropout_mask is not None:
        att = att * dropout_mask.astype(att.dtype) / (1.0 - dropout_p)
    return att.astype(gru_out_bsh.dtype), key[..., 0], query[..., 0]


def laplacian_from_attention(att_bnn):
    """base_model.py:140-147.  Returns (laplacian, attention_sym, degree)."""
    attention = att_bnn.mean(axis=0)
    degree = attention.sum(axis=1)                          # BEFORE symmetrisation (:141)
    attention = 0.5 * (attention + attention.T)
    d_hat = 1.0 / (np.sqrt(degree) + 1e-7)
    lap = d_hat[:, None] * (np.diag(degree) - attention) * d_hat[None, :]
    return lap.astype(att_bnn.dtype), attention.astype(att_bnn.dtype), degree


def cheb_polynomial(lap):
    """base_model.py:121-134: [0, L, 2 L L, 2 L (2 L L) - L] — the first term is ZEROS."""
    first = np.zeros_like(lap)
    second = lap
    third = 2.0 * (lap @ second) - first
    forth = 2.0 * (lap @ third) - second
    return np.stack([first, second, third, forth], axis=0).astype(lap.dtype)


def latent_correlation_layer(x_bwn, p, alpha=0.2, dropout_mask=None, dropout_p=0.5, taps=None):
    """base_model.py:136-149.  x: (B, W, N) -> (mul_L (4,N,N), attention (N,N))."""
    x_seq = np.ascontiguousarray(np.transpose(x_bwn, (2, 0, 1)))      # (N, B, W)
    out = gru_forward(x_seq, p["GRU.weight_ih_l0"], p["GRU.weight_hh_l0"],
                      p["GRU.bias_ih_l0"], p["GRU.bias_hh_l0"])
    out_bsh = np.transpose(out, (1, 0, 2))                              # (B, S, Hd)
    att, key, query = self_graph_attention(out_bsh, p["weight_key"], p["weight_query"], alpha,
                                           dropout_mask, dropout_p)
    lap, attention, degree = laplacian_from_attention(att)
    mul_L = cheb_polynomial(lap)
    if taps is not None:
        taps.update(gru_out=out, key=key, query=query, att_b=att, degree=degree, laplacian=lap)
    return mul_L, attention


# --------------------------------------------------------------------------------------
# spectral block
# --------------------------------------------------------------------------------------
def glu(x, p, prefix):
    """base_model.py:6-13."""
    left = _linear(x, p[prefix + ".linear_left.weight"], p[prefix + ".linear_left.bias"])
    right = _linear(x, p[prefix + ".linear_right.weight"], p[prefix + ".linear_right.bias"])
    return left * _sigmoid(right)


def spe_seq_cell(gfted, p, prefix, taps=None):
    """base_model.py:46-59.  gfted: (B, 4, N, W) -> (B, 4, N, T) with T = multi*W.

    Two-sided DFT over the W time samples; real/imag parts are laid out (B, N, 4*W) with column
    k*W+f; three GLU layers each (even GLUs for real, odd for imag); the (B,N,4*T) result is
    regrouped (B,4,N,T) and inverse-real-transformed using bins 0..T/2 only (Im of bins 0 and
    T/2 ignored) with norm 1/T."""
    B, K, N, W = gfted.shape
    dt = gfted.dtype
    ffted = np.fft.fft(gfted.astype(np.float64), axis=-1)
    real = np.transpose(ffted.real, (0, 2, 1, 3)).reshape(B, N, K * W).astype(dt)
    img = np.transpose(ffted.imag, (0, 2, 1, 3)).reshape(B, N, K * W).astype(dt)
    if taps is not None:
        taps["fft_real"], taps["fft_imag"] = real, img
    for i in range(3):
        real = glu(real, p, f"{prefix}.GLUs.{2 * i}")
        img = glu(img, p, f"{prefix}.GLUs.{2 * i + 1}")
    if taps is not None:
        taps["glu_real"], taps["glu_imag"] = real.astype(dt), img.astype(dt)
    T = real.shape[-1] // K
    real = np.transpose(real.reshape(B, N, K, T), (0, 2, 1, 3))
    img = np.transpose(img.reshape(B, N, K, T), (0, 2, 1, 3))
    spec = (real.astype(np.float64) + 1j * img.astype(np.float64))[..., : T // 2 + 1]
    return np.fft.irfft(spec, n=T, axis=-1).astype(dt)


def stock_block_forward(x_bnw, mul_L, p, prefix, stack_idx, taps=None):
    """base_model.py:61-75.  x: (B, N, W), mul_L: (4, N, N) ->
    (forecast (B,N,W), backcast (B,N,W) or None)."""
    gfted = np.einsum("knm,bmt->bknt", mul_L, x_bnw).astype(x_bnw.dtype)
    iffted = spe_seq_cell(gfted, p, prefix, taps)                       # (B,4,N,T)
    weight = p[prefix + ".weight"][0, :, 0]                              # (4, T, T)
    igfted = np.einsum("bknt,ktu->bnu", iffted, weight).astype(x_bnw.dtype)
    fsrc = _sigmoid(_linear(igfted, p[prefix + ".forecast.weight"], p[prefix + ".forecast.bias"]))
    forecast = _linear(fsrc, p[prefix + ".forecast_result.weight"],
                       p[prefix + ".forecast_result.bias"])
    if stack_idx == 0:
        short = _linear(x_bnw, p[prefix + ".backcast_short_cut.weight"],
                        p[prefix + ".backcast_short_cut.bias"])
        back = _sigmoid(_linear(igfted, p[prefix + ".backcast.weight"],
                                p[prefix + ".backcast.bias"]) - short)
    else:
        back = None
    if taps is not None:
        taps.update(gfted=gfted, iffted=iffted, igfted=igfted)
    return forecast.astype(x_bnw.dtype), (None if back is None else back.astype(x_bnw.dtype))


# --------------------------------------------------------------------------------------
# full model
# --------------------------------------------------------------------------------------
def model_forward(x_bwn, p, stack_cnt=2, alpha=0.2, dropout_mask=None, dropout_p=0.5, taps=None):
    """base_model.py:167-179.  x: (B, W, N) -> (forecast (B,H,N) | (B,1,N) if H==1,
    attention (N,N))."""
    dt = x_bwn.dtype
    p = {k: np.asarray(v, dtype=dt) for k, v in p.items()}
    lat_taps = {} if taps is not None else None
    mul_L, attention = latent_correlation_layer(x_bwn, p, alpha, dropout_mask, dropout_p, lat_taps)
    X = np.ascontiguousarray(np.transpose(x_bwn, (0, 2, 1)))           # (B, N, W)
    results = []
    block_taps = []
    for i in range(stack_cnt):
        bt = {} if taps is not None else None
        fc, X = stock_block_forward(X, mul_L, p, f"stock_block.{i}", i, bt)
        results.append(fc)
        block_taps.append(bt)
    f = results[0] + results[1]                                          # :174 hard-codes 2
    f = _linear(f, p["fc.0.weight"], p["fc.0.bias"])
    f = _leaky_relu(f, 0.01)                                             # nn.LeakyReLU() default
    f = _linear(f, p["fc.2.weight"], p["fc.2.bias"])                     # (B, N, H)
    out = np.ascontiguousarray(np.transpose(f, (0, 2, 1))).astype(dt)    # (B, H, N)
    if taps is not None:
        taps.update(lat_taps)
        taps["mul_L"] = mul_L
        for i, bt in enumerate(block_taps):
            for k, v in bt.items():
                taps[f"block{i}.{k}"] = v
            taps[f"block{i}.forecast"] = results[i]
    return out, attention


def parameter_count(N, W, H, multi=5, stack_cnt=2):
    """SURVEY.md §8(d) P(N); equals sum(p.numel()) of the reference Model."""
    T = multi * W
    d = K_ORDER * T
    per_block = (K_ORDER * T * T + T * T + T + T * W + W + W * W + W
                 + 4 * (4 * W * d + d) + 8 * (d * d + d))
    return (2 * N + 3 * N * N + 3 * N * W + 6 * N + stack_cnt * per_block + (T * W + W)
            + W * W + W + W * H + H)


def forward_flops(B, N, W, H, multi=5, stack_cnt=2):
    """Algorithmic forward flops with dead work removed — SURVEY.md §8(d) formula, the
    numerator of roofline.achieved."""
    T = multi * W
    d = K_ORDER * T
    K = K_ORDER
    gru = 6 * B * N * N * W + 6 * B * N ** 3 + 12 * B * N * N
    att = 10 * B * N * N
    lap = 4 * N ** 3 + 6 * N * N
    blk = (2 * (K - 1) * B * N * N * W + 2 * B * N * d * 2 * (36 + 30) + 2 * B * N * d * d * 4
           + 2 * B * N * d * 2 * (124 + 116) + 2 * K * B * N * T * T + 2 * B * N * (T * T + T * W))
    return (gru + att + lap + stack_cnt * blk + 2 * B * N * (T * W + W * W)
            + 2 * B * N * (W * W + W * H))
