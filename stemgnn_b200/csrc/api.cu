// C ABI of stemgnn_b200 (see include/stemgnn_b200.h): workspace carving + forward orchestration.
// Every kernel is launched on the caller's stream; nothing here allocates device memory.
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "internal.cuh"

namespace sg {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void clear_error() { g_err[0] = 0; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static ProfileHook g_hook = {nullptr, nullptr};
ProfileHook* profile_hook() { return &g_hook; }

// ---- workspace layout (offsets in floats, 256-byte aligned) -------------------------------------
Workspace carve_workspace(const stemgnn_dims_t& dm, int training, float* base) {
  Workspace ws = {};
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += (n + 63) / 64 * 64;
    return p;
  };
  const size_t B = dm.B, N = dm.N, W = dm.W, T = (size_t)dm.multi * dm.W, d = 4 * T, R = B * N;
  ws.xs = take(N * B * W);
  ws.x_bnw = take(R * W);
  ws.key = take(R);
  ws.query = take(R);
  ws.qmax = take(B);
  ws.a_raw = take(N * N);
  ws.deg = take(N);
  ws.mul_L = take(4 * N * N);
  ws.attention = take(N * N);
  ws.gru_scratch = take(2 * R);
  // input projection of the FFMA2 fallback (N, B, 3N) — and, on the tensor-core path, the packed fp16 hi/lo W_hh images,
  // which are LARGER than the projection for tiny batches (B*N of a few rows): size for both
  const size_t gi_img = (gru_tc_image_bytes(dm.N, dm.W) + sizeof(float) - 1) / sizeof(float);
  ws.gi = take(N * R * 3 > gi_img ? N * R * 3 : gi_img);
  ws.skbuf = take(8 * (N * N > 3 * N * B * W ? N * N : 3 * N * B * W));
  const size_t Np = (size_t)pad4(dm.N);
  ws.mul_Lp = take(3 * N * Np);
  ws.x_pad = take(B * W * Np);
  ws.eig_lambda = take(N + 2);
  ws.eig_U = take((N + 2) * (N + 2));
  ws.eig_S = take((N + 2) * (N + 2));
  ws.eig_info = reinterpret_cast<int*>(take(4));
  ws.row_m = training ? take(R) : nullptr;
  ws.row_zinv = training ? take(R) : nullptr;
  ws.h_all = training ? take(N * R) : nullptr;
  ws.g_r = training ? take(N * R) : nullptr;
  ws.g_z = training ? take(N * R) : nullptr;
  ws.g_n = training ? take(N * R) : nullptr;
  ws.g_hn = training ? take(N * R) : nullptr;
  for (int i = 0; i < STEMGNN_MAX_STACK; ++i) {
    BlockWs& b = ws.blk[i];
    b.G = take(R * 4 * W);              // 3W columns on the model path, 4W for the stage API
    b.w1f = take(4 * d * 4 * W);
    b.ic = take(2 * T * T);
    b.ri = take(8 * T * T);
    b.wout = take(8 * T * ((T + W + 15) / 16 * 16));   // woutT: (round16(T+W), 8T), pad rows zero
    b.act1 = take(2 * R * d);
    b.act2 = take(2 * R * d);
    b.act3 = take(R * 2 * d);
    b.pre = take(R * (T + W));
    b.forecast = take(R * W);
    b.bc_bnw = take(R * W);
    b.bc_bwn = take(R * W);
    b.bc_pad = take(B * W * Np);
    for (int g = 0; g < 6; ++g) {
      b.save_l[g] = training ? take(R * d) : nullptr;
      b.save_s[g] = training ? take(R * d) : nullptr;
    }
    b.fs = training ? take(R * T) : nullptr;
    for (int c = 0; c < 2; ++c) b.hscratch[c] = take((glu_chain_h_scratch_halves((int)R, (int)d, (int)(4 * W)) + 1) / 2);
  }
  if (training) {
    BwdWs& w = ws.bwd;
    const size_t H = dm.H;
    w.h_fsum = take(R * W); w.h_act = take(R * W); w.h_dhj = take(R * W); w.h_dout = take(R * H);
    w.d_fsum = take(R * W);
    w.d_pre = take(R * (T + W)); w.negdz = take(R * W); w.d_act3 = take(R * 2 * d);
    w.d_wout = take(8 * T * (T + W)); w.d_ri = take(8 * T * T); w.d_w1f = take(4 * d * 3 * W);
    w.dlrT = take(2 * d * ((R + 3) / 4 * 4)); w.inT = take(d * ((R + 3) / 4 * 4)); w.wsT = take(d * 2 * d);
    w.dlr = take(R * 2 * d); w.d_act[0] = take(R * d); w.d_act[1] = take(R * d);
    w.d_G = take(R * 3 * W); w.d_Gp = take(R * 3 * W);
    w.d_bc = take(R * W); w.d_x0 = take(R * W); w.d_mul_L = take(4 * N * N);
    w.dAsym = take(N * N); w.ddeg = take(N); w.dA = take(N * N);
    w.dots = take(R); w.d_key = take(R); w.d_query = take(R);
    w.dgh = take(N * R * 3); w.dh[0] = take(R); w.dh[1] = take(R); w.d_xs = take(N * B * W);
    w.dghT = take(3 * N * (N * B + 4)); w.hT = take(N * (N * B + 4));
  }
  ws.floats = off;
  return ws;
}

static int check_dims(const stemgnn_dims_t* dm) {
  SG_CHECK(dm != nullptr, "dims is null");
  SG_CHECK(dm->B > 0 && dm->N > 0 && dm->W > 0 && dm->H > 0 && dm->multi > 0,
           "invalid dims B=%d N=%d W=%d H=%d multi=%d", dm->B, dm->N, dm->W, dm->H, dm->multi);
  SG_CHECK(dm->W <= 64 && dm->H <= 64, "W=%d / H=%d above the supported 64", dm->W, dm->H);
  SG_CHECK((long long)dm->B * dm->N * dm->N < (1ll << 40), "problem too large");
  return 0;
}

static int check_ws(const stemgnn_dims_t* dm, int training, void* ws, size_t bytes) {
  SG_CHECK(ws != nullptr, "workspace is null");
  SG_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  const size_t need = stemgnn_workspace_bytes(dm, training);
  SG_CHECK(bytes >= need, "workspace too small: %zu < %zu bytes", bytes, need);
  return 0;
}

// ---- GLU layer dispatch: tcgen05 TF32 or fp32 FFMA2 ------------------------------------------------
int glu_layer(int M, int N, int K, const float* A, int lda, const float* Wl, const float* bl,
              const float* Wr, const float* br, float* out, int ldo, float* save_l, float* save_s,
              int gemm_mode, cudaStream_t st) {
  if (gemm_mode != 1) {
    const int rc = glu_gemm_tc(M, N, K, A, lda, Wl, bl, Wr, br, out, ldo, save_l, save_s, N, st);
    if (rc == 0) return 0;
    if (rc > 0) return rc;
    SG_CHECK(gemm_mode != 2, "tcgen05 GLU GEMM does not support M=%d N=%d K=%d lda=%d", M, N, K, lda);
  }
  GemmOperands g = {A, lda, 0, Wl, K, 0, Wr, M, N, K};
  EpiGlu epi = {out, ldo, 0, bl, br, 0, save_l, save_s, N, 0};
  return launch_sgemm<false, true, true>(g, epi, 1, st, "glu_gemm");
}

// ---- folded weights of one block ----------------------------------------------------------------
int fold_block_weights(const stemgnn_dims_t& dm, const stemgnn_block_params_t& bp, int stack_idx,
                       int kfirst, int nk, const BlockWs& b, cudaStream_t st) {
  const int W = dm.W, T = dm.multi * dm.W, d = 4 * T;
  const int PW = (stack_idx == 0) ? T + W : T;
  for (int c = 0; c < 2; ++c) {
    SG_TRY(launch_fold_in(bp.glu_left_w[c], b.w1f + (size_t)(c * 2 + 0) * d * nk * W, d, W, c, kfirst, nk, st));
    SG_TRY(launch_fold_in(bp.glu_right_w[c], b.w1f + (size_t)(c * 2 + 1) * d * nk * W, d, W, c, kfirst, nk, st));
  }
  SG_TRY(launch_irfft_table(b.ic, T, st));
  for (int c = 0; c < 2; ++c) {   // RI[c][k*T+f][u] = sum_t ic[c][f][t] weight[k][t][u]
    GemmOperands g = {b.ic + (size_t)c * T * T, T, 0, bp.weight, T, (long long)T * T, nullptr, T, T, T};
    EpiAxpby epi = {b.ri + (size_t)c * 4 * T * T, T, (long long)T * T, nullptr, 0, 0, 1.f, 0.f};
    SG_TRY((launch_sgemm<false, false, false>(g, epi, 4, st, "fold_out_ri")));
  }
  {   // woutT[o][c] = sum_u W{f,b}[o][u] RI[c][u]  (rows 0..T-1: forecast.weight, T..T+W-1: backcast.weight)
    const int PWp = (PW + 15) / 16 * 16;
    if (PWp > PW)
      SG_CUDA(cudaMemsetAsync(b.wout + (size_t)PW * 8 * T, 0, (size_t)(PWp - PW) * 8 * T * sizeof(float), st));
    GemmOperands g = {bp.forecast_w, T, 0, b.ri, T, 0, nullptr, T, 8 * T, T};
    EpiAxpby epi = {b.wout, 8 * T, 0, nullptr, 0, 0, 1.f, 0.f};
    SG_TRY((launch_sgemm<false, true, false>(g, epi, 1, st, "fold_out_forecast")));
    if (stack_idx == 0) {
      GemmOperands g2 = {bp.backcast_w, T, 0, b.ri, T, 0, nullptr, W, 8 * T, T};
      EpiAxpby epi2 = {b.wout + (size_t)T * 8 * T, 8 * T, 0, nullptr, 0, 0, 1.f, 0.f};
      SG_TRY((launch_sgemm<false, true, false>(g2, epi2, 1, st, "fold_out_backcast")));
    }
  }
  return 0;
}

// GLU chain on rows G (R x ncol) -> act3 (R x 2d) = [real3 | imag3]
int glu_chain(const stemgnn_dims_t& dm, const stemgnn_block_params_t& bp, int ncol, int gemm_mode,
              const BlockWs& b, cudaStream_t st, int reuse_w = 0, int g_ready = 0) {
  const int R = dm.B * dm.N, T = dm.multi * dm.W, d = 4 * T;
  for (int c = 0; c < 2; ++c) {
    const float* w1l = b.w1f + (size_t)(c * 2 + 0) * d * ncol;
    const float* w1r = b.w1f + (size_t)(c * 2 + 1) * d * ncol;
    float* a1 = b.act1 + (size_t)c * R * d;
    float* a2 = b.act2 + (size_t)c * R * d;
    if (gemm_mode != 1) {   // fused three-layer chain on the tensor cores (activations stay in shared memory)
      const float* const w[3][2] = {{w1l, w1r},
                                    {bp.glu_left_w[2 + c], bp.glu_right_w[2 + c]},
                                    {bp.glu_left_w[4 + c], bp.glu_right_w[4 + c]}};
      const float* const bias[3][2] = {{bp.glu_left_b[c], bp.glu_right_b[c]},
                                       {bp.glu_left_b[2 + c], bp.glu_right_b[2 + c]},
                                       {bp.glu_left_b[4 + c], bp.glu_right_b[4 + c]}};
      const bool keep = b.save_l[c] != nullptr;         // training: the backward needs every layer's tensors
      float* const act[2] = {keep ? a1 : nullptr, keep ? a2 : nullptr};
      float* const sl[3] = {b.save_l[c], b.save_l[2 + c], b.save_l[4 + c]};
      float* const ss[3] = {b.save_s[c], b.save_s[2 + c], b.save_s[4 + c]};
      // default (auto) and bf16 modes: kind::f16 chain — fp16 hi/lo split operands (fp32 parity) or bf16 operands;
      // gemm_mode 2 keeps the round-1 truncated-TF32 chain
      int rc = -1;
      static const bool no_h = getenv("STEMGNN_GLU_NO_F16") != nullptr;
      if ((gemm_mode == 0 || gemm_mode == 3) && !no_h)
        rc = glu_chain_h(gemm_mode == 3 ? 1 : 0, R, d, ncol, b.G, ncol, w, bias, b.act3 + (size_t)c * d, 2 * d, act, sl, ss,
                         reinterpret_cast<unsigned short*>(b.hscratch[c]), reuse_w,
                         c == 1 ? reinterpret_cast<const unsigned short*>(b.hscratch[0]) : nullptr, st, g_ready);
      if (rc < 0) rc = glu_chain_tc(R, d, ncol, b.G, ncol, w, bias, b.act3 + (size_t)c * d, 2 * d, act, sl, ss, st);
      if (rc == 0) continue;
      if (rc > 0) return rc;
    }
    SG_TRY(glu_layer(R, d, ncol, b.G, ncol, w1l, bp.glu_left_b[c], w1r, bp.glu_right_b[c], a1, d,
                     b.save_l[c], b.save_s[c], gemm_mode, st));
    SG_TRY(glu_layer(R, d, d, a1, d, bp.glu_left_w[2 + c], bp.glu_left_b[2 + c], bp.glu_right_w[2 + c],
                     bp.glu_right_b[2 + c], a2, d, b.save_l[2 + c], b.save_s[2 + c], gemm_mode, st));
    SG_TRY(glu_layer(R, d, d, a2, d, bp.glu_left_w[4 + c], bp.glu_left_b[4 + c], bp.glu_right_w[4 + c],
                     bp.glu_right_b[4 + c], b.act3 + (size_t)c * d, 2 * d, b.save_l[4 + c],
                     b.save_s[4 + c], gemm_mode, st));
  }
  return 0;
}

// StockBlockLayer.forward (base_model.py:61-75)
// mul_Lp / x_pad: TMA-able copies ((3N, pad4(N)) = mul_L[1..3], (B*W, pad4(N)) = the block input) or null
int block_forward(const stemgnn_dims_t& dm, const stemgnn_block_params_t& bp, int stack_idx,
                  int gemm_mode, int reuse_folded, const float* x_bnw, const float* x_bwn,
                  const float* mul_L, const BlockWs& b, float* skbuf, cudaStream_t st,
                  const float* mul_Lp = nullptr, const float* x_pad = nullptr) {
  const int B = dm.B, N = dm.N, W = dm.W, T = dm.multi * W, d = 4 * T, R = B * N;
  const int PW = (stack_idx == 0) ? T + W : T;
  if (!reuse_folded) SG_TRY(fold_block_weights(dm, bp, stack_idx, 1, 3, b, st));
  int g_ready = 0;
  static const bool no_h = getenv("STEMGNN_GLU_NO_F16") != nullptr;
  const bool h_chain = (gemm_mode == 0 || gemm_mode == 3) && !no_h && d % 16 == 0 && d <= 256 && 3 * W <= 256;
  unsigned short* g_img = h_chain ? reinterpret_cast<unsigned short*>(b.hscratch[0]) : nullptr;
  const int ldh = (3 * W + 63) / 64 * 64;
  // default / bf16 modes: the graph Fourier transform on tcgen05 (3xTF32 split operands in the default mode), its
  // epilogue writes the rows of G and the 16-bit operand images the kind::f16 GLU chain reads
  const bool tc3 = gemm_mode == 0 || gemm_mode == 3;
  const int split_ops = gemm_mode == 0 ? 1 : 0;
  int rc_gft = -1;
  if (tc3) rc_gft = gft_tc(mul_Lp, pad4(N), x_pad, pad4(N), b.G, g_img, ldh, gemm_mode == 3 ? 1 : 0, B, N, W, split_ops, st);
  if (rc_gft > 0) return rc_gft;
  if (rc_gft == 0) g_ready = g_img != nullptr ? 1 : 0;
  else   // fp32 FFMA2 split-K GEMM; its reduction also emits the operand images
    SG_TRY(launch_gft(mul_L, x_bwn, b.G, skbuf, B, N, W, st, g_img, ldh, gemm_mode == 3 ? 1 : 0, &g_ready));
  SG_TRY(glu_chain(dm, bp, 3 * W, gemm_mode, b, st, reuse_folded, g_ready));
  HeadArgs h = {};
  h.pre = b.pre; h.ldp = PW; h.x_bnw = x_bnw;
  h.bf = bp.forecast_b; h.wfr = bp.forecast_result_w; h.bfr = bp.forecast_result_b;
  h.forecast = b.forecast; h.save_fs = b.fs;
  h.R = R; h.N = N; h.T = T; h.W = W;
  if (stack_idx == 0) {
    h.bb = bp.backcast_b; h.wsc = bp.shortcut_w; h.bsc = bp.shortcut_b;
    h.backcast_bnw = b.bc_bnw; h.backcast_bwn = b.bc_bwn;
  }
  if (tc3) {   // folded output map + sigmoid heads in one launch (the pre-activations never leave tensor memory)
    const int rc = out_head_tc(b.act3, 2 * d, b.wout, (PW + 15) / 16 * 16, h, b.bc_pad, pad4(N), split_ops, st);
    if (rc == 0) return 0;
    if (rc > 0) return rc;
  }
  {   // pre = [real3 | imag3] @ woutT^T : tcgen05 TF32 unless exact fp32 is requested
    int rc = -1;
    if (gemm_mode != 1)
      rc = tc_gemm(R, (PW + 15) / 16 * 16, 2 * d, 1.f, b.act3, 2 * d, b.wout, 2 * d, (PW + 15) / 16 * 16, b.pre,
                   nullptr, 0, PW, PW, 0, 1, st, gemm_mode == 0 ? -1 : 0);
    if (rc > 0) return rc;
    if (rc < 0) {   // exact fp32 requested, or a shape outside the tensor-core kernel's envelope
      GemmOperands g = {b.act3, 2 * d, 0, b.wout, 2 * d, 0, nullptr, R, PW, 2 * d};
      EpiAxpby epi = {b.pre, PW, 0, nullptr, 0, 0, 1.f, 0.f};
      SG_TRY((launch_sgemm<false, true, false>(g, epi, 1, st, "out_gemm")));
    }
  }
  SG_TRY(launch_block_head(h, st));
  if (stack_idx == 0 && tc3)   // the next block's tcgen05 GFT reads the padded copy
    SG_TRY(launch_pad_rows(b.bc_bwn, (long long)B * W, N, N, b.bc_pad, pad4(N), st));
  return 0;
}

static int check_block_params(const stemgnn_block_params_t* bp, int stack_idx) {
  SG_CHECK(bp != nullptr, "block params null");
  SG_CHECK(bp->weight && bp->forecast_w && bp->forecast_b && bp->forecast_result_w &&
               bp->forecast_result_b, "block %d: null parameter pointer", stack_idx);
  if (stack_idx == 0)
    SG_CHECK(bp->backcast_w && bp->backcast_b && bp->shortcut_w && bp->shortcut_b,
             "block 0: null backcast/shortcut parameter");
  for (int g = 0; g < 6; ++g)
    SG_CHECK(bp->glu_left_w[g] && bp->glu_left_b[g] && bp->glu_right_w[g] && bp->glu_right_b[g],
             "block %d: null GLU %d parameter", stack_idx, g);
  return 0;
}

// C[i] = alpha * sum_z P[z][i] + beta * Cin[i]   (fixed summation order)
// C2 (optional): second copy with row pitch ld2 (rows of `ncols` elements)
__global__ void splitk_reduce_kernel(const float* __restrict__ P, int ks, int n, float alpha,
                                     const float* __restrict__ Cin, float beta, float* __restrict__ C,
                                     float* __restrict__ C2, int ncols, int ld2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int z = 0; z < ks; ++z) acc += P[(long long)z * n + i];
  const float v = alpha * acc + (Cin != nullptr ? beta * Cin[i] : 0.f);
  C[i] = v;
  if (C2 != nullptr) C2[(long long)(i / ncols) * ld2 + i % ncols] = v;      // ld2 may carry a row interleave (3 * pitch)
}

int graph_forward(const stemgnn_dims_t& dm, const stemgnn_fwd_opts_t& op, const float* key,
                  const float* query, float* attention, const Workspace& ws, cudaStream_t st) {
  const int B = dm.B, N = dm.N;
  AttnArgs a = {};
  a.key = key; a.query = query; a.qmax = ws.qmax; a.a_raw = ws.a_raw; a.deg = ws.deg;
  a.row_m = ws.row_m; a.row_zinv = ws.row_zinv;
  a.mask = op.dropout_mask; a.seed = op.dropout_seed; a.offset = op.dropout_offset;
  a.offset_dev = op.dropout_offset_dev;
  a.alpha = op.leaky_alpha; a.p = op.dropout_p;
  a.use_dropout = (op.training && (op.dropout_p > 0.f || op.dropout_mask != nullptr)) ? 1 : 0;
  a.B = B; a.N = N;
  SG_CHECK(op.dropout_p >= 0.f && op.dropout_p < 1.f, "dropout_p=%f outside [0,1)", op.dropout_p);
  SG_TRY(launch_attention(a, ws.qmax, st));
  if (op.graph_allreduce != nullptr) {   // data parallel, global-batch graph: mean over ranks of the shard means
    op.graph_allreduce(ws.a_raw, (long long)N * N, op.graph_allreduce_user, st);
    op.graph_allreduce(ws.deg, N, op.graph_allreduce_user, st);
  }
  // ws.mul_Lp: (3N, Np) copy of mul_L[1..3] for the tcgen05 graph Fourier transform, rows interleaved as n*3 + k'
  const int Np = pad4(N);
  SG_TRY(launch_laplacian(ws.a_raw, ws.deg, attention, ws.mul_L, N, st, ws.mul_Lp, Np));
  const size_t nn = (size_t)N * N;
  if (op.graph_mode == 1 && !op.training && N >= 2 && N <= 512) {
    // opt-in eigendecomposition path (north_star; the reference's dead `graph_fft` hook): L = U Lambda U^T by the fused
    // Laplacian + Jacobi kernel, polynomial stack rebuilt as U p(Lambda) U^T
    SG_TRY(laplacian_eig(ws.a_raw, ws.deg, N, ws.eig_lambda, ws.eig_U, ws.eig_info, 30, 1e-6f, st));
    SG_TRY(eig_poly_stack(ws.eig_lambda, ws.eig_U, N, ws.eig_S, ws.mul_L, st));
    return launch_pad_rows(ws.mul_L + nn, 3ll * N, N, N, ws.mul_Lp, Np, st, 3, 0, N);
  }
  // the two N^3 Chebyshev products fill only a few CTAs: deterministic split-K (partials + ordered reduce)
  const int ks = pick_ksplit(N, N, N);
  for (int term = 2; term <= 3; ++term) {
    // term 2: third = 2 L L (base_model.py:131; first_laplacian is zero); term 3: forth = 2 L third - L (:132)
    const float* Bm = ws.mul_L + (size_t)(term - 1) * nn;
    float* Cm = ws.mul_L + (size_t)term * nn;
    const float* Cin = term == 3 ? ws.mul_L + nn : nullptr;
    if (ks > 1 && ks <= 8) {
      GemmOperands g = {ws.mul_L + nn, N, 0, Bm, N, 0, nullptr, N, N, N, ks};
      EpiPartial epi = {ws.skbuf, N, (long long)nn};
      SG_TRY((launch_sgemm<false, false, false>(g, epi, 1, st, "cheb_splitk")));
      splitk_reduce_kernel<<<ceil_div((int)nn, 256), 256, 0, st>>>(ws.skbuf, ks, (int)nn, 2.f, Cin, -1.f, Cm,
                                                                    ws.mul_Lp + (size_t)(term - 1) * Np, N, 3 * Np);
      SG_LAUNCH_CHECK("splitk_reduce_kernel");
    } else {
      GemmOperands g = {ws.mul_L + nn, N, 0, Bm, N, 0, nullptr, N, N, N};
      EpiAxpby epi = {Cm, N, 0, Cin, N, 0, 2.f, Cin != nullptr ? -1.f : 0.f};
      SG_TRY((launch_sgemm<false, false, false>(g, epi, 1, st, "cheb")));
      SG_TRY(launch_pad_rows(Cm, N, N, N, ws.mul_Lp, Np, st, 3, term - 1));
    }
  }
  return 0;
}

}  // namespace sg

using namespace sg;

extern "C" {

int stemgnn_version(void) { return STEMGNN_ABI_VERSION; }
const char* stemgnn_last_error(void) { return g_err; }

long long stemgnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* stemgnn_gru_kernel_name(void) { return gru_tc_kernel_name(); }

void stemgnn_profile_gru(void* start_event, void* stop_event) {
  g_hook.start = static_cast<cudaEvent_t>(start_event);
  g_hook.stop = static_cast<cudaEvent_t>(stop_event);
}

int stemgnn_device_ok(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return p.major == 10 ? 1 : 0;
}

size_t stemgnn_workspace_bytes(const stemgnn_dims_t* dims, int training) {
  if (dims == nullptr || dims->B <= 0 || dims->N <= 0 || dims->W <= 0 || dims->multi <= 0) return 0;
  return carve_workspace(*dims, training, nullptr).floats * sizeof(float);
}

int stemgnn_model_forward(const stemgnn_dims_t* dims, const stemgnn_params_t* p,
                          const stemgnn_fwd_opts_t* opts, const float* x, float* forecast,
                          float* attention, float* mul_L, void* workspace, size_t workspace_bytes,
                          stemgnn_stream_t stream) {
  clear_error();
  SG_TRY(check_dims(dims));
  SG_CHECK(p && opts && x && forecast && attention, "null argument");
  SG_CHECK(p->weight_key && p->weight_query && p->gru_w_ih && p->gru_w_hh && p->gru_b_ih &&
               p->gru_b_hh && p->fc0_w && p->fc0_b && p->fc2_w && p->fc2_b, "null parameter pointer");
  SG_TRY(check_block_params(&p->block[0], 0));
  SG_TRY(check_block_params(&p->block[1], 1));
  SG_TRY(check_ws(dims, opts->training, workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const stemgnn_dims_t& dm = *dims;
  Workspace ws = carve_workspace(dm, opts->training, static_cast<float*>(workspace));

  SG_TRY(launch_prep_layouts(x, ws.xs, ws.x_bnw, dm.B, dm.W, dm.N, st, ws.x_pad, pad4(dm.N)));
  GruArgs ga = {ws.xs, p->gru_w_ih, p->gru_w_hh, p->gru_b_ih, p->gru_b_hh, p->weight_key,
                p->weight_query, ws.key, ws.query, ws.h_all, ws.gi, ws.g_r, ws.g_z, ws.g_n, ws.g_hn, dm.B, dm.N, dm.W};
  ga.tc_reuse = (opts->reuse_folded && !opts->training) ? 1 : 0;
  SG_TRY(gru_keyquery_forward(ga, 0, ws.gru_scratch, st));
  SG_TRY(graph_forward(dm, *opts, ws.key, ws.query, attention, ws, st));
  SG_TRY(block_forward(dm, p->block[0], 0, opts->gemm_mode, opts->reuse_folded && !opts->training, ws.x_bnw, x,
                       ws.mul_L, ws.blk[0], ws.skbuf, st, ws.mul_Lp, ws.x_pad));
  SG_TRY(block_forward(dm, p->block[1], 1, opts->gemm_mode, opts->reuse_folded && !opts->training,
                       ws.blk[0].bc_bnw, ws.blk[0].bc_bwn, ws.mul_L, ws.blk[1], ws.skbuf, st, ws.mul_Lp,
                       ws.blk[0].bc_pad));
  SG_TRY(launch_model_head(ws.blk[0].forecast, ws.blk[1].forecast, p->fc0_w, p->fc0_b, p->fc2_w,
                           p->fc2_b, forecast, dm.B, dm.N, dm.W, dm.H, st));
  if (mul_L != nullptr)
    SG_CUDA(cudaMemcpyAsync(mul_L, ws.mul_L, (size_t)4 * dm.N * dm.N * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
  return 0;
}

int stemgnn_gru_keyquery_forward(const stemgnn_dims_t* dims, const stemgnn_params_t* p,
                                 const float* x, float* key, float* query, float* gru_out, int path,
                                 void* workspace, size_t workspace_bytes, stemgnn_stream_t stream) {
  clear_error();
  SG_TRY(check_dims(dims));
  SG_CHECK(p && x && key && query, "null argument");
  SG_CHECK(p->weight_key && p->weight_query && p->gru_w_ih && p->gru_w_hh && p->gru_b_ih &&
               p->gru_b_hh, "null GRU parameter pointer");
  SG_TRY(check_ws(dims, 0, workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = carve_workspace(*dims, 0, static_cast<float*>(workspace));
  SG_TRY(launch_prep_layouts(x, ws.xs, ws.x_bnw, dims->B, dims->W, dims->N, st));
  GruArgs ga = {ws.xs, p->gru_w_ih, p->gru_w_hh, p->gru_b_ih, p->gru_b_hh, p->weight_key,
                p->weight_query, key, query, gru_out, ws.gi, nullptr, nullptr, nullptr, nullptr, dims->B, dims->N, dims->W};
  return gru_keyquery_forward(ga, path, ws.gru_scratch, st);
}

int stemgnn_graph_forward(const stemgnn_dims_t* dims, const stemgnn_fwd_opts_t* opts,
                          const float* key, const float* query, float* attention, float* mul_L,
                          void* workspace, size_t workspace_bytes, stemgnn_stream_t stream) {
  clear_error();
  SG_TRY(check_dims(dims));
  SG_CHECK(opts && key && query && attention && mul_L, "null argument");
  stemgnn_fwd_opts_t op = *opts;
  SG_TRY(check_ws(dims, 0, workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = carve_workspace(*dims, 0, static_cast<float*>(workspace));
  SG_TRY(graph_forward(*dims, op, key, query, attention, ws, st));
  SG_CUDA(cudaMemcpyAsync(mul_L, ws.mul_L, (size_t)4 * dims->N * dims->N * sizeof(float),
                          cudaMemcpyDeviceToDevice, st));
  return 0;
}

// x_bnw (B,N,W) -> (B,W,N)
__global__ void bnw_to_bwn_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int N,
                                  int W) {
  const long long total = (long long)B * N * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % N), t = (int)((idx / N) % W), b = (int)(idx / ((long long)N * W));
    out[idx] = in[((long long)b * N + n) * W + t];
  }
}

int stemgnn_block_forward(const stemgnn_dims_t* dims, const stemgnn_block_params_t* bp,
                          int stack_idx, int gemm_mode, const float* x_bnw, const float* mul_L,
                          float* forecast, float* backcast, void* workspace, size_t workspace_bytes,
                          stemgnn_stream_t stream) {
  clear_error();
  SG_TRY(check_dims(dims));
  SG_CHECK(stack_idx == 0 || stack_idx == 1, "stack_idx=%d", stack_idx);
  SG_TRY(check_block_params(bp, stack_idx));
  SG_CHECK(x_bnw && mul_L && forecast && (stack_idx == 1 || backcast), "null argument");
  SG_TRY(check_ws(dims, 0, workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = carve_workspace(*dims, 0, static_cast<float*>(workspace));
  const BlockWs& b = ws.blk[stack_idx];
  const long long total = (long long)dims->B * dims->N * dims->W;
  float* x_bwn = ws.blk[1 - stack_idx].bc_bwn;   // scratch from the other block's slot
  bnw_to_bwn_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(x_bnw, x_bwn, dims->B, dims->N, dims->W);
  SG_LAUNCH_CHECK("bnw_to_bwn_kernel");
  // TMA-able copies of the caller's operands, so that the stage API runs the same tcgen05 kernels as the model path
  const int Np = pad4(dims->N);
  float* x_pad = ws.blk[1 - stack_idx].bc_pad;
  SG_TRY(launch_pad_rows(x_bwn, (long long)dims->B * dims->W, dims->N, dims->N, x_pad, Np, st));
  SG_TRY(launch_pad_rows(mul_L + (size_t)dims->N * dims->N, 3ll * dims->N, dims->N, dims->N, ws.mul_Lp, Np, st, 3, 0,
                         dims->N));
  SG_TRY(block_forward(*dims, *bp, stack_idx, gemm_mode, 0, x_bnw, x_bwn, mul_L, b, ws.skbuf, st, ws.mul_Lp, x_pad));
  SG_CUDA(cudaMemcpyAsync(forecast, b.forecast, (size_t)total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (stack_idx == 0)
    SG_CUDA(cudaMemcpyAsync(backcast, b.bc_bnw, (size_t)total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int stemgnn_spe_seq_cell_forward(const stemgnn_dims_t* dims, const stemgnn_block_params_t* bp,
                                 int gemm_mode, const float* gfted, float* iffted, void* workspace,
                                 size_t workspace_bytes, stemgnn_stream_t stream) {
  clear_error();
  SG_TRY(check_dims(dims));
  SG_TRY(check_block_params(bp, 1));
  SG_CHECK(gfted && iffted, "null argument");
  SG_TRY(check_ws(dims, 0, workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = carve_workspace(*dims, 0, static_cast<float*>(workspace));
  const BlockWs& b = ws.blk[0];
  const int W = dims->W, T = dims->multi * W;
  // all four Chebyshev channels are honoured here (a caller may pass a non-zero channel 0)
  SG_TRY(fold_block_weights(*dims, *bp, 1, 0, 4, b, st));
  SG_TRY(launch_gfted_to_rows(gfted, b.G, dims->B, dims->N, W, st));
  SG_TRY(glu_chain(*dims, *bp, 4 * W, gemm_mode, b, st));
  return launch_irfft_rows(b.act3, b.ic, iffted, dims->B, dims->N, T, st);
}

// x[b][t][n] = series[(end_idx[b] - W + t) * N + n] ;  y[b][h][n] = series[(end_idx[b] + h) * N + n]
__global__ void gather_windows_kernel(const float* __restrict__ series, int T, int N,
                                      const int32_t* __restrict__ end_idx, int B, int W, int H,
                                      float* __restrict__ x, float* __restrict__ y) {
  const long long per = (long long)(W + H) * N;
  const long long total = (long long)B * per;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    const long long r = idx - (long long)b * per;
    const int row = (int)(r / N), n = (int)(r % N);
    const long long src_row = (long long)end_idx[b] - W + row;
    const float v = (src_row >= 0 && src_row < T) ? series[src_row * N + n] : 0.f;
    if (row < W) x[((long long)b * W + row) * N + n] = v;
    else y[((long long)b * H + (row - W)) * N + n] = v;
  }
}

int stemgnn_gather_windows(const float* series, int T, int N, const int32_t* end_idx, int B, int W, int H,
                           float* x, float* y, stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(series && end_idx && x && y && T > 0 && N > 0 && B > 0 && W > 0 && H >= 0, "gather_windows: bad arguments");
  const long long total = (long long)B * (W + H) * N;
  const int blocks = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
  gather_windows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(series, T, N, end_idx, B, W, H, x, y);
  SG_LAUNCH_CHECK("gather_windows_kernel");
  return 0;
}

int stemgnn_sgemm(int M, int N, int K, float alpha, const float* A, int lda, int a_kmajor,
                  const float* B, int ldb, int b_nk, float beta, float* C, int ldc,
                  stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(M >= 0 && N >= 0 && K >= 0 && A && B && C, "sgemm: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmOperands g = {A, lda, 0, B, ldb, 0, nullptr, M, N, K};
  EpiAxpby epi = {C, ldc, 0, beta != 0.f ? C : nullptr, ldc, 0, alpha, beta};
  if (!a_kmajor && b_nk) return launch_sgemm<false, true, false>(g, epi, 1, st, "sgemm_nt");
  if (!a_kmajor && !b_nk) return launch_sgemm<false, false, false>(g, epi, 1, st, "sgemm_nn");
  if (a_kmajor && b_nk) return launch_sgemm<true, true, false>(g, epi, 1, st, "sgemm_tt");
  return launch_sgemm<true, false, false>(g, epi, 1, st, "sgemm_tn");
}

int stemgnn_tc_gemm(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                    int split_ops, stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(M > 0 && N > 0 && K > 0 && A && B && C, "tc_gemm: bad arguments");
  const int rc = tc_gemm(M, N, K, 1.f, A, lda, B, ldb, N, C, nullptr, 0, ldc, N, 0, 1, static_cast<cudaStream_t>(stream),
                         split_ops ? 1 : 0);
  SG_CHECK(rc >= 0, "tc_gemm: unsupported shape M=%d N=%d K=%d lda=%d ldb=%d", M, N, K, lda, ldb);
  return rc;
}

int stemgnn_gft_forward(const float* mul_L, const float* x, float* G, int B, int N, int W, float* scratch,
                        stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(mul_L && x && G && scratch && B > 0 && N > 0 && W > 0, "gft_forward: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Np = pad4(N);
  float* Lp = scratch;
  float* xp = scratch + (size_t)4 * N * Np;
  SG_TRY(launch_pad_rows(mul_L + (size_t)N * N, 3ll * N, N, N, Lp, Np, st, 3, 0, N));
  SG_TRY(launch_pad_rows(x, (long long)B * W, N, N, xp, Np, st));
  const int rc = gft_tc(Lp, Np, xp, Np, G, nullptr, 0, 0, B, N, W, 1, st);
  SG_CHECK(rc >= 0, "gft_forward: tcgen05 path unavailable for B=%d N=%d W=%d", B, N, W);
  return rc;
}

int stemgnn_glu_gemm(int M, int N, int K, const float* A, int lda, const float* Wl, const float* bl,
                     const float* Wr, const float* br, float* out, int ldo, int use_tc,
                     stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(M > 0 && N > 0 && K > 0 && A && Wl && bl && Wr && br && out, "glu_gemm: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return glu_layer(M, N, K, A, lda, Wl, bl, Wr, br, out, ldo, nullptr, nullptr, use_tc ? 2 : 1, st);
}

int stemgnn_glu_chain(int M, int N, int K1, const float* G, int ldg, const float* const* weights, const float* const* biases,
                      float* out3, int ldo3, int mode, void* scratch, stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(M > 0 && G && weights && biases && out3 && scratch, "glu_chain: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* const w[3][2] = {{weights[0], weights[1]}, {weights[2], weights[3]}, {weights[4], weights[5]}};
  const float* const b[3][2] = {{biases[0], biases[1]}, {biases[2], biases[3]}, {biases[4], biases[5]}};
  float* const act[2] = {nullptr, nullptr};
  float* const sv[3] = {nullptr, nullptr, nullptr};
  int rc;
  if (mode == 2) rc = glu_chain_tc(M, N, K1, G, ldg, w, b, out3, ldo3, act, sv, sv, st);
  else rc = glu_chain_h(mode == 3 ? 1 : 0, M, N, K1, G, ldg, w, b, out3, ldo3, act, sv, sv,
                        static_cast<unsigned short*>(scratch), 0, nullptr, st);
  SG_CHECK(rc >= 0, "glu_chain: unsupported shape M=%d N=%d K1=%d", M, N, K1);
  return rc;
}

size_t stemgnn_glu_chain_scratch_bytes(int M, int N, int K1) { return glu_chain_h_scratch_halves(M, N, K1) * 2; }

int stemgnn_model_backward(const stemgnn_dims_t* dims, const stemgnn_params_t* params,
                           const stemgnn_fwd_opts_t* opts, const float* x, const float* d_forecast,
                           const float* d_attention, const stemgnn_grads_t* grads, float* d_x,
                           void* workspace, size_t workspace_bytes, stemgnn_stream_t stream) {
  clear_error();
  return model_backward(dims, params, opts, x, d_forecast, d_attention, grads, d_x, workspace,
                        workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
