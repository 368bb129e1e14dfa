// Host-side schedule of one fused-MLP backward: a single C-ABI call issues the whole kernel chain
//   LayerNorm backward -> wgrad L3 -> dgrad L3 (* act'(a2)) -> wgrad L2 -> dgrad L2 (* act'(a1)) -> wgrad L1
//   -> input gradients per segment (optionally accumulated onto a residual)
// so the caller (gnn_fluid_dynamics_b200/training.py) pays one foreign call per MLP instead of ~14 and the
// GPU, not the host, bounds the training step.  No allocation: scratch comes from the caller's workspace.
#include "common.cuh"

namespace gnnfd {
size_t pack_mlp_bytes_tc(int k_in, int hidden, int n_out, int precision);
int ln_backward_launch(const float *g, const float *xhat, const float *rstd, const float *ln_w, int64_t rows,
                       float *dy, float *sums, void *workspace, size_t workspace_bytes, cudaStream_t stream);
size_t ln_backward_ws(int64_t rows);

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BwdLayout {
  size_t dy, da2, da1, sums, lnws, wgws, total;
  size_t pk_chain, pk_w3t, pk_w2t, pk_w1t[3], pk_total;
};

static BwdLayout bwd_layout(const gnnfd_mlp_args *f) {
  BwdLayout L{};
  const size_t mat = al256((size_t)f->rows * 128 * 4);
  size_t o = 0;
  L.dy = o; o += f->has_ln ? mat : 0;
  L.da2 = o; o += mat;
  L.da1 = o; o += mat;
  L.sums = o; o += al256(3 * 128 * 4);
  L.lnws = o; o += al256(ln_backward_ws(f->rows));
  // the weight-gradient GEMMs of one MLP keep their split-K partials side by side (dW3, dW2, dW1 lead + rest: <= 128 +
  // 128 + 128 + 256 columns) and are reduced by ONE launch at the end
  L.wgws = o; o += 4 * (al256(gnnfd_wgrad_workspace_bytes(f->rows, 160)) + 256);   // 4 GEMMs: 640 columns + 4 column-sum regions
  L.total = o;
  size_t p = 0;
  L.pk_chain = p; p += al256(pack_mlp_bytes_tc(f->n_out, 128, 128, f->precision));
  L.pk_w3t = p; p += al256(pack_mlp_bytes_tc(f->n_out, 128, 128, f->precision));
  L.pk_w2t = p; p += al256(pack_mlp_bytes_tc(128, 128, 128, f->precision));
  for (int s = 0; s < 3; ++s) { L.pk_w1t[s] = p; if (s < f->n_seg) p += al256(pack_mlp_bytes_tc(128, 128, 128, f->precision)); }
  L.pk_total = p;
  return L;
}

// one tensor-core Linear  out = (src . Wt^T) (* act'(mul)) (+ residual)   (see gnnfd_mlp_args.n_layers == 1)
static void linear_args(gnnfd_mlp_args &a, const gnnfd_mlp_args *f, const float *src, int src_ld, int k_in,
                        const float *w, int ld_n, int ld_k, int w_rows) {
  a = gnnfd_mlp_args{};
  a.rows = f->rows;
  a.n_seg = 1;
  a.seg[0].src = src; a.seg[0].ld = src_ld; a.seg[0].col = 0; a.seg[0].width = k_in; a.seg[0].mode = GNNFD_SEG_DIRECT;
  a.k_in = k_in; a.hidden = 128; a.n_out = 128;
  a.w1 = w; a.w1_ld_n = ld_n; a.w1_ld_k = ld_k; a.w1_rows = w_rows;
  a.act = GNNFD_ACT_SILU;
  a.precision = f->precision;
  a.n_layers = 1;
}
// the dgrad chain as one 3-layer pass: dy -> (.W3) * act'(a2) -> (.W2) * act'(a1) -> .W1[:, segment 0]
static void chain_args(gnnfd_mlp_args &a, const gnnfd_mlp_args *f, const float *dy) {
  a = gnnfd_mlp_args{};
  a.rows = f->rows;
  a.n_seg = 1;
  a.seg[0].src = dy; a.seg[0].ld = f->n_out; a.seg[0].col = 0; a.seg[0].width = f->n_out; a.seg[0].mode = GNNFD_SEG_DIRECT;
  a.k_in = f->n_out; a.hidden = 128; a.n_out = 128;
  a.w1 = f->w3; a.w1_ld_n = 1; a.w1_ld_k = 128; a.w1_rows = 128;
  a.w2 = f->w2; a.w2_ld_n = 1; a.w2_ld_k = 128;
  a.w3 = f->w1; a.w3_ld_n = 1; a.w3_ld_k = f->k_in; a.w3_rows = f->seg[0].width;
  a.act = f->act;
  a.precision = f->precision;
  a.n_layers = 3;
  a.bwd_chain = 1;
}
}  // namespace gnnfd

using namespace gnnfd;

extern "C" size_t gnnfd_mlp_backward_workspace_bytes(const gnnfd_mlp_args *fwd) {
  if (fwd == nullptr) return 0;
  return bwd_layout(fwd).total + 256;
}

extern "C" size_t gnnfd_pack_mlp_backward_bytes(const gnnfd_mlp_args *fwd) {
  if (fwd == nullptr || fwd->precision == GNNFD_PREC_F32) return 0;
  return bwd_layout(fwd).pk_total + 256;
}

extern "C" int gnnfd_pack_mlp_backward(const gnnfd_mlp_args *f, void *packed_out, int32_t chain, void *stream) {
  GNNFD_CHECK_ARG(f != nullptr && packed_out != nullptr, "null argument");
  GNNFD_CHECK_ARG(f->precision != GNNFD_PREC_F32, "the backward needs a tensor-core precision");
  const BwdLayout L = bwd_layout(f);
  uint8_t *pk = (uint8_t *)packed_out;
  gnnfd_mlp_args a;
  float dummy_out;
  int rc;
  if (chain) {                                                             // one 3-layer pass incl. dIn_0
    chain_args(a, f, f->w3);
    a.hid_mul1 = a.hid_mul2 = f->w3;
    a.out_raw = &dummy_out;
    if ((rc = gnnfd_pack_mlp(&a, pk + L.pk_chain, stream)) != GNNFD_OK) return rc;
  } else {
    linear_args(a, f, f->w3, f->n_out, f->n_out, f->w3, 1, 128, 128);     // dH2 = dy . W3
    a.out_raw = &dummy_out;
    if ((rc = gnnfd_pack_mlp(&a, pk + L.pk_w3t, stream)) != GNNFD_OK) return rc;
    linear_args(a, f, f->w2, 128, 128, f->w2, 1, 128, 128);               // dH1 = dA2 . W2
    a.out_raw = &dummy_out;
    if ((rc = gnnfd_pack_mlp(&a, pk + L.pk_w2t, stream)) != GNNFD_OK) return rc;
  }
  int col0 = 0;
  for (int s = 0; s < f->n_seg; ++s) {                                    // dIn_s = dA1 . W1[:, col0:col0+width]
    if (!(chain && s == 0)) {
      linear_args(a, f, f->w1, 128, 128, f->w1 + col0, 1, f->k_in, f->seg[s].width);
      a.out_raw = &dummy_out;
      if ((rc = gnnfd_pack_mlp(&a, pk + L.pk_w1t[s], stream)) != GNNFD_OK) return rc;
    }
    col0 += f->seg[s].width;
  }
  return GNNFD_OK;
}

extern "C" int gnnfd_mlp_backward(const gnnfd_mlp_backward_args *b, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(b != nullptr, "null args");
  const gnnfd_mlp_args *f = &b->fwd;
  GNNFD_CHECK_ARG(f->precision != GNNFD_PREC_F32, "the backward needs a tensor-core precision");
  GNNFD_CHECK_ARG(f->hidden == 128 && f->n_out >= 1 && f->n_out <= 128, "bad widths");
  GNNFD_CHECK_ARG(f->mul == nullptr, "backward through the `mul` epilogue is not implemented");
  GNNFD_CHECK_ARG(f->n_seg >= 1 && f->n_seg <= 3, "n_seg must be 1..3");
  if (f->rows == 0) {
    // no rows: parameter gradients are zero
    const int k = f->k_in, n = f->n_out;
    if (b->d_w1 && b->skip_wgrad_l1 != 1) GNNFD_CUDA(cudaMemsetAsync(b->d_w1, 0, (size_t)128 * k * 4, stream));
    if (b->d_w2) GNNFD_CUDA(cudaMemsetAsync(b->d_w2, 0, (size_t)128 * 128 * 4, stream));
    if (b->d_w3) GNNFD_CUDA(cudaMemsetAsync(b->d_w3, 0, (size_t)n * 128 * 4, stream));
    if (b->d_b1) GNNFD_CUDA(cudaMemsetAsync(b->d_b1, 0, 128 * 4, stream));
    if (b->d_b2) GNNFD_CUDA(cudaMemsetAsync(b->d_b2, 0, 128 * 4, stream));
    if (b->d_b3) GNNFD_CUDA(cudaMemsetAsync(b->d_b3, 0, n * 4, stream));
    if (b->d_ln_w) GNNFD_CUDA(cudaMemsetAsync(b->d_ln_w, 0, n * 4, stream));
    if (b->d_ln_b) GNNFD_CUDA(cudaMemsetAsync(b->d_ln_b, 0, n * 4, stream));
    return GNNFD_OK;
  }
  GNNFD_CHECK_ARG(b->g && b->a1 && b->a2 && b->packed_bwd && b->workspace, "null pointer");
  GNNFD_CHECK_ARG((b->d_w1 || b->skip_wgrad_l1) && b->d_w2 && b->d_w3, "null weight-gradient output");
  const BwdLayout L = bwd_layout(f);
  if (b->workspace_bytes < L.total) { set_error("gnnfd_mlp_backward: workspace too small"); return GNNFD_E_WORKSPACE; }
  uint8_t *ws = (uint8_t *)b->workspace;
  const uint8_t *pk = (const uint8_t *)b->packed_bwd;
  float *da2 = (float *)(ws + L.da2), *da1 = b->da1_out ? b->da1_out : (float *)(ws + L.da1), *sums = (float *)(ws + L.sums);
  uint8_t *wgws = ws + L.wgws;
  size_t wgws_bytes = L.total - L.wgws;
  WgReduceJob jobs[WG_MAX_REDUCE_JOBS];
  int n_jobs = 0;
  // run one weight-gradient GEMM with its reduction deferred; its partials take the next slice of the workspace
  auto wgrad_deferred = [&](const gnnfd_wgrad_args &w) -> int {
    if (n_jobs >= WG_MAX_REDUCE_JOBS) { set_error("gnnfd_mlp_backward: too many weight-gradient GEMMs"); return GNNFD_E_BADARG; }
    const int r = wgrad_run(&w, wgws, wgws_bytes, stream, &jobs[n_jobs]);
    if (r != GNNFD_OK) return r;
    wgws += jobs[n_jobs].ws_used; wgws_bytes -= jobs[n_jobs].ws_used;
    ++n_jobs;
    return GNNFD_OK;
  };
  const int code = f->act + 1;   // SiLU -> 1, tanh -> 2
  const int n_out = f->n_out;
  int rc;

  // ---- LayerNorm backward
  const float *dy = b->g;
  bool b3_from_ln = false;
  if (f->has_ln) {
    GNNFD_CHECK_ARG(n_out == 128 && b->xhat && b->rstd, "LayerNorm backward needs xhat/rstd and n_out == 128");
    float *dyb = (float *)(ws + L.dy);
    if ((rc = ln_backward_launch(b->g, b->xhat, b->rstd, f->ln_w, f->rows, dyb, sums, ws + L.lnws, L.wgws - L.lnws, stream)) != GNNFD_OK) return rc;
    dy = dyb;
    if (b->d_ln_w) GNNFD_CUDA(cudaMemcpyAsync(b->d_ln_w, sums, 128 * 4, cudaMemcpyDeviceToDevice, stream));
    if (b->d_ln_b) GNNFD_CUDA(cudaMemcpyAsync(b->d_ln_b, sums + 128, 128 * 4, cudaMemcpyDeviceToDevice, stream));
    if (b->d_b3) GNNFD_CUDA(cudaMemcpyAsync(b->d_b3, sums + 256, 128 * 4, cudaMemcpyDeviceToDevice, stream));
    b3_from_ln = true;
  }
  auto direct = [](const float *src, int ld, int width) {
    gnnfd_segment s{};
    s.src = src; s.ld = ld; s.col = 0; s.width = width; s.mode = GNNFD_SEG_DIRECT;
    return s;
  };
  // ---- dgrad chain: dA2 = (dy W3) * act'(a2), dA1 = (dA2 W2) * act'(a1) [, dIn_0 = dA1 W1[:, seg 0] (+ residual)]
  const bool chain = b->din_out[0] != nullptr;
  if (chain) {
    gnnfd_mlp_args a;
    chain_args(a, f, dy);
    a.packed = pk + L.pk_chain;
    a.hid_mul1 = b->a2; a.hid_mul2 = b->a1; a.save_a1 = da2; a.save_a2 = da1;
    if (b->din_residual[0] != nullptr) { a.residual = b->din_residual[0]; a.out_sum = b->din_out[0]; }
    else a.out_raw = b->din_out[0];
    if ((rc = gnnfd_mlp_forward(&a, stream)) != GNNFD_OK) return rc;
  } else {
    gnnfd_mlp_args a;
    linear_args(a, f, dy, n_out, n_out, f->w3, 1, 128, 128);
    a.packed = pk + L.pk_w3t; a.mul = b->a2; a.mul_mode = code; a.out_raw = da2;
    if ((rc = gnnfd_mlp_forward(&a, stream)) != GNNFD_OK) return rc;
    linear_args(a, f, da2, 128, 128, f->w2, 1, 128, 128);
    a.packed = pk + L.pk_w2t; a.mul = b->a1; a.mul_mode = code; a.out_raw = da1;
    if ((rc = gnnfd_mlp_forward(&a, stream)) != GNNFD_OK) return rc;
  }
  // ---- weight gradients.  dW3 = dy^T act(a2), dW2 = dA2^T act(a1) and the leading contiguous block of dW1 = dA1^T In
  //      are direct x direct GEMMs over the same rows: ONE lean launch with three TMEM accumulators.  (skip_wgrad_l1:
  //      1 = all of dW1 / db1 is left to the caller, 2 = only the assembled (gathered) segments are.)
  gnnfd_wgrad_args w3{}, w2{}, w1{};
  {
    w3.rows = f->rows;
    float *cs = (b->d_b3 && !b3_from_ln) ? b->d_b3 : nullptr;
    if (n_out == 128) {
      w3.a = direct(dy, 128, 128); w3.n_b = 1; w3.b[0] = direct(b->a2, 128, 128); w3.b_act = code;
      w3.out = b->d_w3; w3.ld_out = 128; w3.colsum = cs;
    } else {   // narrow head: the 128-wide activation on the M side, transposed store
      w3.a = direct(b->a2, 128, 128); w3.a_act = code; w3.n_b = 1; w3.b[0] = direct(dy, n_out, n_out);
      w3.out = b->d_w3; w3.ld_out = 128; w3.transpose_out = 1; w3.colsum = cs; w3.colsum_of_b = 1;
    }
    w2.rows = f->rows;
    w2.a = direct(da2, 128, 128); w2.n_b = 1; w2.b[0] = direct(b->a1, 128, 128); w2.b_act = code;
    w2.out = b->d_w2; w2.ld_out = 128; w2.colsum = b->d_b2;
  }
  const gnnfd_segment &s0 = f->seg[0];
  const bool lead = b->skip_wgrad_l1 != 1 && s0.mode == GNNFD_SEG_DIRECT && s0.width == 128 && s0.ld == 128 && s0.col == 0 &&
                    (f->n_seg > 1 || b->skip_wgrad_l1 == 0) && (reinterpret_cast<uintptr_t>(s0.src) & 15) == 0 && b->d_w1 != nullptr;
  if (lead) {
    w1.rows = f->rows;
    w1.a = direct(da1, 128, 128); w1.n_b = 1; w1.b[0] = s0;
    w1.out = b->d_w1; w1.ld_out = f->k_in; w1.colsum = b->d_b1;
  }
  {
    const gnnfd_wgrad_args *lean[3];
    int n_lean = 0;
    if (wgrad_is_lean(&w3)) lean[n_lean++] = &w3;
    if (wgrad_is_lean(&w2)) lean[n_lean++] = &w2;
    if (lead && wgrad_is_lean(&w1)) lean[n_lean++] = &w1;
    if (n_lean > 0) {
      if ((rc = wgrad_lean_run(lean, n_lean, wgws, wgws_bytes, stream, &jobs[n_jobs])) != GNNFD_OK) return rc;
      for (int j = 0; j < n_lean; ++j) { wgws += jobs[n_jobs].ws_used; wgws_bytes -= jobs[n_jobs].ws_used; ++n_jobs; }
    }
    if (!wgrad_is_lean(&w3) && (rc = wgrad_deferred(w3)) != GNNFD_OK) return rc;
    if (!wgrad_is_lean(&w2) && (rc = wgrad_deferred(w2)) != GNNFD_OK) return rc;
    if (lead && !wgrad_is_lean(&w1) && (rc = wgrad_deferred(w1)) != GNNFD_OK) return rc;
  }
  // ---- the rest of dW1 (assembled segments: gathers, means; or everything when there is no lean leading block),
  //      remaining segment input gradients
  {
    if (b->skip_wgrad_l1 == 0) {
      gnnfd_wgrad_args w{};
      w.rows = f->rows;
      w.a = direct(da1, 128, 128);
      if (lead) {
        w.n_b = f->n_seg - 1;
        for (int s = 1; s < f->n_seg; ++s) w.b[s - 1] = f->seg[s];
        w.out = b->d_w1 + 128; w.ld_out = f->k_in;
      } else {
        w.n_b = f->n_seg;
        for (int s = 0; s < f->n_seg; ++s) w.b[s] = f->seg[s];
        w.out = b->d_w1; w.ld_out = f->k_in; w.colsum = b->d_b1;
      }
      if (w.n_b > 0 && (rc = wgrad_deferred(w)) != GNNFD_OK) return rc;
    }
    int col0 = 0;
    for (int s = 0; s < f->n_seg; ++s) {
      if (b->din_out[s] != nullptr && !(chain && s == 0)) {
        gnnfd_mlp_args a;
        linear_args(a, f, da1, 128, 128, f->w1 + col0, 1, f->k_in, f->seg[s].width);
        a.packed = pk + L.pk_w1t[s];
        if (b->din_residual[s] != nullptr) { a.residual = b->din_residual[s]; a.out_sum = b->din_out[s]; }
        else a.out_raw = b->din_out[s];
        if ((rc = gnnfd_mlp_forward(&a, stream)) != GNNFD_OK) return rc;
      }
      col0 += f->seg[s].width;
    }
  }
  // ---- all split-K reductions of this MLP in one launch
  return wgrad_reduce_jobs(jobs, n_jobs, stream);
}
