/*
 * gnnfd_b200.h - C ABI of the B200-native message-passing hot path.
 *
 * The reference (aj-dray/gnn-fluid-dynamics) is pure Python: its "operator interface" for this path
 * is a set of library calls made from src/models/*.py.  Each entry point below replaces one group
 * of those call sites (cited per function, paths relative to the reference root).  A maintainer of
 * the reference binds them with ctypes; this repo's own host layer
 * (gnn_fluid_dynamics_b200/_lib.py) does exactly that.
 *
 * Conventions
 *  - extern "C", POD arguments only: device pointers, sizes, enums, a cudaStream_t passed as void*.
 *  - Every function returns 0 on success or a negative GNNFD_E_* code; no C++ exception crosses
 *    the ABI; gnnfd_last_error() returns a static message for the calling thread.
 *  - No allocation inside: outputs and workspaces are caller-owned device memory, sized with the
 *    matching *_workspace_bytes query.  Kernels are stream-ordered, never synchronise the host and
 *    are CUDA-graph capturable.
 *  - All floating-point tensors are fp32 row-major; all index tensors consumed by kernels are
 *    int32 (gnnfd_index_narrow converts the reference's int64 tensors once per mesh).
 */
#ifndef GNNFD_B200_H
#define GNNFD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNNFD_ABI_VERSION 5

enum {
  GNNFD_OK = 0,
  GNNFD_E_BADARG = -1,      /* null pointer, negative size, unsupported width */
  GNNFD_E_UNSUPPORTED = -2, /* combination not implemented by the selected precision */
  GNNFD_E_WORKSPACE = -3,   /* workspace too small */
  GNNFD_E_CUDA = -4,        /* a CUDA runtime call or launch failed */
  GNNFD_E_RANGE = -5        /* an index is out of range (reported through the device flag) */
};

/* arithmetic used for the three GEMMs of an MLP (accumulation, bias, activation, LayerNorm and
 * residual are fp32 in every mode) */
enum {
  GNNFD_PREC_F32 = 0,    /* CUDA-core FFMA, exact fp32 operands */
  GNNFD_PREC_BF16X3 = 1, /* tcgen05 kind::f16, split-bf16 operands: hi*hi + lo*hi + hi*lo */
  GNNFD_PREC_BF16X1 = 2, /* tcgen05 kind::f16, single bf16 pass (fails the 1e-3 parity bar) */
  GNNFD_PREC_FP16X2 = 3, /* tcgen05 kind::f16, fp16 weights, split-fp16 activations */
  GNNFD_PREC_FP16X3 = 4  /* tcgen05 kind::f16, split-fp16 both operands */
};

enum { GNNFD_ACT_SILU = 0, GNNFD_ACT_TANH = 1 };

/* how one K-segment of an MLP's input row r is assembled */
enum {
  GNNFD_SEG_DIRECT = 0, /* src[r, col:col+width]                                             */
  GNNFD_SEG_GATHER = 1, /* src[idx0[r], col:col+width]                 (x[row], x[col])       */
  GNNFD_SEG_SUM2 = 2,   /* src[idx0[r]] + src[idx1[r]]                 (Conservative.py:230)  */
  GNNFD_SEG_DIFF2 = 3,  /* src[idx0[r]] - src[idx1[r]]                 (Conservative.py:621)  */
  GNNFD_SEG_MEAN3 = 4,  /* ((src[idx0[r]] + src[idx1[r]]) + src[idx2[r]]) / 3.0  (Fvgn.py:317-321) */
  /* ABI v4: signed sum of three gathered rows, (s0 src[r0] + s1 src[r1]) + s2 src[r2] - the direct signed edge->cell
   * aggregation of the Conservative models (Conservative.py:243-254) fused into the node MLP's input assembly: a triangle
   * cell has exactly three faces, so its scatter_add row is a fixed-degree sum.  idx entries encode (row, sign):
   * v >= 0: +src[v];  v < 0: -src[~v];  GNNFD_SUM3S_ZERO: no contribution (a boundary face is a self-loop whose +e and
   * -e entries cancel).  Forward only (inference). */
  GNNFD_SEG_SUM3S = 5
};
#define GNNFD_SUM3S_ZERO INT32_MIN

typedef struct {
  const float *src;      /* row-major source matrix */
  const int32_t *idx[3]; /* per-output-row gather indices (NULL where unused) */
  int32_t ld;            /* row stride of src, in floats */
  int32_t col;           /* first source column */
  int32_t width;         /* number of K columns contributed */
  int32_t mode;          /* GNNFD_SEG_* */
  /* Optional (GATHER segments, split tensor-core precisions): a 16-bit "split shadow" of src written by the
   * gnnfd_mlp_forward call that produced src (its out_split): row r = [hi parts of src[r, 0:ld] | lo parts], 2 * ld
   * 16-bit values (bf16 for BF16X3, fp16 for FP16X3), x = hi + lo.  With it the gathered k-blocks of layer 1 are staged
   * by TMA (cp.async.bulk.tensor ... tile::gather4) straight into the UMMA shared-memory image: no load / convert /
   * store work in the SM.  src_rows = number of rows of src (bounds of the TMA tensor map). */
  const void *split;
  int64_t src_rows;
} gnnfd_segment;

/*
 * One fused "assemble input row -> Linear -> act -> Linear -> act -> Linear -> [LayerNorm] ->
 * [* mul] -> [+ residual]" pass over `rows` rows.
 *
 * Replaces, per call: torch.cat + x[row]/x[col] gathers + nn.Sequential(Linear,SiLU,Linear,SiLU,
 * Linear)+LayerNorm + residual add of
 *   Face_Block.forward  src/models/Fvgn.py:292-296, src/models/Mgn.py:234-238,
 *                       src/models/Conservative.py:228-234
 *   Cell_Block.forward  src/models/Fvgn.py:316-323, src/models/Mgn.py:258-265 (MLP part),
 *                       src/models/Conservative.py:251-252
 *   Encoder / Decoder   src/models/Fvgn.py:263-266, 332-333; src/models/Mgn.py:205-208, 274-275
 *   build_mlp           src/models/Model.py:12-40;  build_mlp_antisym src/models/Conservative.py:31-43
 *   residuals           src/models/Fvgn.py:281-282, src/models/Mgn.py:223-224
 */
typedef struct {
  int64_t rows;
  int32_t n_seg;
  gnnfd_segment seg[3];
  int32_t k_in;   /* sum of segment widths == W1.shape[1] */
  int32_t hidden; /* 128 */
  int32_t n_out;  /* 128 (latent) or 1..16 (decoder head) */
  /* fp32 parameters in PyTorch layout: w1[hidden,k_in] w2[hidden,hidden] w3[n_out,hidden];
   * biases and LayerNorm affine may be NULL (bias-free antisym MLP; decoder without LN) */
  const float *w1, *b1, *w2, *b2, *w3, *b3, *ln_w, *ln_b;
  int32_t has_ln; /* LayerNorm(n_out) with eps ln_eps; affine iff ln_w != NULL */
  float ln_eps;
  int32_t act;    /* GNNFD_ACT_* */
  const float *mul;      /* optional [rows,n_out]: out *= mul (Conservative.py:233) */
  const float *residual; /* optional [rows,n_out] */
  float *out_raw;        /* optional [rows,n_out]: MLP(+LN)(*mul) output */
  float *out_sum;        /* optional [rows,n_out]: residual + output */
  /* tensor-core precisions only: operand pack produced by gnnfd_pack_mlp (NULL for F32) */
  const void *packed;
  int32_t precision; /* GNNFD_PREC_* */
  /* --- training support (tensor-core precisions; all optional, zero = inference behaviour) --------
   * n_layers 0|3: the 3-Linear MLP above.  n_layers 1: a single Linear out = In W1^T (+ b1), n_out ==
   * hidden == 128 rows of w1, no activation / LayerNorm - the building block of the backward dgrad chain
   * (dH = dA W, with w1 = W^T).
   * mul_mode: how `mul` is applied: 0 out *= mul, 1 out *= silu'(mul), 2 out *= tanh'(mul)  (mul then
   * holds the saved pre-activation, i.e. dA = dH . act'(A)).
   * save_a1/save_a2 [rows, hidden]: pre-activations (bias included) of the two hidden layers;
   * save_rstd [rows]: LayerNorm 1/sqrt(var + eps) - what the backward needs (autograd stash). */
  int32_t n_layers;
  int32_t mul_mode;
  float *save_a1, *save_a2, *save_rstd;
  float *save_xhat; /* [rows, n_out] normalised rows before the LayerNorm affine (n_out == 128 only) */
  /* n_layers == 1 only: how w1 is addressed - element (output feature n, input feature k) is
   * w1[n * w1_ld_n + k * w1_ld_k]; both 0 = the PyTorch layout [n_out, k_in] (ld_n = k_in, ld_k = 1).
   * dgrad uses a forward weight W[out, in] transposed in place: w1 = W + col0, ld_n = 1, ld_k = in.
   * w1_rows: valid output features (rows >= w1_rows of the 128 are zero); 0 = 128. */
  int32_t w1_ld_n, w1_ld_k, w1_rows;
  /* bwd_chain = 1 (n_layers 0|3, no LayerNorm, n_out == 128, biases ignored): the dgrad chain of an MLP as one
   * pass - hidden epilogue l (l = 1, 2) computes v = acc * act'(hid_mul{l}[r, :]) instead of bias + activation,
   * stores v to save_a{l} (dA2, dA1: the wgrad operands) and feeds it to the next Linear.  All three weights
   * are addressed with explicit strides (element (n, k) at w[n * ld_n + k * ld_k]), w3_rows = valid outputs. */
  int32_t bwd_chain;
  const float *hid_mul1, *hid_mul2;
  int32_t w2_ld_n, w2_ld_k, w3_ld_n, w3_ld_k, w3_rows;
  /* Peer-memory gather (domain-decomposed mesh, one process per GPU): with peer_shift > 0 every GATHER segment
   * reads row (i & ((1 << peer_shift) - 1)) of the matrix at peer_base[i >> peer_shift] - the latent rows of
   * ghost cells are loaded straight from the owning GPU's HBM over NVLink (P2P mapped memory) by the kernel
   * that consumes them, so there is no halo pack / send / receive step (gnn_fluid_dynamics_b200/dist.py).
   * All peer matrices share the segment's ld / col.  Tensor-core precisions only. */
  const float *peer_base[8];
  int32_t peer_shift;
  /* out_split (optional, n_out == 128, split precisions): [rows, 256] 16-bit split shadow (hi | lo, see
   * gnnfd_segment.split) of out_raw (split_of_sum = 0) or of out_sum (split_of_sum = 1) - what the NEXT block's
   * gathers consume.  out_raw itself may then be NULL: the fp32 copy is not needed by anybody in inference.
   * residual == out_sum (the residual stream updated IN PLACE) is allowed and is the fast path: the add is done by
   * the TMA store (cp.reduce.async.bulk.tensor .add) in L2 and the residual is never loaded by the SM. */
  void *out_split;
  int32_t split_of_sum;
  /* ABI v5: training-mode dropout of the two hidden activations (config.training.dropout_rate > 0 puts a Dropout after
   * each SiLU, src/models/Model.py:29-33).  With 0 < dropout_p < 1 hidden unit (row r, column n) of hidden layer l (0, 1)
   * is dropped when gnnfd_dropout_hash(dropout_seed, l, r, n) < dropout_p * 2^32 (a counter-based hash, restated in
   * tests/test_gpu_dropout.py; no generator state on the device).  The kernel replaces a dropped unit's PRE-activation
   * by GNNFD_DROPPED before the stash and the activation: SiLU(-1e30) = -0 and SiLU'(-1e30) = -0 in the kernels' own
   * formulas, so save_a1 / save_a2 carry the mask and gnnfd_mlp_backward needs no mask input.  The 1 / (1 - p) rescale
   * of the kept units is the CALLER's: pass w2 and w3 already divided by (1 - p) - in the forward, the pack and the
   * backward - and scale d_w2 / d_w3 by 1 / (1 - p) (the chain rule of that substitution).  SiLU MLPs at tensor-core
   * precisions only; 0 = no dropout (inference, and every shipped config). */
  float dropout_p;
  uint64_t dropout_seed;
  /* ABI v5: static_operands = 1 promises that the operand pack (`packed`), the bias / LayerNorm vectors and the gather
   * index arrays are NOT written by any kernel that can still be in flight when this one starts (e.g. they were produced
   * before the CUDA graph this launch is captured in).  With programmatic dependent launch (gnnfd_set_launch_overlap) the
   * kernel then streams its weights, loads those vectors and stages the first tiles' indices BEFORE griddepcontrol.wait,
   * i.e. under the previous kernel's tail; everything the previous kernels produce (segment sources, residual, mul) is
   * still read after the wait.  Results are identical; forward / inference launches only (ignored with bwd_chain). */
  int32_t static_operands;
} gnnfd_mlp_args;
#define GNNFD_DROPPED (-1e30f)

int gnnfd_abi_version(void);
const char *gnnfd_last_error(void);

/* The dropout mask function of gnnfd_mlp_args.dropout_p, evaluated on the HOST (no GPU needed): hidden unit
 * (row, col) of hidden layer `layer` (0 | 1) is dropped iff gnnfd_dropout_hash(seed, layer, row, col) < p * 2^32.
 * Lets a binding (and tests/test_abi.py) pin its own restatement of the mask. */
uint32_t gnnfd_dropout_hash(uint64_t seed, int32_t layer, uint32_t row, uint32_t col);

/* Launch policy of every kernel of this library: on != 0 launches them with programmatic stream serialisation
 * (programmatic dependent launch): the next kernel's CTAs become resident as SMs drain and run their on-chip prologue
 * under the previous kernel's tail; every kernel executes griddepcontrol.wait before its first global access, so the
 * results are the stream-ordered ones.  Measured: -4 % on a launch-bound 2k-cell rollout step, -0.8 % on the 242k-face
 * training step, +1.5 % on 200k-cell inference steps - the host side switches it per call site.  Returns the previous
 * setting.  GNNFD_PDL=0/1 in the environment overrides it for the whole process.  Replaces nothing in the reference
 * (its kernels are library launches, src/train.py:253-256). */
int gnnfd_set_launch_overlap(int32_t on);

/* L2 eviction-priority hints of the streaming kernels (bit mask: 1 = training stash stored evict-first, 2 = residual-stream
 * outputs stored evict-first, 4 = contiguous operand rows loaded evict-first, 8 = gathered rows loaded evict-last,
 * 16 = raw outputs stored evict-first; negative = the library's default).  The hint is an operand of the same
 * instructions, so results are bit-identical under every mask.  Returns the mask in force before the call.
 * GNNFD_L2_HINTS=<mask> in the environment overrides it for the whole process.  Replaces nothing in the reference. */
int gnnfd_set_l2_hints(int32_t mask);

/* int64 -> int32 index conversion with range check [0, limit); *err_flag (device int32, caller
 * zeroes it) is set to 1 if any index is out of range.  Replaces nothing in the reference: it is
 * the one-off narrowing of edge_index / face tensors (src/datasets/DataSet.py:212-213 uses long). */
int gnnfd_index_narrow(const int64_t *src, int32_t *dst, int64_t n, int64_t limit,
                       int32_t *err_flag, void *stream);

/* Receiver-sorted CSR of an index vector: perm = stable argsort(index), offsets[r] = #{index < r}.
 * Integer-exact against torch.sort(index, stable=True) + bincount.
 * It is what lets the deterministic segment sums replace torch_scatter.scatter_add's atomics
 * (src/models/Fvgn.py:307-314, src/models/Mgn.py:249-256, src/models/Conservative.py:244-249). */
size_t gnnfd_csr_workspace_bytes(int64_t n, int64_t n_rows);
int gnnfd_csr_build(const int32_t *index, int64_t n, int64_t n_rows, int32_t *offsets /*n_rows+1*/,
                    int32_t *perm /*n*/, void *workspace, size_t workspace_bytes, void *stream);

/* Deterministic segment sum over CSR rows, replacing scatter_add(cat[A;B], cat[i0;i1], dim_size):
 *   out[r, :] = sum over p in perm[offsets[r]:offsets[r+1]] (ascending position p, i.e. the CPU
 *   scatter_add order) of   p < n_half ?  a[p, col_a:col_a+width]
 *                                      :  sign_b * b[p - n_half, col_b:col_b+width]
 * two-hop halves (Fvgn.py:312-314): a=b=e, col_a=0, col_b=H/2, width=H/2, sign_b=+1
 * signed edge->cell (Conservative.py:248-249): a=b=e, cols 0, width=H, sign_b=-1
 * Vertex_Block (VertPot.py:219-221): a=b=e, cols 0, width=H, sign_b=+1 */
int gnnfd_segment_sum(const float *a, const float *b, int32_t ld_a, int32_t ld_b, int32_t col_a,
                      int32_t col_b, int32_t width, float sign_b, int64_t n_half,
                      const int32_t *offsets, const int32_t *perm, int64_t n_rows, float *out,
                      int32_t ld_out, void *stream);

int gnnfd_mlp_forward(const gnnfd_mlp_args *args, void *stream);

/* operand pack for the tensor-core precisions (bf16/fp16 hi+lo parts in the UMMA shared-memory
 * layout, biases and LayerNorm affine appended) */
size_t gnnfd_pack_mlp_bytes(int32_t k_in, int32_t hidden, int32_t n_out, int32_t precision);
int gnnfd_pack_mlp(const gnnfd_mlp_args *args, void *packed_out, void *stream);

/* ------------------------------------------------------------------------------------ training
 * Backward kernels (BASELINE.json north_star (e)).  The reference gets its backward from autograd over
 * the same call sites (src/train.py:256 `losses["total_log_loss"].backward()`); these entry points are
 * what a torch.autograd.Function around gnnfd_mlp_forward calls (gnn_fluid_dynamics_b200/training.py).
 *
 * dgrad of one Linear is gnnfd_mlp_forward with n_layers = 1 and the forward weight addressed
 * transposed (w1_ld_n / w1_ld_k), its activation derivative fused through mul / mul_mode. */

/* LayerNorm backward (nn.LayerNorm of Model.py:39):
 *   dy = rstd * (g*w - mean(g*w) - xhat * mean(g*w*xhat))       rows x 128
 *   sums[0,:] = sum_r g*xhat (d ln_w)   sums[1,:] = sum_r g (d ln_b)   sums[2,:] = sum_r dy (d bias of Linear 3)
 * ln_w may be NULL (no affine).  Deterministic (fixed row ownership, ordered partial sums). */
size_t gnnfd_ln_backward_workspace_bytes(int64_t rows);
int gnnfd_ln_backward(const float *g, const float *xhat, const float *rstd, const float *ln_w, int64_t rows,
                      float *dy, float *sums /*[3,128]*/, void *workspace, size_t workspace_bytes, void *stream);

/* Weight gradient of one Linear: out[m, n] = sum_r A[r, m] * B[r, n]  (tcgen05, MN-major operands - split-bf16
 * kind::f16 by default, single-pass kind::tf32 optionally - split-K over the grid with an ordered reduction).  A is a DIRECT [rows, <=128] matrix, B is assembled from
 * up to three segments exactly like the forward input (so dW1 of a Face_Block reads e, x[row], x[col] in
 * place); a_act / b_act (0 none, 1 SiLU, 2 tanh) turn a saved pre-activation into the hidden activation on
 * load.  colsum (optional) receives the column sums of A (or of B segment 0 when colsum_of_b) - the bias
 * gradient.  transpose_out stores out[n * ld_out + m]. */
typedef struct {
  int64_t rows;
  gnnfd_segment a;
  int32_t a_act;
  int32_t n_b;
  gnnfd_segment b[3];
  int32_t b_act;
  float *out;
  int32_t ld_out;
  int32_t transpose_out;
  float *colsum;
  int32_t colsum_of_b;
  int32_t precision; /* 0 (default): split-bf16 operands, kind::f16 hi*hi + lo*hi + hi*lo (~1e-5);  1: single-pass TF32 (~3e-4) */
} gnnfd_wgrad_args;
size_t gnnfd_wgrad_workspace_bytes(int64_t rows, int32_t n_cols_padded /* sum of B widths, each rounded up to 64 */);
int gnnfd_wgrad(const gnnfd_wgrad_args *args, void *workspace, size_t workspace_bytes, void *stream);

/* Generic transpose of one index half of gnnfd_segment_sum (its autograd backward for arbitrary column windows):
 *   dst[k, col:col+width] += scale * src[idx[k], 0:width]     (dst [rows, >= col+width], in place) */
int gnnfd_gather_cols_add(float *dst, int32_t ld_dst, int32_t col, int32_t width, const float *src, int32_t ld_src,
                          const int32_t *idx, float scale, int64_t rows, void *stream);

/* The whole backward of one fused MLP in ONE call (the host-side schedule lives in the library so the GPU,
 * not the foreign-call overhead, bounds the training step):
 *   LayerNorm backward -> dW3, dA2 = (dy W3) act'(a2) -> dW2, dA1 = (dA2 W2) act'(a1) -> dW1 ->
 *   input gradient of every segment whose din_out is set: din_out[s] = din_residual[s] + dA1 W1[:, segment s]
 * `fwd` repeats the forward call's arguments (segments, parameters, rows, act, has_ln, precision); g is the
 * gradient w.r.t. out_raw; a1/a2/xhat/rstd are the forward's stashes; packed_bwd comes from
 * gnnfd_pack_mlp_backward (transposed-weight operand packs; rebuild when the parameters change).
 * Gradient outputs that are NULL are skipped (d_w1..3 are required).  Segment input gradients are
 * [rows, 128] matrices (columns >= the segment width are zero); gather / mean segments are scattered by
 * the caller with gnnfd_segment_sum3. */
typedef struct {
  gnnfd_mlp_args fwd;
  const float *g;
  const float *a1, *a2, *xhat, *rstd;
  const void *packed_bwd;
  float *d_w1, *d_b1, *d_w2, *d_b2, *d_w3, *d_b3, *d_ln_w, *d_ln_b;
  float *din_out[3];
  const float *din_residual[3];
  void *workspace;
  size_t workspace_bytes;
  /* Optional: da1_out [rows, 128] receives dA1 (the gradient at the first hidden pre-activation) and, with
   * skip_wgrad_l1 = 1, the library leaves dW1 / db1 to the caller; with skip_wgrad_l1 = 2 (ABI v4) it still computes
   * db1 and the leading contiguous 128-column block of dW1 (segment 0 DIRECT: the residual stream) - in the same launch
   * as dW2 / dW3 - and leaves only the assembled (gathered) segments' columns to the caller.  Used when the MLP input gathers rows of a node
   * matrix: by linearity sum_e dA1[e]^T x[row[e]] = (segment-sum of dA1 by row)^T x, so the caller reduces dA1 onto
   * the nodes first and runs the weight-gradient GEMM and the input-gradient Linear over N node rows instead of E
   * gathered edge rows (gnn_fluid_dynamics_b200/training.py). */
  float *da1_out;
  int32_t skip_wgrad_l1;
} gnnfd_mlp_backward_args;
size_t gnnfd_mlp_backward_workspace_bytes(const gnnfd_mlp_args *fwd);
size_t gnnfd_pack_mlp_backward_bytes(const gnnfd_mlp_args *fwd);
int gnnfd_pack_mlp_backward(const gnnfd_mlp_args *fwd, void *packed_out,
                            int32_t chain /* 1: the call will request din_out[0] (fused chain); 0: it will not */,
                            void *stream);
int gnnfd_mlp_backward(const gnnfd_mlp_backward_args *args, void *stream);

/* Transpose of a gather = deterministic segment sum with up to three source parts, a scale and a base:
 *   out[r, :] = base[r, :] + scale * sum over p in perm[offsets[r]:offsets[r+1]] (ascending) of
 *               p < n ? a[p, col_a:+width] : p < 2n ? sign_b * b[p-n, col_b:+width] : c[p-2n, col_c:+width]
 * d x[row], d x[col] of a Face_Block (CSR of cat[row; col]); d vsum of the 3-vertex mean (CSR of
 * cat[vf0; vf1; vf2], scale 1/3).  b, c NULL = a; base NULL = 0. */
int gnnfd_segment_sum3(const float *a, const float *b, const float *c, int32_t ld, int32_t col_a, int32_t col_b,
                       int32_t col_c, int32_t width, float sign_b, int64_t n_part, const int32_t *offsets,
                       const int32_t *perm, int64_t n_rows, float scale, const float *base, int32_t ld_base,
                       float *out, int32_t ld_out, void *stream);

/* Transpose of the edge->vertex / edge->cell segment sums (a scatter_add backward is a gather) over
 * dst[rows, 128], base NULL = 0, base may alias dst:
 *   halves: dst[k, 0:64] = base[k, 0:64] + src[i0[k], 0:64], dst[k, 64:128] = base[k, 64:128] + sign * src[i1[k], 0:64]
 *   else  : dst[k, :] = base[k, :] + src[i0[k], :] + sign * src[i1[k], :] */
int gnnfd_gather_pair_add(float *dst, const float *base, const float *src, int32_t ld_src, const int32_t *i0,
                          const int32_t *i1, float sign, int32_t halves, int64_t rows, void *stream);

/* Halo pack of the domain-decomposed processor (one exchange of ghost-cell latents per GN_Block, SURVEY.md 8e;
 * the reference has no counterpart - it runs one mesh on one GPU): out[r, 0:width] = src[idx[r], 0:width].
 * The receive side needs no unpack: ghost rows are stored contiguously per owner rank. */
int gnnfd_gather_rows(const float *src, int32_t ld, const int32_t *idx, int64_t n, int32_t width, float *out,
                      void *stream);

/* cudaDeviceEnablePeerAccess(peer_device) for the calling thread's current device (already-enabled is not an
 * error): kernels of this library may then dereference pointers into that device's memory (peer_base). */
int gnnfd_enable_peer_access(int32_t peer_device);

/* sizeof(gnnfd_mlp_args) (which = 0), sizeof(gnnfd_wgrad_args) (1), sizeof(gnnfd_segment) (2),
 * sizeof(gnnfd_mlp_backward_args) (3): lets a
 * foreign-function binding verify its struct mirror against the library it loaded. */
size_t gnnfd_struct_size(int32_t which);

/* Diagnostic: cycle counters recorded by CTA 0 of the last tensor-core MLP launch (synchronises).
 * out16[0..6]  MMA issuer: total, w_empty wait, w_full wait, acc_free wait, a_full wait, act_ready
 *              wait, tiles;  [8..10] epilogue: total, hidden acc_full wait, final acc_full wait;
 *              [12..13] producer: total, a_empty wait. */
int gnnfd_tc_profile_read(uint64_t *out16);

/* ------------------------------------------------------------------------- finite-volume glue
 * The tensor code between the decoder and the loss / the next rollout step, one kernel per operation, fixed-degree
 * gathers, deterministic fp64-partial reductions, no host synchronisation (gnn_fluid_dynamics_b200/fvm_ops.py wraps
 * each pair as a torch.autograd.Function).  `workspace` = gnnfd_glue_workspace_bytes() bytes, zeroed once by the
 * caller (every kernel leaves it reusable). */
size_t gnnfd_glue_workspace_bytes(void);

/* out[f] = BatchNorm1d(1)( area[f] * mean(dt) / ((volume[row[f]] + volume[col[f]]) / 2) )
 * Replaces normalize_face_area, src/utils/normalisation.py:325-344 (called from Integrator.forward
 * src/models/Fvgn.py:226 and FvgnA.loss src/models/Fvgn.py:182).  training != 0: batch statistics (written to
 * stats[0] = mean, stats[1] = 1/sqrt(var + eps)) and the running-stat update applied n_updates times (momentum,
 * unbiased variance, num_batches_tracked += n_updates); training == 0: running statistics, stats unused. */
int gnnfd_face_area_norm(const float *area, const float *volume, const int32_t *row, const int32_t *col,
                         const float *dt, int32_t n_dt, int64_t n_faces, const float *bn_weight, const float *bn_bias,
                         float *running_mean, float *running_var, int64_t *num_batches_tracked, int32_t training,
                         float momentum, float eps, int32_t n_updates, float *out, float *stats, void *workspace,
                         size_t workspace_bytes, void *stream);
/* d_weight = sum_f g[f] * xhat[f], d_bias = sum_f g[f]  (the raw face area carries no gradient) */
int gnnfd_face_area_norm_backward(const float *area, const float *volume, const int32_t *row, const int32_t *col,
                                  const float *dt, int32_t n_dt, int64_t n_faces, const float *stats,
                                  const float *running_mean, const float *running_var, float eps, const float *g,
                                  float *d_weight, float *d_bias, void *workspace, size_t workspace_bytes, void *stream);

/* FVM integrator, src/models/Fvgn.py:221-255 (chain_flux_dot_product: src/utils/maths.py:12-20), and the cell
 * divergence of the face velocity, src/utils/fvm.py:26-37.  edge_out rows (stride ld) = (u, v, p, d0, d1);
 * cf0..2[c] = the three face ids of cell c (f_graph.face); normal = c_graph.normal [N, 3, 2].
 *   acc[c] = -(sum_j u_f (u_f . n_cj) a_f) - (sum_j p_f n_cj a_f) / rho + sum_j (d0, d1)_f        (acc may be NULL)
 *   div[c] = sum_j (u_f . n_cj) a_f                                                              (div may be NULL) */
int gnnfd_fvm_integrate(const float *edge_out, int32_t ld, const float *area, const float *normal, const int32_t *cf0,
                        const int32_t *cf1, const int32_t *cf2, int64_t n_cells, float rho, float *acc, float *div,
                        void *stream);
/* out[j][c][0:width] = t[cf_j[c]][0:width], j = 0..2: the three `x[f_graph.face[j]]` gathers that every integrator /
 * divergence of the model zoo starts with (src/models/Fvgn.py:232-246, Flux.py:186-203, VertPot.py:128-146,
 * src/utils/fvm.py:26-37) as one launch; out is [3, n_cells, width] contiguous, width <= 8.  The backward replaces
 * autograd's sort-based index_put: one thread per face adds the gradients of the (cell, slot) pairs that gathered it - its
 * at most two cells row[f], col[f] (c_graph.edge_index) - in a fixed order, no atomics:
 *   d_t[f][0:width] = sum_{(c, j): cf_j[c] = f} g[j][c][0:width]     (every face row is written; ld_d = row stride of d_t) */
int gnnfd_gather3(const float *t, int32_t ld, int32_t width, const int32_t *cf0, const int32_t *cf1, const int32_t *cf2,
                  int64_t n_cells, float *out, void *stream);
int gnnfd_gather3_backward(const float *g, int32_t width, const int32_t *cf0, const int32_t *cf1, const int32_t *cf2,
                           const int32_t *row, const int32_t *col, int64_t n_cells, int64_t n_faces, float *d_t,
                           int32_t ld_d, void *stream);

/* FluxA's integrator, src/models/Flux.py:166-206, on the signed per-cell face flux of face_flux_to_cell_flux,
 * src/utils/fvm.py:96-156 (forward only: evaluation / rollout).  edge_out rows (stride ld) = (u, v, p, phi, d0, d1);
 * row / col = c_graph.edge_index (owner, neighbour; a boundary face is a self-loop or has neighbour -1);
 * coeff = normalize_vol_dt(...) [E], area = normalize_face_area(...) [E] (both static over a rollout).
 *   s_cj = +1 if c == row[f], -1 if f is interior and c == col[f], else 0      (f = cf_j[c])
 *   acc[c] = 1 * (-(sum_j (u, v)_f (phi_f s_cj) coeff_f) - (sum_j p_f n_cj area_f) / rho) + sum_j (d0, d1)_f   (may be NULL)
 *   cell_flux[c, j] = (edge_out[f, flux_col] * flux_scale + flux_shift) s_cj                                    (may be NULL)
 * Rounded operation by operation in the reference's order: bit-identical to the tensor expression. */
int gnnfd_flux_integrate(const float *edge_out, int32_t ld, int32_t flux_col, const float *coeff, const float *area,
                         const float *normal, const int32_t *cf0, const int32_t *cf1, const int32_t *cf2,
                         const int32_t *row, const int32_t *col, int64_t n_cells, float rho, float *acc, float *cell_flux,
                         float flux_scale, float flux_shift, void *stream);
/* transpose: one thread per face gathers from its (at most two) cells row[f], col[f]; n_cols = 5 with g_acc (all of
 * u, v, p, d0, d1), 2 with g_div only; d_area[f] = gradient w.r.t. the normalised face area. */
int gnnfd_fvm_integrate_backward(const float *edge_out, int32_t ld, const float *area, const float *normal,
                                 const int32_t *cf0, const int32_t *cf1, const int32_t *cf2, const int32_t *row,
                                 const int32_t *col, int64_t n_faces, float rho, const float *g_acc, const float *g_div,
                                 float *d_edge_out, int32_t ld_g, int32_t n_cols, float *d_area, void *stream);

/* out2[0] = mean over unmasked rows and all columns of (a - b)^2, out2[1] = number of elements averaged.
 * Replaces MSE_per_element_torch(output, target, mask), src/utils/loss.py:55-60, without output[mask] indexing. */
int gnnfd_masked_mse(const float *a, int32_t ld_a, const float *b, int32_t ld_b, const uint8_t *mask, int64_t rows,
                     int32_t cols, float *out2, void *workspace, size_t workspace_bytes, void *stream);
int gnnfd_masked_mse_backward(const float *a, int32_t ld_a, const float *b, int32_t ld_b, const uint8_t *mask, int64_t rows,
                              int32_t cols, const float *fwd_out2, const float *g, float *d_a, int32_t ld_d, void *stream);

/* In-place per-column affine (de)normalisation of a [rows, ld] matrix (ABI v4):
 *   forward:  t[r, cols[k]] = (t[r, cols[k]] - shift[k]) / scale[k]        inverse:  t * scale[k] + shift[k]
 * for k < n_spec; cols / shift / scale are DEVICE arrays (no host read of the statistics).  Replaces the per-column
 * tensor expressions of Normalizer.input / output (src/utils/normalisation.py:255-322: z_score, mean / std / max scale,
 * min_max all have this form) - ~4 tiny kernels per normalised column - with one launch per tensor, bit-identical
 * (separately rounded subtract / divide, multiply / add). */
int gnnfd_affine_columns(float *t, int64_t rows, int32_t ld, int32_t n_spec, const int32_t *cols, const float *shift,
                         const float *scale, int32_t inverse, void *stream);

/* Rollout state advance: src/rollout.py:336-340 (velocity update), update_features src/models/Fvgn.py:133-148 /
 * src/models/Mgn.py:139-151 and the next step's input z-scoring src/utils/normalisation.py:255-278, in two kernels.
 *   vel = has_change ? x_raw[:, 0:2] + delta : delta;  x_raw[:, 0:2] = vel;  x_norm[:, 0:2] = (vel - mean) / scale
 *   dv = bc_mask[f] ? bc_value[f, 0:2] : vel[row[f]] - vel[col[f]];  f_raw[:, 0:2] = dv;  f_norm[:, 0:2] = z-scored dv
 * cell_mean_std4 / face_mean_std4: HOST pointers to (mean0, scale0, mean1, scale1); optional outputs may be NULL. */
int gnnfd_state_advance(float *x_raw, int32_t ld_x, const float *delta, int32_t ld_d, int32_t has_change, int64_t n_cells,
                        float *x_norm, int32_t ld_xn, const float *cell_mean_std4, const int32_t *row, const int32_t *col,
                        const uint8_t *bc_mask, const float *bc_value, int32_t ld_bc, int64_t n_faces, float *f_raw,
                        int32_t ld_f, float *f_norm, int32_t ld_fn, const float *face_mean_std4, float *vel_out,
                        void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNFD_B200_H */
