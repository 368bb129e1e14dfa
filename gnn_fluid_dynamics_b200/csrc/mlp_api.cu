// C-ABI entry points for the fused MLP block: argument validation + precision dispatch.
#include "common.cuh"

namespace gnnfd {
int mlp_forward_f32(const gnnfd_mlp_args *args, cudaStream_t stream);
int mlp_forward_tc(const gnnfd_mlp_args *args, cudaStream_t stream);
size_t pack_mlp_bytes_tc(int k_in, int hidden, int n_out, int precision);
int pack_mlp_tc(const gnnfd_mlp_args *args, void *packed_out, cudaStream_t stream);
int tc_profile_read(unsigned long long *out16);

int validate_mlp_args(const gnnfd_mlp_args *a) {
  GNNFD_CHECK_ARG(a != nullptr, "null args");
  GNNFD_CHECK_ARG(a->rows >= 0, "negative rows");
  GNNFD_CHECK_ARG(a->hidden == 128, "hidden width must be 128");
  GNNFD_CHECK_ARG(a->n_seg >= 1 && a->n_seg <= 3, "n_seg must be 1..3");
  GNNFD_CHECK_ARG(a->n_out >= 1 && a->n_out <= 128, "n_out out of range");
  GNNFD_CHECK_ARG(a->act == GNNFD_ACT_SILU || a->act == GNNFD_ACT_TANH, "unknown activation");
  int k = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    const gnnfd_segment &sg = a->seg[s];
    GNNFD_CHECK_ARG(sg.width > 0 && sg.ld >= sg.col + sg.width && sg.col >= 0, "bad segment geometry");
    GNNFD_CHECK_ARG(sg.mode >= GNNFD_SEG_DIRECT && sg.mode <= GNNFD_SEG_SUM3S, "bad segment mode");
    if (a->rows > 0) {
      GNNFD_CHECK_ARG(sg.src != nullptr, "null segment source");
      if (sg.mode >= GNNFD_SEG_GATHER) GNNFD_CHECK_ARG(sg.idx[0] != nullptr, "null gather index 0");
      if (sg.mode >= GNNFD_SEG_SUM2) GNNFD_CHECK_ARG(sg.idx[1] != nullptr, "null gather index 1");
      if (sg.mode >= GNNFD_SEG_MEAN3) GNNFD_CHECK_ARG(sg.idx[2] != nullptr, "null gather index 2");
    }
    k += sg.width;
  }
  GNNFD_CHECK_ARG(k == a->k_in, "segment widths do not add up to k_in");
  GNNFD_CHECK_ARG(a->n_layers == 0 || a->n_layers == 1 || a->n_layers == 3, "n_layers must be 0, 1 or 3");
  if (a->n_layers == 1) {
    GNNFD_CHECK_ARG(a->w1 != nullptr, "null weight");
    GNNFD_CHECK_ARG(a->precision != GNNFD_PREC_F32, "n_layers == 1 needs a tensor-core precision");
    GNNFD_CHECK_ARG(a->n_out == 128 && !a->has_ln, "n_layers == 1 needs n_out == 128 and no LayerNorm");
  } else {
    GNNFD_CHECK_ARG(a->w1 && a->w2 && a->w3, "null weight");
  }
  GNNFD_CHECK_ARG(a->mul_mode >= 0 && a->mul_mode <= 2, "bad mul_mode");
  GNNFD_CHECK_ARG(a->peer_shift >= 0 && a->peer_shift < 31, "bad peer_shift");
  if (a->peer_shift > 0) {
    GNNFD_CHECK_ARG(a->precision != GNNFD_PREC_F32, "peer-memory gathers need a tensor-core precision");
    for (int s = 0; s < a->n_seg; ++s)
      if (a->seg[s].mode == GNNFD_SEG_GATHER)
        GNNFD_CHECK_ARG((a->seg[s].width & 63) == 0 && (a->seg[s].ld & 3) == 0 && (a->seg[s].col & 3) == 0,
                        "peer-memory gather segments must be 64-column multiples with 16-byte aligned rows");
      else
        GNNFD_CHECK_ARG(a->seg[s].mode == GNNFD_SEG_DIRECT, "with peer_shift only DIRECT and GATHER segments are supported");
  }
  if (a->bwd_chain) {
    GNNFD_CHECK_ARG(a->n_layers != 1 && a->n_out == 128 && !a->has_ln && a->precision != GNNFD_PREC_F32,
                    "bwd_chain needs the 3-layer tensor-core path, n_out == 128 and no LayerNorm");
    GNNFD_CHECK_ARG(a->hid_mul1 && a->hid_mul2, "bwd_chain needs hid_mul1/hid_mul2");
  }
  if (a->precision == GNNFD_PREC_F32)
    GNNFD_CHECK_ARG(!a->save_a1 && !a->save_a2 && !a->save_rstd && !a->save_xhat && a->mul_mode == 0,
                    "training stashes need a tensor-core precision");
  GNNFD_CHECK_ARG(a->dropout_p >= 0.f && a->dropout_p < 1.f, "dropout_p must be in [0, 1)");
  if (a->dropout_p > 0.f)
    GNNFD_CHECK_ARG(a->precision != GNNFD_PREC_F32 && a->n_layers != 1 && !a->bwd_chain && a->act == GNNFD_ACT_SILU,
                    "dropout needs the 3-layer SiLU MLP at a tensor-core precision");
  GNNFD_CHECK_ARG(!a->save_xhat || a->n_out == 128, "save_xhat needs n_out == 128");
  GNNFD_CHECK_ARG(!a->out_sum || a->residual, "out_sum requires residual");
  GNNFD_CHECK_ARG(a->out_raw || a->out_sum || a->out_split || a->rows == 0, "no output requested");
  if (a->out_split)
    GNNFD_CHECK_ARG((a->precision == GNNFD_PREC_BF16X3 || a->precision == GNNFD_PREC_FP16X3) && a->n_out == 128,
                    "out_split needs a split precision (BF16X3 / FP16X3) and n_out == 128");
  for (int s = 0; s < a->n_seg; ++s)
    if (a->seg[s].split != nullptr) {
      GNNFD_CHECK_ARG(a->seg[s].src_rows > 0, "segment with a split shadow needs src_rows");
      // a source that exists ONLY as its shadow (src aliases split) can be consumed by TMA gathers alone
      if ((const void *)a->seg[s].src == a->seg[s].split)
        GNNFD_CHECK_ARG((a->precision == GNNFD_PREC_BF16X3 || a->precision == GNNFD_PREC_FP16X3) &&
                            a->seg[s].mode == GNNFD_SEG_GATHER && (a->seg[s].width & 63) == 0 && (a->seg[s].col & 63) == 0 &&
                            (a->seg[s].ld & 63) == 0 && a->peer_shift == 0,
                        "a shadow-only segment must be a 64-column-aligned GATHER at a split precision");
    }
  return GNNFD_OK;
}
}  // namespace gnnfd

using namespace gnnfd;

extern "C" int gnnfd_mlp_forward(const gnnfd_mlp_args *args, void *stream) {
  int rc = validate_mlp_args(args);
  if (rc != GNNFD_OK) return rc;
  if (args->rows == 0) return GNNFD_OK;
  switch (args->precision) {
    case GNNFD_PREC_F32:
      return mlp_forward_f32(args, (cudaStream_t)stream);
    case GNNFD_PREC_BF16X3:
    case GNNFD_PREC_BF16X1:
    case GNNFD_PREC_FP16X2:
    case GNNFD_PREC_FP16X3:
      return mlp_forward_tc(args, (cudaStream_t)stream);
    default:
      set_error("gnnfd_mlp_forward: unknown precision %d", args->precision);
      return GNNFD_E_BADARG;
  }
}

extern "C" size_t gnnfd_pack_mlp_bytes(int32_t k_in, int32_t hidden, int32_t n_out, int32_t precision) {
  if (precision == GNNFD_PREC_F32) return 0;
  return pack_mlp_bytes_tc(k_in, hidden, n_out, precision);
}

extern "C" int gnnfd_pack_mlp(const gnnfd_mlp_args *args, void *packed_out, void *stream) {
  int rc = validate_mlp_args(args);
  if (rc != GNNFD_OK) return rc;
  if (args->precision == GNNFD_PREC_F32) return GNNFD_OK;
  GNNFD_CHECK_ARG(packed_out != nullptr, "null pack buffer");
  return pack_mlp_tc(args, packed_out, (cudaStream_t)stream);
}

extern "C" int gnnfd_tc_profile_read(uint64_t *out16) {
  GNNFD_CHECK_ARG(out16 != nullptr, "null output");
  return tc_profile_read((unsigned long long *)out16);
}

extern "C" uint32_t gnnfd_dropout_hash(uint64_t seed, int32_t layer, uint32_t row, uint32_t col) {
  return dropout_hash(dropout_row_hash(dropout_layer_key(seed, layer), row), col);      // host evaluation of common.cuh
}

extern "C" size_t gnnfd_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(gnnfd_mlp_args);
    case 1: return sizeof(gnnfd_wgrad_args);
    case 2: return sizeof(gnnfd_segment);
    case 3: return sizeof(gnnfd_mlp_backward_args);
    default: return 0;
  }
}
