// Deterministic segment sum over receiver-sorted CSR rows (kernel (c) of BASELINE.json: north_star).
//
// Replaces torch_scatter.scatter_add's atomics: every output row is owned by one group of lanes that
// walks the row's contributions in ascending source position - the same order the CPU scatter_add
// uses - so results are reproducible and match the sequential CPU sum bit for bit.
// HBM traffic: each contribution is one coalesced `width`-float read (float4 per lane), each output
// row one coalesced write; the CSR costs 4 B per contribution + 4 B per row.
#include "common.cuh"

namespace gnnfd {

// A row's sum is a chain of three dependent memory latencies (its CSR offsets -> its perm entries -> the source rows),
// and one row per warp left the kernel latency-bound (2.5 TB/s on the [E,128] -> [2N,128] reduction of the training
// step).  So every warp walks its rows in a grid-stride loop as a software pipeline: while the source rows of the current
// row are in flight it loads the perm entries of its next row and the offsets of the one after.  The summation order of a
// row is unchanged (ascending source position), so results stay bit-identical to the sequential CPU scatter_add.
template <int LPR>  // lanes per output row: width = 4 * LPR floats
__global__ void __launch_bounds__(256) segment_sum_kernel(
    const float *__restrict__ a, const float *__restrict__ b, int ld_a, int ld_b, int col_a, int col_b,
    float sign_b, int64_t n_half, const int32_t *__restrict__ offsets, const int32_t *__restrict__ perm,
    int64_t n_rows, float *__restrict__ out, int ld_out) {
  pdl_entry();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;                     // warps in the grid
  const int64_t n_groups = (n_rows + RPW - 1) / RPW;                                 // a warp handles RPW rows at a time
  const float *pa = a + col_a + sub * 4;
  const float *pb = b + col_b + sub * 4;
  auto load_off = [&](int64_t g, int &beg, int &end) {
    const int64_t row = g * RPW + grp;
    beg = end = 0;
    if (g < n_groups && row < n_rows) { beg = __ldg(offsets + row); end = __ldg(offsets + row + 1); }
  };
  auto load_perm = [&](int beg, int end, int (&q)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = (beg + j < end) ? __ldg(perm + beg + j) : -1;
  };
  int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int bc, ec, bn, en, qc[4];
  load_off(g, bc, ec);
  load_off(g + stride, bn, en);
  load_perm(bc, ec, qc);
  for (; g < n_groups; g += stride) {
    float4 v[4];
    float sg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                        // the current row's first four source rows
      sg[j] = 0.f;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qc[j] >= 0) {
        const int64_t q = qc[j];
        if (q < n_half) { v[j] = ldg_f4(pa + q * ld_a); sg[j] = 1.0f; }
        else { v[j] = ldg_f4(pb + (q - n_half) * ld_b); sg[j] = sign_b; }
      }
    }
    int qn[4], bnn, enn;
    load_perm(bn, en, qn);                               // next row's entries, the row after's offsets
    load_off(g + 2 * stride, bnn, enn);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (qc[j] >= 0) {  // sequential order; +-1 scaling is exact
        acc.x += sg[j] * v[j].x; acc.y += sg[j] * v[j].y; acc.z += sg[j] * v[j].z; acc.w += sg[j] * v[j].w;
      }
    }
    for (int p = bc + 4; p < ec; p += 4) {               // rows with more than four contributions
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sg[j] = 0.f;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p + j < ec) {
          const int64_t q = __ldg(perm + p + j);
          if (q < n_half) { v[j] = ldg_f4(pa + q * ld_a); sg[j] = 1.0f; }
          else { v[j] = ldg_f4(pb + (q - n_half) * ld_b); sg[j] = sign_b; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (p + j < ec) {
          acc.x += sg[j] * v[j].x; acc.y += sg[j] * v[j].y; acc.z += sg[j] * v[j].z; acc.w += sg[j] * v[j].w;
        }
      }
    }
    const int64_t row = g * RPW + grp;
    if (row < n_rows) *reinterpret_cast<float4 *>(out + row * (int64_t)ld_out + sub * 4) = acc;
    bc = bn; ec = en; bn = bnn; en = enn;
#pragma unroll
    for (int j = 0; j < 4; ++j) qc[j] = qn[j];
  }
}

}  // namespace gnnfd

using namespace gnnfd;

extern "C" int gnnfd_segment_sum(const float *a, const float *b, int32_t ld_a, int32_t ld_b,
                                 int32_t col_a, int32_t col_b, int32_t width, float sign_b,
                                 int64_t n_half, const int32_t *offsets, const int32_t *perm,
                                 int64_t n_rows, float *out, int32_t ld_out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(n_rows >= 0 && n_half >= 0, "negative size");
  if (n_rows == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(a && b && offsets && out, "null pointer");
  GNNFD_CHECK_ARG(perm || n_half == 0, "null perm");
  GNNFD_CHECK_ARG((ld_a % 4) == 0 && (ld_b % 4) == 0 && (ld_out % 4) == 0 && (col_a % 4) == 0 &&
                      (col_b % 4) == 0,
                  "strides/columns must be multiples of 4 floats");
  GNNFD_CHECK_ARG(sign_b == 1.0f || sign_b == -1.0f, "sign_b must be +-1");
  const int lpr = width / 4;
  GNNFD_CHECK_ARG(width > 0 && (width % 4) == 0 && lpr <= 32 && (32 % lpr) == 0,
                  "width must be 4*2^k <= 128");
  const int rpw = 32 / lpr;
  const int64_t warps = (n_rows + rpw - 1) / rpw;
  int64_t blocks64 = (warps * 32 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 4;             // grid-stride: the four resident blocks per SM (58 registers) walk all rows
  const int blocks = (int)(blocks64 < cap ? blocks64 : cap);
#define LAUNCH(L)                                                                                   \
  launch_pdl(segment_sum_kernel<L>, dim3(blocks), dim3(256), 0, stream, a, b, ld_a, ld_b, col_a, col_b, sign_b, n_half, \
                                                    offsets, perm, n_rows, out, ld_out)
  switch (lpr) {
    case 32: LAUNCH(32); break;
    case 16: LAUNCH(16); break;
    case 8: LAUNCH(8); break;
    case 4: LAUNCH(4); break;
    case 2: LAUNCH(2); break;
    default: LAUNCH(1); break;
  }
#undef LAUNCH
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}
