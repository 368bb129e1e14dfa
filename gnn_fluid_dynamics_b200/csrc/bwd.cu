// Backward-pass kernels that are HBM-bound streaming work (no GEMM): LayerNorm backward with its
// parameter-gradient column sums, and the transposes of the forward gathers / segment sums
// ("the scatter becomes a gather and vice versa", BASELINE.json north_star (e)).
//
// Everything is deterministic: row ownership is fixed by the launch geometry and partial sums are added
// in a fixed order; there are no atomics.
#include "common.cuh"

namespace gnnfd {

constexpr int LNB_THREADS = 256, LNB_WARPS = LNB_THREADS / 32;

// One warp per row (lane = float4 column):  dy = rstd * (g w - mean(g w) - xhat mean(g w xhat))
// Per-CTA column sums of g xhat (d ln_w), g (d ln_b) and dy (d b3) -> partial[cta][3][128].
__global__ void __launch_bounds__(LNB_THREADS) ln_backward_kernel(
    const float *__restrict__ g, const float *__restrict__ xhat, const float *__restrict__ rstd,
    const float *__restrict__ ln_w, int64_t rows, float *__restrict__ dy, float *__restrict__ partial) {
  pdl_entry();
  __shared__ float s_part[LNB_WARPS][3][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 w4 = ln_w ? ldg_f4(ln_w + lane * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 s_gx = make_float4(0.f, 0.f, 0.f, 0.f), s_g = s_gx, s_dy = s_gx;
  const int64_t n_warps = (int64_t)gridDim.x * LNB_WARPS;
  // four rows per warp and iteration: eight 512 B loads in flight per warp (one row at a time left the kernel
  // latency-bound at 3.5 TB/s)
  constexpr int U = 4;
  for (int64_t r0 = ((int64_t)blockIdx.x * LNB_WARPS + warp) * U; r0 < rows; r0 += n_warps * U) {
    float4 g4[U], x4[U];
    float rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = r0 + u < rows;
      g4[u] = ok ? ldg_f4(g + (r0 + u) * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      x4[u] = ok ? ldg_f4(xhat + (r0 + u) * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      rs[u] = ok ? __ldg(rstd + r0 + u) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 gw = make_float4(g4[u].x * w4.x, g4[u].y * w4.y, g4[u].z * w4.z, g4[u].w * w4.w);
      float a = (gw.x + gw.y) + (gw.z + gw.w);
      float b = (gw.x * x4[u].x + gw.y * x4[u].y) + (gw.z * x4[u].z + gw.w * x4[u].w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      a *= (1.0f / 128.0f);
      b *= (1.0f / 128.0f);
      float4 d;
      d.x = rs[u] * (gw.x - a - x4[u].x * b); d.y = rs[u] * (gw.y - a - x4[u].y * b);
      d.z = rs[u] * (gw.z - a - x4[u].z * b); d.w = rs[u] * (gw.w - a - x4[u].w * b);
      if (r0 + u < rows) *reinterpret_cast<float4 *>(dy + (r0 + u) * 128 + lane * 4) = d;
      s_gx.x += g4[u].x * x4[u].x; s_gx.y += g4[u].y * x4[u].y; s_gx.z += g4[u].z * x4[u].z; s_gx.w += g4[u].w * x4[u].w;
      s_g.x += g4[u].x; s_g.y += g4[u].y; s_g.z += g4[u].z; s_g.w += g4[u].w;
      s_dy.x += d.x; s_dy.y += d.y; s_dy.z += d.z; s_dy.w += d.w;
    }
  }
  *reinterpret_cast<float4 *>(&s_part[warp][0][lane * 4]) = s_gx;
  *reinterpret_cast<float4 *>(&s_part[warp][1][lane * 4]) = s_g;
  *reinterpret_cast<float4 *>(&s_part[warp][2][lane * 4]) = s_dy;
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 128; i += LNB_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) s += (&s_part[w][0][0])[i];
    partial[(size_t)blockIdx.x * 384 + i] = s;
  }
}

// out[i] = sum over parts of partial[c][i]: one warp per output, lanes stride over the parts, fixed shuffle tree
__global__ void __launch_bounds__(256) sum_partials_kernel(const float *__restrict__ partial, int n_parts, int width,
                                                           float *__restrict__ out) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= width) return;
  float s = 0.f;
  for (int c = lane; c < n_parts; c += 32) s += partial[(size_t)c * width + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = s;
}

static int lnb_grid(int64_t rows) {
  const int64_t want = (rows + LNB_WARPS * 8 - 1) / (LNB_WARPS * 8);   // >= 8 rows per warp
  const int64_t cap = (int64_t)num_sms() * 2;                          // few partials: the ordered sum stays short
  return (int)(want < 1 ? 1 : want > cap ? cap : want);
}

size_t ln_backward_ws(int64_t rows) { return (size_t)lnb_grid(rows) * 384 * 4 + 256; }

int ln_backward_launch(const float *g, const float *xhat, const float *rstd, const float *ln_w, int64_t rows,
                       float *dy, float *sums, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int grid = lnb_grid(rows);
  if (workspace == nullptr || workspace_bytes < (size_t)grid * 384 * 4) { set_error("ln_backward: workspace too small"); return GNNFD_E_WORKSPACE; }
  launch_pdl(ln_backward_kernel, dim3(grid), dim3(LNB_THREADS), 0, stream, g, xhat, rstd, ln_w, rows, dy, (float *)workspace);
  GNNFD_LAUNCH_CHECK();
  launch_pdl(sum_partials_kernel, dim3((384 * 32 + 255) / 256), dim3(256), 0, stream, (const float *)workspace, grid, 384, sums);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

// Three-part segment sum with scale and base:  out[r] = base[r] + scale * sum_{p in row r} part(p)
// part(p) = p < n ? a[p] : p < 2n ? sign_b * b[p - n] : c[p - 2n]   (ascending p: deterministic)
template <int LPR>
__global__ void __launch_bounds__(256) segment_sum3_kernel(
    const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ c, int ld, int col_a,
    int col_b, int col_c, float sign_b, int64_t n_part, const int32_t *__restrict__ offsets,
    const int32_t *__restrict__ perm, int64_t n_rows, float scale, const float *__restrict__ base, int ld_base,
    float *__restrict__ out, int ld_out) {
  pdl_entry();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp * RPW + lane / LPR;
  if (row >= n_rows) return;
  const int beg = offsets[row], end = offsets[row + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = beg; p < end; p += 4) {
    float4 v[4];
    float s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] = 0.f;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p + j < end) {
        const int64_t q = perm[p + j];
        if (q < n_part) { v[j] = ldg_f4(a + q * ld + col_a + sub * 4); s[j] = 1.0f; }
        else if (q < 2 * n_part) { v[j] = ldg_f4(b + (q - n_part) * ld + col_b + sub * 4); s[j] = sign_b; }
        else { v[j] = ldg_f4(c + (q - 2 * n_part) * ld + col_c + sub * 4); s[j] = 1.0f; }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p + j < end) { acc.x += s[j] * v[j].x; acc.y += s[j] * v[j].y; acc.z += s[j] * v[j].z; acc.w += s[j] * v[j].w; }
    }
  }
  acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
  if (base != nullptr) {
    const float4 b4 = ldg_f4(base + row * (int64_t)ld_base + sub * 4);
    acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
  }
  *reinterpret_cast<float4 *>(out + row * (int64_t)ld_out + sub * 4) = acc;
}

// Transpose of the edge->node segment sums:  one warp per destination row k (128 floats)
//   HALVES: dst[k, 0:64] += src[i0[k], 0:64];  dst[k, 64:128] += sign * src[i1[k], 0:64]     (two-hop halves)
//   else  : dst[k, :]    += src[i0[k], :] + sign * src[i1[k], :]                                (signed edge->cell)
template <bool HALVES>
__global__ void __launch_bounds__(256) gather_pair_add_kernel(float *dst, const float *base, const float *__restrict__ src,
                                                              int ld_src, const int32_t *__restrict__ i0,
                                                              const int32_t *__restrict__ i1, float sign, int64_t rows) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (k >= rows) return;
  float4 d = base ? *reinterpret_cast<const float4 *>(base + k * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (HALVES) {
    const bool second = lane >= 16;
    const int64_t r = second ? __ldg(i1 + k) : __ldg(i0 + k);
    const float4 v = ldg_f4(src + r * ld_src + (lane & 15) * 4);
    const float s = second ? sign : 1.0f;
    d.x += s * v.x; d.y += s * v.y; d.z += s * v.z; d.w += s * v.w;
  } else {
    const float4 u = ldg_f4(src + (int64_t)__ldg(i0 + k) * ld_src + lane * 4);
    const float4 v = ldg_f4(src + (int64_t)__ldg(i1 + k) * ld_src + lane * 4);
    d.x += u.x + sign * v.x; d.y += u.y + sign * v.y; d.z += u.z + sign * v.z; d.w += u.w + sign * v.w;
  }
  *reinterpret_cast<float4 *>(dst + k * 128 + lane * 4) = d;
}

// dst[k, col:col+width] += scale * src[idx[k], 0:width]  - the transpose of one half of any edge->node segment sum
// (generic backward of gnnfd_segment_sum for arbitrary column windows); one warp per row, float4 lanes
__global__ void __launch_bounds__(256) gather_cols_add_kernel(float *__restrict__ dst, int ld_dst, int col, int width,
                                                              const float *__restrict__ src, int ld_src,
                                                              const int32_t *__restrict__ idx, float scale, int64_t rows) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (k >= rows) return;
  const int64_t r = __ldg(idx + k);
  for (int c = lane * 4; c < width; c += 128) {
    float4 d = *reinterpret_cast<const float4 *>(dst + k * ld_dst + col + c);
    const float4 v = ldg_f4(src + r * ld_src + c);
    d.x += scale * v.x; d.y += scale * v.y; d.z += scale * v.z; d.w += scale * v.w;
    *reinterpret_cast<float4 *>(dst + k * ld_dst + col + c) = d;
  }
}

}  // namespace gnnfd

using namespace gnnfd;

extern "C" size_t gnnfd_ln_backward_workspace_bytes(int64_t rows) { return ln_backward_ws(rows); }

extern "C" int gnnfd_ln_backward(const float *g, const float *xhat, const float *rstd, const float *ln_w,
                                 int64_t rows, float *dy, float *sums, void *workspace, size_t workspace_bytes,
                                 void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(rows >= 0, "negative rows");
  GNNFD_CHECK_ARG(sums != nullptr, "null sums");
  if (rows == 0) { GNNFD_CUDA(cudaMemsetAsync(sums, 0, 384 * 4, stream)); return GNNFD_OK; }
  GNNFD_CHECK_ARG(g && xhat && rstd && dy, "null pointer");
  return ln_backward_launch(g, xhat, rstd, ln_w, rows, dy, sums, workspace, workspace_bytes, stream);
}

extern "C" int gnnfd_segment_sum3(const float *a, const float *b, const float *c, int32_t ld, int32_t col_a,
                                  int32_t col_b, int32_t col_c, int32_t width, float sign_b, int64_t n_part,
                                  const int32_t *offsets, const int32_t *perm, int64_t n_rows, float scale,
                                  const float *base, int32_t ld_base, float *out, int32_t ld_out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(n_rows >= 0 && n_part >= 0, "negative size");
  if (n_rows == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(a && offsets && out && (perm || n_part == 0), "null pointer");
  if (b == nullptr) b = a;
  if (c == nullptr) c = a;
  GNNFD_CHECK_ARG((ld % 4) == 0 && (ld_out % 4) == 0 && (ld_base % 4) == 0 && (col_a % 4) == 0 && (col_b % 4) == 0 &&
                      (col_c % 4) == 0, "strides/columns must be multiples of 4 floats");
  const int lpr = width / 4;
  GNNFD_CHECK_ARG(width > 0 && (width % 4) == 0 && lpr <= 32 && (32 % lpr) == 0, "width must be 4*2^k <= 128");
  const int rpw = 32 / lpr;
  const int64_t warps = (n_rows + rpw - 1) / rpw;
  const int blocks = (int)((warps * 32 + 255) / 256);
#define LAUNCH(L)                                                                                            \
  launch_pdl(segment_sum3_kernel<L>, dim3(blocks), dim3(256), 0, stream, a, b, c, ld, col_a, col_b, col_c, sign_b, n_part, offsets, \
                                                     perm, n_rows, scale, base, ld_base, out, ld_out)
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

extern "C" int gnnfd_gather_pair_add(float *dst, const float *base, const float *src, int32_t ld_src, const int32_t *i0,
                                     const int32_t *i1, float sign, int32_t halves, int64_t rows, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(rows >= 0, "negative rows");
  if (rows == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(dst && src && i0 && i1, "null pointer");
  GNNFD_CHECK_ARG((ld_src % 4) == 0, "ld_src must be a multiple of 4 floats");
  const int blocks = (int)((rows * 32 + 255) / 256);
  if (halves) launch_pdl(gather_pair_add_kernel<true>, dim3(blocks), dim3(256), 0, stream, dst, base, src, ld_src, i0, i1, sign, rows);
  else launch_pdl(gather_pair_add_kernel<false>, dim3(blocks), dim3(256), 0, stream, dst, base, src, ld_src, i0, i1, sign, rows);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_gather_cols_add(float *dst, int32_t ld_dst, int32_t col, int32_t width, const float *src,
                                     int32_t ld_src, const int32_t *idx, float scale, int64_t rows, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(rows >= 0, "negative rows");
  if (rows == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(dst && src && idx, "null pointer");
  GNNFD_CHECK_ARG(width > 0 && (width % 4) == 0 && (col % 4) == 0 && (ld_dst % 4) == 0 && (ld_src % 4) == 0,
                  "widths / columns / strides must be multiples of 4 floats");
  const int blocks = (int)((rows * 32 + 255) / 256);
  launch_pdl(gather_cols_add_kernel, dim3(blocks), dim3(256), 0, stream, dst, ld_dst, col, width, src, ld_src, idx, scale, rows);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}
