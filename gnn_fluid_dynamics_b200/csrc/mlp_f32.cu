// Fused MLP block, exact-fp32 CUDA-core variant (GNNFD_PREC_F32).
//
// One CTA owns a tile of 64 rows for the whole chain
//   assemble input row (gather / concat / sum / mean3) -> Linear+act -> Linear+act -> Linear
//   -> LayerNorm -> *mul -> +residual
// so the [rows,384] concat and both hidden activations of the reference (Fvgn.py:294-295,
// Model.py:26-39) never touch HBM: per row the kernel reads its sources once and writes its outputs
// once.  Weights stay in PyTorch [out,in] layout and are streamed through shared memory in 32-wide
// K chunks with cp.async double buffering.  This variant is the exact-arithmetic anchor for the
// tensor-core variants (mlp_tc.cu) and for the backward kernels.
#include "common.cuh"

namespace gnnfd {

constexpr int F32_BM = 64;        // rows per tile
constexpr int F32_THREADS = 256;  // 16 x 16 thread grid, 4 rows x 8 cols per thread
constexpr int F32_H = 128;        // hidden width
constexpr int F32_KC = 32;        // K chunk
constexpr int F32_WS = F32_KC + 4;   // padded weight-chunk row stride (floats)
constexpr int F32_HS = F32_H + 4;    // padded hidden-tile row stride

struct F32Params {
  gnnfd_mlp_args a;
  int kp;        // k_in rounded up to a multiple of F32_KC
  int stride_a;  // kp + 4
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// load W[0:128, k0:k0+32] (row stride K) into sw[128][F32_WS]; zero beyond K
__device__ __forceinline__ void load_w_chunk(const float *__restrict__ W, int K, int k0, float *sw,
                                             bool vec_ok, int n_rows_w) {
  const int tid = threadIdx.x;
  if (vec_ok && k0 + F32_KC <= K) {
#pragma unroll
    for (int it = 0; it < (F32_H * F32_KC / 4) / F32_THREADS; ++it) {
      int idx = tid + it * F32_THREADS;  // 0..1023
      int r = idx >> 3, c4 = idx & 7;
      if (r < n_rows_w)
        cp_async16(sw + r * F32_WS + c4 * 4, W + (size_t)r * K + k0 + c4 * 4);
      else
        *reinterpret_cast<float4 *>(sw + r * F32_WS + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int idx = tid; idx < F32_H * F32_KC; idx += F32_THREADS) {
      int r = idx >> 5, c = idx & 31;
      float v = 0.f;
      if (r < n_rows_w && k0 + c < K) v = __ldg(W + (size_t)r * K + k0 + c);
      sw[r * F32_WS + c] = v;
    }
  }
}

// acc[4][8] += s_in[rows ty*4+i][0:K] . W[cols j*16+tx][0:K]^T
__device__ __forceinline__ void gemm_layer(const float *s_in, int stride_in, int K,
                                           const float *__restrict__ W, float *s_w, float (&acc)[4][8],
                                           int n_rows_w) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool vec_ok = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int n_chunks = (K + F32_KC - 1) / F32_KC;
  load_w_chunk(W, K, 0, s_w, vec_ok, n_rows_w);
  cp_async_commit();
  for (int c = 0; c < n_chunks; ++c) {
    float *cur = s_w + (c & 1) * (F32_H * F32_WS);
    if (c + 1 < n_chunks) {
      load_w_chunk(W, K, (c + 1) * F32_KC, s_w + ((c + 1) & 1) * (F32_H * F32_WS), vec_ok, n_rows_w);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float *a_base = s_in + (ty * 4) * stride_in + c * F32_KC;
    const float *w_base = cur + tx * F32_WS;
#pragma unroll
    for (int kk = 0; kk < F32_KC; kk += 4) {
      float4 a4[4], w4[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a4[i] = *reinterpret_cast<const float4 *>(a_base + i * stride_in + kk);
#pragma unroll
      for (int j = 0; j < 8; ++j) w4[j] = *reinterpret_cast<const float4 *>(w_base + j * 16 * F32_WS + kk);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(a4[i].x, w4[j].x, acc[i][j]);
          acc[i][j] = fmaf(a4[i].y, w4[j].y, acc[i][j]);
          acc[i][j] = fmaf(a4[i].z, w4[j].z, acc[i][j]);
          acc[i][j] = fmaf(a4[i].w, w4[j].w, acc[i][j]);
        }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// hidden epilogue: s_out[row][col] = act(acc + bias[col])
__device__ __forceinline__ void store_hidden(const float (&acc)[4][8], const float *__restrict__ bias,
                                             int act, float *s_out, bool apply_act) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int col = j * 16 + tx;
    float b = bias ? __ldg(bias + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = acc[i][j] + b;
      s_out[(ty * 4 + i) * F32_HS + col] = apply_act ? act_f(v, act) : v;
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// assemble the tile's input rows into s_a[64][stride_a]; zero padding beyond k_in and beyond `rows`
__device__ void assemble_input(const F32Params &p, int64_t row0, float *s_a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const gnnfd_mlp_args &a = p.a;
  for (int r = warp; r < F32_BM; r += F32_THREADS / 32) {
    int64_t g = row0 + r;
    float *dst = s_a + r * p.stride_a;
    int k = 0;
    if (g < a.rows) {
      for (int s = 0; s < a.n_seg; ++s) {
        const gnnfd_segment &sg = a.seg[s];
        const float *b0, *b1 = nullptr, *b2 = nullptr;
        float s0 = 1.f, s1 = 1.f, s2 = 1.f;     // SUM3S signs
        if (sg.mode == GNNFD_SEG_DIRECT) {
          b0 = sg.src + g * sg.ld + sg.col;
        } else if (sg.mode == GNNFD_SEG_SUM3S) {
          int32_t r0, r1, r2;
          sum3s_decode(__ldg(sg.idx[0] + g), r0, s0);
          sum3s_decode(__ldg(sg.idx[1] + g), r1, s1);
          sum3s_decode(__ldg(sg.idx[2] + g), r2, s2);
          b0 = sg.src + (int64_t)r0 * sg.ld + sg.col;
          b1 = sg.src + (int64_t)r1 * sg.ld + sg.col;
          b2 = sg.src + (int64_t)r2 * sg.ld + sg.col;
        } else {
          b0 = sg.src + (int64_t)__ldg(sg.idx[0] + g) * sg.ld + sg.col;
          if (sg.mode >= GNNFD_SEG_SUM2) b1 = sg.src + (int64_t)__ldg(sg.idx[1] + g) * sg.ld + sg.col;
          if (sg.mode == GNNFD_SEG_MEAN3) b2 = sg.src + (int64_t)__ldg(sg.idx[2] + g) * sg.ld + sg.col;
        }
        const bool vec = ((sg.width & 3) == 0) && ((sg.ld & 3) == 0) && ((sg.col & 3) == 0) &&
                         ((k & 3) == 0) && ((reinterpret_cast<uintptr_t>(sg.src) & 15) == 0);
        if (vec) {
          for (int c = lane * 4; c < sg.width; c += 128) {
            float4 v = ldg_f4(b0 + c);
            if (sg.mode == GNNFD_SEG_SUM2) {
              float4 w = ldg_f4(b1 + c);
              v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            } else if (sg.mode == GNNFD_SEG_DIFF2) {
              float4 w = ldg_f4(b1 + c);
              v.x -= w.x; v.y -= w.y; v.z -= w.z; v.w -= w.w;
            } else if (sg.mode == GNNFD_SEG_MEAN3) {
              float4 w = ldg_f4(b1 + c), u = ldg_f4(b2 + c);
              v.x = ((v.x + w.x) + u.x) / 3.0f; v.y = ((v.y + w.y) + u.y) / 3.0f;
              v.z = ((v.z + w.z) + u.z) / 3.0f; v.w = ((v.w + w.w) + u.w) / 3.0f;
            } else if (sg.mode == GNNFD_SEG_SUM3S) {
              float4 w = ldg_f4(b1 + c), u = ldg_f4(b2 + c);
              v.x = (s0 * v.x + s1 * w.x) + s2 * u.x; v.y = (s0 * v.y + s1 * w.y) + s2 * u.y;
              v.z = (s0 * v.z + s1 * w.z) + s2 * u.z; v.w = (s0 * v.w + s1 * w.w) + s2 * u.w;
            }
            *reinterpret_cast<float4 *>(dst + k + c) = v;
          }
        } else {
          for (int c = lane; c < sg.width; c += 32) {
            float v = __ldg(b0 + c);
            if (sg.mode == GNNFD_SEG_SUM2) v += __ldg(b1 + c);
            else if (sg.mode == GNNFD_SEG_DIFF2) v -= __ldg(b1 + c);
            else if (sg.mode == GNNFD_SEG_MEAN3) v = ((v + __ldg(b1 + c)) + __ldg(b2 + c)) / 3.0f;
            else if (sg.mode == GNNFD_SEG_SUM3S) v = (s0 * v + s1 * __ldg(b1 + c)) + s2 * __ldg(b2 + c);
            dst[k + c] = v;
          }
        }
        k += sg.width;
      }
    }
    for (int c = k + lane; c < p.kp; c += 32) dst[c] = 0.f;
  }
}

__global__ void __launch_bounds__(F32_THREADS, 1) mlp_f32_kernel(const F32Params p) {
  pdl_entry();
  extern __shared__ __align__(16) float smem[];
  const gnnfd_mlp_args &a = p.a;
  float *s_a = smem;                                // [64][stride_a]
  float *s_h1 = s_a + F32_BM * p.stride_a;          // [64][132]
  float *s_h2 = s_h1 + F32_BM * F32_HS;             // [64][132]
  float *s_w = s_h2 + F32_BM * F32_HS;              // 2 x [128][36]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.rows + F32_BM - 1) / F32_BM;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * F32_BM;
    assemble_input(p, row0, s_a);
    __syncthreads();

    float acc[4][8];
    zero_acc(acc);
    gemm_layer(s_a, p.stride_a, a.k_in, a.w1, s_w, acc, F32_H);
    store_hidden(acc, a.b1, a.act, s_h1, true);
    __syncthreads();
    zero_acc(acc);
    gemm_layer(s_h1, F32_HS, F32_H, a.w2, s_w, acc, F32_H);
    store_hidden(acc, a.b2, a.act, s_h2, true);
    __syncthreads();

    if (a.n_out == F32_H) {
      zero_acc(acc);
      gemm_layer(s_h2, F32_HS, F32_H, a.w3, s_w, acc, F32_H);
      store_hidden(acc, a.b3, a.act, s_h1, false);  // y3 tile
      __syncthreads();
      // LayerNorm + mul + residual, one warp per row, float4 per lane
      const float4 g4 = (a.has_ln && a.ln_w) ? ldg_f4(a.ln_w + lane * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 be4 = (a.has_ln && a.ln_b) ? ldg_f4(a.ln_b + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = warp; r < F32_BM; r += F32_THREADS / 32) {
        int64_t g = row0 + r;
        if (g >= a.rows) break;
        float4 v = *reinterpret_cast<const float4 *>(s_h1 + r * F32_HS + lane * 4);
        if (a.has_ln) {
          float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / F32_H);
          float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
          float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / F32_H);
          float rstd = 1.0f / sqrtf(var + a.ln_eps);
          v.x = dx * rstd * g4.x + be4.x; v.y = dy * rstd * g4.y + be4.y;
          v.z = dz * rstd * g4.z + be4.z; v.w = dw * rstd * g4.w + be4.w;
        }
        size_t off = (size_t)g * F32_H + lane * 4;
        if (a.mul) {
          float4 m = ldg_f4(a.mul + off);
          v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
        }
        if (a.out_raw) *reinterpret_cast<float4 *>(a.out_raw + off) = v;
        if (a.out_sum) {
          float4 q = ldg_f4(a.residual + off);
          q.x += v.x; q.y += v.y; q.z += v.z; q.w += v.w;
          *reinterpret_cast<float4 *>(a.out_sum + off) = q;
        }
      }
    } else {
      // narrow head (decoder): n_out <= 16, no LayerNorm.  4 threads per row, each a K quarter.
      for (int idx = tid; idx < a.n_out * F32_H; idx += F32_THREADS) s_w[idx] = __ldg(a.w3 + idx);
      __syncthreads();
      const int r = tid >> 2, q = tid & 3;
      const float *hrow = s_h2 + r * F32_HS + q * 32;
      int64_t g = row0 + r;
      for (int o = 0; o < a.n_out; ++o) {
        const float *wrow = s_w + o * F32_H + q * 32;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) s = fmaf(hrow[k], wrow[k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == (o & 3) && g < a.rows) {
          float v = s + (a.b3 ? __ldg(a.b3 + o) : 0.f);
          size_t off = (size_t)g * a.n_out + o;
          if (a.mul) v *= __ldg(a.mul + off);
          if (a.out_raw) a.out_raw[off] = v;
          if (a.out_sum) a.out_sum[off] = __ldg(a.residual + off) + v;
        }
      }
    }
    __syncthreads();
  }
}

int mlp_forward_f32(const gnnfd_mlp_args *args, cudaStream_t stream) {
  F32Params p;
  p.a = *args;
  p.kp = ((args->k_in + F32_KC - 1) / F32_KC) * F32_KC;
  p.stride_a = p.kp + 4;
  if (args->n_out != F32_H && (args->n_out < 1 || args->n_out > 16 || args->has_ln)) {
    set_error("mlp_forward_f32: n_out must be 128, or 1..16 without LayerNorm (got %d)", args->n_out);
    return GNNFD_E_UNSUPPORTED;
  }
  size_t smem = (size_t)(F32_BM * p.stride_a + 2 * F32_BM * F32_HS + 2 * F32_H * F32_WS) * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("mlp_forward_f32: k_in=%d needs %zu B of shared memory (> 227 KB)", args->k_in, smem);
    return GNNFD_E_UNSUPPORTED;
  }
  static bool attr_set[GNNFD_MAX_DEVICES] = {false};
  if (!attr_set[current_device()]) {
    GNNFD_CUDA(cudaFuncSetAttribute(mlp_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[current_device()] = true;
  }
  int64_t n_tiles = (args->rows + F32_BM - 1) / F32_BM;
  int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  launch_pdl(mlp_f32_kernel, dim3(grid), dim3(F32_THREADS), smem, stream, p);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

}  // namespace gnnfd
