// Fused MLP block on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   assemble input rows (gather / concat / sum / mean3)            -> bf16|fp16 hi/lo parts, SMEM
//   3 x [ tcgen05.mma kind::f16, M=128 N=128 K=16, fp32 accum in TMEM ]
//   epilogues: tcgen05.ld -> +bias, SiLU/Tanh -> hi/lo split -> SMEM A operand of the next layer
//   final epilogue: +bias -> LayerNorm -> *mul -> coalesced (+residual) stores
//
// One CTA (256 threads) owns a 128-row tile end to end; no intermediate touches HBM.  Operand
// precision is a template: split operands (x = hi + lo, products hi*hi + lo*hi + hi*lo) restore
// ~fp32 accuracy on the bf16/fp16 tensor pipe (single-pass bf16 fails the 1e-3 bar).
//
// Shared-memory operand layout: canonical UMMA K-major SWIZZLE_128B - a k-block is 64 elements
// (128 B per row), rows in groups of 8 (1024 B atoms), 16-byte chunk c of row r stored at chunk
// c ^ (r & 7).  Weights are pre-packed by gnnfd_pack_mlp into exactly this image per (layer,
// k-block, part), so one cp.async.bulk (TMA bulk copy, mbarrier complete_tx) lands a stage.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gnnfd {

constexpr int TC_BM = 128;            // rows per tile == UMMA M
constexpr int TC_H = 128;             // hidden width == UMMA N
constexpr int TC_THREADS = 256;       // 8 warps: (warp & 3) = TMEM lane quarter, (warp >> 2) = column half
constexpr int TC_KB = 64;             // elements per k-block (128 B of 16-bit operands)
constexpr int TC_IMG = TC_BM * 128;   // bytes of one [128 x 64] operand image = 16 KB
constexpr int TC_STAGE = 4 * TC_IMG;  // A_hi, A_lo, W_hi, W_lo
constexpr int TC_NSTAGE = 2;
constexpr int TC_ACT = 4 * TC_IMG;    // hidden activation operand: 2 k-blocks x (hi, lo)
constexpr int TC_OUT_STRIDE = 132;    // fp32 staging of the output tile (aliases the stages)
constexpr int TC_SMEM = TC_NSTAGE * TC_STAGE + TC_ACT + 3 * 3 * TC_BM * 4 + 2 * TC_BM * 8 + 256 + 1024;
constexpr int TC_TMEM_COLS = 128;

// ---------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // LBO = 16 B (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO = 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a/b format, K-major both, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(int fmt /*0 f16, 1 bf16*/, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TC_BM >> 4) << 24);
}

// -------------------------------------------------------------------------- operand conversion
template <bool FP16>
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  if constexpr (FP16) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
  }
}

// byte offset of 16-byte chunk `c` of row `r` inside a [rows x 64] SWIZZLE_128B image
__device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

struct TcParams {
  gnnfd_mlp_args a;
  int kb1;       // k-blocks of layer 1
  int ksteps1;   // K=16 steps in the LAST k-block of layer 1 (1..4)
  int n3;        // UMMA N of layer 3: 128, or 16 for a narrow head
  uint32_t w_block_bytes;   // bytes of one packed 128-row k-block (all parts)
  uint32_t w3_block_bytes;  // bytes of one packed layer-3 k-block
};

// -------------------------------------------------------------------------------------- kernel
template <bool FP16, int NA, int NW>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const gnnfd_mlp_args &a = p.a;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *s_stage = smem;                                  // 2 x 64 KB
  uint8_t *s_act = smem + TC_NSTAGE * TC_STAGE;             // 64 KB
  int32_t *s_idx = (int32_t *)(s_act + TC_ACT);             // [3 seg][3][128]
  float2 *s_stat = (float2 *)(s_idx + 3 * 3 * TC_BM);       // [2 halves][128]
  uint64_t *s_bar = (uint64_t *)(s_stat + 2 * TC_BM);       // full[2], empty[2], acc
  uint32_t *s_tmem = (uint32_t *)(s_bar + 8);
  float *s_out = (float *)s_stage;                          // [128][132] fp32, aliases the stages

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t *bar_full = s_bar, *bar_empty = s_bar + 2, *bar_acc = s_bar + 4;

  if (tid == 0) {
    mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(s_tmem, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  constexpr uint32_t IDESC_H = make_idesc(FP16 ? 0 : 1, TC_H);
  const uint32_t idesc3 = make_idesc(FP16 ? 0 : 1, p.n3);
  const uint8_t *wpack = (const uint8_t *)a.packed;
  const uint8_t *w2pack = wpack + (size_t)p.kb1 * p.w_block_bytes;
  const uint8_t *w3pack = w2pack + (size_t)2 * p.w_block_bytes;

  // pipeline state (identical in every thread)
  uint32_t it = 0;         // stage-use counter
  uint32_t acc_phase = 0;  // parity of bar_acc

  const int64_t n_tiles = (a.rows + TC_BM - 1) / TC_BM;
  // epilogue ownership: this thread reads TMEM lane `erow`, columns [64*ehalf, 64*ehalf + 64)
  const int erow = (warp & 3) * 32 + lane;
  const int ehalf = warp >> 2;
  const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TC_BM;
    // gather indices of this tile -> shared
    for (int s = 0; s < a.n_seg; ++s) {
      const gnnfd_segment &sg = a.seg[s];
      int n_idx = sg.mode == GNNFD_SEG_DIRECT ? 0 : sg.mode == GNNFD_SEG_GATHER ? 1 : sg.mode == GNNFD_SEG_MEAN3 ? 3 : 2;
      for (int q = tid; q < n_idx * TC_BM; q += TC_THREADS) {
        int j = q / TC_BM, r = q % TC_BM;
        int64_t g = row0 + r;
        s_idx[(s * 3 + j) * TC_BM + r] = g < a.rows ? __ldg(sg.idx[j] + g) : 0;
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- layer 1: K streamed
    int seg = 0, seg_k0 = 0;  // segment containing the current k-block
    for (int kb = 0; kb < p.kb1; ++kb, ++it) {
      const int st = it & 1;
      uint8_t *sA = s_stage + st * TC_STAGE;
      uint8_t *sW = sA + 2 * TC_IMG;
      if (it >= 2) mbar_wait(&bar_empty[st], ((it >> 1) - 1) & 1);   // MMAs of the previous use are done
      if (tid == 0) {
        mbar_expect_tx(&bar_full[st], p.w_block_bytes);
        bulk_g2s(sW, wpack + (size_t)kb * p.w_block_bytes, p.w_block_bytes, &bar_full[st]);
      }
      const int k0 = kb * TC_KB;
      while (seg + 1 < a.n_seg && k0 >= seg_k0 + a.seg[seg].width) { seg_k0 += a.seg[seg].width; ++seg; }
      const gnnfd_segment &sg = a.seg[seg];
      const int kloc = k0 - seg_k0;                   // first column of this k-block inside the segment
      const int kvalid = min(TC_KB, sg.width - kloc);  // valid columns in this k-block
      const int ksteps = (kb == p.kb1 - 1) ? p.ksteps1 : 4;
      const int32_t *ix = s_idx + seg * 3 * TC_BM;
      const bool vec = ((sg.ld & 3) == 0) && (((sg.col + kloc) & 3) == 0) && ((kvalid & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(sg.src) & 15) == 0);
      // thread -> (row = tid/16 + 16 j, float4 column f4 = tid % 16)
      const int f4 = tid & 15;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = (tid >> 4) + 16 * j;
        const int64_t g = row0 + r;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < a.rows && f4 * 4 < ksteps * 16) {
          const int64_t i0 = sg.mode == GNNFD_SEG_DIRECT ? g : (int64_t)ix[r];
          const float *b0 = sg.src + i0 * sg.ld + sg.col + kloc + f4 * 4;
          if (vec && f4 * 4 < kvalid) {
            x = ldg_f4(b0);
            if (sg.mode >= GNNFD_SEG_SUM2) {
              float4 y = ldg_f4(sg.src + (int64_t)ix[TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4);
              if (sg.mode == GNNFD_SEG_DIFF2) { x.x -= y.x; x.y -= y.y; x.z -= y.z; x.w -= y.w; }
              else { x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
              if (sg.mode == GNNFD_SEG_MEAN3) {
                float4 z = ldg_f4(sg.src + (int64_t)ix[2 * TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4);
                x.x = (x.x + z.x) / 3.0f; x.y = (x.y + z.y) / 3.0f;
                x.z = (x.z + z.z) / 3.0f; x.w = (x.w + z.w) / 3.0f;
              }
            }
          } else if (!vec) {
            float t4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float t = 0.f;
              if (f4 * 4 + q < kvalid) {
                t = __ldg(b0 + q);
                if (sg.mode >= GNNFD_SEG_SUM2) {
                  float y = __ldg(sg.src + (int64_t)ix[TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4 + q);
                  t = sg.mode == GNNFD_SEG_DIFF2 ? t - y : t + y;
                  if (sg.mode == GNNFD_SEG_MEAN3)
                    t = (t + __ldg(sg.src + (int64_t)ix[2 * TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4 + q)) / 3.0f;
                }
              }
              t4[q] = t;
            }
            x = make_float4(t4[0], t4[1], t4[2], t4[3]);
          }
        }
        v[j] = x;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = (tid >> 4) + 16 * j;
        if (f4 * 4 < ksteps * 16) {
          uint32_t h0, l0, h1, l1;
          split2<FP16>(v[j].x, v[j].y, h0, l0);
          split2<FP16>(v[j].z, v[j].w, h1, l1);
          const uint32_t off = sw128(r, f4 >> 1) + (f4 & 1) * 8;
          *reinterpret_cast<uint2 *>(sA + off) = make_uint2(h0, h1);
          if (NA == 2) *reinterpret_cast<uint2 *>(sA + TC_IMG + off) = make_uint2(l0, l1);
        }
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(&bar_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t aH = smem_u32(sA), aL = aH + TC_IMG, wH = smem_u32(sW), wL = wH + TC_IMG;
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t ko = k * 32;  // 16 elements = 32 bytes along K inside the swizzled row
          umma_f16(tmem_base, make_desc(aH + ko), make_desc(wH + ko), IDESC_H, (kb | k) != 0);
          if (NA == 2) umma_f16(tmem_base, make_desc(aL + ko), make_desc(wH + ko), IDESC_H, 1);
          if (NW == 2) umma_f16(tmem_base, make_desc(aH + ko), make_desc(wL + ko), IDESC_H, 1);
        }
        umma_commit(&bar_empty[st]);
        if (kb == p.kb1 - 1) umma_commit(bar_acc);
      }
    }

    // ---------------------------------------------------------- layers 2 and 3: A = activations
    for (int layer = 2; layer <= 3; ++layer) {
      // epilogue of the previous layer: TMEM -> +bias, act -> hi/lo -> s_act (this thread: one k-block
      // of its row).  Meanwhile the first weight block of this layer streams in.
      const uint8_t *wl = layer == 2 ? w2pack : w3pack;
      const uint32_t wbytes = layer == 2 ? p.w_block_bytes : p.w3_block_bytes;
      {
        const int st = it & 1;
        if (it >= 2) mbar_wait(&bar_empty[st], ((it >> 1) - 1) & 1);
        if (tid == 0) {
          mbar_expect_tx(&bar_full[st], wbytes);
          bulk_g2s(s_stage + st * TC_STAGE + 2 * TC_IMG, wl, wbytes, &bar_full[st]);
        }
        const int st2 = (it + 1) & 1;
        if (it + 1 >= 2) mbar_wait(&bar_empty[st2], (((it + 1) >> 1) - 1) & 1);
        if (tid == 0) {
          mbar_expect_tx(&bar_full[st2], wbytes);
          bulk_g2s(s_stage + st2 * TC_STAGE + 2 * TC_IMG, wl + wbytes, wbytes, &bar_full[st2]);
        }
      }
      mbar_wait(bar_acc, acc_phase);
      acc_phase ^= 1;
      __syncwarp();
      tc_fence_after();
      const float *bias = layer == 2 ? a.b1 : a.b2;
#pragma unroll
      for (int h32 = 0; h32 < 2; ++h32) {
        float acc[32];
        tmem_ld32(t_lane + ehalf * 64 + h32 * 32, acc);
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 4 chunks of 8 values
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = ehalf * 64 + h32 * 32 + c * 8 + q * 2;
            float x0 = acc[c * 8 + q * 2] + (bias ? __ldg(bias + col) : 0.f);
            float x1 = acc[c * 8 + q * 2 + 1] + (bias ? __ldg(bias + col + 1) : 0.f);
            if (a.act == GNNFD_ACT_SILU) { x0 = silu_fast(x0); x1 = silu_fast(x1); }
            else { x0 = tanhf(x0); x1 = tanhf(x1); }
            split2<FP16>(x0, x1, hi[q], lo[q]);
          }
          const uint32_t off = (uint32_t)ehalf * (2 * TC_IMG) + sw128(erow, h32 * 4 + c);
          *reinterpret_cast<uint4 *>(s_act + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (NA == 2) *reinterpret_cast<uint4 *>(s_act + TC_IMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = layer == 2 ? IDESC_H : idesc3;
        const uint32_t part = layer == 2 ? (uint32_t)TC_IMG : p.w3_block_bytes / NW;
        for (int kb = 0; kb < 2; ++kb) {
          const int st = (it + kb) & 1;
          mbar_wait(&bar_full[st], ((it + kb) >> 1) & 1);
          tc_fence_after();
          const uint32_t aH = smem_u32(s_act) + kb * 2 * TC_IMG, aL = aH + TC_IMG;
          const uint32_t wH = smem_u32(s_stage + st * TC_STAGE + 2 * TC_IMG), wL = wH + part;
          for (int k = 0; k < 4; ++k) {
            const uint32_t ko = k * 32;
            umma_f16(tmem_base, make_desc(aH + ko), make_desc(wH + ko), idesc, (kb | k) != 0);
            if (NA == 2) umma_f16(tmem_base, make_desc(aL + ko), make_desc(wH + ko), idesc, 1);
            if (NW == 2) umma_f16(tmem_base, make_desc(aH + ko), make_desc(wL + ko), idesc, 1);
          }
          umma_commit(&bar_empty[st]);
        }
        umma_commit(bar_acc);
      }
      it += 2;
    }

    // ------------------------------------------------------------------------ final epilogue
    mbar_wait(bar_acc, acc_phase);
    acc_phase ^= 1;
    __syncwarp();
    tc_fence_after();
    // both stages are free once bar_acc fired (all MMAs that read them completed): s_out may alias
    if (a.n_out == TC_H) {
      float y[64];
#pragma unroll
      for (int h32 = 0; h32 < 2; ++h32) {
        float acc[32];
        tmem_ld32(t_lane + ehalf * 64 + h32 * 32, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = ehalf * 64 + h32 * 32 + i;
          y[h32 * 32 + i] = acc[i] + (a.b3 ? __ldg(a.b3 + col) : 0.f);
        }
      }
      tc_fence_before();
      if (a.has_ln) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) s += y[i];
        s_stat[ehalf * TC_BM + erow].x = s;
        __syncthreads();
        const float mean = (s_stat[erow].x + s_stat[TC_BM + erow].x) * (1.0f / TC_H);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) { const float d = y[i] - mean; q += d * d; }
        s_stat[ehalf * TC_BM + erow].y = q;
        __syncthreads();
        const float var = (s_stat[erow].y + s_stat[TC_BM + erow].y) * (1.0f / TC_H);
        const float rstd = rsqrtf(var + a.ln_eps);
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const int col = ehalf * 64 + i;
          const float g = a.ln_w ? __ldg(a.ln_w + col) : 1.f, b = a.ln_b ? __ldg(a.ln_b + col) : 0.f;
          y[i] = (y[i] - mean) * rstd * g + b;
        }
      }
#pragma unroll
      for (int i = 0; i < 64; i += 4)
        *reinterpret_cast<float4 *>(s_out + erow * TC_OUT_STRIDE + ehalf * 64 + i) =
            make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      __syncthreads();
      // coalesced copy-out: one warp per row, float4 per lane
      for (int r = warp; r < TC_BM; r += TC_THREADS / 32) {
        const int64_t g = row0 + r;
        if (g >= a.rows) break;
        float4 o = *reinterpret_cast<const float4 *>(s_out + r * TC_OUT_STRIDE + lane * 4);
        const size_t off = (size_t)g * TC_H + lane * 4;
        if (a.mul) { const float4 m = ldg_f4(a.mul + off); o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w; }
        if (a.out_raw) *reinterpret_cast<float4 *>(a.out_raw + off) = o;
        if (a.out_sum) {
          float4 q = ldg_f4(a.residual + off);
          q.x += o.x; q.y += o.y; q.z += o.z; q.w += o.w;
          *reinterpret_cast<float4 *>(a.out_sum + off) = q;
        }
      }
    } else {
      // narrow head: columns [0, n_out) of a 16-wide accumulator; column half 0 threads only
      if (ehalf == 0) {
        uint32_t r16[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r16[0]), "=r"(r16[1]), "=r"(r16[2]), "=r"(r16[3]), "=r"(r16[4]), "=r"(r16[5]), "=r"(r16[6]),
              "=r"(r16[7]), "=r"(r16[8]), "=r"(r16[9]), "=r"(r16[10]), "=r"(r16[11]), "=r"(r16[12]),
              "=r"(r16[13]), "=r"(r16[14]), "=r"(r16[15])
            : "r"(t_lane)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int64_t g = row0 + erow;
        if (g < a.rows) {
#pragma unroll
          for (int o = 0; o < 16; ++o) {
            if (o < a.n_out) {
              float vv = __uint_as_float(r16[o]) + (a.b3 ? __ldg(a.b3 + o) : 0.f);
              const size_t off = (size_t)g * a.n_out + o;
              if (a.mul) vv *= __ldg(a.mul + off);
              if (a.out_raw) a.out_raw[off] = vv;
              if (a.out_sum) a.out_sum[off] = __ldg(a.residual + off) + vv;
            }
          }
        }
      }
      tc_fence_before();
    }
    fence_proxy_async();
    __syncthreads();   // s_out / s_idx / TMEM accumulator are reused by the next tile
    tc_fence_after();
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ------------------------------------------------------------------------------ weight packing
// image element (n, k) of one part: 16-bit value at sw128(n, (k % 64) / 8) + (k % 8) * 2
template <bool FP16>
__global__ void pack_weights_kernel(const float *__restrict__ w, int n_rows_w, int K, int n_img_rows,
                                    int n_kblocks, int nw, uint8_t *__restrict__ out, uint32_t block_bytes) {
  // one thread per (k-block, image row, 16-byte chunk)
  const int total = n_kblocks * n_img_rows * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 7, n = (i >> 3) % n_img_rows, kb = (i >> 3) / n_img_rows;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = kb * TC_KB + c * 8 + q * 2;
      const float x0 = (n < n_rows_w && k < K) ? w[(size_t)n * K + k] : 0.f;
      const float x1 = (n < n_rows_w && k + 1 < K) ? w[(size_t)n * K + k + 1] : 0.f;
      split2<FP16>(x0, x1, hi[q], lo[q]);
    }
    uint8_t *blk = out + (size_t)kb * block_bytes;
    const uint32_t off = sw128(n, c);
    *reinterpret_cast<uint4 *>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (nw == 2) *reinterpret_cast<uint4 *>(blk + block_bytes / 2 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct TcMode { bool fp16; int na, nw; };
static bool tc_mode(int precision, TcMode &m) {
  switch (precision) {
    case GNNFD_PREC_BF16X3: m = {false, 2, 2}; return true;
    case GNNFD_PREC_BF16X1: m = {false, 1, 1}; return true;
    case GNNFD_PREC_FP16X2: m = {true, 2, 1}; return true;
    case GNNFD_PREC_FP16X3: m = {true, 2, 2}; return true;
  }
  return false;
}

static int tc_geometry(const gnnfd_mlp_args *a, const TcMode &m, TcParams &p) {
  if (a->hidden != TC_H) return GNNFD_E_UNSUPPORTED;
  const int kpad = ((a->k_in + 15) / 16) * 16;
  p.kb1 = (kpad + TC_KB - 1) / TC_KB;
  p.ksteps1 = (kpad - (p.kb1 - 1) * TC_KB) / 16;
  p.n3 = a->n_out == TC_H ? TC_H : 16;
  if (a->n_out != TC_H && (a->n_out > 16 || a->has_ln)) return GNNFD_E_UNSUPPORTED;
  p.w_block_bytes = (uint32_t)(m.nw * TC_IMG);
  p.w3_block_bytes = (uint32_t)(m.nw * p.n3 * 128);
  return GNNFD_OK;
}

size_t pack_mlp_bytes_tc(int k_in, int hidden, int n_out, int precision) {
  TcMode m;
  if (!tc_mode(precision, m)) return 0;
  gnnfd_mlp_args a{};
  a.k_in = k_in; a.hidden = hidden; a.n_out = n_out;
  TcParams p;
  if (tc_geometry(&a, m, p) != GNNFD_OK) return 0;
  return (size_t)(p.kb1 + 2) * p.w_block_bytes + (size_t)2 * p.w3_block_bytes;
}

int pack_mlp_tc(const gnnfd_mlp_args *a, void *packed_out, cudaStream_t stream) {
  TcMode m;
  if (!tc_mode(a->precision, m)) { set_error("pack_mlp_tc: bad precision"); return GNNFD_E_BADARG; }
  TcParams p;
  if (tc_geometry(a, m, p) != GNNFD_OK) { set_error("pack_mlp_tc: unsupported shape"); return GNNFD_E_UNSUPPORTED; }
  uint8_t *out = (uint8_t *)packed_out;
  uint8_t *o2 = out + (size_t)p.kb1 * p.w_block_bytes;
  uint8_t *o3 = o2 + (size_t)2 * p.w_block_bytes;
#define PACK(FP)                                                                                              \
  do {                                                                                                        \
    pack_weights_kernel<FP><<<64, 256, 0, stream>>>(a->w1, TC_H, a->k_in, TC_H, p.kb1, m.nw, out, p.w_block_bytes); \
    pack_weights_kernel<FP><<<32, 256, 0, stream>>>(a->w2, TC_H, TC_H, TC_H, 2, m.nw, o2, p.w_block_bytes);   \
    pack_weights_kernel<FP><<<32, 256, 0, stream>>>(a->w3, a->n_out, TC_H, p.n3, 2, m.nw, o3, p.w3_block_bytes); \
  } while (0)
  if (m.fp16) PACK(true); else PACK(false);
#undef PACK
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

int mlp_forward_tc(const gnnfd_mlp_args *a, cudaStream_t stream) {
  TcMode m;
  if (!tc_mode(a->precision, m)) { set_error("mlp_forward_tc: bad precision"); return GNNFD_E_BADARG; }
  TcParams p;
  p.a = *a;
  if (tc_geometry(a, m, p) != GNNFD_OK) {
    set_error("mlp_forward_tc: unsupported shape (hidden=%d n_out=%d)", a->hidden, a->n_out);
    return GNNFD_E_UNSUPPORTED;
  }
  if (!a->packed) { set_error("mlp_forward_tc: packed operand buffer is NULL (call gnnfd_pack_mlp)"); return GNNFD_E_BADARG; }
  if (a->n_seg > 1)
    for (int s = 0; s < a->n_seg; ++s)
      if (a->seg[s].width % TC_KB != 0) {
        set_error("mlp_forward_tc: with several segments every width must be a multiple of 64");
        return GNNFD_E_UNSUPPORTED;
      }
  const int64_t n_tiles = (a->rows + TC_BM - 1) / TC_BM;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
#define LAUNCH(FP, NA_, NW_)                                                                              \
  do {                                                                                                    \
    static bool attr = false;                                                                             \
    if (!attr) {                                                                                          \
      GNNFD_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<FP, NA_, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)); \
      attr = true;                                                                                        \
    }                                                                                                     \
    mlp_tc_kernel<FP, NA_, NW_><<<grid, TC_THREADS, TC_SMEM, stream>>>(p);                                \
  } while (0)
  if (!m.fp16 && m.na == 2 && m.nw == 2) LAUNCH(false, 2, 2);
  else if (!m.fp16 && m.na == 1) LAUNCH(false, 1, 1);
  else if (m.fp16 && m.nw == 1) LAUNCH(true, 2, 1);
  else LAUNCH(true, 2, 2);
#undef LAUNCH
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

}  // namespace gnnfd
