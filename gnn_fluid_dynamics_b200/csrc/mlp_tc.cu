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
constexpr int TC_KB = 64;             // elements per k-block (128 B of 16-bit operands)
constexpr int TC_IMG = TC_BM * 128;   // bytes of one [128 x 64] operand image = 16 KB
// warp roles: 0-7 epilogue (warp & 3 = TMEM lane quarter, warp >> 2 = column half),
//             8-15 producers (gather -> split -> swizzled A stage), 16 = TMA + MMA issuer
constexpr int TC_EPI_WARPS = 8, TC_PROD_WARPS = 8;
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32, TC_PROD_THREADS = TC_PROD_WARPS * 32;
constexpr int TC_MMA_WARP = TC_EPI_WARPS + TC_PROD_WARPS;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;   // 544
constexpr int TC_A_STAGES = 2;        // A ring: {A_hi, A_lo} images per stage
constexpr int TC_W_SLOTS = 4;         // W ring: one 16 KB image (hi or lo part of a k-block) per slot
constexpr int TC_A_BYTES = TC_A_STAGES * 2 * TC_IMG;   // 64 KB
constexpr int TC_W_BYTES = TC_W_SLOTS * TC_IMG;        // 64 KB
constexpr int TC_ACT = 4 * TC_IMG;    // hidden activation operand: 2 k-blocks x (hi, lo); also output staging
constexpr int TC_STG_STRIDE = 36;     // floats per row of a warp's 32x32 output staging block
constexpr int TC_IDX_SLOT = 9 * TC_BM;               // ints: [3 seg][3 idx][128 rows]
constexpr int TC_SMEM = TC_A_BYTES + TC_W_BYTES + TC_ACT + 2 * TC_IDX_SLOT * 4 + 5 * TC_H * 4 +
                        2 * TC_BM * 8 + 256 + 1024;
constexpr int TC_TMEM_COLS = 256;     // two accumulators: tile j uses columns (j & 1) * 128

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

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

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
// Weight units (one 16 KB image each: the hi or lo part of a 64-wide k-block) are consumed in the MMA
// issue order   L1(0); for j: L2(j), L1(j+1), L3(j)   - L1 of the next tile is queued between L2 and L3
// of the current one so the tensor pipe works on it while the epilogue warps turn L2's accumulator into
// L3's operand.  WSeq enumerates that order for the loader cursor.
// Diagnostic cycle counters of CTA 0 (role wait times), read back with gnnfd_tc_profile_read.
__device__ unsigned long long g_tc_prof[16];
#define PROF_WAIT(slot, stmt)                                  \
  do {                                                         \
    const long long t0_ = clock64();                           \
    stmt;                                                      \
    if (blockIdx.x == 0) prof[slot] += clock64() - t0_;        \
  } while (0)

template <int NW>
struct WSeq {
  int T, j, step, kb, part, kb1;
  bool valid;
  __device__ void init(int tiles, int kb1_) {
    T = tiles; kb1 = kb1_; j = -1; step = 1; kb = 0; part = 0; valid = tiles > 0;
  }
  __device__ int blocks() const { return step == 1 ? kb1 : 2; }
  __device__ void advance() {
    if (++part < NW) return;
    part = 0;
    if (++kb < blocks()) return;
    kb = 0;
    // next segment
    if (j < 0) { j = 0; step = 0; return; }
    ++step;
    if (step == 1 && j + 1 >= T) ++step;
    if (step == 3) { step = 0; ++j; if (j >= T) valid = false; }
  }
};

template <bool FP16, int NA, int NW>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const gnnfd_mlp_args &a = p.a;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *s_a = smem;                                     // A ring: 2 x {hi, lo}
  uint8_t *s_w = s_a + TC_A_BYTES;                         // W ring: 4 x 16 KB
  uint8_t *s_act = s_w + TC_W_BYTES;                       // activations / output staging
  int32_t *s_idx = (int32_t *)(s_act + TC_ACT);            // [2 tiles][3 seg][3][128]
  float *s_vec = (float *)(s_idx + 2 * TC_IDX_SLOT);       // b1, b2, b3, ln_w, ln_b
  float2 *s_stat = (float2 *)(s_vec + 5 * TC_H);           // [2 halves][128]
  uint64_t *s_bar = (uint64_t *)(s_stat + 2 * TC_BM);
  uint64_t *a_full = s_bar, *a_empty = s_bar + 2, *w_full = s_bar + 4, *w_empty = s_bar + 8;
  uint64_t *acc_full = s_bar + 12, *acc_free = s_bar + 14, *act_ready = s_bar + 16;
  uint32_t *s_tmem = (uint32_t *)(s_bar + 18);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], TC_PROD_THREADS);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], TC_EPI_THREADS);
    }
    for (int i = 0; i < TC_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(act_ready, TC_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 5 * TC_H; i += TC_THREADS) {
    const int v = i / TC_H, c = i % TC_H;
    const float *src = v == 0 ? a.b1 : v == 1 ? a.b2 : v == 2 ? a.b3 : v == 3 ? a.ln_w : a.ln_b;
    const int n = (v == 2 || v >= 3) ? a.n_out : TC_H;
    s_vec[i] = (src && c < n) ? __ldg(src + c) : (v == 3 ? 1.f : 0.f);
  }
  if (warp == TC_MMA_WARP) tmem_alloc(s_tmem, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int64_t n_tiles = (a.rows + TC_BM - 1) / TC_BM;
  const int T = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA

  if (warp >= TC_EPI_WARPS && warp < TC_MMA_WARP) {
    // =============================================================================== producers
    const int pt = tid - TC_EPI_THREADS;   // 0..255
    const int f4 = pt & 15;                // float4 column inside the 64-wide k-block
    const int rbase = pt >> 4;             // rows rbase + 16 j

    auto stage_idx = [&](int j) {          // async copy of tile j's gather indices into slot j & 1
      const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)j * gridDim.x) * TC_BM;
      int32_t *dst = s_idx + (j & 1) * TC_IDX_SLOT;
      for (int s = 0; s < a.n_seg; ++s) {
        const gnnfd_segment &sg = a.seg[s];
        const int n_idx = sg.mode == GNNFD_SEG_DIRECT ? 0 : sg.mode == GNNFD_SEG_GATHER ? 1
                          : sg.mode == GNNFD_SEG_MEAN3 ? 3 : 2;
        for (int q = pt; q < n_idx * TC_BM; q += TC_PROD_THREADS) {
          const int ji = q / TC_BM, r = q % TC_BM;
          const int64_t g = row0 + r;
          int32_t *d = dst + (s * 3 + ji) * TC_BM + r;
          if (g < a.rows) cp_async4(d, sg.idx[ji] + g); else *d = 0;
        }
      }
    };

    float4 vn[8];
    auto load_block = [&](int j, int kb) {   // issue the global loads of k-block kb of tile j into vn
      const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)j * gridDim.x) * TC_BM;
      int seg = 0, seg_k0 = 0;
      const int k0 = kb * TC_KB;
      while (seg + 1 < a.n_seg && k0 >= seg_k0 + a.seg[seg].width) { seg_k0 += a.seg[seg].width; ++seg; }
      const gnnfd_segment &sg = a.seg[seg];
      const int kloc = k0 - seg_k0;
      const int kvalid = min(TC_KB, sg.width - kloc);
      const int ksteps = (kb == p.kb1 - 1) ? p.ksteps1 : 4;
      const int32_t *ix = s_idx + (j & 1) * TC_IDX_SLOT + seg * 3 * TC_BM;
      const bool vec = ((sg.ld & 3) == 0) && (((sg.col + kloc) & 3) == 0) && ((kvalid & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(sg.src) & 15) == 0);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int r = rbase + 16 * jj;
        const int64_t g = row0 + r;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < a.rows && f4 * 4 < ksteps * 16) {
          const int64_t i0 = sg.mode == GNNFD_SEG_DIRECT ? g : (int64_t)ix[r];
          const float *b0 = sg.src + i0 * sg.ld + sg.col + kloc + f4 * 4;
          if (vec) {
            if (f4 * 4 < kvalid) {
              x = ldg_f4(b0);
              if (sg.mode >= GNNFD_SEG_SUM2) {
                const float4 y = ldg_f4(sg.src + (int64_t)ix[TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4);
                if (sg.mode == GNNFD_SEG_DIFF2) { x.x -= y.x; x.y -= y.y; x.z -= y.z; x.w -= y.w; }
                else { x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
                if (sg.mode == GNNFD_SEG_MEAN3) {
                  const float4 z = ldg_f4(sg.src + (int64_t)ix[2 * TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4);
                  x.x = (x.x + z.x) / 3.0f; x.y = (x.y + z.y) / 3.0f;
                  x.z = (x.z + z.z) / 3.0f; x.w = (x.w + z.w) / 3.0f;
                }
              }
            }
          } else {
            float t4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float t = 0.f;
              if (f4 * 4 + q < kvalid) {
                t = __ldg(b0 + q);
                if (sg.mode >= GNNFD_SEG_SUM2) {
                  const float y = __ldg(sg.src + (int64_t)ix[TC_BM + r] * sg.ld + sg.col + kloc + f4 * 4 + q);
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
        vn[jj] = x;
      }
    };

    if (T > 0) {
      stage_idx(0);
      cp_async_commit_wait_all();
      named_bar_sync(1, TC_PROD_THREADS);
      if (T > 1) stage_idx(1);
      load_block(0, 0);
    }
    uint32_t pa = 0;
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
    for (int j = 0; j < T; ++j) {
      for (int kb = 0; kb < p.kb1; ++kb, ++pa) {
        float4 vc[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) vc[jj] = vn[jj];
        // issue the next block's loads before touching shared memory (keeps HBM requests in flight)
        if (kb + 1 < p.kb1) {
          load_block(j, kb + 1);
        } else if (j + 1 < T) {
          cp_async_commit_wait_all();                 // indices of tile j+1 have landed
          named_bar_sync(1, TC_PROD_THREADS);         // ... for every producer thread; slot j&1 is free
          if (j + 2 < T) stage_idx(j + 2);
          load_block(j + 1, 0);
        }
        const int st = pa & 1;
        if (pa >= 2) PROF_WAIT(0, mbar_wait(&a_empty[st], ((pa >> 1) - 1) & 1));
        uint8_t *sA = s_a + st * 2 * TC_IMG;
        const int ksteps = (kb == p.kb1 - 1) ? p.ksteps1 : 4;
        if (f4 * 4 < ksteps * 16) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int r = rbase + 16 * jj;
            uint32_t h0, l0, h1, l1;
            split2<FP16>(vc[jj].x, vc[jj].y, h0, l0);
            split2<FP16>(vc[jj].z, vc[jj].w, h1, l1);
            const uint32_t off = sw128(r, f4 >> 1) + (f4 & 1) * 8;
            *reinterpret_cast<uint2 *>(sA + off) = make_uint2(h0, h1);
            if (NA == 2) *reinterpret_cast<uint2 *>(sA + TC_IMG + off) = make_uint2(l0, l1);
          }
        }
        fence_proxy_async();
        mbar_arrive(&a_full[st]);
      }
    }
    if (blockIdx.x == 0 && pt == 0) {
      g_tc_prof[12] = clock64() - t_begin;
      g_tc_prof[13] = prof[0];
    }
  } else if (warp == TC_MMA_WARP) {
    // ======================================================================= TMA + MMA issuer
    if (lane == 0 && T > 0) {
      constexpr uint32_t IDESC_H = make_idesc(FP16 ? 0 : 1, TC_H);
      const uint32_t idesc3 = make_idesc(FP16 ? 0 : 1, p.n3);
      const uint8_t *w1p = (const uint8_t *)a.packed;
      const uint8_t *w2p = w1p + (size_t)p.kb1 * p.w_block_bytes;
      const uint8_t *w3p = w2p + (size_t)2 * p.w_block_bytes;
      const uint32_t w3_part = p.w3_block_bytes / NW;
      WSeq<NW> seq;
      seq.init(T, p.kb1);
      uint32_t wl = 0, wc = 0, ca = 0;
      unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_begin = clock64();

      auto prefetch = [&](uint32_t upto) {      // issue weight units [wl, upto)
        while (wl < upto && seq.valid) {
          const int slot = wl & (TC_W_SLOTS - 1);
          if (wl >= TC_W_SLOTS) PROF_WAIT(1, mbar_wait(&w_empty[slot], ((wl / TC_W_SLOTS) - 1) & 1));
          const uint8_t *src;
          uint32_t bytes;
          if (seq.step == 1) { src = w1p + (size_t)seq.kb * p.w_block_bytes + seq.part * TC_IMG; bytes = TC_IMG; }
          else if (seq.step == 0) { src = w2p + (size_t)seq.kb * p.w_block_bytes + seq.part * TC_IMG; bytes = TC_IMG; }
          else { src = w3p + (size_t)seq.kb * p.w3_block_bytes + seq.part * w3_part; bytes = w3_part; }
          mbar_expect_tx(&w_full[slot], bytes);
          bulk_g2s(s_w + slot * TC_IMG, src, bytes, &w_full[slot]);
          seq.advance();
          ++wl;
        }
      };
      // one k-block: A images at a_hi / a_lo, weight parts from the ring; `first` = start of a layer
      auto mma_block = [&](uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, int ksteps, uint32_t idesc, bool first) {
        prefetch(wc + 3);
        int slot = wc & (TC_W_SLOTS - 1);
        PROF_WAIT(2, mbar_wait(&w_full[slot], (wc / TC_W_SLOTS) & 1));
        tc_fence_after();
        uint32_t wb = smem_u32(s_w + slot * TC_IMG);
        for (int k = 0; k < ksteps; ++k) {
          umma_f16(d_tmem, make_desc(a_hi + k * 32), make_desc(wb + k * 32), idesc, !(first && k == 0));
          if (NA == 2) umma_f16(d_tmem, make_desc(a_lo + k * 32), make_desc(wb + k * 32), idesc, 1);
        }
        umma_commit(&w_empty[slot]);
        ++wc;
        if (NW == 2) {
          prefetch(wc + 3);
          slot = wc & (TC_W_SLOTS - 1);
          PROF_WAIT(2, mbar_wait(&w_full[slot], (wc / TC_W_SLOTS) & 1));
          tc_fence_after();
          wb = smem_u32(s_w + slot * TC_IMG);
          for (int k = 0; k < ksteps; ++k)
            umma_f16(d_tmem, make_desc(a_hi + k * 32), make_desc(wb + k * 32), idesc, 1);
          umma_commit(&w_empty[slot]);
          ++wc;
        }
      };
      auto layer1 = [&](int j) {
        const int b = j & 1;
        if (j >= 2) { PROF_WAIT(3, mbar_wait(&acc_free[b], ((j >> 1) - 1) & 1)); tc_fence_after(); }
        const uint32_t d = tmem_base + b * TC_H;
        for (int kb = 0; kb < p.kb1; ++kb, ++ca) {
          const int st = ca & 1;
          PROF_WAIT(4, mbar_wait(&a_full[st], (ca >> 1) & 1));
          tc_fence_after();
          const uint32_t ah = smem_u32(s_a + st * 2 * TC_IMG);
          mma_block(d, ah, ah + TC_IMG, (kb == p.kb1 - 1) ? p.ksteps1 : 4, IDESC_H, kb == 0);
          umma_commit(&a_empty[st]);
        }
        umma_commit(&acc_full[b]);
      };
      auto layer23 = [&](int j, int layer) {
        const int b = j & 1;
        PROF_WAIT(5, mbar_wait(act_ready, layer == 2 ? 0 : 1));
        tc_fence_after();
        const uint32_t d = tmem_base + b * TC_H;
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t ah = smem_u32(s_act) + kb * 2 * TC_IMG;
          mma_block(d, ah, ah + TC_IMG, 4, layer == 2 ? IDESC_H : idesc3, kb == 0);
        }
        umma_commit(&acc_full[b]);
      };

      layer1(0);
      for (int j = 0; j < T; ++j) {
        layer23(j, 2);
        if (j + 1 < T) layer1(j + 1);
        layer23(j, 3);
      }
      if (blockIdx.x == 0) {
        g_tc_prof[0] = clock64() - t_begin;
        for (int i = 1; i < 6; ++i) g_tc_prof[i] = prof[i];
        g_tc_prof[6] = (unsigned long long)T;
      }
    }
    __syncwarp();
  } else {
    // ================================================================================ epilogue
    const int q4 = warp & 3, ehalf = warp >> 2;
    const int erow = q4 * 32 + lane;
    float *stg = (float *)s_act + warp * (32 * TC_STG_STRIDE);   // this warp's 32x32 staging block
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
    for (int j = 0; j < T; ++j) {
      const int b = j & 1;
      const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)j * gridDim.x) * TC_BM;
      const uint32_t t_acc = tmem_base + b * TC_H + ((uint32_t)(q4 * 32) << 16) + ehalf * 64;
      const uint32_t ph0 = 3u * (uint32_t)(j >> 1);
      // ---- hidden layers: accumulator -> +bias, act -> hi/lo -> next layer's A operand in s_act
      for (int layer = 0; layer < 2; ++layer) {
        PROF_WAIT(0, mbar_wait(&acc_full[b], (ph0 + layer) & 1));
        __syncwarp();
        tc_fence_after();
        const float *bias = s_vec + layer * TC_H;
#pragma unroll
        for (int h32 = 0; h32 < 2; ++h32) {
          float acc[32];
          tmem_ld32(t_acc + h32 * 32, acc);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = ehalf * 64 + h32 * 32 + c * 8 + q * 2;
              float x0 = acc[c * 8 + q * 2] + bias[col];
              float x1 = acc[c * 8 + q * 2 + 1] + bias[col + 1];
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
        mbar_arrive(act_ready);
      }
      // ---- final epilogue
      PROF_WAIT(1, mbar_wait(&acc_full[b], (ph0 + 2) & 1));
      __syncwarp();
      tc_fence_after();
      if (a.n_out == TC_H) {
        float mean = 0.f, rstd = 1.f;
        if (a.has_ln) {
          float s = 0.f;
#pragma unroll
          for (int h32 = 0; h32 < 2; ++h32) {
            float acc[32];
            tmem_ld32(t_acc + h32 * 32, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) s += acc[i] + s_vec[2 * TC_H + ehalf * 64 + h32 * 32 + i];
          }
          s_stat[ehalf * TC_BM + erow].x = s;
          named_bar_sync(2, TC_EPI_THREADS);
          mean = (s_stat[erow].x + s_stat[TC_BM + erow].x) * (1.0f / TC_H);
          float qv = 0.f;
#pragma unroll
          for (int h32 = 0; h32 < 2; ++h32) {
            float acc[32];
            tmem_ld32(t_acc + h32 * 32, acc);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = acc[i] + s_vec[2 * TC_H + ehalf * 64 + h32 * 32 + i] - mean;
              qv += d * d;
            }
          }
          s_stat[ehalf * TC_BM + erow].y = qv;
          named_bar_sync(2, TC_EPI_THREADS);
          rstd = rsqrtf((s_stat[erow].y + s_stat[TC_BM + erow].y) * (1.0f / TC_H) + a.ln_eps);
        }
#pragma unroll
        for (int h32 = 0; h32 < 2; ++h32) {
          float acc[32];
          tmem_ld32(t_acc + h32 * 32, acc);
          if (h32 == 1) {   // last TMEM read of this tile: the accumulator may be overwritten
            tc_fence_before();
            mbar_arrive(&acc_free[b]);
          }
          const int cbase = ehalf * 64 + h32 * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o;
            o.x = (acc[i] + s_vec[2 * TC_H + cbase + i] - mean) * rstd * s_vec[3 * TC_H + cbase + i] + s_vec[4 * TC_H + cbase + i];
            o.y = (acc[i + 1] + s_vec[2 * TC_H + cbase + i + 1] - mean) * rstd * s_vec[3 * TC_H + cbase + i + 1] + s_vec[4 * TC_H + cbase + i + 1];
            o.z = (acc[i + 2] + s_vec[2 * TC_H + cbase + i + 2] - mean) * rstd * s_vec[3 * TC_H + cbase + i + 2] + s_vec[4 * TC_H + cbase + i + 2];
            o.w = (acc[i + 3] + s_vec[2 * TC_H + cbase + i + 3] - mean) * rstd * s_vec[3 * TC_H + cbase + i + 3] + s_vec[4 * TC_H + cbase + i + 3];
            *reinterpret_cast<float4 *>(stg + lane * TC_STG_STRIDE + i) = o;
          }
          __syncwarp();
          // coalesced copy-out of the 32x32 block: each instruction covers 4 rows x 128 B
          const int rr = lane >> 3, c4 = lane & 7;
#pragma unroll
          for (int jr = 0; jr < 8; ++jr) {
            const int rl = jr * 4 + rr;
            const int64_t g = row0 + q4 * 32 + rl;
            if (g < a.rows) {
              float4 o = *reinterpret_cast<const float4 *>(stg + rl * TC_STG_STRIDE + c4 * 4);
              const size_t off = (size_t)g * TC_H + cbase + c4 * 4;
              if (a.mul) { const float4 m = ldg_f4(a.mul + off); o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w; }
              if (a.out_raw) *reinterpret_cast<float4 *>(a.out_raw + off) = o;
              if (a.out_sum) {
                float4 r4 = ldg_f4(a.residual + off);
                r4.x += o.x; r4.y += o.y; r4.z += o.z; r4.w += o.w;
                *reinterpret_cast<float4 *>(a.out_sum + off) = r4;
              }
            }
          }
          __syncwarp();
        }
        // staging aliases s_act: the next tile's hidden epilogue writes it only after every epilogue
        // thread is done reading its staging block
        named_bar_sync(2, TC_EPI_THREADS);
      } else {
        if (ehalf == 0) {
          uint32_t r16[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r16[0]), "=r"(r16[1]), "=r"(r16[2]), "=r"(r16[3]), "=r"(r16[4]), "=r"(r16[5]), "=r"(r16[6]),
                "=r"(r16[7]), "=r"(r16[8]), "=r"(r16[9]), "=r"(r16[10]), "=r"(r16[11]), "=r"(r16[12]),
                "=r"(r16[13]), "=r"(r16[14]), "=r"(r16[15])
              : "r"(t_acc)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int64_t g = row0 + erow;
          if (g < a.rows) {
#pragma unroll
            for (int o = 0; o < 16; ++o) {
              if (o < a.n_out) {
                float vv = __uint_as_float(r16[o]) + s_vec[2 * TC_H + o];
                const size_t off = (size_t)g * a.n_out + o;
                if (a.mul) vv *= __ldg(a.mul + off);
                if (a.out_raw) a.out_raw[off] = vv;
                if (a.out_sum) a.out_sum[off] = __ldg(a.residual + off) + vv;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_free[b]);
      }
    }
    if (blockIdx.x == 0 && tid == 0) {
      g_tc_prof[8] = clock64() - t_begin;
      g_tc_prof[9] = prof[0];
      g_tc_prof[10] = prof[1];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, TC_TMEM_COLS);
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

int tc_profile_read(unsigned long long *out16) {
  GNNFD_CUDA(cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(unsigned long long) * 16));
  return GNNFD_OK;
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
