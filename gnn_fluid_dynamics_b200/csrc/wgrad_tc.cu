// Weight-gradient GEMM of the fused MLP block on tcgen05 (fp32 accumulation in TMEM), sm_100a.
// Two operand precisions (template MODE):
//   0  split-bf16 (default): every operand staged as hi = bf16(x), lo = bf16(x - hi); kind::f16 products
//      hi*hi + lo*hi + hi*lo (~1e-5 on dW, same scheme as the forward) - 4 bytes of shared memory per element,
//      i.e. exactly the footprint of one TF32 image, and the bf16 pipe runs at twice the TF32 rate;
//   1  single-pass TF32 (operands rounded to nearest; ~3e-4 on dW, up to ~1e-3 for cancellation-heavy sums).
//
//   D[128, n] = sum over rows r of  A[r, 0:128]^T  B[r, 0:n]          (dW = dA^T . H, reduction over E or N rows)
//
// Both operands are row-major [rows, cols] with the REDUCTION dimension as the row, i.e. "MN-major" UMMA
// operands: a stage is a [32 rows x cols] fp32 slab whose shared-memory image is the canonical MN-major
// layout for 32-bit operands, SWIZZLE_128B_BASE32B (the only one tcgen05 accepts for MN-major tf32):
// 32-column (128 B) x 4-row atoms of 512 B, 32-byte chunk c of row r stored at chunk c ^ (r & 3), atoms
// ordered [k-atom][mn-atom] (LBO = 512 B between MN atoms, SBO = n_atoms * 512 B between K atoms).
// The image is written by the producer warps with conflict-free 16-byte stores in the same orientation as
// global memory (no transposition anywhere); one tcgen05.mma M128 x N<=256 x K8 consumes two K atoms.
//
// Producers assemble B on the fly exactly like the forward kernel assembles its input rows (direct /
// gather / sum2 / diff2 / mean3 segments) and can apply the activation to a saved pre-activation
// (H = SiLU(a)), so neither the concatenated MLP input nor the hidden activations are ever materialised.
// Column sums of one operand (the bias gradient) are accumulated by the producers in registers.
//
// Split-K over a persistent grid: CTA i reduces rows [i * rows_per_cta, ...) into partial[i][128][n_pad];
// reduce_partials_kernel adds the partials in CTA order (deterministic, no atomics).
#include "tc_common.cuh"

namespace gnnfd {

constexpr int WG_KR = 32;          // rows per stage = 8 K atoms
constexpr int WG_PROD_WARPS = 16;  // warp w owns stage rows {w, w + 16}
constexpr int WG_THREADS = (WG_PROD_WARPS + 1) * 32;   // + MMA issuer warp
constexpr int WG_MAX_STAGES = 6;
constexpr int WG_MAX_SLOTS = 8;    // gather-index slots prefetched one stage ahead

struct WgPiece {           // piece 0 = A (128 columns); pieces 1.. = consecutive column blocks of B
  const float *src;
  int32_t ld, col, width, mode, act, vec;
  int32_t atom0;           // first MN atom of this piece inside its operand image
  int32_t slot0;           // first gather-index slot
};

struct WgParams {
  WgPiece pc[4];
  const int32_t *slot[WG_MAX_SLOTS];
  int n_pieces, n_slots;
  int64_t rows;
  int n_pad;               // B columns, multiple of 32
  int stages;
  int64_t rows_per_cta;    // multiple of WG_KR
  float *partial;          // [grid][128][n_pad]
  float *colsum;           // [grid][128] or nullptr
  int colsum_piece;
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(512 >> 4) << 16;         // LBO: stride between MN atoms
  d |= (uint64_t)(sbo_bytes >> 4) << 32;   // SBO: stride between K atoms (4 rows each)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
  return d;
}
// MN-major SWIZZLE_128B descriptor for 16-bit operands: 64-column (128 B) x 8-row atoms of 1024 B, 16-byte chunk c of
// row r at c ^ (r & 7); LBO = 1024 B between MN atoms, SBO between K atoms (8 rows each)
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// c = F32, a = b = BF16, both MN-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// c = F32, a = b = TF32, both MN-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {   // round to nearest (the tensor core itself truncates)
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float wg_silu(float v) { return v * rcp_ftz(1.0f + ex2_ftz(v * -1.4426950408889634f)); }
// the same on a pair, with the packed two-wide fp32 operations (half the issue slots; tc_common.cuh)
__device__ __forceinline__ void wg_silu2(float &a, float &b) {
  float e0 = a, e1 = b;
  mul2(e0, e1, -1.4426950408889634f, -1.4426950408889634f);
  e0 = ex2_ftz(e0); e1 = ex2_ftz(e1);
  add2(e0, e1, 1.0f, 1.0f);
  mul2(a, b, rcp_ftz(e0), rcp_ftz(e1));
}

// one (row, piece) item: the lane's float4 of the assembled operand row (zero outside the matrix / the width)
__device__ __forceinline__ float4 wg_load_item(const WgPiece &pc, int lane, bool row_ok, int64_t g, int32_t i0,
                                               int32_t i1, int32_t i2) {
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok || lane * 4 >= pc.width) return t;
  const int64_t r0 = pc.mode == GNNFD_SEG_DIRECT ? g : (int64_t)i0;
  if (pc.vec) {
    const float *b = pc.src + pc.col + lane * 4;
    t = ldg_f4(b + r0 * pc.ld);
    if (pc.mode >= GNNFD_SEG_SUM2) {
      const float4 y = ldg_f4(b + (int64_t)i1 * pc.ld);
      if (pc.mode == GNNFD_SEG_DIFF2) { t.x -= y.x; t.y -= y.y; t.z -= y.z; t.w -= y.w; }
      else { t.x += y.x; t.y += y.y; t.z += y.z; t.w += y.w; }
      if (pc.mode == GNNFD_SEG_MEAN3) {
        const float4 z = ldg_f4(b + (int64_t)i2 * pc.ld);
        constexpr float third = 1.0f / 3.0f;
        t.x = (t.x + z.x) * third; t.y = (t.y + z.y) * third; t.z = (t.z + z.z) * third; t.w = (t.w + z.w) * third;
      }
    }
  } else {   // narrow / unaligned sources (encoder inputs, decoder heads)
    float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane * 4 + q;
      if (c < pc.width) {
        float t0 = __ldg(pc.src + r0 * pc.ld + pc.col + c);
        if (pc.mode >= GNNFD_SEG_SUM2) {
          const float y = __ldg(pc.src + (int64_t)i1 * pc.ld + pc.col + c);
          t0 = pc.mode == GNNFD_SEG_DIFF2 ? t0 - y : t0 + y;
          if (pc.mode == GNNFD_SEG_MEAN3) t0 = (t0 + __ldg(pc.src + (int64_t)i2 * pc.ld + pc.col + c)) * (1.0f / 3.0f);
        }
        e4[q] = t0;
      }
    }
    t = make_float4(e4[0], e4[1], e4[2], e4[3]);
  }
  return t;
}

// NP = pieces (A + NP - 1 column blocks of B); MODE 0 split-bf16 (two 16-bit images per operand), 1 TF32.
// Both modes stage 32 rows x 4 bytes per element: A 16 KB + B n_pad * 128 B per stage.
template <int NP, int MODE>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr bool BF = MODE == 0;
  // MN atoms: TF32 32 columns x 4 rows (512 B), bf16 64 columns x 8 rows (1024 B)
  const int nb_atoms = BF ? p.n_pad >> 6 : p.n_pad >> 5;
  const uint32_t a_img = 32 * 128 * 2;                          // bf16: one [32 rows x 128] image = 8 KB
  const uint32_t b_img = 32 * (uint32_t)p.n_pad * 2;            // bf16: one [32 rows x n_pad] image
  const uint32_t a_bytes = 16 * 1024;
  const uint32_t b_bytes = 32 * (uint32_t)p.n_pad * 4;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t *s_tail = smem + (size_t)p.stages * stage_bytes;
  float *s_cs = (float *)s_tail;                                // [16 warps][128] column-sum scratch
  uint64_t *s_bar = (uint64_t *)(s_tail + WG_PROD_WARPS * 128 * 4);
  uint64_t *full = s_bar, *empty = s_bar + WG_MAX_STAGES, *done = s_bar + 2 * WG_MAX_STAGES;
  uint32_t *s_tmem = (uint32_t *)(s_bar + 2 * WG_MAX_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], WG_PROD_WARPS); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_PROD_WARPS) tmem_alloc(s_tmem, 512);
  pdl_wait();   // everything above is on-chip set-up that may run under the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r_end = min(p.rows, r_begin + p.rows_per_cta);
  const int n_stage_iters = (int)((r_end - r_begin + WG_KR - 1) / WG_KR);

  if (warp < WG_PROD_WARPS) {
    // =============================================================================== producers
    // warp w owns rows {w, w + 16} of every stage; lane l owns float4 column l of each piece
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    // gather indices of a stage, one per lane: lane = slot * 2 + row slot (fetched one stage before use)
    auto load_idx = [&](int it) -> int32_t {
      const int j = lane & 1, q = lane >> 1;
      const int64_t g = r_begin + (int64_t)it * WG_KR + warp + 16 * j;
      return (q < p.n_slots && it < n_stage_iters && g < r_end) ? __ldg(p.slot[q] + g) : 0;
    };
    // per-piece constants hoisted out of the stage loop (pi is a compile-time index after unrolling)
    const float *base[NP];
    int32_t ldp[NP];
    uint32_t off0[NP];
    bool lane_on[NP], fast[NP];
#pragma unroll
    for (int pi = 0; pi < NP; ++pi) {
      const WgPiece &pc = p.pc[pi];
      const int natoms = pi == 0 ? 4 : nb_atoms;
      if (NP < 4) { base[pi] = pc.src + pc.col + lane * 4; ldp[pi] = pc.ld; }   // NP == 4: recomputed on use (register budget)
      lane_on[pi] = lane * 4 < (BF ? ((pc.width + 63) & ~63) : ((pc.width + 31) & ~31));
      fast[pi] = pc.vec && pc.mode <= GNNFD_SEG_GATHER && (pc.width & 31) == 0;   // 16-byte loads, no column tail
      if (BF) {
        // bf16 image: stage row r = warp + 16 j -> K atom (warp >> 3) + 2 j, row warp & 7 inside the atom; the lane's 4
        // columns are 8 bytes: MN atom lane >> 4, 16-byte chunk (lane & 15) >> 1 stored at chunk ^ (r & 7), half lane & 1
        const int na = pi == 0 ? 2 : nb_atoms, a0 = pc.atom0;
        off0[pi] = (uint32_t)(((warp >> 3) * na + a0 + (lane >> 4)) * 1024 + (warp & 7) * 128 +
                              (((((lane & 15) >> 1) ^ (warp & 7))) << 4) + ((lane & 1) << 3));
      } else {
        // tf32 image: stage row r = warp + 16 j: r & 3 == warp & 3, r >> 2 == (warp >> 2) + 4 j
        off0[pi] = (uint32_t)(((warp >> 2) * natoms + pc.atom0 + (lane >> 3)) * 512 + (warp & 3) * 128 +
                              (((((lane & 7) >> 1) ^ (warp & 3))) << 5) + ((lane & 1) << 4));
      }
    }
    auto load_stage = [&](int it, int32_t idx, float4(&v)[NP][2]) {
#pragma unroll
      for (int pi = 0; pi < NP; ++pi) {
        const WgPiece &pc = p.pc[pi];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int64_t g = r_begin + (int64_t)it * WG_KR + warp + 16 * j;
          const bool row_ok = it < n_stage_iters && g < r_end;
          if (fast[pi]) {
            int64_t r0 = g;
            if (pc.mode == GNNFD_SEG_GATHER) r0 = __shfl_sync(0xffffffffu, idx, (pc.slot0 & 15) * 2 + j);
            const float *bp = NP < 4 ? base[pi] : pc.src + pc.col + lane * 4;
            const int64_t ldv = NP < 4 ? ldp[pi] : pc.ld;
            v[pi][j] = (row_ok && lane * 4 < pc.width) ? ldg_f4(bp + r0 * ldv) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            int32_t i0 = 0, i1 = 0, i2 = 0;
            if (pc.mode != GNNFD_SEG_DIRECT) {     // uniform per piece
              i0 = __shfl_sync(0xffffffffu, idx, (pc.slot0 & 15) * 2 + j);
              if (pc.mode >= GNNFD_SEG_SUM2) i1 = __shfl_sync(0xffffffffu, idx, ((pc.slot0 + 1) & 15) * 2 + j);
              if (pc.mode == GNNFD_SEG_MEAN3) i2 = __shfl_sync(0xffffffffu, idx, ((pc.slot0 + 2) & 15) * 2 + j);
            }
            v[pi][j] = wg_load_item(pc, lane, row_ok, g, i0, i1, i2);
          }
        }
      }
    };
    // activation / tf32 rounding / swizzled store of stage `it`, then hand the stage to the MMA issuer
    // (stage, round) of the next stage to store, kept as running counters: `it % p.stages` is an integer division by a
    // kernel parameter per stage, 17 % of this kernel's stall samples before (ncu source view)
    const int n_stages = p.stages;
    int st = 0;
    uint32_t st_round = 0;
    auto store_stage = [&](int it, const float4(&v)[NP][2]) {
      uint8_t *sA = smem + (size_t)st * stage_bytes, *sB = sA + a_bytes;
      if (st_round > 0) mbar_wait(&empty[st], (st_round - 1) & 1);
#pragma unroll
      for (int pi = 0; pi < NP; ++pi) {
        if (lane_on[pi]) {
          const int act = p.pc[pi].act;
          uint8_t *img = (pi == 0 ? sA : sB) + off0[pi];
          float4 t0 = v[pi][0], t1 = v[pi][1];
          if (pi == p.colsum_piece) {
            cs.x += t0.x + t1.x; cs.y += t0.y + t1.y; cs.z += t0.z + t1.z; cs.w += t0.w + t1.w;
          }
          if (act == 1) {        // SiLU of a saved pre-activation (uniform branch, 8 independent MUFU chains)
            t0.x = wg_silu(t0.x); t0.y = wg_silu(t0.y); t0.z = wg_silu(t0.z); t0.w = wg_silu(t0.w);
            t1.x = wg_silu(t1.x); t1.y = wg_silu(t1.y); t1.z = wg_silu(t1.z); t1.w = wg_silu(t1.w);
          } else if (act == 2) {
            t0.x = tanhf(t0.x); t0.y = tanhf(t0.y); t0.z = tanhf(t0.z); t0.w = tanhf(t0.w);
            t1.x = tanhf(t1.x); t1.y = tanhf(t1.y); t1.z = tanhf(t1.z); t1.w = tanhf(t1.w);
          }
          if (BF) {
            // hi / lo bf16 pairs of the lane's 4 columns: 8-byte stores into the hi image and the lo image
            const uint32_t part = pi == 0 ? a_img : b_img;
            const uint32_t jstep = (uint32_t)(2 * (pi == 0 ? 2 : nb_atoms) * 1024);     // row + 16 = 2 K atoms further
            uint32_t h0, l0, h1, l1;
            split2<false>(t0.x, t0.y, h0, l0); split2<false>(t0.z, t0.w, h1, l1);
            *reinterpret_cast<uint2 *>(img) = make_uint2(h0, h1);
            *reinterpret_cast<uint2 *>(img + part) = make_uint2(l0, l1);
            split2<false>(t1.x, t1.y, h0, l0); split2<false>(t1.z, t1.w, h1, l1);
            *reinterpret_cast<uint2 *>(img + jstep) = make_uint2(h0, h1);
            *reinterpret_cast<uint2 *>(img + jstep + part) = make_uint2(l0, l1);
          } else {
            t0.x = to_tf32(t0.x); t0.y = to_tf32(t0.y); t0.z = to_tf32(t0.z); t0.w = to_tf32(t0.w);
            t1.x = to_tf32(t1.x); t1.y = to_tf32(t1.y); t1.z = to_tf32(t1.z); t1.w = to_tf32(t1.w);
            *reinterpret_cast<float4 *>(img) = t0;
            *reinterpret_cast<float4 *>(img + (pi == 0 ? 4 : nb_atoms) * 2048) = t1;   // row + 16 = 4 K atoms further
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
      if (++st == n_stages) { st = 0; ++st_round; }
    };
    // two stages of loads in flight per warp: the loads of stage it + 1 are issued before stage it is converted
    float4 v0[NP][2], v1[NP][2];
    int32_t ix1 = load_idx(1);
    load_stage(0, load_idx(0), v0);
    for (int it = 0; it < n_stage_iters; it += 2) {
      const int32_t ix2 = load_idx(it + 2);
      load_stage(it + 1, ix1, v1);
      store_stage(it, v0);
      const int32_t ix3 = load_idx(it + 3);
      load_stage(it + 2, ix2, v0);
      if (it + 1 < n_stage_iters) store_stage(it + 1, v1);
      ix1 = ix3;
    }
    if (p.colsum != nullptr) {   // column sums: warps reduced in fixed order
      *reinterpret_cast<float4 *>(s_cs + warp * 128 + lane * 4) = cs;
      named_bar_sync(1, WG_PROD_WARPS * 32);
      if (tid < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < WG_PROD_WARPS; ++w) s += s_cs[w * 128 + tid];
        p.colsum[(size_t)blockIdx.x * 128 + tid] = s;
      }
    }
    // ================================================================================ epilogue
    if (warp < 4) {
      mbar_wait(done, 0);
      tc_fence_after();
      float *dst = p.partial + ((size_t)blockIdx.x * 128 + warp * 32 + lane) * p.n_pad;
      for (int c = 0; c < (p.n_pad >> 5); ++c) {
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, acc);
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4 *>(dst + c * 32 + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
      tc_fence_before();
    }
  } else {
    // ================================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t sbo_a = BF ? 2 * 1024 : 4 * 512, sbo_b = (uint32_t)nb_atoms * (BF ? 1024 : 512);
      const int n_stages = p.stages;
      int st = 0;
      uint32_t st_round = 0;
      for (int it = 0; it < n_stage_iters; ++it) {
        mbar_wait(&full[st], st_round & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + (size_t)st * stage_bytes), sB = sA + a_bytes;
        if (BF) {
#pragma unroll 1
          for (int ks = 0; ks < WG_KR / 16; ++ks) {          // K = 16 rows = 2 K atoms per MMA
            const uint64_t ah = make_desc_mn16(sA + 2 * ks * sbo_a, sbo_a), al = make_desc_mn16(sA + a_img + 2 * ks * sbo_a, sbo_a);
            for (int n0 = 0; n0 < p.n_pad; n0 += 256) {
              const int n = min(256, p.n_pad - n0);
              const uint32_t idesc = make_idesc_bf16_mn(n);
              const uint32_t boff = 2 * ks * sbo_b + (n0 >> 6) * 1024;
              const uint64_t bh = make_desc_mn16(sB + boff, sbo_b), bl = make_desc_mn16(sB + b_img + boff, sbo_b);
              umma_ss(tmem_base + n0, ah, bh, idesc, (it | ks) != 0);
              umma_ss(tmem_base + n0, al, bh, idesc, 1);
              umma_ss(tmem_base + n0, ah, bl, idesc, 1);
            }
          }
        } else {
#pragma unroll 1
          for (int ka = 0; ka < WG_KR / 8; ++ka) {
            const uint64_t ad = make_desc_mn(sA + 2 * ka * sbo_a, sbo_a);
            const uint32_t acc = (it | ka) != 0;
            for (int n0 = 0; n0 < p.n_pad; n0 += 256) {
              const int n = min(256, p.n_pad - n0);
              const uint64_t bd = make_desc_mn(sB + 2 * ka * sbo_b + (n0 >> 5) * 512, sbo_b);
              umma_tf32_ss(tmem_base + n0, ad, bd, make_idesc_tf32_mn(n), acc);
            }
          }
        }
        umma_commit(&empty[st]);
        if (++st == n_stages) { st = 0; ++st_round; }
      }
      umma_commit(done);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_PROD_WARPS) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ lean kernel
// The case that carries almost all of a training step's weight-gradient bytes: A and B both DIRECT, contiguous
// [rows, 128] matrices, B optionally a saved pre-activation (act: 0 none, 1 SiLU, 2 tanh), split-bf16.  The producers run
// a loop specialised for it - running row pointers instead of per-load 64-bit index arithmetic, no per-piece mode / width
// predicates, THREE stages of loads in flight per warp: ~120 instructions per warp and stage against ~340 of the generic
// assembly loop, which paced the kernel (issue-bound at 2.7 TB/s; 4.4 TB/s now).
// Up to THREE such GEMMs over the same rows run in ONE launch (dW3, dW2 and the leading block of dW1 of an MLP backward):
// job j accumulates into TMEM columns [128 j, 128 j + 128); the stage ring and its barriers simply continue from job to
// job, so the launch / prologue / partial-write overhead (~19 us per GEMM against 56 us of streaming) is paid once.
constexpr int WG_LEAN_MAX_JOBS = 3;
struct WgLeanJob {
  const float *a, *b;
  int32_t act;
  float *partial;   // [grid][128][128]
  float *colsum;    // [grid][128] column sums of A, or nullptr
};
struct WgLeanParams {
  WgLeanJob job[WG_LEAN_MAX_JOBS];
  int n_jobs, stages;
  int64_t rows, rows_per_cta;
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_lean_kernel(const __grid_constant__ WgLeanParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t a_bytes = 16 * 1024, stage_bytes = 32 * 1024;      // A hi | lo (8 KB each), B hi | lo
  constexpr uint32_t JSTEP = 4096, PART = 8192;
  uint8_t *s_tail = smem + (size_t)p.stages * stage_bytes;
  float *s_cs = (float *)s_tail;                                // [16 warps][128] column-sum scratch
  uint64_t *s_bar = (uint64_t *)(s_tail + WG_PROD_WARPS * 128 * 4);
  uint64_t *full = s_bar, *empty = s_bar + WG_MAX_STAGES, *done = s_bar + 2 * WG_MAX_STAGES;
  uint32_t *s_tmem = (uint32_t *)(s_bar + 2 * WG_MAX_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], WG_PROD_WARPS); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_PROD_WARPS) tmem_alloc(s_tmem, 512);
  pdl_wait();   // everything above is on-chip set-up that may run under the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r_end = min(p.rows, r_begin + p.rows_per_cta);
  const int n_local = (int)(r_end - r_begin);                     // rows of this CTA (> 0)
  const int n_it = (n_local + WG_KR - 1) / WG_KR;
  const int n_stages = p.stages;

  if (warp < WG_PROD_WARPS) {
    // =============================================================================== producers
    // image offset of the lane's 4 columns of stage row `warp`: K atom warp >> 3 (x 2 MN atoms), MN atom lane >> 4, row
    // warp & 7 inside the atom, 16-byte chunk ((lane & 15) >> 1) ^ (row & 7), half lane & 1; row + 16 = 2 K atoms further
    const uint32_t off = (uint32_t)(((warp >> 3) * 2 + (lane >> 4)) * 1024 + (warp & 7) * 128 +
                                    (((((lane & 15) >> 1) ^ (warp & 7))) << 4) + ((lane & 1) << 3));
    int st = 0;
    uint32_t st_round = 0;
    for (int jb = 0; jb < p.n_jobs; ++jb) {
      const WgLeanJob &job = p.job[jb];
      const int act = job.act;
      const bool want_cs = job.colsum != nullptr;
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
      const float *la = job.a + (r_begin + warp) * 128 + lane * 4;      // load pointers of the next stage to fetch
      const float *lb = job.b + (r_begin + warp) * 128 + lane * 4;
      int lrow = warp;                                                  // its first row, relative to r_begin
      auto fetch = [&](float4(&a)[2], float4(&b)[2]) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool ok = lrow + 16 * j < n_local;
          a[j] = ok ? ldg_f4(la + j * 16 * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
          b[j] = ok ? ldg_f4(lb + j * 16 * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        la += WG_KR * 128; lb += WG_KR * 128; lrow += WG_KR;
      };
      auto put = [&](const float4(&a)[2], const float4(&b)[2]) {
        const uint32_t sA = smem_u32(smem + (size_t)st * stage_bytes) + off, sB = sA + a_bytes;
        if (st_round > 0) mbar_wait(&empty[st], (st_round - 1) & 1);
        if (want_cs) {
          cs.x += a[0].x + a[1].x; cs.y += a[0].y + a[1].y; cs.z += a[0].z + a[1].z; cs.w += a[0].w + a[1].w;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t h0, l0, h1, l1;
          split2<false>(a[j].x, a[j].y, h0, l0); split2<false>(a[j].z, a[j].w, h1, l1);
          sts_u2(sA + j * JSTEP, h0, h1);
          sts_u2(sA + j * JSTEP + PART, l0, l1);
          float4 t = b[j];
          if (act == 1) { wg_silu2(t.x, t.y); wg_silu2(t.z, t.w); }
          else if (act == 2) { t.x = tanhf(t.x); t.y = tanhf(t.y); t.z = tanhf(t.z); t.w = tanhf(t.w); }
          split2<false>(t.x, t.y, h0, l0); split2<false>(t.z, t.w, h1, l1);
          sts_u2(sB + j * JSTEP, h0, h1);
          sts_u2(sB + j * JSTEP + PART, l0, l1);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
        if (++st == n_stages) { st = 0; ++st_round; }
      };
      float4 a0[2], b0[2], a1[2], b1[2], a2[2], b2[2];
      fetch(a0, b0);
      fetch(a1, b1);
      for (int it = 0; it < n_it; it += 3) {
        fetch(a2, b2);
        put(a0, b0);
        fetch(a0, b0);
        if (it + 1 < n_it) put(a1, b1);
        fetch(a1, b1);
        if (it + 2 < n_it) put(a2, b2);
      }
      if (want_cs) {   // column sums of A: warps reduced in fixed order (uniform branch: every producer takes it)
        *reinterpret_cast<float4 *>(s_cs + warp * 128 + lane * 4) = cs;
        named_bar_sync(1, WG_PROD_WARPS * 32);
        if (tid < 128) {
          float sacc = 0.f;
#pragma unroll
          for (int w = 0; w < WG_PROD_WARPS; ++w) sacc += s_cs[w * 128 + tid];
          job.colsum[(size_t)blockIdx.x * 128 + tid] = sacc;
        }
        named_bar_sync(1, WG_PROD_WARPS * 32);      // the scratch is free for the next job
      }
    }
    // ================================================================================ epilogue
    if (warp < 4) {
      mbar_wait(done, 0);
      tc_fence_after();
      for (int jb = 0; jb < p.n_jobs; ++jb) {
        float *dst = p.job[jb].partial + ((size_t)blockIdx.x * 128 + warp * 32 + lane) * 128;
        for (int c = 0; c < 4; ++c) {
          float acc[32];
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + jb * 128 + c * 32, acc);
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4 *>(dst + c * 32 + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
        }
      }
      tc_fence_before();
    }
  } else {
    // ================================================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t sbo = 2 * 1024;
      const uint32_t idesc = make_idesc_bf16_mn(128);
      int st = 0;
      uint32_t st_round = 0;
      for (int jb = 0; jb < p.n_jobs; ++jb) {
        const uint32_t d = tmem_base + jb * 128;
        for (int it = 0; it < n_it; ++it) {
          mbar_wait(&full[st], st_round & 1);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + (size_t)st * stage_bytes), sB = sA + a_bytes;
#pragma unroll
          for (int ks = 0; ks < WG_KR / 16; ++ks) {          // K = 16 rows = 2 K atoms per MMA
            const uint64_t ah = make_desc_mn16(sA + 2 * ks * sbo, sbo), al = make_desc_mn16(sA + PART + 2 * ks * sbo, sbo);
            const uint64_t bh = make_desc_mn16(sB + 2 * ks * sbo, sbo), bl = make_desc_mn16(sB + PART + 2 * ks * sbo, sbo);
            umma_ss(d, ah, bh, idesc, (it | ks) != 0);
            umma_ss(d, al, bh, idesc, 1);
            umma_ss(d, ah, bl, idesc, 1);
          }
          umma_commit(&empty[st]);
          if (++st == n_stages) { st = 0; ++st_round; }
        }
      }
      umma_commit(done);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_PROD_WARPS) tmem_dealloc(tmem_base, 512);
}

// out[m, j] = sum over the split-K partials part[c][m][j] (+ column sums), optionally stored transposed.
// 32 outputs x 16 part-groups per block: group y adds parts y, y + 16, ... (coalesced across the 32 outputs: ~10 loads in
// flight per thread instead of ~37 dependent-latency-bound ones), the sixteen group sums are combined by a fixed tree -
// deterministic.  (This launch follows every weight-gradient GEMM: ~110 per training step.)
constexpr int RP_OUT = 32, RP_GROUPS = 16;
__global__ void __launch_bounds__(RP_OUT * RP_GROUPS) reduce_partials_kernel(const float *__restrict__ part, int n_parts, int n_pad,
                                                              int m_valid, int n_valid, float *__restrict__ out,
                                                              int ld_out, int transpose, const float *__restrict__ cs_part,
                                                              float *__restrict__ cs_out, int cs_valid) {
  pdl_entry();
  __shared__ float s_sum[RP_GROUPS][RP_OUT];
  const int total = m_valid * n_valid;
  const int tx = threadIdx.x & (RP_OUT - 1), ty = threadIdx.x / RP_OUT;
  const int i = blockIdx.x * RP_OUT + tx;
  float s = 0.f;
  if (i < total) {
    const int m = i / n_valid, j = i % n_valid;
    const float *src = part + (size_t)m * n_pad + j;
    for (int c = ty; c < n_parts; c += RP_GROUPS) s += src[(size_t)c * 128 * n_pad];
  } else if (cs_out != nullptr && i - total < cs_valid) {
    for (int c = ty; c < n_parts; c += RP_GROUPS) s += cs_part[(size_t)c * 128 + (i - total)];
  }
  s_sum[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float r[RP_GROUPS];
#pragma unroll
    for (int g = 0; g < RP_GROUPS; ++g) r[g] = s_sum[g][tx];
#pragma unroll
    for (int w = RP_GROUPS / 2; w > 0; w >>= 1)
#pragma unroll
      for (int g = 0; g < w; ++g) r[g] += r[g + w];
    if (i < total) {
      const int m = i / n_valid, j = i % n_valid;
      if (transpose) out[(size_t)j * ld_out + m] = r[0]; else out[(size_t)m * ld_out + j] = r[0];
    } else if (cs_out != nullptr && i - total < cs_valid) {
      cs_out[i - total] = r[0];
    }
  }
}

struct WgReduceJobs { WgReduceJob job[WG_MAX_REDUCE_JOBS]; int first_block[WG_MAX_REDUCE_JOBS + 1]; int n; };
// the same reduction for several GEMMs in one launch: block -> (job, block of the job)
__global__ void __launch_bounds__(RP_OUT * RP_GROUPS) reduce_partials_multi_kernel(const __grid_constant__ WgReduceJobs jobs) {
  pdl_entry();
  __shared__ float s_sum[RP_GROUPS][RP_OUT];
  int q = 0;
  while (q + 1 < jobs.n && (int)blockIdx.x >= jobs.first_block[q + 1]) ++q;
  const WgReduceJob &jb = jobs.job[q];
  const int total = jb.m_valid * jb.n_valid;
  const int tx = threadIdx.x & (RP_OUT - 1), ty = threadIdx.x / RP_OUT;
  const int i = ((int)blockIdx.x - jobs.first_block[q]) * RP_OUT + tx;
  float s = 0.f;
  if (i < total) {
    const int m = i / jb.n_valid, j = i % jb.n_valid;
    const float *src = jb.part + (size_t)m * jb.n_pad + j;
    for (int c = ty; c < jb.n_parts; c += RP_GROUPS) s += src[(size_t)c * 128 * jb.n_pad];
  } else if (jb.cs_out != nullptr && i - total < jb.cs_valid) {
    for (int c = ty; c < jb.n_parts; c += RP_GROUPS) s += jb.cs_part[(size_t)c * 128 + (i - total)];
  }
  s_sum[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float r[RP_GROUPS];
#pragma unroll
    for (int g = 0; g < RP_GROUPS; ++g) r[g] = s_sum[g][tx];
#pragma unroll
    for (int w = RP_GROUPS / 2; w > 0; w >>= 1)
#pragma unroll
      for (int g = 0; g < w; ++g) r[g] += r[g + w];
    if (i < total) {
      const int m = i / jb.n_valid, j = i % jb.n_valid;
      if (jb.transpose) jb.out[(size_t)j * jb.ld_out + m] = r[0]; else jb.out[(size_t)m * jb.ld_out + j] = r[0];
    } else if (jb.cs_out != nullptr && i - total < jb.cs_valid) {
      jb.cs_out[i - total] = r[0];
    }
  }
}

int wgrad_reduce_jobs(const WgReduceJob *jobs, int n_jobs, cudaStream_t stream) {
  if (n_jobs <= 0) return GNNFD_OK;
  if (n_jobs > WG_MAX_REDUCE_JOBS) { set_error("wgrad_reduce_jobs: too many jobs"); return GNNFD_E_BADARG; }
  WgReduceJobs js{};
  js.n = n_jobs;
  int blocks = 0;
  for (int q = 0; q < n_jobs; ++q) {
    js.job[q] = jobs[q];
    js.first_block[q] = blocks;
    const int total = jobs[q].m_valid * jobs[q].n_valid + (jobs[q].cs_out ? jobs[q].cs_valid : 0);
    blocks += (total + RP_OUT - 1) / RP_OUT;
  }
  js.first_block[n_jobs] = blocks;
  if (blocks == 0) return GNNFD_OK;
  launch_pdl(reduce_partials_multi_kernel, dim3(blocks), dim3(RP_OUT * RP_GROUPS), 0, stream, js);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

static int64_t wg_rows_per_cta(int64_t rows, int &grid) {
  const int sms = num_sms();
  int64_t per = (rows + sms - 1) / sms;
  per = ((per + WG_KR - 1) / WG_KR) * WG_KR;
  if (per < 4 * WG_KR) per = 4 * WG_KR;                   // tiny inputs: fewer CTAs, fewer partials
  grid = (int)((rows + per - 1) / per);
  return per;
}

bool wgrad_is_lean(const gnnfd_wgrad_args *a) {
  const gnnfd_segment &b0 = a->b[0];
  return a->precision == 0 && a->n_b == 1 && a->rows > 0 && a->a.mode == GNNFD_SEG_DIRECT && b0.mode == GNNFD_SEG_DIRECT &&
         a->a.width == 128 && b0.width == 128 && a->a.ld == 128 && b0.ld == 128 && a->a.col == 0 && b0.col == 0 &&
         a->a_act == 0 && !a->colsum_of_b && a->a.src != nullptr && b0.src != nullptr && a->out != nullptr &&
         ((reinterpret_cast<uintptr_t>(a->a.src) | reinterpret_cast<uintptr_t>(b0.src)) & 15) == 0;
}

int wgrad_lean_run(const gnnfd_wgrad_args *const *args, int n, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   WgReduceJob *jobs_out) {
  if (n < 1 || n > WG_LEAN_MAX_JOBS) { set_error("wgrad_lean_run: 1..3 GEMMs per launch"); return GNNFD_E_BADARG; }
  for (int j = 0; j < n; ++j)
    if (!wgrad_is_lean(args[j]) || args[j]->rows != args[0]->rows) {
      set_error("wgrad_lean_run: GEMM %d is not a lean direct x direct GEMM over the same rows", j);
      return GNNFD_E_BADARG;
    }
  WgLeanParams p{};
  int grid;
  p.rows = args[0]->rows;
  p.rows_per_cta = wg_rows_per_cta(p.rows, grid);
  p.n_jobs = n;
  p.stages = WG_MAX_STAGES;
  const size_t per_job = (((size_t)grid * 128 * 128 * 4 + (size_t)grid * 128 * 4) + 255) & ~(size_t)255;
  if (workspace == nullptr || workspace_bytes < per_job * n) { set_error("wgrad_lean_run: workspace too small"); return GNNFD_E_WORKSPACE; }
  for (int j = 0; j < n; ++j) {
    const gnnfd_wgrad_args *a = args[j];
    float *part = (float *)((uint8_t *)workspace + per_job * j);
    float *cs_part = part + (size_t)grid * 128 * 128;
    p.job[j].a = a->a.src; p.job[j].b = a->b[0].src; p.job[j].act = a->b_act;
    p.job[j].partial = part; p.job[j].colsum = a->colsum ? cs_part : nullptr;
    WgReduceJob &r = jobs_out[j];
    r = WgReduceJob{};
    r.part = part; r.cs_part = cs_part; r.out = a->out; r.cs_out = a->colsum;
    r.n_parts = grid; r.n_pad = 128; r.m_valid = 128; r.n_valid = 128;
    r.ld_out = a->ld_out; r.transpose = a->transpose_out; r.cs_valid = a->colsum ? 128 : 0;
    r.ws_used = per_job;
  }
  const int smem = p.stages * 32 * 1024 + WG_PROD_WARPS * 128 * 4 + (2 * WG_MAX_STAGES + 1) * 8 + 64 + 1024;
  static bool attr[GNNFD_MAX_DEVICES] = {false};
  if (!attr[current_device()]) {
    GNNFD_CUDA(cudaFuncSetAttribute(wgrad_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr[current_device()] = true;
  }
  launch_pdl(wgrad_lean_kernel, dim3(grid), dim3(WG_THREADS), smem, stream, p);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

}  // namespace gnnfd

using namespace gnnfd;

static int wg_n_pad(const gnnfd_wgrad_args *a) {   // split-bf16 pads to 64-column atoms, TF32 to 32-column atoms
  const int g = a->precision == 1 ? 31 : 63;
  int n = 0;
  for (int s = 0; s < a->n_b; ++s) n += (a->b[s].width + g) & ~g;
  return n;
}

extern "C" size_t gnnfd_wgrad_workspace_bytes(int64_t rows, int32_t n_cols_padded /* to 64 */) {
  if (rows <= 0) return 256;
  int grid;
  wg_rows_per_cta(rows, grid);
  return (size_t)grid * 128 * (size_t)n_cols_padded * 4 + (size_t)grid * 128 * 4 + 256;
}

extern "C" int gnnfd_wgrad(const gnnfd_wgrad_args *a, void *workspace, size_t workspace_bytes, void *stream_) {
  return gnnfd::wgrad_run(a, workspace, workspace_bytes, (cudaStream_t)stream_, nullptr);
}

// `defer` != nullptr: the split-K partials stay in the workspace (defer->ws_used bytes) and their reduction is described
// in *defer for a later wgrad_reduce_jobs launch; nullptr: reduced right away.
int gnnfd::wgrad_run(const gnnfd_wgrad_args *a, void *workspace, size_t workspace_bytes, cudaStream_t stream, WgReduceJob *defer) {
  if (defer != nullptr) *defer = WgReduceJob{};
  if (a != nullptr && wgrad_is_lean(a)) {     // the dominant case: contiguous [rows, 128] x [rows, 128]
    WgReduceJob job;
    const int rc = wgrad_lean_run(&a, 1, workspace, workspace_bytes, stream, &job);
    if (rc != GNNFD_OK) return rc;
    if (defer != nullptr) { *defer = job; return GNNFD_OK; }
    return wgrad_reduce_jobs(&job, 1, stream);
  }
  GNNFD_CHECK_ARG(a != nullptr && a->out != nullptr, "null args/out");
  GNNFD_CHECK_ARG(a->rows >= 0, "negative rows");
  GNNFD_CHECK_ARG(a->n_b >= 1 && a->n_b <= 3, "n_b must be 1..3");
  GNNFD_CHECK_ARG(a->a.mode == GNNFD_SEG_DIRECT && a->a.width > 0 && a->a.width <= 128, "A must be a DIRECT segment of <= 128 columns");
  const int n_pad = wg_n_pad(a);
  GNNFD_CHECK_ARG(n_pad <= 384, "B wider than 384 columns");
  int n_valid = 0;
  for (int s = 0; s < a->n_b; ++s) {
    GNNFD_CHECK_ARG(a->b[s].width > 0 && a->b[s].width <= 128, "B segment width must be 1..128");
    GNNFD_CHECK_ARG(s == a->n_b - 1 || (a->b[s].width & 63) == 0, "only the last B segment may be narrower than a multiple of 64");
    n_valid += a->b[s].width;
  }
  const int cs_valid = a->colsum ? (a->colsum_of_b ? a->b[0].width : a->a.width) : 0;
  if (a->rows == 0) {
    const int m_valid = a->a.width;
    if (a->transpose_out) { for (int j = 0; j < n_valid; ++j) GNNFD_CUDA(cudaMemsetAsync(a->out + (size_t)j * a->ld_out, 0, m_valid * 4, stream)); }
    else { for (int m = 0; m < m_valid; ++m) GNNFD_CUDA(cudaMemsetAsync(a->out + (size_t)m * a->ld_out, 0, n_valid * 4, stream)); }
    if (a->colsum) GNNFD_CUDA(cudaMemsetAsync(a->colsum, 0, cs_valid * 4, stream));
    return GNNFD_OK;
  }
  WgParams p{};
  int grid;
  p.rows_per_cta = wg_rows_per_cta(a->rows, grid);
  p.rows = a->rows;
  p.n_pad = n_pad;
  const size_t need = (size_t)grid * 128 * (size_t)n_pad * 4 + (size_t)grid * 128 * 4;
  if (workspace == nullptr || workspace_bytes < need) { set_error("gnnfd_wgrad: workspace too small"); return GNNFD_E_WORKSPACE; }
  p.partial = (float *)workspace;
  float *cs_part = p.partial + (size_t)grid * 128 * n_pad;
  p.colsum = a->colsum ? cs_part : nullptr;
  p.colsum_piece = a->colsum ? (a->colsum_of_b ? 1 : 0) : -1;
  p.n_pieces = 1 + a->n_b;
  int atom = 0, slot = 0;
  for (int pi = 0; pi < p.n_pieces; ++pi) {
    const gnnfd_segment &sg = pi == 0 ? a->a : a->b[pi - 1];
    WgPiece &pc = p.pc[pi];
    GNNFD_CHECK_ARG(sg.src != nullptr, "null segment source");
    GNNFD_CHECK_ARG(sg.mode >= GNNFD_SEG_DIRECT && sg.mode <= GNNFD_SEG_MEAN3, "bad segment mode");
    pc.src = sg.src; pc.ld = sg.ld; pc.col = sg.col; pc.width = sg.width; pc.mode = sg.mode;
    pc.act = pi == 0 ? a->a_act : a->b_act;
    pc.vec = ((sg.ld & 3) == 0) && ((sg.col & 3) == 0) && ((sg.width & 3) == 0) &&
             ((reinterpret_cast<uintptr_t>(sg.src) & 15) == 0);
    pc.atom0 = pi == 0 ? 0 : atom;
    if (pi > 0) atom += a->precision == 1 ? ((sg.width + 31) & ~31) >> 5 : ((sg.width + 63) & ~63) >> 6;
    const int n_idx = sg.mode == GNNFD_SEG_DIRECT ? 0 : sg.mode == GNNFD_SEG_GATHER ? 1 : sg.mode == GNNFD_SEG_MEAN3 ? 3 : 2;
    pc.slot0 = slot;
    for (int q = 0; q < n_idx; ++q) {
      GNNFD_CHECK_ARG(sg.idx[q] != nullptr, "null gather index");
      GNNFD_CHECK_ARG(slot < WG_MAX_SLOTS, "too many gather indices");
      p.slot[slot++] = sg.idx[q];
    }
  }
  p.n_slots = slot;
  const uint32_t stage_bytes = 16 * 1024 + (uint32_t)n_pad * 128;
  int stages = (int)((200u * 1024u) / stage_bytes);
  p.stages = stages > WG_MAX_STAGES ? WG_MAX_STAGES : stages;
  const int smem = p.stages * (int)stage_bytes + WG_PROD_WARPS * 128 * 4 + (2 * WG_MAX_STAGES + 1) * 8 + 64 + 1024;
#define WG_LAUNCH(NP_, MD_)                                                                                     \
  do {                                                                                                         \
    static bool attr[GNNFD_MAX_DEVICES] = {false};                                                                                  \
    if (!attr[current_device()]) {                                                                                               \
      GNNFD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<NP_, MD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr[current_device()] = true;                                                                                             \
    }                                                                                                          \
    launch_pdl(wgrad_tc_kernel<NP_, MD_>, dim3(grid), dim3(WG_THREADS), smem, stream, p);                            \
  } while (0)
  if (a->precision == 1) { if (p.n_pieces == 2) WG_LAUNCH(2, 1); else if (p.n_pieces == 3) WG_LAUNCH(3, 1); else WG_LAUNCH(4, 1); }
  else { if (p.n_pieces == 2) WG_LAUNCH(2, 0); else if (p.n_pieces == 3) WG_LAUNCH(3, 0); else WG_LAUNCH(4, 0); }
#undef WG_LAUNCH
  GNNFD_LAUNCH_CHECK();
  const int m_valid = a->a.width;
  const int total = m_valid * n_valid + cs_valid;
  if (defer != nullptr) {
    defer->part = p.partial; defer->cs_part = cs_part; defer->out = a->out; defer->cs_out = a->colsum;
    defer->n_parts = grid; defer->n_pad = n_pad; defer->m_valid = m_valid; defer->n_valid = n_valid;
    defer->ld_out = a->ld_out; defer->transpose = a->transpose_out; defer->cs_valid = cs_valid;
    defer->ws_used = (need + 255) & ~(size_t)255;
    return GNNFD_OK;
  }
  launch_pdl(reduce_partials_kernel, dim3((total + RP_OUT - 1) / RP_OUT), dim3(RP_OUT * RP_GROUPS), 0, stream, p.partial, grid, n_pad, m_valid, n_valid, a->out,
                                                                  a->ld_out, a->transpose_out, cs_part, a->colsum, cs_valid);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}
