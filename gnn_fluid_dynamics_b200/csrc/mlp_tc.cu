// Fused MLP block on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   assemble input rows (gather / concat / sum / mean3)            -> bf16|fp16 hi/lo parts, SMEM
//   layer 1: tcgen05.mma kind::f16 SS (A, W from shared memory), M=128 N=128 K=16, fp32 accum in TMEM
//   hidden epilogues: tcgen05.ld -> +bias, SiLU/Tanh -> hi/lo split -> tcgen05.st IN PLACE
//   layers 2, 3: tcgen05.mma TS (A operand read straight from TMEM, W from shared memory)
//   final epilogue: +bias -> LayerNorm -> *mul -> coalesced (+residual) stores
//
// One persistent CTA per SM keeps THREE 128-row tiles in flight.  TMEM (512 columns) = three X regions (tile j % 3)
// + one Y region;  L1: D=X | E1: X -> X (in place) | L2: A=X, D=Y | E2: Y -> Y | L3: A=Y, D=X | final: X -> HBM.
// Y is handed from tile to tile by the in-order tensor pipe (L2 of tile j + 1 is issued after L3 of tile j by the
// same thread).  No intermediate touches HBM or shared memory.
// In-place layout: the 16 fp32 columns of group c become 8 columns of packed hi pairs + 8 of lo pairs.
//
// Operand precision is a template: split operands (x = hi + lo, products hi*hi + lo*hi + hi*lo) restore
// ~fp32 accuracy on the bf16/fp16 tensor pipe (single-pass bf16 fails the 1e-3 bar).
//
// Shared-memory operand layout: canonical UMMA K-major SWIZZLE_128B - a k-block is 64 elements
// (128 B per row), rows in groups of 8 (1024 B atoms), 16-byte chunk c of row r stored at chunk
// c ^ (r & 7).  Weights are pre-packed by gnnfd_pack_mlp into exactly this image per (layer,
// k-block, part), so one cp.async.bulk (TMA bulk copy, mbarrier complete_tx) lands a unit.
#include <stdlib.h>
#include <cuda.h>   // CUtensorMap (types only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include "tc_common.cuh"

#ifndef GNNFD_ABL
#define GNNFD_ABL 0   // timing ablations (scripts/abl_edge.py): results are WRONG with any value but 0
#endif

namespace gnnfd {

// warp roles: 0-7 epilogue of even local tiles, 8-15 epilogue of odd local tiles - the two groups convert different
//             tiles CONCURRENTLY (warp & 3 = TMEM lane quarter, thread = row, (warp >> 2) & 1 = column half),
//             16-23 producers (gather -> split -> swizzled A stage), 24 = layer-1 MMA issuer, 25 = layer-2/3 MMA
//             issuer, 26 = layer-1 weight loader, 27 = layer-2/3 weight loader (TMA bulk copies).
// The two issuers run as a dataflow: nothing orders layer 1 of tile j + 1 against layers 2/3 of tile j except the
// mbarriers that carry the data, so neither a slow epilogue nor a slow gather stalls the other chain.
// TMEM: three X regions (layer-1 accumulator / hidden-1 operand / layer-3 accumulator of tiles j % 3) and ONE Y region
// (layer-2 accumulator / hidden-2 operand), which the in-order tensor pipe hands from tile to tile.
// Registers: 896 threads launch at 72 per thread; the epilogue works in 16-column groups to live within that;
// warpgroup 24-27 shrinks to 40 (setmaxnreg.dec) and what it releases goes to the producers (forward: 88) or to the
// epilogue warps (backward chain: 80) (setmaxnreg.inc; only registers released inside the CTA can be claimed).
constexpr int TC_EPI_WARPS = 16, TC_PROD_WARPS = 8;
constexpr int TC_EPI_GROUP = 8;       // epilogue warps per tile
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32, TC_PROD_THREADS = TC_PROD_WARPS * 32;
constexpr int TC_MMA_WARP = TC_EPI_WARPS + TC_PROD_WARPS;   // layer-1 issuer (also owns the TMEM allocation)
constexpr int TC_MMA23_WARP = TC_MMA_WARP + 1;
constexpr int TC_WLD_WARP = TC_MMA_WARP + 2;
constexpr int TC_WLD23_WARP = TC_MMA_WARP + 3;
constexpr int TC_THREADS = (TC_MMA_WARP + 4) * 32;   // 896
constexpr int TC_A_STAGES_MAX = 3;    // A ring: {A_hi, A_lo} images per stage; TcParams::a_stages = 2 (register-staged
                                      // producers only) or 3 (launches with TMA-gathered k-blocks: no register staging)
// W rings (layer 1 | layers 2/3): one 16 KB image (hi or lo part of a k-block) per slot, TcParams::w_slots slots each.
// Large launches run TWO slots per ring: measured FASTER than three, because the 32 KB not carved out of the L1 serve
// the producers' gathers (203 -> 194 us on the edge block).  Launches of a tile or two per CTA are a pure latency
// chain and get THREE (the 2k-cell rollout step: 1.00 -> 0.9x ms).  The rings sit at the END of shared memory so the
// dynamic allocation (which sets the L1 carve-out) only covers the slots in use.
constexpr int TC_W_SLOTS_MAX = 3;
constexpr int TC_X_SLOTS = 3;         // X regions in TMEM
constexpr int TC_STAGE_BYTES = 2 * TC_IMG;   // 32 KB per A stage
constexpr int TC_STG_BYTES = TC_EPI_WARPS * 32 * 16 * 4;   // 32 KB: one XOR-swizzled 32 x 16 fp32 staging block per epilogue warp
constexpr int TC_IDX_SLOTS = 4;
constexpr int TC_IDX_SLOT = 9 * TC_BM;               // ints: [3 seg][3 idx][128 rows]
constexpr int TC_NBAR = 32;
constexpr int TC_SMEM_FIXED = TC_STG_BYTES + TC_IDX_SLOTS * TC_IDX_SLOT * 4 +
                              5 * TC_H * 4 + 4 * TC_BM * 8 + TC_NBAR * 8 + 64;
constexpr int tc_smem_bytes(int w_slots, int a_stages) {      // + alignment slack before s_a and before the rings
  return TC_SMEM_FIXED + a_stages * TC_STAGE_BYTES + 2 * w_slots * TC_IMG + 2 * 1024;
}
static_assert(tc_smem_bytes(2, TC_A_STAGES_MAX) <= 227 * 1024 && tc_smem_bytes(TC_W_SLOTS_MAX, 2) <= 227 * 1024 &&
              tc_smem_bytes(2, 2) + TC_STG_BYTES <= 227 * 1024, "shared memory");
constexpr int TC_TMEM_COLS = 512;     // X0 | X1 | X2 | Y, 128 columns each
constexpr uint32_t TC_Y_COL = TC_X_SLOTS * 128;
constexpr int TC_MAX_KB = 8;

// how the producers assemble one 64-wide k-block of layer 1's input
struct KbDesc {
  const float *src;   // segment source matrix
  int32_t ld;         // row stride (floats)
  int32_t colk;       // first source column of this k-block
  int32_t mode;       // GNNFD_SEG_*
  int32_t kvalid;     // valid columns in this k-block (<= 64)
  int32_t seg;        // segment number (index slot)
  int32_t vec;        // 16-byte vector loads are legal
  int32_t tma;        // 1: GATHER k-block staged by TMA gather4 from the segment's split shadow (tm_seg[seg])
  int32_t lo_col;     // tma: column offset of the lo parts inside a shadow row (= ld of the fp32 source)
};

// fast final epilogue (template EPI = 1): what leaves the CTA and how
// (template EPI = 2: the generic epilogue with training-mode dropout of the hidden activations - its own instantiation so
//  that the hidden epilogue of every other launch carries no mask code)
enum {
  EPI_ST_RAW = 1,    // staging = out          -> TMA store to tm_raw
  EPI_RED_SUM = 2,   // staging = out          -> TMA reduce-add into tm_sum (residual == out_sum, updated in place)
  EPI_LDRES = 4,     // staging = out + residual (loaded thread-per-row) -> TMA store to tm_sum
  EPI_SPLIT = 8,     // 16-bit split shadow of out (or of out + residual with EPI_LDRES) -> out_split
};

struct TcParams {
  gnnfd_mlp_args a;
  int kb1;       // k-blocks of layer 1
  int ksteps1;   // K=16 steps in the LAST k-block of layer 1 (1..4)
  int n3;        // UMMA N of layer 3: 128, or 16 for a narrow head
  int nl;        // layers: 3 (MLP) or 1 (single Linear)
  int w_slots;   // slots per weight ring: 2 (large launches) or 3 (latency-bound small ones)
  uint32_t w_block_bytes;   // bytes of one packed 128-row k-block (all parts)
  uint32_t w3_block_bytes;  // bytes of one packed layer-3 k-block
  int64_t direct_tile_bytes;   // bytes of one tile's rows of a contiguous DIRECT segment 0 (0: no L2 prefetch)
  KbDesc kb[TC_MAX_KB];
  int a_stages;  // stages of the A ring (2 | 3)
  int epi;       // EPI_* bits (kernels instantiated with EPI = 1)
  int n_tma;     // k-blocks staged by TMA
  int l1_depth;  // layer-1 k-blocks whose MMAs may be queued in the tensor pipe at once (0 = unlimited)
  alignas(64) CUtensorMap tm_seg[3];   // split shadows of the gathered segments ([src_rows, 2 * ld] 16-bit, box 64 x 1)
  alignas(64) CUtensorMap tm_raw;      // out_raw / out_sum as [rows, 128] fp32, box 16 x 32, SWIZZLE_64B
  alignas(64) CUtensorMap tm_sum;
  alignas(64) CUtensorMap tm_save[2];  // save_a1 / save_a2 (training stash; backward chain: dA2 / dA1), same geometry
  alignas(64) CUtensorMap tm_split;    // out_split as [rows, 256] 16-bit, box 16 x 32, no swizzle
  int save_tma;  // the hidden epilogues write their stash through the staging block + TMA tensor stores
  int split_tma; // the split shadow leaves through a SECOND staging block (after the weight rings) + TMA tensor stores
  L2Policies pol;   // L2 eviction-priority operands of the global accesses (common.cuh)
  uint32_t drop_thresh;    // dropout (template EPI = 2): a hidden unit is dropped when its hash < drop_thresh
  uint32_t drop_key[2];    // per hidden layer (common.cuh: dropout_layer_key)
};

// Diagnostic cycle counters of CTA 0 (role wait times), read back with gnnfd_tc_profile_read; only
// recorded when the library is built with -DGNNFD_TC_PROF.
__device__ unsigned long long g_tc_prof[16];
#ifdef GNNFD_TC_PROF
#define PROF_DECL unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long t_begin = clock64()
#define PROF_WAIT(slot, stmt)                                  \
  do {                                                         \
    const long long t0_ = clock64();                           \
    stmt;                                                      \
    if (blockIdx.x == 0) prof[slot] += clock64() - t0_;        \
  } while (0)
#else
#define PROF_DECL do {} while (0)
#define PROF_WAIT(slot, stmt) stmt
#endif

// ------------------------------------------------------------------------------------ producers
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// generic (slow) path: partial k-blocks, unaligned sources
__device__ __noinline__ void tc_load_block_generic(const TcParams &p, const int32_t *ix, int64_t row0, int kb,
                                                   int rbase, int f4, float4 (&v)[8]) {
  const KbDesc &d = p.kb[kb];
  const int ksteps = (kb == p.kb1 - 1) ? p.ksteps1 : 4;
  const int32_t *ixs = ix + d.seg * 3 * TC_BM;
  const bool active = f4 * 4 < ksteps * 16;
  const int64_t rows = p.a.rows;
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int r = rbase + 16 * jj;
    const int64_t g = row0 + r;
    float t4[4] = {0.f, 0.f, 0.f, 0.f};
    if (active && g < rows) {
      if (d.mode == GNNFD_SEG_SUM3S) {     // signed sum of three gathered rows
        int32_t r3[3];
        float s3[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) sum3s_decode(ixs[u * TC_BM + r], r3[u], s3[u]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (f4 * 4 + q < d.kvalid) {
            const int64_t c = d.colk + f4 * 4 + q;
            t4[q] = (s3[0] * __ldg(d.src + (int64_t)r3[0] * d.ld + c) + s3[1] * __ldg(d.src + (int64_t)r3[1] * d.ld + c)) +
                    s3[2] * __ldg(d.src + (int64_t)r3[2] * d.ld + c);
          }
        }
        v[jj] = make_float4(t4[0], t4[1], t4[2], t4[3]);
        continue;
      }
      const int64_t i0 = d.mode == GNNFD_SEG_DIRECT ? g : (int64_t)ixs[r];
      const float *b0 = d.src + i0 * d.ld + d.colk + f4 * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t = 0.f;
        if (f4 * 4 + q < d.kvalid) {
          t = __ldg(b0 + q);
          if (d.mode >= GNNFD_SEG_SUM2) {
            const float y = __ldg(d.src + (int64_t)ixs[TC_BM + r] * d.ld + d.colk + f4 * 4 + q);
            t = d.mode == GNNFD_SEG_DIFF2 ? t - y : t + y;
            if (d.mode == GNNFD_SEG_MEAN3)
              t = (t + __ldg(d.src + (int64_t)ixs[2 * TC_BM + r] * d.ld + d.colk + f4 * 4 + q)) / 3.0f;
          }
        }
        t4[q] = t;
      }
    }
    v[jj] = make_float4(t4[0], t4[1], t4[2], t4[3]);
  }
}

// Issue the global loads of k-block kb of the tile starting at row0 into v.  Rows past the end of the
// matrix read a clamped (valid) row: their results are never stored.
// LEAN (template): the launch has only full, 16-byte aligned k-blocks of DIRECT / GATHER / MEAN3 segments and no peer
// matrices (host-checked in mlp_forward_tc) - the other assembly modes are not compiled into the producers' loop.
template <int LEAN>
__device__ __forceinline__ void tc_load_block(const TcParams &p, const int32_t *ix, int64_t row0, int kb,
                                              int rbase, int f4, float4 (&v)[8]) {
  const KbDesc &d = p.kb[kb];
  if constexpr (!LEAN) {
    if (!d.vec || d.kvalid != TC_KB) {
      float4 t[8];   // only this temporary gets a stack home; v stays in registers
      tc_load_block_generic(p, ix, row0, kb, rbase, f4, t);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) v[jj] = t[jj];
      return;
    }
  }
  const int32_t *ixs = ix + d.seg * 3 * TC_BM + rbase;
  const uint32_t ixa = smem_u32(ixs);           // the index slot lives in shared memory: ld.shared, not generic loads
  const float *base = d.src + d.colk + f4 * 4;
  const int64_t ld = d.ld;
  if (d.mode == GNNFD_SEG_DIRECT) {
    // one 64-bit tile base, then 32-bit row offsets (a tile spans < 2^31 floats: 128 rows x ld)
    const float *tb = base + row0 * ld;
    const int last = (int)min((int64_t)TC_BM, p.a.rows - row0) - 1;
    const uint32_t ldu = (uint32_t)d.ld;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) v[jj] = ldg_f4_hint(tb + (uint32_t)min(rbase + 16 * jj, last) * ldu, p.pol.ld_stream);
  } else if (d.mode == GNNFD_SEG_GATHER) {
    if (!LEAN && p.a.peer_shift > 0) {   // rows of ghost cells come straight from the owning GPU's HBM (P2P over NVLink)
      const uint32_t mask = (1u << p.a.peer_shift) - 1u;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const uint32_t i = (uint32_t)lds_s32(ixa + 64 * jj);
        v[jj] = ldg_f4(p.a.peer_base[i >> p.a.peer_shift] + d.colk + f4 * 4 + (int64_t)(i & mask) * ld);
      }
    } else {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) v[jj] = ldg_f4_hint(base + (int64_t)lds_s32(ixa + 64 * jj) * ld, p.pol.ld_keep);
    }
  } else if (!LEAN && d.mode == GNNFD_SEG_SUM3S) {
    // (s0 a + s1 b) + s2 c: three gathered rows with the signs decoded from the index entries
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float4 y[2], z[2];
      float sg[2][3];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jj = h * 2 + u;
        int32_t r0, r1, r2;
        sum3s_decode(lds_s32(ixa + 64 * jj), r0, sg[u][0]);
        sum3s_decode(lds_s32(ixa + (TC_BM + 16 * jj) * 4), r1, sg[u][1]);
        sum3s_decode(lds_s32(ixa + (2 * TC_BM + 16 * jj) * 4), r2, sg[u][2]);
        v[jj] = ldg_f4_hint(base + (int64_t)r0 * ld, p.pol.ld_keep);
        y[u] = ldg_f4_hint(base + (int64_t)r1 * ld, p.pol.ld_keep);
        z[u] = ldg_f4_hint(base + (int64_t)r2 * ld, p.pol.ld_keep);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 a = v[h * 2 + u];
        const float s0 = sg[u][0], s1 = sg[u][1], s2 = sg[u][2];
        v[h * 2 + u] = make_float4((s0 * a.x + s1 * y[u].x) + s2 * z[u].x, (s0 * a.y + s1 * y[u].y) + s2 * z[u].y,
                                   (s0 * a.z + s1 * y[u].z) + s2 * z[u].z, (s0 * a.w + s1 * y[u].w) + s2 * z[u].w);
      }
    }
  } else if (LEAN || d.mode == GNNFD_SEG_MEAN3) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float4 y[2], z[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jj = h * 2 + u;
        v[jj] = ldg_f4_hint(base + (int64_t)lds_s32(ixa + 64 * jj) * ld, p.pol.ld_keep);
        y[u] = ldg_f4_hint(base + (int64_t)lds_s32(ixa + (TC_BM + 16 * jj) * 4) * ld, p.pol.ld_keep);
        z[u] = ldg_f4_hint(base + (int64_t)lds_s32(ixa + (2 * TC_BM + 16 * jj) * 4) * ld, p.pol.ld_keep);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 t = f4_add(f4_add(v[h * 2 + u], y[u]), z[u]);
        constexpr float third = 1.0f / 3.0f;   // operands are rounded to 2 x bf16 right after: 1 ulp is immaterial
        v[h * 2 + u] = make_float4(t.x * third, t.y * third, t.z * third, t.w * third);
      }
    }
  } else {
    const bool diff = d.mode == GNNFD_SEG_DIFF2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = h * 4 + u;
        v[jj] = ldg_f4_hint(base + (int64_t)lds_s32(ixa + 64 * jj) * ld, p.pol.ld_keep);
        y[u] = ldg_f4_hint(base + (int64_t)lds_s32(ixa + (TC_BM + 16 * jj) * 4) * ld, p.pol.ld_keep);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[h * 4 + u] = diff ? f4_sub(v[h * 4 + u], y[u]) : f4_add(v[h * 4 + u], y[u]);
    }
  }
}

template <bool FP16, int NA>
__device__ __forceinline__ void tc_store_block(uint32_t sA, const float4 (&v)[8], uint32_t off0, bool active) {
  // off0 = sw128(rbase, f4 >> 1) + (f4 & 1) * 8: rows rbase + 16 jj share (r & 7), so row jj is off0 + jj * 2048
  if (active) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      uint32_t h0, l0, h1, l1;
      split2<FP16>(v[jj].x, v[jj].y, h0, l0);
      split2<FP16>(v[jj].z, v[jj].w, h1, l1);
      sts_u2(sA + off0 + jj * 2048, h0, h1);
      if (NA == 2) sts_u2(sA + TC_IMG + off0 + jj * 2048, l0, l1);
    }
  }
}

// -------------------------------------------------------------------------------------- kernel
template <bool FP16, int NA, int NW, bool BWD, int EPI, int LEAN = 0>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const __grid_constant__ TcParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const gnnfd_mlp_args &a = p.a;
  const int a_stages = p.a_stages;
  // LEAN instantiation (host-checked): 3-layer SiLU MLP, 128 outputs, no `mul`, no peer matrices, full aligned k-blocks of
  // DIRECT / GATHER / MEAN3 segments, stash (if any) through TMA - the rarely used variants are not compiled into the
  // roles' hot loops, which together overflow the instruction cache otherwise (no_inst stalls, ncu)
  // (LEAN = 2: the same for a single Linear - n_layers = 1, contiguous operand rows, optional act'(mul) / residual: the
  //  dgrad launches of the backward)
  const int nl = LEAN == 1 ? 3 : LEAN == 2 ? 1 : p.nl;
  // MMA issuer loops: ROLLED in the training / all-purpose instantiations (one thread issues an MMA every 64+ cycles, so
  // loop overhead is free there, while the unrolled form was 20 KB of code in the instruction cache the roles share:
  // forward + stash 194 -> 188 us), UNROLLED in the inference instantiation, whose single-tile launches are a latency chain
  // (2k-cell rollout 0.671 ms/step unrolled vs 0.683 rolled) and which is not instruction-fetch bound
  constexpr int U4 = EPI == 1 ? 4 : 1, U2 = EPI == 1 ? 2 : 1;
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *s_a = smem;                                     // A ring
  float *s_stg = (float *)(s_a + a_stages * TC_STAGE_BYTES);   // output staging, one swizzled 32x16 block per epilogue warp
  int32_t *s_idx = (int32_t *)((uint8_t *)s_stg + TC_STG_BYTES);   // [4 tiles][3 seg][3][128]
  float *s_vec = (float *)(s_idx + TC_IDX_SLOTS * TC_IDX_SLOT);    // b1, b2, b3, ln_w, ln_b
  float2 *s_stat = (float2 *)(s_vec + 5 * TC_H);                   // LayerNorm partials [2 slots][2 halves][128 rows]
  uint64_t *s_bar = (uint64_t *)(s_stat + 4 * TC_BM);
  uint64_t *a_full = s_bar, *a_empty = s_bar + 3, *w_full = s_bar + 6, *w_empty = s_bar + 9;
  uint64_t *w23_full = s_bar + 12, *w23_empty = s_bar + 15;
  uint64_t *acc_full = s_bar + 18, *acc_free = s_bar + 21, *hid_ready = s_bar + 24;   // hid_ready[x_slot*2 + half]
  uint32_t *s_tmem = (uint32_t *)(s_bar + TC_NBAR);
  const int w_slots = p.w_slots;
  uint8_t *s_w = (uint8_t *)(((uintptr_t)(s_tmem + 16) + 1023) & ~(uintptr_t)1023);   // layer-1 W ring
  uint8_t *s_w23 = s_w + w_slots * TC_IMG;                                            // layer-2/3 W ring

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < TC_A_STAGES_MAX; ++i) { mbar_init(&a_full[i], TC_PROD_WARPS); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < TC_W_SLOTS_MAX; ++i) {
      mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
      mbar_init(&w23_full[i], 1); mbar_init(&w23_empty[i], 1);
    }
    for (int i = 0; i < TC_X_SLOTS; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], TC_EPI_GROUP); }
    for (int i = 0; i < 2 * TC_X_SLOTS; ++i) mbar_init(&hid_ready[i], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_MMA_WARP) tmem_alloc(s_tmem, TC_TMEM_COLS);
  if (tid == 32) {
    for (int i = 0; i < a.n_seg; ++i)
      if (a.seg[i].split != nullptr) prefetch_tmap(&p.tm_seg[i]);
    if (EPI == 1) { prefetch_tmap(&p.tm_raw); prefetch_tmap(&p.tm_sum); }
    if (p.save_tma) { prefetch_tmap(&p.tm_save[0]); prefetch_tmap(&p.tm_save[1]); }
    if (p.split_tma) prefetch_tmap(&p.tm_split);
  }
  // programmatic dependent launch (common.cuh): everything above is on-chip set-up that may run under the previous
  // kernel's tail; from here on global memory is read
  // (static_operands: the weights, the bias / LayerNorm vectors and the gather indices were produced long before this
  //  launch, so they too are fetched under the previous kernel's tail and only the roles that touch activations wait)
  const bool early = !BWD && a.static_operands != 0;
  if (!early) pdl_wait();
  for (int i = tid; i < 5 * TC_H; i += TC_THREADS) {
    const int v = i / TC_H, c = i % TC_H;
    const float *src = v == 0 ? a.b1 : v == 1 ? a.b2 : v == 2 ? (nl == 1 ? a.b1 : a.b3) : v == 3 ? a.ln_w : a.ln_b;
    if (BWD && v < 3) src = nullptr;
    const int n = (v == 2 || v >= 3) ? a.n_out : TC_H;
    s_vec[i] = (src && c < n) ? __ldg(src + c) : (v == 3 ? 1.f : 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int64_t n_tiles = (a.rows + TC_BM - 1) / TC_BM;
  const int T = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA
  auto tile_row0 = [&](int j) { return ((int64_t)blockIdx.x + (int64_t)j * gridDim.x) * TC_BM; };

  if (warp >= TC_EPI_WARPS && warp < TC_MMA_WARP) {
    // =============================================================================== producers
    // forward: gathers keep two k-blocks of rows in flight per thread and need the registers; backward chain: the
    // hidden epilogues (saved pre-activation prefetch) need them more than the contiguous dA loads do
    // forward: the gather path wants the registers (72 -> 88); backward chain: the hidden epilogues (saved
    // pre-activation prefetch) want them more than the contiguous dA loads do, so the producers stay at 72 there
    // (measured: edge forward 163 us with 72/88 vs 185 us with 80/72; dgrad chain 242 us with 80/72 vs 253 us)
    // (launches whose gathers are staged by TMA leave the registers to the epilogue instead: the producers then only
    //  load the contiguous DIRECT segment)
    if (!BWD && !(EPI == 1 && p.n_tma > 0)) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int pt = tid - TC_EPI_THREADS;   // 0..255
    const int f4 = pt & 15;                // float4 column inside the 64-wide k-block
    const int rbase = pt >> 4;             // rows rbase + 16 j
    const int NB = T * p.kb1;              // k-blocks this CTA produces, in MMA consumption order
    const uint32_t sa_u32 = smem_u32(s_a), off0 = sw128(rbase, f4 >> 1) + (f4 & 1) * 8;

    auto stage_idx = [&](int j) {          // async copy of tile j's gather indices into slot j & 3
      if (j < T) {
        const int64_t row0 = tile_row0(j);
        int32_t *dst = s_idx + (j & (TC_IDX_SLOTS - 1)) * TC_IDX_SLOT;
        for (int s = 0; s < a.n_seg; ++s) {
          const gnnfd_segment &sg = a.seg[s];
          const int n_idx = sg.mode == GNNFD_SEG_DIRECT ? 0 : sg.mode == GNNFD_SEG_GATHER ? 1
                            : sg.mode >= GNNFD_SEG_MEAN3 ? 3 : 2;
          for (int q = pt; q < n_idx * TC_BM; q += TC_PROD_THREADS) {
            const int ji = q / TC_BM, r = q % TC_BM;
            const int64_t g = row0 + r;
            int32_t *d = dst + (s * 3 + ji) * TC_BM + r;
            if (g < a.rows) cp_async4(d, sg.idx[ji] + g); else *d = 0;
          }
        }
      }
      // L2 prefetch ONE tile ahead of the tile whose loads are about to be issued: close enough that the lines
      // are still resident when read (three tiles ahead, the whole grid streams more than the L2 holds in
      // between and every prefetched line was fetched from DRAM twice - ncu dram__bytes_read)
      const int jp = j - 2;
      if (jp >= 0 && jp < T && !(early && j < 3)) {      // (the first three calls run before griddepcontrol.wait)
        const int64_t prow0 = tile_row0(jp);
        const int64_t nrow = min((int64_t)TC_BM, a.rows - prow0);
        if (pt == 0 && p.direct_tile_bytes > 0)
          bulk_prefetch_l2(a.seg[0].src + prow0 * a.seg[0].ld, (uint32_t)(nrow * a.seg[0].ld * 4));
        // (the saved pre-activations read by the backward chain's hidden epilogues are NOT prefetched: measured
        //  7 % slower with the prefetch - 375 vs 349 us on the edge chain - and 240 MB of extra DRAM reads)
      }
      cp_async_commit();
    };
    // the loads of block b + 1 are issued right after block b is converted; (ij, ikb) tracks the next block to issue
    int ij = 0, ikb = 0;
    auto issue = [&](float4(&v)[8]) {
#if GNNFD_ABL == 2 || GNNFD_ABL == 3   // ablation: no global loads in the producers
      return;
#endif
      while (ij < T && p.kb[ikb].tma) {   // TMA-staged k-blocks have no register stage: skip to the next loaded one
        if (++ikb == p.kb1) { ikb = 0; ++ij; }
      }
      if (ij < T) {
        tc_load_block<LEAN>(p, s_idx + (ij & (TC_IDX_SLOTS - 1)) * TC_IDX_SLOT, tile_row0(ij), ikb, rbase, f4, v);
        if (++ikb == p.kb1) { ikb = 0; ++ij; }
      }
    };
    PROF_DECL;
    int sj = 0, skb = 0, st = 0;
    uint32_t sphase = 1;   // parity to wait on a_empty for: first pass over the ring needs no wait
    auto step = [&](float4(&v)[8]) {
      if (skb == 0) {
        // tile boundary: every index copy issued so far has landed (tiles <= sj + 2) and every producer
        // has issued its loads of tile sj - 1, so that tile's index slot can be refilled with tile sj + 3
        if (sj > 0) { PROF_WAIT(4, cp_async_wait_all(); named_bar_sync(1, TC_PROD_THREADS)); }
        stage_idx(sj + 3);
      }
      // one warp polls the stage's barrier, the others block on a named barrier (see the epilogue's group_wait)
      PROF_WAIT(0, if (warp == TC_EPI_WARPS) mbar_wait(&a_empty[st], sphase); named_bar_sync(12, TC_PROD_THREADS));
      const KbDesc &dk = p.kb[skb];
      if (dk.tma) {
        // TMA gather: 2 parts x 32 groups of 4 rows = 64 gather4 copies per k-block, 8 per producer warp (lanes 0-7):
        // warp pw stages part pw >> 2, row groups (pw & 3) * 8 + lane, straight into the SWIZZLE_128B A image (the
        // tensor map carries the same swizzle; the image is 1024 B aligned and a row group is 4 x 128 B).  Each warp
        // announces its own 8 x 512 B on the stage's barrier, which is also its arrival.
        const int pw = warp - TC_EPI_WARPS;
        // ONE lane issues the warp's 8 copies in a loop: the tensor-map pointer, column and barrier then live in uniform
        // registers and only the four row indices are moved per copy (8 divergent lanes made the compiler serialise the
        // lanes with ~24 instructions per copy: 11.6 % of the kernel's executed instructions, ncu source view)
#if GNNFD_ABL == 8      // ablation: no TMA gathers (stale operands)
        if (lane == 0) mbar_arrive(&a_full[st]);
#else
        if (lane == 0) {
          mbar_expect_tx(&a_full[st], 8 * 512);
          const int part = pw >> 2, rg0 = (pw & 3) * 8;
          const uint32_t ixa = smem_u32(s_idx + (sj & (TC_IDX_SLOTS - 1)) * TC_IDX_SLOT + dk.seg * 3 * TC_BM + rg0 * 4);
          const uint32_t dst = sa_u32 + st * TC_STAGE_BYTES + part * TC_IMG + rg0 * 512;
          const int colp = dk.colk + part * dk.lo_col;
          const CUtensorMap *tm = &p.tm_seg[dk.seg];
#pragma unroll 1
          for (int g = 0; g < 8; ++g) {
            const int4 ix = lds_s32x4(ixa + g * 16);
            tma_gather4_hint(dst + g * 512, tm, colp, ix.x, ix.y, ix.z, ix.w, &a_full[st], p.pol.ld_keep);
          }
        }
        __syncwarp();
#endif
        if (++skb == p.kb1) { skb = 0; ++sj; }
        if (++st == a_stages) { st = 0; sphase ^= 1; }
        return;
      }
      const int ksteps = (skb == p.kb1 - 1) ? p.ksteps1 : 4;
#if GNNFD_ABL != 3        // ablation 3: no conversion / shared-memory stores either
      PROF_WAIT(1, (tc_store_block<FP16, NA>(sa_u32 + st * TC_STAGE_BYTES, v, off0, f4 * 4 < ksteps * 16)));
#endif
      PROF_WAIT(2, fence_proxy_async(); __syncwarp(); if (lane == 0) mbar_arrive(&a_full[st]));
      if (++skb == p.kb1) { skb = 0; ++sj; }
      if (++st == a_stages) { st = 0; sphase ^= 1; }
      PROF_WAIT(3, issue(v));
    };

    // ONE block of loads in flight per thread, one copy of the load / convert / store code: the version that kept two
    // blocks in flight (two unrolled copies, 64 payload registers) measured 13 % SLOWER on the edge block (194 vs 170 us)
    // - the roles' hot loops then exceed the 32 KB L1.5 instruction cache and instruction fetch, not gather latency,
    // paces the producers (profiles/r01_mlp_tc_role_stalls.txt)
    // (Also tried: two HALF blocks of 4 rows per thread, the loads of half h of block b + 1 issued right after half h of
    //  block b is converted, same code size and registers as this - 204 us vs 163 us on the edge block, reverted.)
    float4 v0[8];
#if GNNFD_ABL == 2 || GNNFD_ABL == 3
    for (int i = 0; i < 8; ++i) v0[i] = make_float4(1.f, 2.f, 3.f, 4.f);
#endif
    stage_idx(0); stage_idx(1); stage_idx(2);
    if (early) pdl_wait();      // from here on the segment sources are read
    cp_async_wait_all();
    named_bar_sync(1, TC_PROD_THREADS);
    issue(v0);
#pragma unroll 1
    for (int b = 0; b < NB; ++b) step(v0);
    cp_async_wait_all();
#ifdef GNNFD_TC_PROF
    if (blockIdx.x == 0 && pt == 0) {
      g_tc_prof[12] = clock64() - t_begin; g_tc_prof[13] = prof[0];
      g_tc_prof[14] = prof[1]; g_tc_prof[15] = prof[2]; g_tc_prof[7] = prof[3]; g_tc_prof[11] = prof[4];
    }
#endif
  } else if (warp == TC_WLD_WARP || warp == TC_WLD23_WARP) {
    // ============================================================================ weight loaders
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0 && T > 0) {
      const uint8_t *w1p = (const uint8_t *)a.packed;
      const uint8_t *w2p = w1p + (size_t)p.kb1 * p.w_block_bytes;
      const uint8_t *w3p = w2p + (size_t)2 * p.w_block_bytes;
      const uint32_t w3_part = p.w3_block_bytes / NW;
      const bool first = warp == TC_WLD_WARP;
      const int n_slots = w_slots;
      uint64_t *full = first ? w_full : w23_full, *empty = first ? w_empty : w23_empty;
      uint8_t *ring = first ? s_w : s_w23;
      int slot = 0;
      uint32_t round = 0;
      auto load_unit = [&](const uint8_t *src, uint32_t bytes) {
        if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
#if GNNFD_ABL == 1   // ablation: weights streamed for the first ring pass only
        if (round > 0) { mbar_arrive(&full[slot]); if (++slot == n_slots) { slot = 0; ++round; } return; }
#endif
        mbar_expect_tx(&full[slot], bytes);
        bulk_g2s(ring + slot * TC_IMG, src, bytes, &full[slot]);
        if (++slot == n_slots) { slot = 0; ++round; }
      };
      // same order as the issuer that drains this ring (see there)
      if (first) {
        for (int j = 0; j < T; ++j)
          for (int kb = 0; kb < p.kb1; ++kb)
            for (int part = 0; part < NW; ++part) load_unit(w1p + (size_t)kb * p.w_block_bytes + part * TC_IMG, TC_IMG);
      } else if (nl > 1) {
        for (int j = 0; j < T; ++j) {
          for (int kb = 0; kb < 2; ++kb)
            for (int part = 0; part < NW; ++part) load_unit(w2p + (size_t)kb * p.w_block_bytes + part * TC_IMG, TC_IMG);
          for (int kb = 0; kb < 2; ++kb)
            for (int part = 0; part < NW; ++part) load_unit(w3p + (size_t)kb * p.w3_block_bytes + part * w3_part, w3_part);
        }
      }
    }
    __syncwarp();
  } else if (warp == TC_MMA_WARP) {
    // ====================================================================== layer-1 MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0 && T > 0) {
      constexpr uint32_t IDESC_H = make_idesc(FP16 ? 0 : 1, TC_H);
      constexpr uint32_t idesc1 = IDESC_H;
      int st = 0, ws = 0;
      uint32_t a_round = 0, w_round = 0;
      PROF_DECL;
      // consume one weight unit: returns its descriptor; release with umma_commit(&w_empty[slot])
      auto w_acquire = [&](int &slot) {
        slot = ws;
        PROF_WAIT(2, mbar_wait(&w_full[slot], w_round & 1));
        tc_fence_after();
        if (++ws == w_slots) { ws = 0; ++w_round; }
        return make_desc(smem_u32(s_w + slot * TC_IMG));
      };
      // The tensor pipe runs MMAs in issue order and layers 2 / 3 of the tiles ahead sit on the epilogue's critical path
      // (and hold the single Y region): layer 1 therefore keeps at most l1_depth k-blocks of MMAs queued, so a layer-2 / 3
      // k-block issued by the other issuer waits behind <= l1_depth x 12 MMAs instead of a whole tile's 72.
      const int depth = p.l1_depth;
      int pst = 0;               // stage of the k-block issued `depth` k-blocks ago
      uint32_t p_round = 0, n_issued = 0;
      for (int j = 0; j < T; ++j) {
        const int xs = j % TC_X_SLOTS, n = j / TC_X_SLOTS;
        if (n >= 1) { PROF_WAIT(3, mbar_wait(&acc_free[xs], (n - 1) & 1)); tc_fence_after(); }
        const uint32_t d = tmem_base + xs * 128;
        for (int kb = 0; kb < p.kb1; ++kb) {
          if (depth > 0 && n_issued >= (uint32_t)depth) {
            mbar_wait(&a_empty[pst], p_round & 1);       // its MMAs have completed (commit on a_empty)
            if (++pst == a_stages) { pst = 0; ++p_round; }
          }
          ++n_issued;
          PROF_WAIT(4, mbar_wait(&a_full[st], a_round & 1));
          tc_fence_after();
          const int ksteps = (kb == p.kb1 - 1) ? p.ksteps1 : 4;
          const uint64_t ah = make_desc(smem_u32(s_a + st * TC_STAGE_BYTES)), al = ah + (TC_IMG >> 4);
          int slot;
          uint64_t wb = w_acquire(slot);
#if GNNFD_ABL != 4
#pragma unroll U4
          for (int k = 0; k < ksteps; ++k) {
            umma_ss(d, ah + 2 * k, wb + 2 * k, idesc1, (kb | k) != 0);
            if (NA == 2) umma_ss(d, al + 2 * k, wb + 2 * k, idesc1, 1);
          }
#endif
          umma_commit(&w_empty[slot]);
          if (NW == 2) {
            wb = w_acquire(slot);
#if GNNFD_ABL != 4
#pragma unroll U4
            for (int k = 0; k < ksteps; ++k) umma_ss(d, ah + 2 * k, wb + 2 * k, idesc1, 1);
#endif
            umma_commit(&w_empty[slot]);
          }
          umma_commit(&a_empty[st]);
          if (++st == a_stages) { st = 0; ++a_round; }
        }
        umma_commit(&acc_full[xs]);
      }
#ifdef GNNFD_TC_PROF
      if (blockIdx.x == 0) {
        g_tc_prof[0] = clock64() - t_begin;
        for (int i = 2; i < 5; ++i) g_tc_prof[i] = prof[i];
        g_tc_prof[6] = (unsigned long long)T;
      }
#endif
    }
    __syncwarp();
  } else if (warp == TC_MMA23_WARP) {
    // ==================================================================== layer-2/3 MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0 && T > 0 && nl > 1) {
      constexpr uint32_t IDESC_H = make_idesc(FP16 ? 0 : 1, TC_H);
      const uint32_t idesc3 = make_idesc(FP16 ? 0 : 1, LEAN ? TC_H : p.n3);
      int ws = 0;
      uint32_t w_round = 0;
      PROF_DECL;
      auto w_acquire = [&](int &slot) {
        slot = ws;
        PROF_WAIT(1, mbar_wait(&w23_full[slot], w_round & 1));
        tc_fence_after();
        if (++ws == w_slots) { ws = 0; ++w_round; }
        return make_desc(smem_u32(s_w23 + slot * TC_IMG));
      };
      // A = the in-place converted accumulator region: per 16 fp32 columns, 8 columns of hi pairs | 8 of lo pairs.
      // Layer 2 of tile j overwrites Y after layer 3 of tile j - 1 has read it: same thread, in-order tensor pipe.
      // (loop unrolling: see U4 / U2 above)
#pragma unroll 1
      for (int j = 0; j < T; ++j) {
        const int xs = j % TC_X_SLOTS;
        const uint32_t xr = tmem_base + xs * 128, yr = tmem_base + TC_Y_COL;
#pragma unroll U2
        for (int layer = 2; layer <= 3; ++layer) {
          const uint32_t a_reg = layer == 2 ? xr : yr, d = layer == 2 ? yr : xr;
          const uint32_t idesc = layer == 2 ? IDESC_H : idesc3;
#pragma unroll U2
          for (int kb = 0; kb < 2; ++kb) {
            PROF_WAIT(5, mbar_wait(&hid_ready[xs * 2 + kb], layer == 2 ? 0 : 1));
            tc_fence_after();
            int slot;
            uint64_t wb = w_acquire(slot);
#if GNNFD_ABL != 4
#pragma unroll U4
            for (int k = 0; k < 4; ++k) {
              const uint32_t ta = a_reg + (kb * 4 + k) * 16;
              umma_ts(d, ta, wb + 2 * k, idesc, (kb | k) != 0);
              if (NA == 2) umma_ts(d, ta + 8, wb + 2 * k, idesc, 1);
            }
#endif
            umma_commit(&w23_empty[slot]);
            if (NW == 2) {
              wb = w_acquire(slot);
#if GNNFD_ABL != 4
#pragma unroll U4
              for (int k = 0; k < 4; ++k) umma_ts(d, a_reg + (kb * 4 + k) * 16, wb + 2 * k, idesc, 1);
#endif
              umma_commit(&w23_empty[slot]);
            }
          }
          umma_commit(&acc_full[xs]);
        }
      }
#ifdef GNNFD_TC_PROF
      if (blockIdx.x == 0) { g_tc_prof[1] = prof[1]; g_tc_prof[5] = prof[5]; }
#endif
    }
    __syncwarp();
  } else {
    // ================================================================================ epilogue
    // warp = (tile group grp, column half eh, lane quarter q4): thread = row, 64 of the 128 columns of every tile of its
    // group; half eh of a hidden layer's output is exactly k-block eh of the next layer's operand.  All TMEM traffic is
    // in 16-column groups: 16 fp32 accumulator columns are replaced in place by 8 columns of hi pairs + 8 of lo pairs.
    if (BWD || (EPI == 1 && p.n_tma > 0)) asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    if (early) pdl_wait();      // residual / mul loads and every store of this role follow
    const int grp = warp >> 3, q4 = warp & 3, eh = (warp >> 2) & 1;
    const int erow = q4 * 32 + lane;                       // row of the tile owned by this thread
    const uint32_t stg = smem_u32(s_stg + warp * (32 * 16));   // this warp's 32 x 16 staging block (XOR-swizzled)
    // second block (split shadow: 32 rows x 32 B of hi pairs | 32 x 32 B of lo pairs), behind the weight rings
    const uint32_t stg2 = smem_u32(s_w23 + w_slots * TC_IMG) + warp * 2048;
    const uint32_t vec = smem_u32(s_vec), stat = smem_u32(s_stat);
    const int rr = lane >> 2, c4 = lane & 3;               // copy-out mapping: 8 rows x 64 B per instruction
    // Waiting on an mbarrier polls (try_wait wakes every ~100 cycles: 4 instructions per poll per warp - with 8 warps of
    // a group, 8 producer warps and the issuers all polling, 29 % of the kernel's executed instructions were polls, ncu
    // source view).  So ONE warp of the group polls and the other seven block on a named barrier, which costs no issue
    // slots at all.
    auto group_wait = [&](uint64_t *bar, uint32_t parity) {
      if ((warp & 7) == 0) mbar_wait(bar, parity);
      named_bar_sync(10 + grp, TC_EPI_GROUP * 32);
    };
    PROF_DECL;
    for (int j = grp; j < T; j += 2) {
      const int xs = j % TC_X_SLOTS, n = j / TC_X_SLOTS;
      const int64_t row0 = tile_row0(j);
      const uint32_t lanes = (uint32_t)(q4 * 32) << 16;
      const uint32_t xr = tmem_base + lanes + xs * 128 + eh * 64, yr = tmem_base + lanes + TC_Y_COL + eh * 64;
      // (L2 prefetch of the rows this tile's later phases read - the second hidden layer's saved pre-activations, the
      //  residual - issued from here measured no gain in an A/B on one box: edge MLP backward 859 vs 834 us)
      // ---- hidden layers: accumulator -> +bias, act -> hi/lo pairs, written back in place
      for (int layer = 0; layer < nl - 1; ++layer) {
        const uint32_t reg = layer == 0 ? xr : yr;
        const uint32_t bias = vec + (layer * TC_H + eh * 64) * 4;
        float *save = EPI == 1 ? nullptr : (layer == 0 ? a.save_a1 : a.save_a2);   // EPI 1: no training stash
        const bool save_tma = EPI != 1 && p.save_tma && save != nullptr;              // uniform
        if (save != nullptr) save = (row0 + erow < a.rows) ? save + (size_t)(row0 + erow) * TC_H + eh * 64 : nullptr;
        // backward chain: the saved pre-activation of this layer, fetched one 16-column group ahead of its use.  The
        // loads are COALESCED (8 rows x 64 B per warp instruction, like the final copy-out) and transposed to
        // thread = row through the warp's staging block: a thread-per-row load touches 32 cache lines per instruction
        // and the LSU, not HBM, then paces the epilogue.  Rows past the end read a clamped row; never stored.
        const float *hm = nullptr;
        // 16-column group cg of this layer's saved pre-activations -> the warp's second staging block, as asynchronous
        // 16-byte copies (no registers in flight: with register prefetch the compiler spilled the loaded values, and a
        // spill store waits for its load - the whole DRAM latency sat right behind the prefetch)
        auto hm_copy = [&](int cg) {
#pragma unroll
          for (int jr = 0; jr < 4; ++jr) {
            const int rl = jr * 8 + rr;
            cp_async16_hint(stg2 + (rl * 16 + ((c4 ^ ((rl >> 1) & 3)) << 2)) * 4,
                            hm + cg * 16 + (size_t)min(row0 + q4 * 32 + rl, a.rows - 1) * TC_H, p.pol.ld_stream);
          }
          cp_async_commit();
        };
        if (BWD) {
          hm = (layer == 0 ? a.hid_mul1 : a.hid_mul2) + eh * 64 + c4 * 4;
          hm_copy(0);
        }
        PROF_WAIT(0, group_wait(&acc_full[xs], (3 * n + layer) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[16];
          if (BWD) {
            // this group's copies have landed (every lane waits for its own, then the warp meets): read the thread's own
            // row, and start the next group's copies into the same block once every lane has read
            cp_async_wait_all();
            __syncwarp();
            float4 mm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) mm[i] = lds_f4(stg2 + (lane * 16 + ((i ^ ((lane >> 1) & 3)) << 2)) * 4);
            __syncwarp();
            if (c < 3) hm_copy(c + 1);
            tmem_ld16(reg + c * 16, v);
            const bool silu = LEAN || a.act == GNNFD_ACT_SILU;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 m = mm[i];
              float d0, d1, d2, d3;
              if (silu) { dsilu2(m.x, m.y, d0, d1); dsilu2(m.z, m.w, d2, d3); }
              else { d0 = dtanh(m.x); d1 = dtanh(m.y); d2 = dtanh(m.z); d3 = dtanh(m.w); }
              mul2(v[4 * i], v[4 * i + 1], d0, d1);
              mul2(v[4 * i + 2], v[4 * i + 3], d2, d3);
            }
          } else {
            tmem_ld16(reg + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = lds_f4(bias + (c * 16 + i) * 4);
              add2(v[i], v[i + 1], b4.x, b4.y);
              add2(v[i + 2], v[i + 3], b4.z, b4.w);
            }
            if constexpr (EPI == 2) {
              // training-mode dropout: a dropped unit's PRE-activation becomes GNNFD_DROPPED, whose SiLU and SiLU' are
              // both -0 - the stash written below then carries the mask into the backward kernels
              const uint32_t rh = dropout_row_hash(p.drop_key[layer], (uint32_t)(row0 + erow));
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (dropout_hash(rh, (uint32_t)(eh * 64 + c * 16 + i)) < p.drop_thresh) v[i] = GNNFD_DROPPED;
            }
          }
          if (save_tma) {
            // training stash (pre-activation / dA) of this 32 x 16 block: staged in the warp's SWIZZLE_64B block and
            // written by ONE TMA tensor store (rows past the end of the matrix are clipped by the tensor map)
            if (lane == 0) bulk_wait_read0();      // the previous group's TMA store has finished reading the block
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts_f4(stg + (lane * 16 + ((i ^ ((lane >> 1) & 3)) << 2)) * 4,
                     make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d_hint(&p.tm_save[layer], stg, eh * 64 + c * 16, (int)(row0 + q4 * 32), p.pol.st_stash);
              bulk_commit();
            }
          } else if (!LEAN && save != nullptr) {   // (fallback: thread = row, 64 B per store group)
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4 *>(save + c * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
#if GNNFD_ABL == 5
          if (true) {}
#else
          if (BWD) { /* linear chain: no activation */ }
#endif
          else if (LEAN || a.act == GNNFD_ACT_SILU) act16<GNNFD_ACT_SILU>(v); else act16<GNNFD_ACT_TANH>(v);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) split2<FP16>(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
          tmem_st8(reg + c * 16, hi);
          if (NA == 2) tmem_st8(reg + c * 16 + 8, lo);
        }
        // this half's 64 columns = one k-block of the next layer's operand
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hid_ready[xs * 2 + eh]);
      }
      // ---- final epilogue
      PROF_WAIT(1, group_wait(&acc_full[xs], (nl * n + nl - 1) & 1));
      tc_fence_after();
      if constexpr (EPI == 1) {
        // ---- fast final epilogue (inference: no stash, no mul, n_out = 128).  Thread = row: bias -> LayerNorm -> affine in
        // registers, the 32 x 16 block of the warp staged in the SWIZZLE_64B layout and written by ONE TMA tensor store
        // (or reduce-add: the residual add is then done by the memory system and the residual never enters the SM).
        const int epi = p.epi;
        float mean = 0.f, rstd = 1.f;
        if (a.has_ln) {
          float shift = 0.f, s4[4] = {0.f, 0.f, 0.f, 0.f}, q4s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            float acc[16];
            tmem_ld16(xr + c * 16, acc);
            const uint32_t b3 = vec + (2 * TC_H + eh * 64 + c * 16) * 4;
            if (c == 0) shift = acc[0] + lds_f4(b3).x;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4 b4 = lds_f4(b3 + i * 4);
              sub2(b4.x, b4.y, shift, shift); sub2(b4.z, b4.w, shift, shift);
              float d0 = acc[i], d1 = acc[i + 1], d2 = acc[i + 2], d3 = acc[i + 3];
              add2(d0, d1, b4.x, b4.y); add2(d2, d3, b4.z, b4.w);
              add2(s4[0], s4[1], d0, d1); add2(s4[2], s4[3], d2, d3);
              float t0 = d0, t1 = d1, t2 = d2, t3 = d3;
              fma2(t0, t1, d0, d1, q4s[0], q4s[1]); fma2(t2, t3, d2, d3, q4s[2], q4s[3]);
              q4s[0] = t0; q4s[1] = t1; q4s[2] = t2; q4s[3] = t3;
            }
          }
          const float sh = (s4[0] + s4[1]) + (s4[2] + s4[3]), qh = (q4s[0] + q4s[1]) + (q4s[2] + q4s[3]);
          const float md = sh * (1.0f / 64.0f);
          const float mean_h = shift + md, m2_h = fmaxf(qh - sh * md, 0.f);
          sts_f2(stat + ((grp * 2 + eh) * TC_BM + erow) * 8, make_float2(mean_h, m2_h));
          named_bar_sync(2 + grp * 4 + q4, 64);         // the two warps that share these 32 rows of this tile
          const float2 o = lds_f2(stat + ((grp * 2 + (eh ^ 1)) * TC_BM + erow) * 8);
          const float dm = mean_h - o.x;
          mean = 0.5f * (mean_h + o.x);
          rstd = rsqrtf((m2_h + o.y + dm * dm * 32.0f) * (1.0f / TC_H) + a.ln_eps);
        }
        const float nmr = -mean * rstd;
        const int64_t grow = row0 + erow;
        const int64_t lrow = grow < a.rows ? grow : a.rows - 1;      // clamped (valid) row for loads
        const int trow = (int)(row0 + q4 * 32);                      // first row of this warp's block
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col0 = eh * 64 + c * 16;
          float4 res[4];
          if (epi & EPI_LDRES) {
#pragma unroll
            for (int i = 0; i < 4; ++i) res[i] = ldg_f4_hint(a.residual + (size_t)lrow * TC_H + col0 + i * 4, p.pol.ld_stream);
          }
          float acc[16];
          tmem_ld16(xr + c * 16, acc);
          if (c == 3) {   // last TMEM read of this tile: the slot may be overwritten by the next L1
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[xs]);
          }
          const uint32_t b3 = vec + (2 * TC_H + col0) * 4, lw = vec + (3 * TC_H + col0) * 4, lb = vec + (4 * TC_H + col0) * 4;
          float o[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = lds_f4(b3 + i * 16), w4 = lds_f4(lw + i * 16), g4 = lds_f4(lb + i * 16);
            float x0 = acc[4 * i], x1 = acc[4 * i + 1], x2 = acc[4 * i + 2], x3 = acc[4 * i + 3];
            add2(x0, x1, b4.x, b4.y); add2(x2, x3, b4.z, b4.w);
            fma2(x0, x1, rstd, rstd, nmr, nmr); fma2(x2, x3, rstd, rstd, nmr, nmr);
            fma2(x0, x1, w4.x, w4.y, g4.x, g4.y); fma2(x2, x3, w4.z, w4.w, g4.z, g4.w);
            o[4 * i] = x0; o[4 * i + 1] = x1; o[4 * i + 2] = x2; o[4 * i + 3] = x3;
          }
          if (epi & EPI_LDRES) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              add2(o[4 * i], o[4 * i + 1], res[i].x, res[i].y);
              add2(o[4 * i + 2], o[4 * i + 3], res[i].z, res[i].w);
            }
          }
          const bool both = (epi & EPI_SPLIT) && p.split_tma && (epi & (EPI_ST_RAW | EPI_RED_SUM | EPI_LDRES));
          if ((epi & EPI_SPLIT) && p.split_tma) {
            // 16-bit hi | lo shadow for the next block's TMA gathers: staged row-major (32 B of hi pairs per row, then the lo
            // pairs) and written by two TMA tensor stores - thread-per-row 16-byte stores at a 512 B stride cost 32 cache
            // lines per instruction and made the LSU pace this epilogue.  Bulk groups alternate between the two staging
            // blocks, so "at most one group still reading" means the block about to be overwritten is free.
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split2<FP16>(o[2 * i], o[2 * i + 1], hi[i], lo[i]);
            if (lane == 0) { if (both) bulk_wait_read1(); else bulk_wait_read0(); }
            __syncwarp();
            const uint32_t sp = stg2 + lane * 32;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sp), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sp + 16), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sp + 1024), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sp + 1040), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tm_split, stg2, col0, trow);
              tma_store_2d(&p.tm_split, stg2 + 1024, TC_H + col0, trow);
              bulk_commit();
            }
          } else if ((epi & EPI_SPLIT) && grow < a.rows) {   // (fallback: thread = row stores)
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split2<FP16>(o[2 * i], o[2 * i + 1], hi[i], lo[i]);
            uint16_t *sp = (uint16_t *)a.out_split + (size_t)grow * (2 * TC_H) + col0;
            *reinterpret_cast<uint4 *>(sp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(sp + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4 *>(sp + TC_H) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4 *>(sp + TC_H + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
#if GNNFD_ABL == 9      // ablation: nothing leaves the fast epilogue through the staging block
          if (false) {
#else
          if (epi & (EPI_ST_RAW | EPI_RED_SUM | EPI_LDRES)) {
#endif
            // the previous group's TMA copy must have finished READING the staging block before it is overwritten
            if (lane == 0) { if (both) bulk_wait_read1(); else bulk_wait_read0(); }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts_f4(stg + (lane * 16 + ((i ^ ((lane >> 1) & 3)) << 2)) * 4,
                     make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (epi & EPI_ST_RAW) tma_store_2d_hint(&p.tm_raw, stg, col0, trow, p.pol.st_raw);
              if (epi & EPI_RED_SUM) tma_reduce_add_2d_hint(&p.tm_sum, stg, col0, trow, p.pol.st_out);
              if (epi & EPI_LDRES) tma_store_2d_hint(&p.tm_sum, stg, col0, trow, p.pol.st_out);
              bulk_commit();
            }
          }
        }
      } else
#if GNNFD_ABL == 7
      if (true) { __syncwarp(); if (lane == 0) mbar_arrive(&acc_free[xs]); } else
#endif
      if (LEAN || a.n_out == TC_H) {
        if (p.save_tma) {   // the last stash store must have finished reading the staging block the copy-out reuses
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        float mean = 0.f, rstd = 1.f;
        if (a.has_ln) {
          // this half: shifted single pass over 64 columns -> (mean_h, M2_h); halves merged with Chan's formula
          float shift = 0.f, s4[4] = {0.f, 0.f, 0.f, 0.f}, q4s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            float acc[16];
            tmem_ld16(xr + c * 16, acc);
            const uint32_t b3 = vec + (2 * TC_H + eh * 64 + c * 16) * 4;
            if (c == 0) shift = acc[0] + lds_f4(b3).x;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = lds_f4(b3 + i * 4);
              float d0 = acc[i], d1 = acc[i + 1], d2 = acc[i + 2], d3 = acc[i + 3];
              add2(d0, d1, b4.x, b4.y); add2(d2, d3, b4.z, b4.w);
              sub2(d0, d1, shift, shift); sub2(d2, d3, shift, shift);
              add2(s4[0], s4[1], d0, d1); add2(s4[2], s4[3], d2, d3);
              float t0 = d0, t1 = d1, t2 = d2, t3 = d3;
              fma2(t0, t1, d0, d1, q4s[0], q4s[1]); fma2(t2, t3, d2, d3, q4s[2], q4s[3]);
              q4s[0] = t0; q4s[1] = t1; q4s[2] = t2; q4s[3] = t3;
            }
          }
          const float sh = (s4[0] + s4[1]) + (s4[2] + s4[3]), qh = (q4s[0] + q4s[1]) + (q4s[2] + q4s[3]);
          const float md = sh * (1.0f / 64.0f);
          const float mean_h = shift + md, m2_h = fmaxf(qh - sh * md, 0.f);
          sts_f2(stat + ((grp * 2 + eh) * TC_BM + erow) * 8, make_float2(mean_h, m2_h));
          named_bar_sync(2 + grp * 4 + q4, 64);         // the two warps that share these 32 rows of this tile
          const float2 o = lds_f2(stat + ((grp * 2 + (eh ^ 1)) * TC_BM + erow) * 8);
          const float dm = mean_h - o.x;
          mean = 0.5f * (mean_h + o.x);
          rstd = rsqrtf((m2_h + o.y + dm * dm * 32.0f) * (1.0f / TC_H) + a.ln_eps);
          if (a.save_rstd != nullptr && eh == 0 && row0 + erow < a.rows) a.save_rstd[row0 + erow] = rstd;
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col0 = eh * 64 + c * 16;
          // residual rows of this group, coalesced mapping (8 rows x 64 B), requested before the TMEM read
          float4 res[4];
#if GNNFD_ABL == 6
          if (false) {
#else
          if (a.out_sum) {
#endif
#pragma unroll
            for (int jr = 0; jr < 4; ++jr) {
              const int64_t g = min(row0 + q4 * 32 + jr * 8 + rr, a.rows - 1);
              res[jr] = ldg_f4_hint(a.residual + (size_t)g * TC_H + col0 + c4 * 4, p.pol.ld_stream);
            }
          }
          float acc[16];
          tmem_ld16(xr + c * 16, acc);
          if (c == 3) {   // last TMEM read of this tile: the slot may be overwritten by the next L1
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[xs]);
          }
          // staged value = normalised row (x-hat); the LayerNorm affine is applied in the coalesced copy-out, where the
          // training stash of x-hat is also written.  Staging block: row r, 16-byte chunk i at chunk i ^ ((r >> 1) & 3).
          const uint32_t b3 = vec + (2 * TC_H + col0) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = lds_f4(b3 + i * 16);
            float4 o = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            add2(o.x, o.y, b4.x, b4.y); add2(o.z, o.w, b4.z, b4.w);
            sub2(o.x, o.y, mean, mean); sub2(o.z, o.w, mean, mean);
            mul2(o.x, o.y, rstd, rstd); mul2(o.z, o.w, rstd, rstd);
            sts_f4(stg + (lane * 16 + ((i ^ ((lane >> 1) & 3)) << 2)) * 4, o);
          }
          __syncwarp();
          const float4 w4 = lds_f4(vec + (3 * TC_H + col0 + c4 * 4) * 4), g4 = lds_f4(vec + (4 * TC_H + col0 + c4 * 4) * 4);
#pragma unroll
          for (int jr = 0; jr < 4; ++jr) {
            const int rl = jr * 8 + rr;
            const int64_t g = row0 + q4 * 32 + rl;
#if GNNFD_ABL == 6
            if (g < 0) {
#else
            if (g < a.rows) {
#endif
              float4 o = lds_f4(stg + (rl * 16 + ((c4 ^ ((rl >> 1) & 3)) << 2)) * 4);
              const size_t off = (size_t)g * TC_H + col0 + c4 * 4;
              if (a.save_xhat) stg_f4_hint(a.save_xhat + off, o, p.pol.st_stash);
              fma2(o.x, o.y, w4.x, w4.y, g4.x, g4.y);
              fma2(o.z, o.w, w4.z, w4.w, g4.z, g4.w);
              if (LEAN != 1 && a.mul) {
                float4 m = ldg_f4(a.mul + off);
                if (a.mul_mode == 1) { m.x = dsilu(m.x); m.y = dsilu(m.y); m.z = dsilu(m.z); m.w = dsilu(m.w); }
                else if (a.mul_mode == 2) { m.x = dtanh(m.x); m.y = dtanh(m.y); m.z = dtanh(m.z); m.w = dtanh(m.w); }
                o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
              }
              if (a.out_raw) stg_f4_hint(a.out_raw + off, o, p.pol.st_raw);
              if (a.out_sum) {
                float4 r4 = res[jr];
                add2(r4.x, r4.y, o.x, o.y); add2(r4.z, r4.w, o.z, o.w);
                stg_f4_hint(a.out_sum + off, r4, p.pol.st_out);
              }
            }
          }
          __syncwarp();
        }
      } else {
        if (eh == 0) {
          float acc[16];
          tmem_ld16(xr, acc);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[xs]);
          const int64_t g = row0 + erow;
          if (g < a.rows) {
#pragma unroll
            for (int o = 0; o < 16; ++o) {
              if (o < a.n_out) {
                float vv = acc[o] + s_vec[2 * TC_H + o];
                const size_t off = (size_t)g * a.n_out + o;
                if (a.mul) vv *= __ldg(a.mul + off);
                if (a.out_raw) a.out_raw[off] = vv;
                if (a.out_sum) a.out_sum[off] = __ldg(a.residual + off) + vv;
              }
            }
          }
        } else {
          if (lane == 0) mbar_arrive(&acc_free[xs]);
        }
      }
    }
    if ((EPI == 1 || p.save_tma) && lane == 0) bulk_wait0();      // all TMA stores of this warp have completed
#ifdef GNNFD_TC_PROF
    if (blockIdx.x == 0 && tid == 0) {
      g_tc_prof[8] = clock64() - t_begin;
      g_tc_prof[9] = prof[0];
      g_tc_prof[10] = prof[1];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ------------------------------------------------------------------------------ weight packing
// image element (n, k) of one part: 16-bit value at sw128(n, (k % 64) / 8) + (k % 8) * 2
struct PackJob {   // one weight matrix -> its packed operand images
  const float *w;
  int ld_n, ld_k, n_rows_w, K, n_img_rows, n_kblocks;
  uint8_t *out;
  uint32_t block_bytes;
};
struct PackJobs { PackJob job[3]; int n_jobs; int nw; };

// image element (n, k) of one part: 16-bit value at sw128(n, (k % 64) / 8) + (k % 8) * 2.
// One launch packs every matrix of an MLP: one thread per (matrix, k-block, image row, 16-byte chunk).
template <bool FP16>
__global__ void pack_weights_kernel(const __grid_constant__ PackJobs jobs) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int q = 0; q < jobs.n_jobs; ++q) {
    const PackJob &jb = jobs.job[q];
    const int total = jb.n_kblocks * jb.n_img_rows * 8;
    if (i < total) {
      const int c = i & 7, n = (i >> 3) % jb.n_img_rows, kb = (i >> 3) / jb.n_img_rows;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = kb * TC_KB + c * 8 + t * 2;
        const float x0 = (n < jb.n_rows_w && k < jb.K) ? jb.w[(size_t)n * jb.ld_n + (size_t)k * jb.ld_k] : 0.f;
        const float x1 = (n < jb.n_rows_w && k + 1 < jb.K) ? jb.w[(size_t)n * jb.ld_n + (size_t)(k + 1) * jb.ld_k] : 0.f;
        split2<FP16>(x0, x1, hi[t], lo[t]);
      }
      uint8_t *blk = jb.out + (size_t)kb * jb.block_bytes;
      const uint32_t off = sw128(n, c);
      *reinterpret_cast<uint4 *>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (jobs.nw == 2) *reinterpret_cast<uint4 *>(blk + jb.block_bytes / 2 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      return;
    }
    i -= total;
  }
}

int tc_profile_read(unsigned long long *out16) {
#ifndef GNNFD_TC_PROF
  set_error("gnnfd_tc_profile_read: library built without -DGNNFD_TC_PROF");
  return GNNFD_E_UNSUPPORTED;
#else
  GNNFD_CUDA(cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(unsigned long long) * 16));
  return GNNFD_OK;
#endif
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
  p.nl = a->n_layers == 1 ? 1 : 3;
  if (a->n_layers != 0 && a->n_layers != 1 && a->n_layers != 3) return GNNFD_E_UNSUPPORTED;
  if (p.nl == 1 && (a->n_out != TC_H || a->has_ln)) return GNNFD_E_UNSUPPORTED;
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
  const bool strided = (p.nl == 1 || a->bwd_chain) && (a->w1_ld_n != 0 || a->w1_ld_k != 0);
  const int ld_n1 = strided ? a->w1_ld_n : a->k_in, ld_k1 = strided ? a->w1_ld_k : 1;
  const int rows1 = ((p.nl == 1 || a->bwd_chain) && a->w1_rows > 0) ? a->w1_rows : TC_H;
  const int ld_n2 = a->bwd_chain ? a->w2_ld_n : TC_H, ld_k2 = a->bwd_chain ? a->w2_ld_k : 1;
  const int ld_n3 = a->bwd_chain ? a->w3_ld_n : TC_H, ld_k3 = a->bwd_chain ? a->w3_ld_k : 1;
  const int rows3 = (a->bwd_chain && a->w3_rows > 0) ? a->w3_rows : a->n_out;
  PackJobs jobs{};
  jobs.nw = m.nw;
  jobs.job[0] = PackJob{a->w1, ld_n1, ld_k1, rows1, a->k_in, TC_H, p.kb1, out, p.w_block_bytes};
  jobs.n_jobs = 1;
  if (p.nl == 3) {
    jobs.job[1] = PackJob{a->w2, ld_n2, ld_k2, TC_H, TC_H, TC_H, 2, o2, p.w_block_bytes};
    jobs.job[2] = PackJob{a->w3, ld_n3, ld_k3, rows3, TC_H, p.n3, 2, o3, p.w3_block_bytes};
    jobs.n_jobs = 3;
  }
  int total = 0;
  for (int q = 0; q < jobs.n_jobs; ++q) total += jobs.job[q].n_kblocks * jobs.job[q].n_img_rows * 8;
  if (m.fp16) launch_pdl(pack_weights_kernel<true>, dim3((total + 255) / 256), dim3(256), 0, stream, jobs);
  else launch_pdl(pack_weights_kernel<false>, dim3((total + 255) / 256), dim3(256), 0, stream, jobs);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

// ------------------------------------------------------------------------------------ tensor maps
// cuTensorMapEncodeTiled is a pure host-side encoder in libcuda; it is resolved through the runtime so the library
// keeps linking against cudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}
// 2-D row-major tensor [outer, inner] of `esize`-byte elements, row stride `stride_bytes`, box [box_outer, box_inner]
static int make_tmap_2d(CUtensorMap *m, CUtensorMapDataType dt, const void *base, uint64_t inner, uint64_t outer,
                        uint64_t stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return GNNFD_E_CUDA; }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return GNNFD_E_CUDA; }
  return GNNFD_OK;
}

int mlp_forward_tc(const gnnfd_mlp_args *a, cudaStream_t stream) {
  TcMode m;
  if (!tc_mode(a->precision, m)) { set_error("mlp_forward_tc: bad precision"); return GNNFD_E_BADARG; }
  TcParams p{};
  p.a = *a;
  p.pol = l2_policies();
  if (a->dropout_p > 0.f) {
    const double t = (double)a->dropout_p * 4294967296.0;
    p.drop_thresh = t >= 4294967295.0 ? 4294967295u : (t < 1.0 ? 1u : (uint32_t)t);
    p.drop_key[0] = dropout_layer_key(a->dropout_seed, 0);
    p.drop_key[1] = dropout_layer_key(a->dropout_seed, 1);
  }
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
  if (p.kb1 > TC_MAX_KB) { set_error("mlp_forward_tc: k_in too large"); return GNNFD_E_UNSUPPORTED; }
  for (int kb = 0, seg = 0, seg_k0 = 0; kb < p.kb1; ++kb) {
    const int k0 = kb * TC_KB;
    while (seg + 1 < a->n_seg && k0 >= seg_k0 + a->seg[seg].width) { seg_k0 += a->seg[seg].width; ++seg; }
    const gnnfd_segment &sg = a->seg[seg];
    KbDesc &d = p.kb[kb];
    const int kloc = k0 - seg_k0;
    d.src = sg.src; d.ld = sg.ld; d.colk = sg.col + kloc; d.mode = sg.mode; d.seg = seg;
    d.kvalid = sg.width - kloc < TC_KB ? sg.width - kloc : TC_KB;
    d.vec = ((sg.ld & 3) == 0) && ((d.colk & 3) == 0) && ((d.kvalid & 3) == 0) &&
            ((reinterpret_cast<uintptr_t>(sg.src) & 15) == 0);
    // gathered k-block with a split shadow of its source: staged by TMA gather4 (split precisions, full k-blocks)
    d.tma = (sg.mode == GNNFD_SEG_GATHER && sg.split != nullptr && m.na == 2 && d.kvalid == TC_KB && a->peer_shift == 0 &&
             (sg.ld % TC_KB) == 0 && sg.src_rows > 0 && (reinterpret_cast<uintptr_t>(sg.split) & 15) == 0) ? 1 : 0;
    d.lo_col = sg.ld;
    p.n_tma += d.tma;
    if (sg.split != nullptr && (const void *)sg.src == sg.split && !d.tma) {
      set_error("mlp_forward_tc: segment %d exists only as a split shadow but cannot be staged by TMA here", seg);
      return GNNFD_E_UNSUPPORTED;
    }
  }
  for (int s = 0; s < a->n_seg && p.n_tma > 0; ++s) {
    const gnnfd_segment &sg = a->seg[s];
    bool used = false;
    for (int kb = 0; kb < p.kb1; ++kb) used |= p.kb[kb].tma && p.kb[kb].seg == s;
    if (!used) continue;
    const int rc = make_tmap_2d(&p.tm_seg[s], m.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                                sg.split, (uint64_t)2 * sg.ld, (uint64_t)sg.src_rows, (uint64_t)4 * sg.ld, TC_KB, 1,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != GNNFD_OK) return rc;
  }
  p.a_stages = p.n_tma > 0 ? 3 : 2;
  {
    static int depth = -1;      // experiment knob: GNNFD_L1_DEPTH (0 = unlimited)
    if (depth < 0) {
      const char *e = getenv("GNNFD_L1_DEPTH");
      depth = e ? atoi(e) : 0;
    }
    p.l1_depth = depth;
  }
  // fast final epilogue (TMA tensor stores) whenever the call is plain inference on 128 outputs
  const bool fast = !a->bwd_chain && p.nl == 3 && a->n_out == TC_H && a->mul == nullptr && a->save_a1 == nullptr &&
                    a->save_a2 == nullptr && a->save_xhat == nullptr && a->save_rstd == nullptr && m.na == 2 && m.nw == 2 &&
                    a->dropout_p == 0.f &&
                    (a->out_raw != nullptr || a->out_sum != nullptr || a->out_split != nullptr) &&
                    !(a->out_raw != nullptr && a->out_sum != nullptr && (a->residual != a->out_sum || a->split_of_sum)) &&
                    (a->out_sum == nullptr || a->residual != nullptr) &&
                    ((reinterpret_cast<uintptr_t>(a->out_raw) | reinterpret_cast<uintptr_t>(a->out_sum) |
                      reinterpret_cast<uintptr_t>(a->out_split) | reinterpret_cast<uintptr_t>(a->residual)) & 15) == 0;
  if (a->out_split != nullptr && !fast) {
    set_error("mlp_forward_tc: out_split needs the inference epilogue (n_out = 128, split precision, no stash / mul)");
    return GNNFD_E_UNSUPPORTED;
  }
  if (fast && a->rows > 0) {
    if (a->out_sum != nullptr) {
      const bool in_regs = a->residual != a->out_sum || (a->out_split != nullptr && a->split_of_sum);
      p.epi |= in_regs ? EPI_LDRES : EPI_RED_SUM;
      const int rc = make_tmap_2d(&p.tm_sum, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a->out_sum, TC_H, (uint64_t)a->rows,
                                  TC_H * 4, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != GNNFD_OK) return rc;
    }
    if (a->out_raw != nullptr) {
      p.epi |= EPI_ST_RAW;
      const int rc = make_tmap_2d(&p.tm_raw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a->out_raw, TC_H, (uint64_t)a->rows,
                                  TC_H * 4, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != GNNFD_OK) return rc;
    }
    if (a->out_split != nullptr) p.epi |= EPI_SPLIT;
  }
  int extra_smem = 0;
  if (fast && a->out_split != nullptr && p.a_stages == 2 && (a->rows + TC_BM - 1) / TC_BM > 2 * (int64_t)num_sms()) {
    // room for the second staging block: two-slot weight rings (186 KB + 32 KB); launches of a tile or two per CTA are
    // a latency chain and keep the third weight slot instead
    const int rc = make_tmap_2d(&p.tm_split, m.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                                a->out_split, 2 * TC_H, (uint64_t)a->rows, 4 * TC_H, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != GNNFD_OK) return rc;
    p.split_tma = 1;
    extra_smem = TC_STG_BYTES;
  }
  if (!fast && p.nl == 3 && a->rows > 0 && (a->save_a1 != nullptr || a->save_a2 != nullptr) &&
      ((reinterpret_cast<uintptr_t>(a->save_a1) | reinterpret_cast<uintptr_t>(a->save_a2)) & 15) == 0) {
    float *const sv[2] = {a->save_a1, a->save_a2};
    for (int l = 0; l < 2; ++l) {
      if (sv[l] == nullptr) continue;
      const int rc = make_tmap_2d(&p.tm_save[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sv[l], TC_H, (uint64_t)a->rows, TC_H * 4,
                                  16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != GNNFD_OK) return rc;
    }
    p.save_tma = 1;
  }
  {
    const gnnfd_segment &s0 = a->seg[0];
    const bool contig = s0.mode == GNNFD_SEG_DIRECT && s0.col == 0 && s0.width == s0.ld && (s0.ld & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(s0.src) & 15) == 0;
    p.direct_tile_bytes = contig ? (int64_t)TC_BM * s0.ld * 4 : 0;
  }
  if (a->rows == 0) return GNNFD_OK;
  const int64_t n_tiles = (a->rows + TC_BM - 1) / TC_BM;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  if (a->bwd_chain) extra_smem = TC_STG_BYTES;      // second staging block: asynchronous copies of the saved pre-activations
  p.w_slots = (n_tiles <= 2 * (int64_t)num_sms() && extra_smem == 0) ? TC_W_SLOTS_MAX : 2;
  // LEAN instantiations (BF16X3 only; see the kernel): the common launches of the hot path - full aligned k-blocks of
  // DIRECT / GATHER / MEAN3 (or TMA-staged) segments, SiLU, 128 outputs, no `mul`, no peer matrices, stash through TMA
  bool lean = a->n_out == TC_H && a->peer_shift == 0 && !m.fp16 && m.na == 2 && m.nw == 2 &&
              (p.nl == 1 ? (a->save_a1 == nullptr && a->save_a2 == nullptr && !a->has_ln && !a->bwd_chain)
                         : (a->act == GNNFD_ACT_SILU && a->mul == nullptr &&
                            ((a->save_a1 == nullptr && a->save_a2 == nullptr) || p.save_tma)));
  for (int kb = 0; kb < p.kb1 && lean; ++kb) {
    const KbDesc &d = p.kb[kb];
    lean = d.tma || (d.vec && d.kvalid == TC_KB &&
                     (d.mode == GNNFD_SEG_DIRECT || d.mode == GNNFD_SEG_GATHER || d.mode == GNNFD_SEG_MEAN3));
  }
  {
    static int lean_on = -1;      // A/B knob: GNNFD_LEAN=0 runs the all-purpose instantiations everywhere
    if (lean_on < 0) {
      const char *e = getenv("GNNFD_LEAN");
      lean_on = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    lean = lean && lean_on;
  }
#define LAUNCH1(FP, NA_, NW_, BW, EP, LN)                                                                 \
  do {                                                                                                    \
    static bool attr[GNNFD_MAX_DEVICES] = {false};                                                        \
    if (!attr[current_device()]) {                                                                        \
      GNNFD_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<FP, NA_, NW_, BW, EP, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      227 * 1024));                                                       \
      attr[current_device()] = true;                                                                      \
    }                                                                                                     \
    launch_pdl(mlp_tc_kernel<FP, NA_, NW_, BW, EP, LN>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(p.w_slots, p.a_stages) + extra_smem, stream, p);   \
  } while (0)
#define LAUNCH(FP, NA_, NW_)                                                                              \
  do {                                                                                                    \
    if (a->bwd_chain) LAUNCH1(FP, NA_, NW_, true, 0, 0);                                                  \
    else if (p.drop_thresh != 0u) LAUNCH1(FP, NA_, NW_, false, 2, 0);                                     \
    else LAUNCH1(FP, NA_, NW_, false, 0, 0);                                                              \
  } while (0)
  if (p.a_stages == 3) p.w_slots = 2;      // 3 A stages + 3-slot weight rings do not fit in 227 KB
  if (lean && p.nl == 1) LAUNCH1(false, 2, 2, false, 0, 2);
  else if (lean && fast) LAUNCH1(false, 2, 2, false, 1, 1);
  else if (lean && a->bwd_chain) LAUNCH1(false, 2, 2, true, 0, 1);
  else if (lean && p.drop_thresh == 0u) LAUNCH1(false, 2, 2, false, 0, 1);
  else if (fast && !m.fp16) LAUNCH1(false, 2, 2, false, 1, 0);
  else if (fast) LAUNCH1(true, 2, 2, false, 1, 0);
  else if (!m.fp16 && m.na == 2 && m.nw == 2) LAUNCH(false, 2, 2);
  else if (!m.fp16 && m.na == 1) LAUNCH(false, 1, 1);
  else if (m.fp16 && m.nw == 1) LAUNCH(true, 2, 1);
  else LAUNCH(true, 2, 2);
#undef LAUNCH
#undef LAUNCH1
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

}  // namespace gnnfd
