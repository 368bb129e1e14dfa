// tcgen05 / TMEM / mbarrier / TMA-bulk PTX wrappers and operand-layout helpers shared by the tensor-core
// kernels (mlp_tc.cu, wgrad_tc.cu).  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gnnfd {

constexpr int TC_BM = 128;            // rows per tile == UMMA M
constexpr int TC_H = 128;             // hidden width == UMMA N
constexpr int TC_KB = 64;             // elements per k-block (128 B of 16-bit operands)
constexpr int TC_IMG = TC_BM * 128;   // bytes of one [128 x 64] operand image = 16 KB

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680)   // suspend-time hint: sleep in hardware until the phase completes, not in a spin loop
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t smem_addr, const void *gmem, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T   (A: lane = row, 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------ TMA tensor copies (cp.async.bulk.tensor)
// gather4: four rows (row coordinates r0..r3) x one box of columns starting at `col` of a 2-D tensor land as four
// consecutive box-rows at `dst`; completes box-bytes x 4 on `bar`
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void *tmap, int col, int r0, int r1, int r2, int r3,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_gather4_hint(uint32_t dst, const void *tmap, int col, int r0, int r1, int r2, int r3,
                                                 uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;" ::"r"(dst),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// shared -> global tile store / reduce-add (fp32 add performed by the memory system), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const void *tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void *tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// the same with an L2 eviction-priority operand (common.cuh: L2_EVICT_*)
__device__ __forceinline__ void tma_store_2d_hint(const void *tmap, uint32_t src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(tmap),
               "r"(src), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_hint(const void *tmap, uint32_t src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                   tmap),
               "r"(src), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const void *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64)
// Advancing K by 16 elements (32 B) inside the 128 B swizzle row adds 2 to the start field.
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

// ------------------------------------------------------------------------------ packed fp32 pairs
// sm_100 has two-wide fp32 arithmetic on register pairs (SASS FADD2 / FMUL2 / FFMA2).  Measured on B200
// (profiles/r02_ffma2_issue_rate.log): the same FP32 lanes per clock as the scalar forms but HALF the issue slots - and
// the epilogue warps of the MLP kernel are bound by their dependent instruction stream, not by the FP32 pipe.  Each
// operation is separately rounded (rn), so results are bit-identical to the scalar expressions.
__device__ __forceinline__ void add2(float &a, float &b, float c, float d) {      // (a, b) += (c, d)
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b) : "f"(c), "f"(d));
}
__device__ __forceinline__ void sub2(float &a, float &b, float c, float d) {      // (a, b) -= (c, d)
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tsub.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b) : "f"(c), "f"(d));
}
__device__ __forceinline__ void mul2(float &a, float &b, float c, float d) {      // (a, b) *= (c, d)
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tmul.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b) : "f"(c), "f"(d));
}
__device__ __forceinline__ void fma2(float &a, float &b, float c, float d, float e, float f) {   // (a, b) = (a, b) * (c, d) + (e, f)
  asm("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tmov.b64 z, {%4, %5};\n\t"
      "fma.rn.f32x2 x, x, y, z;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a), "+f"(b) : "f"(c), "f"(d), "f"(e), "f"(f));
}

// -------------------------------------------------------------------------- operand conversion
template <bool FP16>
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  if constexpr (FP16) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    sub2(a, b, hf.x, hf.y);
    __half2 l = __floats2half2_rn(a, b);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    float2 hf = __bfloat1622float2(h);
    sub2(a, b, hf.x, hf.y);
    __nv_bfloat162 l = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = *reinterpret_cast<uint32_t *>(&l);
  }
}

__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int32_t lds_s32(uint32_t saddr) {
  int32_t v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ int4 lds_s32x4(uint32_t saddr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u2(uint32_t saddr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t saddr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// derivatives of the activations at a saved pre-activation x (backward: dA = dH * act'(A))
__device__ __forceinline__ float dsilu(float x) {
  const float sg = rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f));   // same fast sigmoid as the forward's SiLU
  return sg * fmaf(x, 1.0f - sg, 1.0f);
}
// two derivatives at once (same operations, pairwise)
__device__ __forceinline__ void dsilu2(float x0, float x1, float &d0, float &d1) {
  float e0 = x0, e1 = x1;
  mul2(e0, e1, -1.4426950408889634f, -1.4426950408889634f);
  e0 = ex2_ftz(e0); e1 = ex2_ftz(e1);
  add2(e0, e1, 1.0f, 1.0f);
  const float s0 = rcp_ftz(e0), s1 = rcp_ftz(e1);       // sigmoid
  float t0 = 1.0f, t1 = 1.0f;
  sub2(t0, t1, s0, s1);                                   // 1 - sg
  fma2(t0, t1, x0, x1, 1.0f, 1.0f);                       // x (1 - sg) + 1
  mul2(t0, t1, s0, s1);
  d0 = t0; d1 = t1;
}
__device__ __forceinline__ float dtanh(float x) {
  const float t = tanhf(x);
  return 1.0f - t * t;
}
// 16 activations at a time, written stage by stage so the MUFU latencies of independent elements overlap
template <int ACT>
__device__ __forceinline__ void act16(float (&v)[16]) {
  if constexpr (ACT == GNNFD_ACT_SILU) {
    float e[16];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      e[i] = v[i]; e[i + 1] = v[i + 1];
      mul2(e[i], e[i + 1], -1.4426950408889634f, -1.4426950408889634f);
      e[i] = ex2_ftz(e[i]); e[i + 1] = ex2_ftz(e[i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      add2(e[i], e[i + 1], 1.0f, 1.0f);
      e[i] = rcp_ftz(e[i]); e[i + 1] = rcp_ftz(e[i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) mul2(v[i], v[i + 1], e[i], e[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
  }
}

// byte offset of 16-byte chunk `c` of row `r` inside a [rows x 64] SWIZZLE_128B image
__device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

}  // namespace gnnfd
