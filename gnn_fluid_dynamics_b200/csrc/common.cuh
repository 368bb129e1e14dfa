// Shared helpers for the gnnfd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gnnfd_b200.h"

namespace gnnfd {

void set_error(const char *fmt, ...);

#define GNNFD_CHECK_ARG(cond, msg)                 \
  do {                                             \
    if (!(cond)) {                                 \
      gnnfd::set_error("%s: %s", __func__, msg);   \
      return GNNFD_E_BADARG;                       \
    }                                              \
  } while (0)

#define GNNFD_CUDA(call)                                                               \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      gnnfd::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_));  \
      return GNNFD_E_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define GNNFD_LAUNCH_CHECK()                                                           \
  do {                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                               \
    if (e_ != cudaSuccess) {                                                           \
      gnnfd::set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e_));     \
      return GNNFD_E_CUDA;                                                             \
    }                                                                                  \
  } while (0)

// Per-device caches (SM count, "dynamic shared memory opt-in done" flags) are keyed by the CURRENT device ordinal:
// one process may drive several GPUs, and cudaFuncSetAttribute / the SM count are per device.
constexpr int GNNFD_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < GNNFD_MAX_DEVICES) ? dev : 0;
}
inline int num_sms() {
  static int n[GNNFD_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// A step of the hot path is a chain of several hundred DEPENDENT launches on one stream.  Every kernel of the chain
// is launched with programmaticStreamSerializationAllowed, raises `griddepcontrol.launch_dependents` as its first
// instruction and executes `griddepcontrol.wait` before its first access to global memory: the next kernel's CTAs
// become resident on an SM as soon as the previous kernel's CTA there has retired, run their on-chip prologue (barrier
// initialisation, TMEM allocation, tensor-map prefetch) under the previous kernel's tail, and then block in hardware
// until the previous grid has completed and its writes are visible.  Nothing before the wait touches global memory, so
// the result is the stream-ordered one.  A kernel launched WITHOUT the attribute sees both instructions as no-ops.
// GNNFD_PDL=0 in the environment switches the attribute off (plain stream serialisation).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// the two together: kernels without an on-chip prologue
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }

bool pdl_enabled();

// launches `kernel` (which MUST execute pdl_wait() before touching global memory) with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// One deferred split-K reduction of a weight-gradient GEMM (wgrad_tc.cu): gnnfd_mlp_backward runs its GEMMs back to
// back and adds ALL their partials in one launch at the end instead of one small launch after each GEMM.
struct WgReduceJob {
  const float *part;      // [n_parts][128][n_pad]
  const float *cs_part;   // [n_parts][128] column-sum partials (or nullptr)
  float *out, *cs_out;
  int n_parts, n_pad, m_valid, n_valid, ld_out, transpose, cs_valid;
  size_t ws_used;         // bytes of the workspace this GEMM's partials occupy
};
constexpr int WG_MAX_REDUCE_JOBS = 6;
int wgrad_run(const gnnfd_wgrad_args *a, void *workspace, size_t workspace_bytes, cudaStream_t stream, WgReduceJob *defer);
int wgrad_reduce_jobs(const WgReduceJob *jobs, int n_jobs, cudaStream_t stream);
// lean direct x direct GEMMs (contiguous [rows, 128] operands, split-bf16): is `a` one, and up to three of them over the
// same rows in ONE launch with their reductions deferred into jobs_out[0 .. n)
bool wgrad_is_lean(const gnnfd_wgrad_args *a);
int wgrad_lean_run(const gnnfd_wgrad_args *const *args, int n, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   WgReduceJob *jobs_out);

// exact-path activations (IEEE expf / tanhf); the tensor-core path uses the fast variants
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float act_f(float v, int act) {
  return act == GNNFD_ACT_SILU ? silu_f(v) : tanhf(v);
}
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// ---- dropout mask (gnnfd_mlp_args.dropout_p): a counter-based hash of (seed, hidden layer, row, column) --------------
// lowbias32 integer finaliser; the unit is dropped when dropout_hash(...) < p * 2^32
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_layer_key(uint64_t seed, int layer) {
  return hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(layer + 1)));
}
__host__ __device__ __forceinline__ uint32_t dropout_row_hash(uint32_t key, uint32_t row) { return hash32(row * 0x9E3779B1u ^ key); }
__host__ __device__ __forceinline__ uint32_t dropout_hash(uint32_t row_hash, uint32_t col) { return hash32(row_hash + col); }

// GNNFD_SEG_SUM3S index entry -> (row, sign)
__device__ __forceinline__ void sum3s_decode(int32_t v, int32_t &row, float &sign) {
  const bool zero = v == INT32_MIN;
  row = zero ? 0 : (v >= 0 ? v : ~v);
  sign = zero ? 0.f : (v >= 0 ? 1.f : -1.f);
}

__device__ __forceinline__ float4 ldg_f4(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}

// ---- L2 eviction-priority hints ----------------------------------------------------------------------------------
// A step streams several hundred MB through a 126 MB L2 per launch; most of it is written once and read much later (the
// training stash, the residual streams) while the gathered cell / vertex latents (<= 82 MB) are re-read ~6 times by
// faces that are far apart in face order.  Write-once streams are stored with an evict-first policy and the gathered
// rows loaded with evict-last, so the streams stop displacing the rows that are re-read.  The policy is an OPERAND (a
// 64-bit descriptor, the encodings createpolicy.fractional.L2::evict_* produces at fraction 1.0): the same instruction
// with L2_EVICT_NORMAL is the unhinted access, so the hints cost no code and can be A/B-ed at run time
// (GNNFD_L2_HINTS, l2_hint_mask()).
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull,
                   L2_EVICT_LAST = 0x14F0000000000000ull;
enum {
  L2H_ST_STASH = 1,    // training stash (pre-activations, x-hat, dA): evict-first stores
  L2H_ST_OUT = 2,      // residual-stream outputs (out_sum, split shadow): evict-first stores
  L2H_LD_STREAM = 4,   // contiguous (DIRECT) operand rows, saved pre-activations, residual rows: evict-first loads
  L2H_LD_KEEP = 8,     // gathered rows: evict-last loads
  L2H_ST_RAW = 16,     // raw outputs (consumed by the next launch): evict-first stores
};
constexpr int L2_HINT_DEFAULT = 0;
int l2_hint_mask();    // GNNFD_L2_HINTS / gnnfd_set_l2_hints (csr.cu)
struct L2Policies { uint64_t st_stash, st_out, st_raw, ld_stream, ld_keep; };
inline L2Policies l2_policies() {
  const int m = l2_hint_mask();
  L2Policies p;
  p.st_stash = (m & L2H_ST_STASH) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  p.st_out = (m & L2H_ST_OUT) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  p.st_raw = (m & L2H_ST_RAW) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  p.ld_stream = (m & L2H_LD_STREAM) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  p.ld_keep = (m & L2H_LD_KEEP) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
  return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float *p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_f4_hint(float *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol) : "memory");
}

}  // namespace gnnfd
