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

// GNNFD_SEG_SUM3S index entry -> (row, sign)
__device__ __forceinline__ void sum3s_decode(int32_t v, int32_t &row, float &sign) {
  const bool zero = v == INT32_MIN;
  row = zero ? 0 : (v >= 0 ? v : ~v);
  sign = zero ? 0.f : (v >= 0 ? 1.f : -1.f);
}

__device__ __forceinline__ float4 ldg_f4(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}

}  // namespace gnnfd
