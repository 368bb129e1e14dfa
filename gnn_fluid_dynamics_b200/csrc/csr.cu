// Receiver-sorted CSR build (kernel (a) of BASELINE.json: north_star) + index narrowing.
//
// perm = stable argsort(index), offsets[r] = #{index < r}: integer-identical to
// torch.sort(index, stable=True) + bincount.  Counting sort:
//   histogram -> exclusive scan -> unordered bucket fill -> per-row sort of the (distinct) source
//   positions.  The per-row sort makes the result independent of the atomics' arrival order, so the
//   output is deterministic and bit-exact.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace gnnfd {

static thread_local char g_err[512] = "ok";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// launch overlap (programmatic dependent launch, common.cuh): -1 = not chosen yet (the environment decides, default
// off); GNNFD_PDL = 0 / 1 in the environment pins it for the whole process
static std::atomic<int> g_pdl{-1};
static int pdl_env() {
  static const int v = [] {
    const char *e = getenv("GNNFD_PDL");
    return e == nullptr ? -1 : (e[0] == '0' ? 0 : 1);
  }();
  return v;
}
bool pdl_enabled() {
  const int env = pdl_env();
  if (env >= 0) return env != 0;
  return g_pdl.load(std::memory_order_relaxed) > 0;
}

// L2 eviction-priority hints (common.cuh): GNNFD_L2_HINTS=<mask> in the environment pins the mask for the process,
// else gnnfd_set_l2_hints chooses it; default L2_HINT_DEFAULT
static std::atomic<int> g_l2_hints{-1};
int l2_hint_mask() {
  static const int env = [] {
    const char *e = getenv("GNNFD_L2_HINTS");
    return e == nullptr ? -1 : atoi(e);
  }();
  if (env >= 0) return env;
  const int v = g_l2_hints.load(std::memory_order_relaxed);
  return v >= 0 ? v : L2_HINT_DEFAULT;
}

__global__ void narrow_kernel(const int64_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n,
                              int64_t limit, int32_t *err) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t v = src[i];
    if (v < 0 || v >= limit) {
      *err = 1;
      v = 0;
    }
    dst[i] = (int32_t)v;
  }
}

__global__ void hist_kernel(const int32_t *__restrict__ index, int64_t n, int32_t *counts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) atomicAdd(&counts[index[i]], 1);
}

// ---- exclusive scan: tile scan (1024 threads x 4) -> scan of tile sums -> add -----------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  int res = inc - v + warp_sums[warp];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const int32_t *__restrict__ in,
                                                                  int32_t *__restrict__ out,
                                                                  int64_t n, int32_t *tile_sums) {
  __shared__ int total;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    s += v[j];
  }
  int off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = off;
    off += v[j];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int32_t *tile_sums, int64_t n_tiles) {
  __shared__ int total;
  int carry = 0;
  for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    int v = i < n_tiles ? tile_sums[i] : 0;
    int off = block_exclusive_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = off + carry;
    carry += total;
    __syncthreads();
  }
}

__global__ void add_tile_offsets_kernel(int32_t *out, int64_t n, const int32_t *__restrict__ tile_sums) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] += tile_sums[i / SCAN_TILE];
}

__global__ void fill_kernel(const int32_t *__restrict__ index, int64_t n, int32_t *cursor,
                            int32_t *__restrict__ tmp) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int pos = atomicAdd(&cursor[index[i]], 1);
    tmp[pos] = (int32_t)i;
  }
}

// one warp per row: sort the row's source positions ascending (all distinct)
__global__ void sort_rows_kernel(const int32_t *__restrict__ offsets, const int32_t *__restrict__ tmp,
                                 int32_t *__restrict__ perm, int64_t n_rows) {
  int lane = threadIdx.x & 31;
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  int beg = offsets[row], end = offsets[row + 1];
  int d = end - beg;
  if (d <= 0) return;
  if (d <= 32) {
    int v = lane < d ? tmp[beg + lane] : 0x7fffffff;
    // bitonic sort across the warp
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        int other = __shfl_xor_sync(0xffffffffu, v, j);
        bool up = ((lane & k) == 0);
        bool lower = ((lane & j) == 0);
        int mn = min(v, other), mx = max(v, other);
        v = (lower == up) ? mn : mx;
      }
    }
    if (lane < d) perm[beg + lane] = v;
  } else {
    for (int i = lane; i < d; i += 32) {
      int v = tmp[beg + i];
      int rank = 0;
      for (int j = 0; j < d; ++j) rank += (tmp[beg + j] < v);
      perm[beg + rank] = v;
    }
  }
}

static inline int grid_for(int64_t n, int threads, int max_blocks) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace gnnfd

using namespace gnnfd;

extern "C" int gnnfd_abi_version(void) { return GNNFD_ABI_VERSION; }
extern "C" const char *gnnfd_last_error(void) { return g_err; }
extern "C" int gnnfd_set_launch_overlap(int32_t on) {
  const int prev = g_pdl.exchange(on != 0 ? 1 : 0, std::memory_order_relaxed);
  return prev > 0 ? 1 : 0;
}

extern "C" int gnnfd_set_l2_hints(int32_t mask) {
  const int prev = l2_hint_mask();
  g_l2_hints.store(mask < 0 ? -1 : mask, std::memory_order_relaxed);
  return prev;
}

extern "C" int gnnfd_index_narrow(const int64_t *src, int32_t *dst, int64_t n, int64_t limit,
                                  int32_t *err_flag, void *stream) {
  GNNFD_CHECK_ARG(n >= 0 && limit >= 0 && limit <= 0x7fffffffLL, "bad sizes");
  if (n == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(src && dst && err_flag, "null pointer");
  narrow_kernel<<<grid_for(n, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(src, dst, n, limit,
                                                                                 err_flag);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t gnnfd_csr_workspace_bytes(int64_t n, int64_t n_rows) {
  if (n < 0 || n_rows < 0) return 0;
  size_t n_tiles = (size_t)((n_rows + 1 + SCAN_TILE - 1) / SCAN_TILE);
  return align256((size_t)(n_rows + 1) * 4) + align256((size_t)n * 4) + align256((n_tiles + 1) * 4) + 256;
}

extern "C" int gnnfd_csr_build(const int32_t *index, int64_t n, int64_t n_rows, int32_t *offsets,
                               int32_t *perm, void *workspace, size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNNFD_CHECK_ARG(n >= 0 && n_rows >= 0 && n <= 0x7fffffffLL && n_rows < 0x7fffffffLL, "bad sizes");
  GNNFD_CHECK_ARG(offsets, "null offsets");
  if (workspace_bytes < gnnfd_csr_workspace_bytes(n, n_rows)) {
    set_error("gnnfd_csr_build: workspace too small");
    return GNNFD_E_WORKSPACE;
  }
  int64_t m = n_rows + 1;
  if (n == 0) {
    GNNFD_CUDA(cudaMemsetAsync(offsets, 0, (size_t)m * 4, stream));
    return GNNFD_OK;
  }
  GNNFD_CHECK_ARG(index && perm && workspace, "null pointer");
  char *ws = (char *)workspace;
  int32_t *counts = (int32_t *)ws;
  ws += align256((size_t)m * 4);
  int32_t *tmp = (int32_t *)ws;
  ws += align256((size_t)n * 4);
  int32_t *tile_sums = (int32_t *)ws;
  int64_t n_tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
  const int maxb = num_sms() * 16;

  GNNFD_CUDA(cudaMemsetAsync(counts, 0, (size_t)m * 4, stream));
  hist_kernel<<<grid_for(n, 256, maxb), 256, 0, stream>>>(index, n, counts);
  GNNFD_LAUNCH_CHECK();
  scan_tiles_kernel<<<(int)n_tiles, SCAN_THREADS, 0, stream>>>(counts, offsets, m, tile_sums);
  GNNFD_LAUNCH_CHECK();
  scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(tile_sums, n_tiles);
  GNNFD_LAUNCH_CHECK();
  add_tile_offsets_kernel<<<(int)((m + 255) / 256), 256, 0, stream>>>(offsets, m, tile_sums);
  GNNFD_LAUNCH_CHECK();
  // cursor = copy of offsets (first n_rows entries)
  GNNFD_CUDA(cudaMemcpyAsync(counts, offsets, (size_t)m * 4, cudaMemcpyDeviceToDevice, stream));
  fill_kernel<<<grid_for(n, 256, maxb), 256, 0, stream>>>(index, n, counts, tmp);
  GNNFD_LAUNCH_CHECK();
  int64_t threads_total = n_rows * 32;
  if (n_rows > 0) {
    sort_rows_kernel<<<(int)((threads_total + 255) / 256), 256, 0, stream>>>(offsets, tmp, perm, n_rows);
    GNNFD_LAUNCH_CHECK();
  }
  return GNNFD_OK;
}
