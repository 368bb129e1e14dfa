// Halo exchange support for the domain-decomposed processor: pack the rows of boundary cells that a peer
// rank needs into one contiguous send buffer (the receive side lands straight in the ghost rows, which are
// laid out contiguously per owner rank, so there is no unpack pass).  The transfer itself is NCCL P2P
// (torch.distributed batch_isend_irecv: grouped ncclSend / ncclRecv) over NVLink - gnn_fluid_dynamics_b200/dist.py.
#include "common.cuh"

namespace gnnfd {
// one warp per packed row; width = 4 * lanes used
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, int ld,
                                                          const int32_t *__restrict__ idx, int64_t n, int width,
                                                          float *__restrict__ out) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const int64_t s = __ldg(idx + r);
  if ((width & 3) == 0 && (ld & 3) == 0) {
    for (int c = lane * 4; c < width; c += 128)
      *reinterpret_cast<float4 *>(out + r * width + c) = ldg_f4(src + s * ld + c);
  } else {
    for (int c = lane; c < width; c += 32) out[r * width + c] = __ldg(src + s * ld + c);
  }
}
}  // namespace gnnfd

using namespace gnnfd;

extern "C" int gnnfd_gather_rows(const float *src, int32_t ld, const int32_t *idx, int64_t n, int32_t width,
                                 float *out, void *stream) {
  GNNFD_CHECK_ARG(n >= 0 && width > 0 && ld >= width, "bad sizes");
  if (n == 0) return GNNFD_OK;
  GNNFD_CHECK_ARG(src && idx && out, "null pointer");
  GNNFD_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "buffers must be 16-byte aligned");
  const int blocks = (int)((n * 32 + 255) / 256);
  launch_pdl(gather_rows_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, src, ld, idx, n, width, out);
  GNNFD_LAUNCH_CHECK();
  return GNNFD_OK;
}

extern "C" int gnnfd_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  GNNFD_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return GNNFD_OK;
  int can = 0;
  GNNFD_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) { set_error("gnnfd_enable_peer_access: device %d cannot access device %d", dev, peer_device); return GNNFD_E_UNSUPPORTED; }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return GNNFD_OK; }
  if (e != cudaSuccess) { set_error("gnnfd_enable_peer_access: %s", cudaGetErrorString(e)); return GNNFD_E_CUDA; }
  return GNNFD_OK;
}
