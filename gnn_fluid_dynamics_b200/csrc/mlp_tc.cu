// placeholder until the tcgen05 variant lands
#include "common.cuh"
namespace gnnfd {
int mlp_forward_tc(const gnnfd_mlp_args *, cudaStream_t) {
  set_error("tensor-core precision not built yet");
  return GNNFD_E_UNSUPPORTED;
}
size_t pack_mlp_bytes_tc(int, int, int, int) { return 0; }
int pack_mlp_tc(const gnnfd_mlp_args *, void *, cudaStream_t) { return GNNFD_E_UNSUPPORTED; }
}  // namespace gnnfd
