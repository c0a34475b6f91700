// placeholder until residual_tensor.cu lands: PDEIP_PATH_TENSOR fails loudly.
#include "common.cuh"
#include "residual_common.cuh"
namespace pdeip {
#ifndef PDEIP_HAVE_TENSOR_PATH
int mlp_residual_accumulate_tensor(int, const ResidualArgs&, int, cudaStream_t) {
  set_error("PDEIP_PATH_TENSOR is not built in this library");
  return PDEIP_ERR_UNSUPPORTED;
}
#endif
}  // namespace pdeip
