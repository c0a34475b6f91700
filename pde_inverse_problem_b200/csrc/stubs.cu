// stubs.cu — entry points that are declared in pdeip.h but not built yet fail loudly.
#include "common.cuh"
#include "residual_common.cuh"

namespace pdeip {

int mlp_residual_accumulate_tensor(int, const ResidualArgs&, int, cudaStream_t) {
  set_error("PDEIP_PATH_TENSOR is not built in this library");
  return PDEIP_ERR_UNSUPPORTED;
}
int param_residual_accumulate(int, int, const ResidualArgs&, int, cudaStream_t) {
  set_error("parametric residual is not built in this library");
  return PDEIP_ERR_UNSUPPORTED;
}
int param_eval(int, const float*, int, int, const float*, const float*, float*, float*, float*, float*, int64_t,
               cudaStream_t) {
  set_error("parametric model_eval is not built in this library");
  return PDEIP_ERR_UNSUPPORTED;
}
int kmv_mean_grad(int, const float*, int, int, int, const float*, int64_t, int, float*, float*, const float*,
                  cudaStream_t) {
  set_error("KMV residual is not built in this library");
  return PDEIP_ERR_UNSUPPORTED;
}
int kmv_residual_accumulate(int, const ResidualArgs&, int, cudaStream_t) {
  set_error("KMV residual is not built in this library");
  return PDEIP_ERR_UNSUPPORTED;
}

}  // namespace pdeip
