// gmm.cu — K2: stand-alone GMM potential value / gradient and the linear drift.
//
// Replaces core/potential.py:32-61 (gmm_V, g_gmm_V = jax.grad(gmm_V), vg_gmm_V, GMMPotential) and
// QuadraticPotential.gradient (core/potential.py:20-24, mu = 0).  One thread per point; the centres
// sit in shared memory and are broadcast to the warp; the softmax is evaluated online in fp32 with
// direct differences (x - mu_k), so there is no |x|^2 - 2 x.mu + |mu|^2 cancellation.
#include "common.cuh"
#include "drift.cuh"

namespace pdeip {

template <int DP>
__global__ void __launch_bounds__(128) gmm_value_grad_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ mus, int K,
                                                             float inv_sigma2, float* __restrict__ out_value,
                                                             float* __restrict__ out_grad, int64_t n, int d) {
  extern __shared__ __align__(16) float smem[];
  load_padded<DP>(smem, mus, K, d, threadIdx.x, blockDim.x);
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float xr[DP], g[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) xr[i] = (i < d) ? x[p * d + i] : 0.0f;
  float val;
  gmm_grad_thread<DP>(xr, smem, K, inv_sigma2, g, &val);
  if (out_value) out_value[p] = val;
  if (out_grad) {
#pragma unroll
    for (int i = 0; i < DP; ++i)
      if (i < d) out_grad[p * d + i] = g[i];
  }
}

template <int DP>
__global__ void __launch_bounds__(128) linear_grad_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ A,
                                                          float* __restrict__ out, int64_t n, int d) {
  __shared__ __align__(16) float A_s[DP * DP];
  load_padded<DP>(A_s, A, d, d, threadIdx.x, blockDim.x);
  for (int idx = d * DP + threadIdx.x; idx < DP * DP; idx += blockDim.x) A_s[idx] = 0.0f;
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float xr[DP], g[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) xr[i] = (i < d) ? x[p * d + i] : 0.0f;
  linear_grad_thread<DP>(xr, A_s, nullptr, g);
#pragma unroll
  for (int i = 0; i < DP; ++i)
    if (i < d) out[p * d + i] = g[i];
}

}  // namespace pdeip

using namespace pdeip;

extern "C" int pdeip_gmm_value_grad(const float* x, const float* mus, int n_gaussian, float sigma,
                                    float* out_value, float* out_grad, int64_t n, int d, void* stream) {
  PDEIP_REQUIRE(x && mus, PDEIP_ERR_INVALID_ARG, "x / mus is NULL");
  PDEIP_REQUIRE(d >= 1 && d <= 32, PDEIP_ERR_UNSUPPORTED, "gmm_value_grad supports 1 <= d <= 32 (got %d)", d);
  PDEIP_REQUIRE(n_gaussian >= 1 && sigma > 0.0f && n >= 0, PDEIP_ERR_INVALID_ARG, "bad n_gaussian / sigma / n");
  if (n == 0) return PDEIP_OK;
  const float inv_sigma2 = 1.0f / (sigma * sigma);
  const unsigned grid = (unsigned)((n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_GMM(DP)                                                                               \
  do {                                                                                               \
    const size_t smem = sizeof(float) * (size_t)n_gaussian * DP;                                     \
    PDEIP_REQUIRE(smem <= 48 * 1024, PDEIP_ERR_UNSUPPORTED, "GMM centres exceed 48 KB of shared memory"); \
    gmm_value_grad_kernel<DP><<<grid, 128, smem, st>>>(x, mus, n_gaussian, inv_sigma2, out_value,   \
                                                        out_grad, n, d);                            \
  } while (0)
  if (d <= 2) LAUNCH_GMM(2);
  else if (d <= 4) LAUNCH_GMM(4);
  else if (d <= 8) LAUNCH_GMM(8);
  else if (d <= 16) LAUNCH_GMM(16);
  else LAUNCH_GMM(32);
#undef LAUNCH_GMM
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

extern "C" int pdeip_linear_grad(const float* x, const float* A, float* out, int64_t n, int d, void* stream) {
  PDEIP_REQUIRE(x && A && out, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(d >= 1 && d <= 32, PDEIP_ERR_UNSUPPORTED, "linear_grad supports 1 <= d <= 32 (got %d)", d);
  if (n <= 0) return n == 0 ? PDEIP_OK : PDEIP_ERR_INVALID_ARG;
  const unsigned grid = (unsigned)((n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (d <= 2) linear_grad_kernel<2><<<grid, 128, 0, st>>>(x, A, out, n, d);
  else if (d <= 4) linear_grad_kernel<4><<<grid, 128, 0, st>>>(x, A, out, n, d);
  else if (d <= 8) linear_grad_kernel<8><<<grid, 128, 0, st>>>(x, A, out, n, d);
  else if (d <= 16) linear_grad_kernel<16><<<grid, 128, 0, st>>>(x, A, out, n, d);
  else linear_grad_kernel<32><<<grid, 128, 0, st>>>(x, A, out, n, d);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}
