// optim.cu — K6 (Adam + L2 [+ EMA] + norms), K7 (ensemble moments) and the offline 0T gather.
//
// K6 replaces main.py:20-26 (optax.chain(add_decayed_weights, adam(b1=0.9, eps=1e-4))),
//    core/trainer.py:61-70 (step, optional EMA) and :110 (params_norm), utils/common_utils.py:74-76.
// K7 computes sum z and sum z z^T for the validation against the Lyapunov moments
//    (example_problems/kinetic_fokker_planck_example_OU.py:73-93).
// gather replaces the strided time / random trajectory sub-sampling of methods/consistency.py:102-118.
#include <math.h>

#include "common.cuh"

namespace pdeip {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  t = warp_sum(t);
  return __shfl_sync(0xffffffffu, t, 0);
}

__global__ void __launch_bounds__(1024) adam_l2_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       float* __restrict__ ema, int64_t n, float lr, float b1,
                                                       float b2, float eps, float wd, float inv_bc1,
                                                       float inv_bc2, float grad_scale, int use_ema,
                                                       float ema_decay, float* __restrict__ norms) {
  __shared__ float red[32];
  float gn = 0.f, pn = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float pi = p[i];
    const float gi = grad_scale * g[i];
    gn = fmaf(gi, gi, gn);
    const float gd = gi + wd * pi;                       // add_decayed_weights (before Adam)
    const float mi = b1 * m[i] + (1.f - b1) * gd;        // scale_by_adam
    const float vi = b2 * v[i] + (1.f - b2) * gd * gd;
    m[i] = mi;
    v[i] = vi;
    const float mhat = mi * inv_bc1, vhat = vi * inv_bc2;
    float pnew = pi - lr * (mhat / (sqrtf(vhat) + eps));  // eps outside the sqrt (eps_root = 0)
    if (use_ema) {                                        // trainer.py:67-69: params <- raw ema
      const float e = ema_decay * ema[i] + (1.f - ema_decay) * pnew;
      ema[i] = e;
      pnew = e;
    }
    p[i] = pnew;
    pn = fmaf(pnew, pnew, pn);
  }
  gn = block_sum(gn, red);
  pn = block_sum(pn, red);
  if (threadIdx.x == 0 && norms) {
    norms[0] = sqrtf(gn);
    norms[1] = sqrtf(pn);
  }
}

// ---- K7 ------------------------------------------------------------------------------------------
constexpr int kMomWarps = 8;

template <int DIMP>
__global__ void __launch_bounds__(kMomWarps * 32) moments_kernel(const float* __restrict__ z, int64_t n, int dim,
                                                                 int layout, float* __restrict__ ws) {
  // dynamic smem: the point tiles [warps][32][DIMP+1], later reused as the per-warp partials
  // [warps][DIMP*(DIMP+1)]
  extern __shared__ __align__(16) float msm[];
  float (*tile)[32][DIMP + 1] = reinterpret_cast<float (*)[32][DIMP + 1]>(msm);
  float (*part)[DIMP * (DIMP + 1)] = reinterpret_cast<float (*)[DIMP * (DIMP + 1)]>(msm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[(DIMP + 31) / 32][DIMP];
  float accs[(DIMP + 31) / 32];
#pragma unroll
  for (int q = 0; q < (DIMP + 31) / 32; ++q) {
    accs[q] = 0.f;
#pragma unroll
    for (int c = 0; c < DIMP; ++c) acc[q][c] = 0.f;
  }
  const int64_t tile_pts = kMomWarps * 32;
  const int64_t n_tiles = (n + tile_pts - 1) / tile_pts;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t p = t * tile_pts + threadIdx.x;
    __syncwarp();
    for (int c = 0; c < DIMP; ++c)
      tile[warp][lane][c] = (p < n && c < dim) ? z[elem_index(layout, p, c, n, dim)] : 0.f;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < (DIMP + 31) / 32; ++q) {
      const int r = lane + 32 * q;
      if (r < DIMP) {
        for (int pp = 0; pp < 32; ++pp) {
          const float a = tile[warp][pp][r];
          accs[q] += a;
#pragma unroll
          for (int c = 0; c < DIMP; ++c) acc[q][c] = fmaf(a, tile[warp][pp][c], acc[q][c]);
        }
      }
    }
  }
  __syncthreads();  // all warps are done with the tiles before they are reused as partials
#pragma unroll
  for (int q = 0; q < (DIMP + 31) / 32; ++q) {
    const int r = lane + 32 * q;
    if (r < DIMP) {
      part[warp][r] = accs[q];
#pragma unroll
      for (int c = 0; c < DIMP; ++c) part[warp][DIMP + r * DIMP + c] = acc[q][c];
    }
  }
  __syncthreads();
  float* out = ws + (int64_t)blockIdx.x * (dim + dim * dim);
  for (int idx = threadIdx.x; idx < dim + dim * dim; idx += blockDim.x) {
    int src;
    if (idx < dim) src = idx;
    else {
      const int r = (idx - dim) / dim, c = (idx - dim) - r * dim;
      src = DIMP + r * DIMP + c;
    }
    float s = 0.f;
    for (int w = 0; w < kMomWarps; ++w) s += part[w][src];
    out[idx] = s;
  }
}

__global__ void moments_finalize_kernel(const float* __restrict__ ws, int grid_ctas, int len,
                                        float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= len) return;
  double s = 0.0;
  for (int b = 0; b < grid_ctas; ++b) s += (double)ws[(int64_t)b * len + idx];
  out[idx] = (float)s;
}

__global__ void gather_0T_kernel(const float* __restrict__ dataset, int n_time, int dim,
                                 const int64_t* __restrict__ sample_index, int64_t n_sel, int interval, int shift,
                                 int n_time_sel, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n_sel * n_time_sel * dim;
  if (idx >= total) return;
  const int c = (int)(idx % dim);
  const int64_t row = idx / dim;
  const int k = (int)(row % n_time_sel);
  const int64_t s = row / n_time_sel;
  out[idx] = dataset[(sample_index[s] * n_time + (int64_t)k * interval + shift) * dim + c];
}

}  // namespace pdeip

using namespace pdeip;

extern "C" int pdeip_adam_l2_step(float* params, const float* grad, float* m, float* v, float* ema, int64_t n,
                                  float lr, float b1, float b2, float eps, float weight_decay, int64_t count,
                                  float grad_scale, int use_ema, float ema_decay, float* norms, void* stream) {
  PDEIP_REQUIRE(params && grad && m && v, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(n >= 0 && count >= 1, PDEIP_ERR_INVALID_ARG, "n must be >= 0 and count >= 1");
  PDEIP_REQUIRE(!use_ema || ema, PDEIP_ERR_INVALID_ARG, "use_ema set but ema is NULL");
  const float inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, (double)count)));
  const float inv_bc2 = (float)(1.0 / (1.0 - pow((double)b2, (double)count)));
  adam_l2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grad, m, v, ema, n, lr, b1, b2, eps, weight_decay,
                                                      inv_bc1, inv_bc2, grad_scale, use_ema, ema_decay, norms);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

extern "C" size_t pdeip_moments_workspace_bytes(int dim) {
  if (dim < 1) return 0;
  return sizeof(float) * (size_t)sm_count() * (size_t)(dim + dim * dim);
}

extern "C" int pdeip_ensemble_moments(const float* z, int64_t n, int dim, int layout, float* out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  PDEIP_REQUIRE(z && out, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(dim >= 1 && dim <= 64, PDEIP_ERR_UNSUPPORTED, "ensemble_moments supports 1 <= dim <= 64 (got %d)", dim);
  PDEIP_REQUIRE(layout == PDEIP_LAYOUT_AOS || layout == PDEIP_LAYOUT_SOA, PDEIP_ERR_INVALID_ARG, "bad layout");
  PDEIP_REQUIRE(workspace && workspace_bytes >= pdeip_moments_workspace_bytes(dim), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", pdeip_moments_workspace_bytes(dim), workspace_bytes);
  PDEIP_REQUIRE(n >= 0, PDEIP_ERR_INVALID_ARG, "n < 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = sm_count();
  float* ws = (float*)workspace;
#define LAUNCH_MOM(DIMP)                                                                                 \
  do {                                                                                                   \
    const size_t a_ = sizeof(float) * kMomWarps * 32 * (DIMP + 1);                                       \
    const size_t b_ = sizeof(float) * kMomWarps * DIMP * (DIMP + 1);                                     \
    const size_t smem_ = a_ > b_ ? a_ : b_;                                                              \
    auto kern = moments_kernel<DIMP>;                                                                    \
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));  \
    kern<<<grid, kMomWarps * 32, smem_, st>>>(z, n, dim, layout, ws);                                    \
  } while (0)
  if (dim <= 8) LAUNCH_MOM(8);
  else if (dim <= 16) LAUNCH_MOM(16);
  else if (dim <= 32) LAUNCH_MOM(32);
  else LAUNCH_MOM(64);
#undef LAUNCH_MOM
  PDEIP_LAUNCH_OK();
  const int len = dim + dim * dim;
  moments_finalize_kernel<<<(len + 127) / 128, 128, 0, st>>>(ws, grid, len, out);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

extern "C" int pdeip_gather_0T(const float* dataset, int64_t n_traj, int n_time, int dim,
                               const int64_t* sample_index, int64_t n_sel, int interval, int shift, int n_time_sel,
                               float* out, void* stream) {
  PDEIP_REQUIRE(dataset && sample_index && out, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(n_traj >= 0 && n_time >= 1 && dim >= 1 && interval >= 1 && shift >= 0 && n_time_sel >= 0 &&
                    n_sel >= 0,
                PDEIP_ERR_INVALID_ARG, "bad sizes");
  PDEIP_REQUIRE(n_time_sel == 0 || (int64_t)(n_time_sel - 1) * interval + shift < n_time, PDEIP_ERR_INVALID_ARG,
                "time index out of range: (n_time_sel-1)*interval+shift >= n_time");
  const int64_t total = n_sel * n_time_sel * dim;
  if (total == 0) return PDEIP_OK;
  gather_0T_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      dataset, n_time, dim, sample_index, n_sel, interval, shift, n_time_sel, out);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}
