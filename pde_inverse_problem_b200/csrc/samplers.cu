// samplers.cu — exact Gaussian samplers that feed the residual in "online / exact" mode.
//
//  gaussian_sample_grouped : one (mean, cov_half) per group of samples — the kinetic-OU sampler
//      (example_problems/kinetic_fokker_planck_example_OU.py:140-190: Gaussian(mean_t, cov_t).sample per time).
//  ou_exact_sample : per-sample random time of the overdamped OU law in the eigenbasis of F
//      (example_problems/fokker_planck_example.py:48-55,84-96):
//         m_t = U e^{-St} U^T m_0,  P_t = U (E B_0 E + B_S - E B_S E) U^T,  E = e^{-St}, B_S = B / (s_i + s_j)
//      each thread factors its own d x d covariance (Cholesky) and draws x = U (E mt0 + chol xi).
#include "common.cuh"
#include "philox.cuh"

namespace pdeip {

constexpr uint32_t kTagGrouped = 0xFFFFFFFDu;
constexpr uint32_t kTagOuTime = 0xFFFFFFFCu;
constexpr uint32_t kTagOuX = 0xFFFFFFFBu;

__global__ void gaussian_grouped_kernel(float* __restrict__ out, int64_t n_groups, int per_group, int dim,
                                        const float* __restrict__ mus, const float* __restrict__ halves,
                                        uint64_t seed, uint64_t particle_offset) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_groups * per_group) return;
  const int64_t g = p / per_group;
  float xi[64];
  for (int j = 0; j < (dim + 3) / 4; ++j) {
    float r4[4];
    philox_normal4(seed, particle_offset + (uint64_t)p, kTagGrouped, (uint32_t)j, r4);
    for (int k = 0; k < 4; ++k) xi[4 * j + k] = r4[k];
  }
  const float* mu = mus + g * dim;
  const float* H = halves + g * dim * dim;
  for (int i = 0; i < dim; ++i) {
    float s = mu[i];
    for (int k = 0; k < dim; ++k) s = fmaf(H[i * dim + k], xi[k], s);
    out[p * dim + i] = s;
  }
}

constexpr int kOuMaxD = 16;

__global__ void ou_exact_kernel(float* __restrict__ out, float* __restrict__ out_t, int64_t n, int d,
                                const float* __restrict__ U, const float* __restrict__ s,
                                const float* __restrict__ B0, const float* __restrict__ B,
                                const float* __restrict__ mt0, float t_min, float t_max, uint64_t seed,
                                uint64_t particle_offset) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint64_t pid = particle_offset + (uint64_t)p;
  const float t = t_min + (t_max - t_min) * philox_uniform01(seed, pid, kTagOuTime);
  if (out_t) out_t[p] = t;
  float e[kOuMaxD], M[kOuMaxD * kOuMaxD], xi[kOuMaxD], y[kOuMaxD];
  for (int i = 0; i < d; ++i) e[i] = expf(-t * s[i]);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      const float bs = B[i * d + j] / (s[i] + s[j]);
      M[i * d + j] = e[i] * B0[i * d + j] * e[j] + bs - e[i] * bs * e[j];
    }
  // in-place Cholesky M = C C^T (lower)
  for (int j = 0; j < d; ++j) {
    float dg = M[j * d + j];
    for (int k = 0; k < j; ++k) dg -= M[j * d + k] * M[j * d + k];
    dg = sqrtf(fmaxf(dg, 0.f));
    M[j * d + j] = dg;
    const float inv = dg > 0.f ? 1.f / dg : 0.f;
    for (int i = j + 1; i < d; ++i) {
      float v = M[i * d + j];
      for (int k = 0; k < j; ++k) v -= M[i * d + k] * M[j * d + k];
      M[i * d + j] = v * inv;
    }
  }
  for (int j = 0; j < (d + 3) / 4; ++j) {
    float r4[4];
    philox_normal4(seed, pid, kTagOuX, (uint32_t)j, r4);
    for (int k = 0; k < 4; ++k)
      if (4 * j + k < d) xi[4 * j + k] = r4[k];
  }
  for (int i = 0; i < d; ++i) {
    float v = e[i] * mt0[i];
    for (int k = 0; k <= i; ++k) v = fmaf(M[i * d + k], xi[k], v);
    y[i] = v;
  }
  for (int i = 0; i < d; ++i) {
    float v = 0.f;
    for (int k = 0; k < d; ++k) v = fmaf(U[i * d + k], y[k], v);
    out[p * d + i] = v;
  }
}

}  // namespace pdeip

using namespace pdeip;

extern "C" int pdeip_gaussian_sample_grouped(float* out, int64_t n_groups, int per_group, int dim, const float* mus,
                                             const float* cov_halves, uint64_t seed, uint64_t particle_offset,
                                             void* stream) {
  PDEIP_REQUIRE(out && mus && cov_halves, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(dim >= 1 && dim <= 64, PDEIP_ERR_UNSUPPORTED, "1 <= dim <= 64 supported (got %d)", dim);
  PDEIP_REQUIRE(n_groups >= 0 && per_group >= 1, PDEIP_ERR_INVALID_ARG, "bad group sizes");
  const int64_t n = n_groups * per_group;
  if (n == 0) return PDEIP_OK;
  gaussian_grouped_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      out, n_groups, per_group, dim, mus, cov_halves, seed, particle_offset);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

extern "C" int pdeip_ou_exact_sample(float* out, float* out_t, int64_t n, int d, const float* U, const float* s,
                                     const float* B0, const float* B, const float* mt0, float t_min, float t_max,
                                     uint64_t seed, uint64_t particle_offset, void* stream) {
  PDEIP_REQUIRE(out && U && s && B0 && B && mt0, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(d >= 1 && d <= kOuMaxD, PDEIP_ERR_UNSUPPORTED, "1 <= d <= %d supported (got %d)", kOuMaxD, d);
  if (n <= 0) return n == 0 ? PDEIP_OK : PDEIP_ERR_INVALID_ARG;
  ou_exact_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      out, out_t, n, d, U, s, B0, B, mt0, t_min, t_max, seed, particle_offset);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}
