// residual_common.cuh — workspace protocol shared by the residual kernels (MLP and parametric).
//
// Workspace = G partial vectors of PSTRIDE floats, one per persistent CTA:
//     [0, P)               partial parameter gradient of that CTA (flat layout of pdeip.h)
//     [P, P + NUM_SUMS)    partial loss sums (PDEIP_SUM_*)
// `begin` zeroes it, every `accumulate` launch adds CTA b's contribution to partial b (only CTA b
// ever touches it, and launches are stream-ordered, so there are no atomics and the result is
// bit-reproducible), `finalize` reduces over b in a fixed order.
#pragma once

#include "common.cuh"

namespace pdeip {

inline int residual_grid() { return sm_count(); }
inline int64_t residual_pstride(int64_t n_params) { return ((n_params + PDEIP_NUM_SUMS + 31) / 32) * 32; }
inline size_t residual_ws_bytes(int64_t n_params) {
  return sizeof(float) * (size_t)residual_grid() * (size_t)residual_pstride(n_params);
}

// true-gradient specification carried into the kernels
struct TrueGrad {
  int kind;             // PDEIP_DRIFT_NONE / LINEAR / GMM
  const float* params;  // A [d][d] or mus [K][d]
  int n_gaussian;
  float inv_sigma2;
};

struct ResidualArgs {
  const float* params;
  const float* points;
  int64_t n_points;
  int layout;
  int d;
  int layers;  // LH
  float weight;
  float coef;
  TrueGrad tg;
  float* ws;       // this launch's workspace
  int64_t pstride;
  // KMV pairs: Delta = x[j,t] - ref[i,t], j < kmv_n, i < (n_points / (kmv_n * kmv_nt)); ref == NULL: the batch itself
  const float* G;
  const float* c;
  const float* ref;
  int64_t kmv_n;
  int kmv_nt;
  // tensor path, FP 0T set: fp_dirs = d direction tiles (x, e_i) per point tile (0: kinetic point set)
  int fp_dirs;
};

// grad V_true at x (runtime d), parameters in shared memory (unpadded)
__device__ __forceinline__ void true_grad_thread(const TrueGrad& tg, const float* __restrict__ tp_s, int d,
                                                 const float* x, float* gt) {
  if (tg.kind == PDEIP_DRIFT_LINEAR) {
    for (int i = 0; i < d; ++i) {
      float s = 0.f;
      for (int k = 0; k < d; ++k) s = fmaf(tp_s[i * d + k], x[k], s);
      gt[i] = s;
    }
  } else if (tg.kind == PDEIP_DRIFT_GMM) {
    // two-pass max / exp-sum, as jax.scipy.special.logsumexp does (core/potential.py:34-35)
    float m = -INFINITY;
    for (int k = 0; k < tg.n_gaussian; ++k) {
      float s = 0.f;
      for (int i = 0; i < d; ++i) {
        const float r = x[i] - tp_s[k * d + i];
        s = fmaf(r, r, s);
      }
      m = fmaxf(m, -0.5f * tg.inv_sigma2 * s);
    }
    float se = 0.f;
    for (int i = 0; i < d; ++i) gt[i] = 0.f;
    for (int k = 0; k < tg.n_gaussian; ++k) {
      float s = 0.f;
      for (int i = 0; i < d; ++i) {
        const float r = x[i] - tp_s[k * d + i];
        s = fmaf(r, r, s);
      }
      const float e = expf(-0.5f * tg.inv_sigma2 * s - m);
      se += e;
      for (int i = 0; i < d; ++i) gt[i] = fmaf(e, tp_s[k * d + i], gt[i]);
    }
    const float inv = 1.f / se;
    for (int i = 0; i < d; ++i) gt[i] = (x[i] - gt[i] * inv) * tg.inv_sigma2;
  } else {
    for (int i = 0; i < d; ++i) gt[i] = 0.f;
  }
}

inline int true_grad_floats(const TrueGrad& tg, int d) {
  if (tg.kind == PDEIP_DRIFT_LINEAR) return d * d;
  if (tg.kind == PDEIP_DRIFT_GMM) return tg.n_gaussian * d;
  return 0;
}

}  // namespace pdeip
