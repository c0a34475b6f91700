// residual_kmv.cu — KMV pairwise residual (kinetic_mckean_vlasov.py:11-120), phase 1 and the G-dependent sums.
//
// The reference evaluates Phi, grad Phi and v'H v on the full pair tensor Delta[m,n,nt,d] = x_j - ref_i with the
// batch as its own reference set (m = n, :20-23).  loss_nabla = mean_j |mean_i grad Phi(Delta_ij)|^2 is not a
// plain sum over pairs, so the work is split:
//   phase 1 (here)   G[j,t] = mean_i grad Phi(Delta_ij)  (and the same for Phi_true = Delta' F Delta / 2)
//   phase 2          pair set PDEIP_SET_KMV_PAIRS of the residual kernels with the extra direction u = G_j
//                    (d|G_j|^2 = (2/m) sum_i d D_{G_j} Phi(Delta_ij), G_j held constant) and kappa = 2 c_j.
#include "mlp_thread.cuh"
#include "residual_common.cuh"

namespace pdeip {

constexpr int kKmvMaxChunks = 64;

inline int kmv_chunks(int64_t n) {
  int64_t c = (n + 255) / 256;
  return (int)(c < 1 ? 1 : (c > kKmvMaxChunks ? kKmvMaxChunks : c));
}

// one thread per (j,t); blockIdx.y = chunk of the reference index i
template <int MODEL, int H, int LHMAX>
__global__ void __launch_bounds__(128) kmv_phase1_kernel(const float* __restrict__ params, int d, int layers,
                                                         const float* __restrict__ xv, int64_t n, int nt,
                                                         const float* __restrict__ ref, int64_t m,
                                                         int n_chunks, float* __restrict__ ws) {
  extern __shared__ __align__(16) float smem[];
  const MlpShape<H> sh{d, layers};
  const int P = (MODEL == PDEIP_MODEL_MLP) ? sh.num_params() : d * d + d;
  for (int i = threadIdx.x; i < P; i += blockDim.x) smem[i] = params[i];
  __syncthreads();
  const int64_t jt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (jt >= n * nt) return;
  const int t = (int)(jt % nt);
  const int chunk = blockIdx.y;
  const int64_t i_lo = (m * chunk) / n_chunks, i_hi = (m * (chunk + 1)) / n_chunks;  // reference samples of this chunk
  float xj[kDMax], acc[kDMax], accd[kDMax];
  for (int c = 0; c < d; ++c) {
    xj[c] = xv[jt * 2 * d + c];
    acc[c] = 0.f;
    accd[c] = 0.f;
  }
  PointState<H, LHMAX> st;
  MlpThread<H, LHMAX> net(sh, smem, st);
  for (int64_t i = i_lo; i < i_hi; ++i) {
    const float* xi = ref + (i * nt + t) * 2 * d;
    float g[kDMax];
    for (int c = 0; c < d; ++c) {
      st.x[c] = xj[c] - xi[c];
      accd[c] += st.x[c];
    }
    if (MODEL == PDEIP_MODEL_MLP) {
      net.primal_forward();
      net.input_gradient(g);
    } else {  // quadratic: (W + W^T) y + b
      for (int a = 0; a < d; ++a) {
        float s = smem[d * d + a];
        for (int b = 0; b < d; ++b) s = fmaf(smem[a * d + b] + smem[b * d + a], st.x[b], s);
        g[a] = s;
      }
    }
    for (int c = 0; c < d; ++c) acc[c] += g[c];
  }
  float* o = ws + ((int64_t)chunk * n * nt + jt) * 2 * d;
  for (int c = 0; c < d; ++c) {
    o[c] = acc[c];
    o[d + c] = accd[c];
  }
}

__global__ void kmv_phase1_reduce_kernel(const float* __restrict__ ws, int64_t n, int nt, int d, int64_t m, int n_chunks,
                                         const float* __restrict__ true_A, float* __restrict__ out_G,
                                         float* __restrict__ out_Gtrue) {
  const int64_t jt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (jt >= n * nt) return;
  float g[kDMax], sd[kDMax];
  for (int c = 0; c < d; ++c) { g[c] = 0.f; sd[c] = 0.f; }
  for (int ch = 0; ch < n_chunks; ++ch) {
    const float* o = ws + ((int64_t)ch * n * nt + jt) * 2 * d;
    for (int c = 0; c < d; ++c) { g[c] += o[c]; sd[c] += o[d + c]; }
  }
  const float inv = 1.f / (float)m;
  for (int c = 0; c < d; ++c) out_G[jt * d + c] = g[c] * inv;
  if (out_Gtrue && true_A) {
    for (int a = 0; a < d; ++a) {
      float s = 0.f;
      for (int b = 0; b < d; ++b) s = fmaf(true_A[a * d + b], sd[b] * inv, s);
      out_Gtrue[jt * d + a] = s;
    }
  }
}

// adds  w * sum|G|^2, w * sum|Gtrue|^2, w * sum|Gtrue - G|^2  to the sums of workspace partial 0
__global__ void __launch_bounds__(256) kmv_g_sums_kernel(const float* __restrict__ G, const float* __restrict__ Gtrue,
                                                         int64_t n_rows, int d, float w, float* __restrict__ part_sums) {
  __shared__ float red[3][8];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int64_t idx = threadIdx.x; idx < n_rows * d; idx += blockDim.x) {
    const float g = G[idx];
    const float gt = Gtrue ? Gtrue[idx] : 0.f;
    s0 = fmaf(g, g, s0);
    s1 = fmaf(gt, gt, s1);
    s2 = fmaf(gt - g, gt - g, s2);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; red[2][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int k = 0; k < 8; ++k) { a += red[0][k]; b += red[1][k]; c += red[2][k]; }
    part_sums[PDEIP_SUM_G2] += w * a;
    part_sums[PDEIP_SUM_GTRUE2] += w * b;
    part_sums[PDEIP_SUM_GT] += w * c;
    part_sums[PDEIP_SUM_LOSS] += w * (a + b);
  }
}

size_t kmv_ws_bytes(int64_t n, int nt, int d, int64_t m) {
  return sizeof(float) * (size_t)kmv_chunks(m) * (size_t)n * nt * 2 * d;
}

int kmv_mean_grad(int model_kind, const float* params, int d, int hidden, int layers, const float* xv, int64_t n,
                  int nt, const float* ref, int64_t m, float* out_G, float* out_Gtrue, const float* true_A,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  PDEIP_REQUIRE(d >= 1 && d <= kDMax, PDEIP_ERR_UNSUPPORTED, "1 <= d <= %d supported", kDMax);
  PDEIP_REQUIRE(workspace && workspace_bytes >= kmv_ws_bytes(n, nt, d, m), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", kmv_ws_bytes(n, nt, d, m), workspace_bytes);
  const int n_chunks = kmv_chunks(m);
  dim3 grid((unsigned)((n * nt + 127) / 128), n_chunks);
  float* ws = (float*)workspace;
  if (model_kind == PDEIP_MODEL_MLP) {
    PDEIP_REQUIRE(hidden == 32 && layers >= 1 && layers <= 8, PDEIP_ERR_UNSUPPORTED,
                  "KMV MLP path needs hidden_dim == 32 and 1 <= layers <= 8");
    const MlpShape<32> sh{d, layers};
    const size_t smem = sizeof(float) * sh.num_params();
    if (layers <= 2) kmv_phase1_kernel<PDEIP_MODEL_MLP, 32, 2><<<grid, 128, smem, st>>>(params, d, layers, xv, n, nt, ref, m, n_chunks, ws);
    else if (layers <= 4) kmv_phase1_kernel<PDEIP_MODEL_MLP, 32, 4><<<grid, 128, smem, st>>>(params, d, layers, xv, n, nt, ref, m, n_chunks, ws);
    else kmv_phase1_kernel<PDEIP_MODEL_MLP, 32, 8><<<grid, 128, smem, st>>>(params, d, layers, xv, n, nt, ref, m, n_chunks, ws);
  } else if (model_kind == PDEIP_MODEL_QUADRATIC) {
    kmv_phase1_kernel<PDEIP_MODEL_QUADRATIC, 32, 2><<<grid, 128, sizeof(float) * (d * d + d), st>>>(
        params, d, 1, xv, n, nt, ref, m, n_chunks, ws);
  } else {
    PDEIP_REQUIRE(false, PDEIP_ERR_UNSUPPORTED, "KMV residual supports the MLP and quadratic models");
  }
  PDEIP_LAUNCH_OK();
  kmv_phase1_reduce_kernel<<<(unsigned)((n * nt + 127) / 128), 128, 0, st>>>(ws, n, nt, d, m, n_chunks, true_A, out_G,
                                                                             out_Gtrue);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

// ---- moment closure of the quadratic interaction model ------------------------------------------
// Phi(D) = D'W D + b.D.  With ybar = x_j - rbar_t and C_t = mean_i (r_i - rbar_t)(r_i - rbar_t)':
//   mean_i grad Phi(D_ij) = (W + W') ybar + b,   mean_i v'H v = 2 v'W v,   mean_i Phi(D_ij) = Phi(ybar) + tr(W C_t),
// so the O(n m) pair set collapses to ONE pseudo-pair per sample against the reference mean (the pair kernel with
// ref = rbar, m = 1) plus this correction:  loss += sum_jt kappa_jt tr(W C_t),  dW += sum_jt kappa_jt C_t,
// kappa_jt = 2 c_jt weight  (kinetic_mckean_vlasov.py:84-92: the value term is the only one that sees C_t).
__global__ void __launch_bounds__(256) kmv_closure_correction_kernel(const float* __restrict__ W, int d, int64_t n, int nt,
                                                                     const float* __restrict__ c,
                                                                     const float* __restrict__ cov, float weight,
                                                                     float* __restrict__ part) {
  __shared__ float ksum[64];  // sum_j kappa_jt per time stamp (nt <= 64)
  __shared__ float red[8];
  for (int t = 0; t < nt; ++t) {
    float s = 0.f;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s += c[j * nt + t];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
      for (int k = 0; k < 8; ++k) a += red[k];
      ksum[t] = 2.f * weight * a;
    }
    __syncthreads();
  }
  float tr = 0.f;
  for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
    float g = 0.f;
    for (int t = 0; t < nt; ++t) g = fmaf(ksum[t], cov[(int64_t)t * d * d + idx], g);
    part[idx] += g;  // flat layout of the quadratic model: W [d][d] first
    tr = fmaf(W[idx], g, tr);
  }
  tr = warp_sum(tr);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tr;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k < 8; ++k) a += red[k];
    const int P = d * d + d;
    part[P + PDEIP_SUM_D1] += 0.5f * a;
    part[P + PDEIP_SUM_LOSS] += a;
  }
}

int kmv_closure_correction(const float* W, int d, int64_t n, int nt, const float* c, const float* cov, float weight,
                           float* part, cudaStream_t st) {
  PDEIP_REQUIRE(nt >= 1 && nt <= 64, PDEIP_ERR_UNSUPPORTED, "closure correction supports 1 <= nt <= 64 (got %d)", nt);
  kmv_closure_correction_kernel<<<1, 256, 0, st>>>(W, d, n, nt, c, cov, weight, part);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

// ---- time derivatives of the log density of the Gaussian x-marginal, on the device --------------------------------
// example_problems/kinetic_mckean_vlasov_example_quadratic.py:51-69 (d_s log rho) and :120-177 (d_s^2 log rho) are both
// quadratic forms in diff = mean1 - x whose coefficients depend on the time stamp only.  coef [nt][KC], KC = 3 d + 2 +
// 2 d^2: [mean1 (d) | a1 (d) | a2 (d) | k1 | k2 | M1 (d*d) | M2 (d*d)] with
//   d_s  log rho = -a1.diff + k1 - diff'M1 diff / 2,   d_ss log rho = -a2.diff + k2 - diff'M2 diff / 2
// (host float64, once per time stamp: utils/lyapunov.kmv_density_coefficients).  out[t * n + j] = d_ss + d_s^2 + gamma d_s
// at (tau_t, x[j, t]): the [nt, n] array the reference then reshapes to [n, nt] (kinetic_mckean_vlasov.py:57-72).
__global__ void __launch_bounds__(128) kmv_density_terms_kernel(const float* __restrict__ xv, int64_t n, int nt, int d,
                                                                const float* __restrict__ coef, float gamma,
                                                                float* __restrict__ out_c, float* __restrict__ out_ps,
                                                                float* __restrict__ out_ps2) {
  const int64_t jt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (jt >= n * nt) return;
  const int64_t j = jt / nt;
  const int t = (int)(jt - j * nt);
  const int KC = 3 * d + 2 + 2 * d * d;
  const float* ct = coef + (int64_t)t * KC;
  const float* M1 = ct + 3 * d + 2;
  const float* M2 = M1 + d * d;
  float diff[kDMax];
  for (int i = 0; i < d; ++i) diff[i] = __ldg(ct + i) - xv[jt * 2 * d + i];
  float l1 = 0.f, l2 = 0.f, q1 = 0.f, q2 = 0.f;
  for (int i = 0; i < d; ++i) {
    l1 = fmaf(__ldg(ct + d + i), diff[i], l1);
    l2 = fmaf(__ldg(ct + 2 * d + i), diff[i], l2);
    float r1 = 0.f, r2 = 0.f;
    for (int k = 0; k < d; ++k) {
      r1 = fmaf(__ldg(M1 + i * d + k), diff[k], r1);
      r2 = fmaf(__ldg(M2 + i * d + k), diff[k], r2);
    }
    q1 = fmaf(diff[i], r1, q1);
    q2 = fmaf(diff[i], r2, q2);
  }
  const float ps = -l1 + __ldg(ct + 3 * d) - 0.5f * q1;
  const float ps2 = -l2 + __ldg(ct + 3 * d + 1) - 0.5f * q2;
  const int64_t o = (int64_t)t * n + j;
  if (out_ps) out_ps[o] = ps;
  if (out_ps2) out_ps2[o] = ps2;
  if (out_c) out_c[o] = ps2 + ps * ps + gamma * ps;
}

int kmv_density_terms(const float* xv, int64_t n, int nt, int d, const float* coef, float gamma, float* out_c,
                      float* out_ps, float* out_ps2, cudaStream_t st) {
  PDEIP_REQUIRE(d >= 1 && d <= kDMax, PDEIP_ERR_UNSUPPORTED, "1 <= d <= %d supported", kDMax);
  kmv_density_terms_kernel<<<(unsigned)((n * nt + 127) / 128), 128, 0, st>>>(xv, n, nt, d, coef, gamma, out_c, out_ps,
                                                                             out_ps2);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

int kmv_g_sums(const float* G, const float* Gtrue, int64_t n_rows, int d, float w, float* part_sums,
               cudaStream_t st) {
  kmv_g_sums_kernel<<<1, 256, 0, st>>>(G, Gtrue, n_rows, d, w, part_sums);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

}  // namespace pdeip
