// parametric.cu — K5: closed-form residual (loss terms + parameter gradient) of the parametric models.
//
//   GMM model        V(y) = -logsumexp_k(-|y - mu_k|^2 / 2), learnable mus[K][d]
//                    (example_problems/kinetic_fokker_planck_example_GMM.py:214-234)
//   quadratic model  V(y) = y.(yW + b), Flax Dense kernel W[in][out], bias b
//                    (example_problems/kinetic_fokker_planck_example_OU.py:209-220,
//                     example_problems/kinetic_mckean_vlasov_example_quadratic.py:205-216)
// They replace jax.value_and_grad of the same losses (kinetic_fokker_planck.py:33-61,
// kinetic_mckean_vlasov.py:74-112) by the closed forms of SURVEY.md §9.5 (checked against autodiff in
// oracle/taylor.py).  Per point  l = alpha D_v^2 V + beta D_v V + beta2 D_u V + kappa V + c_g |grad V|^2 ;
// the parameter gradient is a sum of batch-reduced outer products  sum_p a_p z_p^T , accumulated per warp
// through two shared-memory tiles exactly like dW in the MLP kernel.
#include "common.cuh"
#include "residual_common.cuh"

namespace pdeip {

constexpr int kPD = 32;      // max d
constexpr int kPK = 64;      // max number of Gaussians
constexpr int kPNW = 8;      // warps per CTA
constexpr int kPZS = 36;     // tileZ row stride (floats), 16-byte aligned and bank-conflict free

struct ParamScratch {
  float* acc;    // per-warp accumulators [n_cols][RS] (+ extra)
  float* tileA;  // [32][rows_max + 1]
  float* tileZ;  // [32][kPZS]
};

// acc[c * RS + r] += sum_{p in warp} a_p[r] * z_p[c],  r < n_rows, c < d.  If bias_off >= 0 also
// acc[bias_off + r] += sum_p a_p[r].
template <int RS>
__device__ __forceinline__ void warp_outer(const ParamScratch& ws, int a_stride, int n_rows, int d, int acc_off,
                                           int bias_off, const float* a_local, const float* z_local) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (int i = 0; i < n_rows; ++i) ws.tileA[lane * a_stride + i] = a_local[i];
  for (int c = 0; c < kPD; ++c) ws.tileZ[lane * kPZS + c] = c < d ? z_local[c] : 0.f;
  __syncwarp();
  for (int r = lane; r < n_rows; r += 32) {
    float acc[kPD];
#pragma unroll
    for (int c = 0; c < kPD; ++c) acc[c] = 0.f;
    float asum = 0.f;
    for (int p = 0; p < 32; ++p) {
      const float a = ws.tileA[p * a_stride + r];
      asum += a;
      const float4* zr = reinterpret_cast<const float4*>(ws.tileZ + p * kPZS);
#pragma unroll
      for (int c4 = 0; c4 < kPD / 4; ++c4) {
        const float4 z = zr[c4];
        acc[4 * c4 + 0] = fmaf(a, z.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(a, z.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(a, z.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(a, z.w, acc[4 * c4 + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < kPD; ++c)
      if (c < d) ws.acc[acc_off + c * RS + r] += acc[c];
    if (bias_off >= 0) ws.acc[bias_off + r] += asum;
  }
}

// ------------------------------------------------------------------------------------------------
// GMM parametric model
// ------------------------------------------------------------------------------------------------
struct GmmPoint {
  float w[kPK], c[kPK], cg[kPK];
  float g[kPD];
  float D1, D2, value;
};

// forward quantities at (y, v): softmax weights, c_k = r_k.v, g = E[r], cg_k = r_k.g
__device__ __forceinline__ void gmm_point_forward(const float* __restrict__ mus_s, int K, int d, const float* y,
                                                  const float* v, GmmPoint& o) {
  float m = -INFINITY;
  for (int k = 0; k < K; ++k) {
    float s = 0.f, cv = 0.f;
    for (int i = 0; i < d; ++i) {
      const float r = y[i] - mus_s[k * d + i];
      s = fmaf(r, r, s);
      cv = fmaf(r, v[i], cv);
    }
    o.w[k] = -0.5f * s;
    o.c[k] = cv;
    m = fmaxf(m, o.w[k]);
  }
  float se = 0.f;
  for (int k = 0; k < K; ++k) {
    o.w[k] = expf(o.w[k] - m);
    se += o.w[k];
  }
  const float inv = 1.f / se;
  for (int i = 0; i < d; ++i) o.g[i] = 0.f;
  float Ec = 0.f, Ec2 = 0.f;
  for (int k = 0; k < K; ++k) {
    o.w[k] *= inv;
    Ec = fmaf(o.w[k], o.c[k], Ec);
    Ec2 = fmaf(o.w[k] * o.c[k], o.c[k], Ec2);
    for (int i = 0; i < d; ++i) o.g[i] = fmaf(o.w[k], y[i] - mus_s[k * d + i], o.g[i]);
  }
  for (int k = 0; k < K; ++k) {
    float s = 0.f;
    for (int i = 0; i < d; ++i) s = fmaf(y[i] - mus_s[k * d + i], o.g[i], s);
    o.cg[k] = s;
  }
  float v2 = 0.f;
  for (int i = 0; i < d; ++i) v2 = fmaf(v[i], v[i], v2);
  o.D1 = Ec;
  o.D2 = v2 - (Ec2 - Ec * Ec);
  o.value = -(m + logf(se));
}

template <int SET>
__global__ void __launch_bounds__(kPNW * 32, 1) gmm_param_residual_kernel(const ResidualArgs a, int K) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RS = 64;
  const int d = a.d;
  const int P = K * d;
  float* mus_s = smem;
  float* tp = mus_s + ((P + 3) & ~3);
  int ntg = 0;
  if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = d * d;
  else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * d;
  float* warp_base = tp + ((ntg + 3) & ~3);
  const int acc_sz = kPD * RS + RS;  // [d cols][RS rows] + srsum[RS]
  const int a_stride = kPK + 1;
  const int per_warp = acc_sz + 32 * a_stride + 32 * kPZS + 1;  // +1 keeps tileZ 16B aligned? handled below
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ParamScratch ws;
  const int per_warp_al = (per_warp + 3) & ~3;
  ws.acc = warp_base + warp * per_warp_al;
  ws.tileZ = ws.acc + acc_sz;               // acc_sz is a multiple of 4 -> 16B aligned
  ws.tileA = ws.tileZ + 32 * kPZS;
  for (int i = threadIdx.x; i < P; i += blockDim.x) mus_s[i] = a.params[i];
  for (int i = threadIdx.x; i < ntg; i += blockDim.x) tp[i] = a.tg.params[i];
  for (int i = lane; i < acc_sz; i += 32) ws.acc[i] = 0.f;
  __syncthreads();

  float sums[PDEIP_NUM_SUMS];
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;
  const int64_t tile_pts = kPNW * 32;
  const int64_t n_tiles = (a.n_points + tile_pts - 1) / tile_pts;
  GmmPoint pt;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * tile_pts + threadIdx.x;
    const bool valid = p < a.n_points;
    const float wt = valid ? a.weight : 0.f;
    float y[kPD], v[kPD];
    const int dimw = 2 * d + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? d : 0);
    for (int i = 0; i < d; ++i) {
      y[i] = valid ? a.points[elem_index(a.layout, p, i, a.n_points, dimw)] : 0.f;
      v[i] = valid ? a.points[elem_index(a.layout, p, d + i, a.n_points, dimw)] : 0.f;
    }
    gmm_point_forward(mus_s, K, d, y, v, pt);
    float alpha, beta, cg;
    if (SET == PDEIP_SET_KFP_0T) {
      const float gamma = a.coef;
      alpha = -2.f * wt; beta = 2.f * gamma * wt; cg = wt;
      float gt[kPD];
      if (a.tg.kind == PDEIP_DRIFT_IN_POINTS) {
        for (int i = 0; i < d; ++i) gt[i] = valid ? a.points[elem_index(a.layout, p, 2 * d + i, a.n_points, dimw)] : 0.f;
      } else {
        true_grad_thread(a.tg, tp, d, y, gt);
      }
      float g2 = 0.f, gt2 = 0.f, gd2 = 0.f;
      for (int i = 0; i < d; ++i) {
        g2 = fmaf(pt.g[i], pt.g[i], g2);
        gt2 = fmaf(gt[i], gt[i], gt2);
        const float df = gt[i] - pt.g[i];
        gd2 = fmaf(df, df, gd2);
      }
      sums[PDEIP_SUM_G2] += wt * g2;
      sums[PDEIP_SUM_D2] += wt * pt.D2;
      sums[PDEIP_SUM_D1] += wt * pt.D1;
      sums[PDEIP_SUM_GTRUE2] += wt * gt2;
      sums[PDEIP_SUM_GT] += wt * gd2;
      sums[PDEIP_SUM_LOSS] += wt * (g2 - 2.f * pt.D2 + 2.f * gamma * pt.D1 + gt2);
    } else {
      alpha = 0.f; beta = a.coef * wt; cg = 0.f;
      sums[PDEIP_SUM_BOUNDARY] += wt * a.coef * pt.D1;
      sums[PDEIP_SUM_LOSS] += wt * a.coef * pt.D1;
    }
    // coefficients of g, v and r_k = y - mu_k in d l / d mu_k   
    const float beta_g = 2.f * cg;
    float Ec = 0.f, Ec2 = 0.f, Ecg = 0.f;
    for (int k = 0; k < K; ++k) {
      Ec = fmaf(pt.w[k], pt.c[k], Ec);
      Ec2 = fmaf(pt.w[k] * pt.c[k], pt.c[k], Ec2);
      Ecg = fmaf(pt.w[k], pt.cg[k], Ecg);
    }
    float sg[kPK], sv[kPK], sr[kPK];
    for (int k = 0; k < K; ++k) {
      const float w = pt.w[k], c = pt.c[k];
      sg[k] = -beta_g * w;
      sv[k] = -beta * w + alpha * (2.f * w * c - 2.f * Ec * w);
      sr[k] = beta_g * w * (pt.cg[k] - Ecg) + beta * w * (c - Ec) +
              alpha * (-w * (c * c - Ec2) + 2.f * Ec * w * (c - Ec));
    }
    warp_outer<RS>(ws, a_stride, K, d, 0, -1, sg, pt.g);
    warp_outer<RS>(ws, a_stride, K, d, 0, -1, sv, v);
    warp_outer<RS>(ws, a_stride, K, d, 0, kPD * RS, sr, y);
  }
  // ---- CTA reduction: dmus[k][i] = acc[i][k] - mu[k][i] * srsum[k] ------------------------------------
  __syncthreads();
  float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
  for (int idx = threadIdx.x; idx < P; idx += blockDim.x) {
    const int k = idx / d, i = idx - k * d;
    float s = 0.f, srs = 0.f;
    for (int w = 0; w < kPNW; ++w) {
      s += warp_base[w * per_warp_al + i * RS + k];
      srs += warp_base[w * per_warp_al + kPD * RS + k];
    }
    part[idx] += s - mus_s[idx] * srs;
  }
  __syncthreads();
  float* red = warp_base + acc_sz;  // warp 0 tileZ
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
    const float s = warp_sum(sums[k]);
    if (lane == 0) red[warp * PDEIP_NUM_SUMS + k] = s;
  }
  __syncthreads();
  if (threadIdx.x < PDEIP_NUM_SUMS) {
    float s = 0.f;
    for (int w = 0; w < kPNW; ++w) s += red[w * PDEIP_NUM_SUMS + threadIdx.x];
    part[P + threadIdx.x] += s;
  }
}

// ------------------------------------------------------------------------------------------------
// quadratic parametric model (also the KMV pair set)
// ------------------------------------------------------------------------------------------------
// g = (W + W^T) y + b ; D_vV = g.v ; D_v^2 V = 2 v^T W v ; V = y^T W y + b.y
__device__ __forceinline__ void quad_point_forward(const float* __restrict__ W_s, const float* __restrict__ b_s,
                                                   int d, const float* y, const float* v, float* g, float& D1,
                                                   float& D2, float& value) {
  float vWv = 0.f, yWy = 0.f, by = 0.f;
  for (int a_ = 0; a_ < d; ++a_) {
    float s = b_s[a_], t = 0.f, u = 0.f;
    for (int b_ = 0; b_ < d; ++b_) {
      s = fmaf(W_s[a_ * d + b_] + W_s[b_ * d + a_], y[b_], s);
      t = fmaf(W_s[a_ * d + b_], v[b_], t);
      u = fmaf(W_s[a_ * d + b_], y[b_], u);
    }
    g[a_] = s;
    vWv = fmaf(v[a_], t, vWv);
    yWy = fmaf(y[a_], u, yWy);
    by = fmaf(b_s[a_], y[a_], by);
  }
  float d1 = 0.f;
  for (int i = 0; i < d; ++i) d1 = fmaf(g[i], v[i], d1);
  D1 = d1;
  D2 = 2.f * vWv;
  value = yWy + by;
}

template <int SET>
__global__ void __launch_bounds__(kPNW * 32, 1) quad_param_residual_kernel(const ResidualArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RS = 32;
  const int d = a.d;
  const int P = d * d + d;
  float* W_s = smem;
  float* b_s = W_s + d * d;
  float* tp = smem + ((P + 3) & ~3);
  int ntg = 0;
  if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = d * d;
  else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * d;
  float* warp_base = tp + ((ntg + 3) & ~3);
  const int acc_sz = kPD * RS + RS;  // dW as [j cols][RS rows i] + db[RS]
  const int a_stride = kPD + 1;
  const int per_warp_al = (acc_sz + 32 * kPZS + 32 * a_stride + 3) & ~3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ParamScratch ws;
  ws.acc = warp_base + warp * per_warp_al;
  ws.tileZ = ws.acc + acc_sz;
  ws.tileA = ws.tileZ + 32 * kPZS;
  for (int i = threadIdx.x; i < P; i += blockDim.x) smem[i] = a.params[i];
  for (int i = threadIdx.x; i < ntg; i += blockDim.x) tp[i] = a.tg.params[i];
  for (int i = lane; i < acc_sz; i += 32) ws.acc[i] = 0.f;
  __syncthreads();

  float sums[PDEIP_NUM_SUMS];
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;
  const int64_t tile_pts = kPNW * 32;
  const int64_t n_tiles = (a.n_points + tile_pts - 1) / tile_pts;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * tile_pts + threadIdx.x;
    const bool valid = p < a.n_points;
    const float wt = valid ? a.weight : 0.f;
    float y[kPD], v[kPD], u[kPD], g[kPD];
    const int dimw = 2 * d + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? d : 0);
    float alpha = 0.f, beta = 0.f, beta2 = 0.f, kappa = 0.f, cg = 0.f;
    if (SET == PDEIP_SET_KMV_PAIRS) {
      // pair index p -> (i, j, t): Delta = x[j,t] - x[i,t], v = v[j,t], u = G[j,t], kappa = 2 c[j,t]
      const int64_t n = a.kmv_n;
      const int nt = a.kmv_nt;
      const int64_t pp = valid ? p : 0;
      const int t = (int)(pp % nt);
      const int64_t ij = pp / nt;
      const int64_t j = ij % n, i = ij / n;
      const float* rj = a.points + (j * nt + t) * 2 * d;
      const float* ri = (a.ref ? a.ref : a.points) + (i * nt + t) * 2 * d;
      for (int c = 0; c < d; ++c) {
        y[c] = valid ? rj[c] - ri[c] : 0.f;
        v[c] = valid ? rj[d + c] : 0.f;
        u[c] = (valid && a.G) ? a.G[(j * nt + t) * d + c] : 0.f;
      }
      alpha = -2.f * wt;
      beta2 = a.G ? 2.f * wt : 0.f;
      kappa = valid ? 2.f * a.c[j * nt + t] * wt : 0.f;
    } else {
      for (int i = 0; i < d; ++i) {
        y[i] = valid ? a.points[elem_index(a.layout, p, i, a.n_points, dimw)] : 0.f;
        v[i] = valid ? a.points[elem_index(a.layout, p, d + i, a.n_points, dimw)] : 0.f;
        u[i] = 0.f;
      }
    }
    float D1, D2, value;
    quad_point_forward(W_s, b_s, d, y, v, g, D1, D2, value);
    if (SET == PDEIP_SET_KFP_0T) {
      const float gamma = a.coef;
      alpha = -2.f * wt; beta = 2.f * gamma * wt; cg = wt;
      float gt[kPD];
      if (a.tg.kind == PDEIP_DRIFT_IN_POINTS) {
        for (int i = 0; i < d; ++i) gt[i] = valid ? a.points[elem_index(a.layout, p, 2 * d + i, a.n_points, dimw)] : 0.f;
      } else {
        true_grad_thread(a.tg, tp, d, y, gt);
      }
      float g2 = 0.f, gt2 = 0.f, gd2 = 0.f;
      for (int i = 0; i < d; ++i) {
        g2 = fmaf(g[i], g[i], g2);
        gt2 = fmaf(gt[i], gt[i], gt2);
        const float df = gt[i] - g[i];
        gd2 = fmaf(df, df, gd2);
      }
      sums[PDEIP_SUM_G2] += wt * g2;
      sums[PDEIP_SUM_D2] += wt * D2;
      sums[PDEIP_SUM_D1] += wt * D1;
      sums[PDEIP_SUM_GTRUE2] += wt * gt2;
      sums[PDEIP_SUM_GT] += wt * gd2;
      sums[PDEIP_SUM_LOSS] += wt * (g2 - 2.f * D2 + 2.f * gamma * D1 + gt2);
    } else if (SET == PDEIP_SET_KFP_BOUNDARY) {
      beta = a.coef * wt;
      sums[PDEIP_SUM_BOUNDARY] += wt * a.coef * D1;
      sums[PDEIP_SUM_LOSS] += wt * a.coef * D1;
    } else {  // KMV pairs: -2 v'Hv + 2 c Phi
      sums[PDEIP_SUM_D2] += wt * D2;
      sums[PDEIP_SUM_D1] += 0.5f * kappa * value;
      sums[PDEIP_SUM_LOSS] += -2.f * wt * D2 + kappa * value;
    }
    // dW += cg 2 (g y^T + y g^T) + 2 alpha v v^T + beta (v y^T + y v^T) + beta2 (u y^T + y u^T) + kappa y y^T
    // db += 2 cg g + beta v + beta2 u + kappa y
    float a1[kPD], a2[kPD], a3[kPD], a4[kPD];
    for (int i = 0; i < d; ++i) {
      a1[i] = 2.f * cg * g[i] + beta * v[i] + beta2 * u[i] + kappa * y[i];
      a2[i] = 2.f * cg * y[i];
      a3[i] = 2.f * alpha * v[i] + beta * y[i];
      a4[i] = beta2 * y[i];
    }
    warp_outer<RS>(ws, a_stride, d, d, 0, kPD * RS, a1, y);
    if (SET == PDEIP_SET_KFP_0T) warp_outer<RS>(ws, a_stride, d, d, 0, -1, a2, g);
    warp_outer<RS>(ws, a_stride, d, d, 0, -1, a3, v);
    if (SET == PDEIP_SET_KMV_PAIRS) warp_outer<RS>(ws, a_stride, d, d, 0, -1, a4, u);
  }
  __syncthreads();
  float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
  for (int idx = threadIdx.x; idx < P; idx += blockDim.x) {
    float s = 0.f;
    int src;
    if (idx < d * d) {
      const int i = idx / d, j = idx - i * d;
      src = j * RS + i;
    } else {
      src = kPD * RS + (idx - d * d);
    }
    for (int w = 0; w < kPNW; ++w) s += warp_base[w * per_warp_al + src];
    part[idx] += s;
  }
  __syncthreads();
  float* red = warp_base + acc_sz;
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
    const float s = warp_sum(sums[k]);
    if (lane == 0) red[warp * PDEIP_NUM_SUMS + k] = s;
  }
  __syncthreads();
  if (threadIdx.x < PDEIP_NUM_SUMS) {
    float s = 0.f;
    for (int w = 0; w < kPNW; ++w) s += red[w * PDEIP_NUM_SUMS + threadIdx.x];
    part[P + threadIdx.x] += s;
  }
}

// ------------------------------------------------------------------------------------------------
// evaluation kernels (value / grad / v'Hv / Laplacian) of the parametric models
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gmm_param_eval_kernel(const float* __restrict__ params, int K, int d,
                                                             const float* __restrict__ x, const float* __restrict__ v,
                                                             float* out_value, float* out_grad, float* out_vHv,
                                                             float* out_lap, int64_t n) {
  extern __shared__ __align__(16) float smem[];
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) smem[i] = params[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float y[kPD], vv[kPD];
  for (int i = 0; i < d; ++i) {
    y[i] = x[p * d + i];
    vv[i] = v ? v[p * d + i] : 0.f;
  }
  GmmPoint pt;
  gmm_point_forward(smem, K, d, y, vv, pt);
  if (out_value) out_value[p] = pt.value;
  if (out_grad)
    for (int i = 0; i < d; ++i) out_grad[p * d + i] = pt.g[i];
  if (out_vHv) out_vHv[p] = pt.D2;
  if (out_lap) {  // tr H = d - sum_k w_k |r_k|^2 + |E[r]|^2
    float er2 = 0.f, g2 = 0.f;
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int i = 0; i < d; ++i) {
        const float r = y[i] - smem[k * d + i];
        s = fmaf(r, r, s);
      }
      er2 = fmaf(pt.w[k], s, er2);
    }
    for (int i = 0; i < d; ++i) g2 = fmaf(pt.g[i], pt.g[i], g2);
    out_lap[p] = (float)d - (er2 - g2);
  }
}

__global__ void __launch_bounds__(128) quad_param_eval_kernel(const float* __restrict__ params, int d,
                                                              const float* __restrict__ x, const float* __restrict__ v,
                                                              float* out_value, float* out_grad, float* out_vHv,
                                                              float* out_lap, int64_t n) {
  extern __shared__ __align__(16) float smem[];
  for (int i = threadIdx.x; i < d * d + d; i += blockDim.x) smem[i] = params[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float y[kPD], vv[kPD], g[kPD];
  for (int i = 0; i < d; ++i) {
    y[i] = x[p * d + i];
    vv[i] = v ? v[p * d + i] : 0.f;
  }
  float D1, D2, value;
  quad_point_forward(smem, smem + d * d, d, y, vv, g, D1, D2, value);
  if (out_value) out_value[p] = value;
  if (out_grad)
    for (int i = 0; i < d; ++i) out_grad[p * d + i] = g[i];
  if (out_vHv) out_vHv[p] = D2;
  if (out_lap) {
    float tr = 0.f;
    for (int i = 0; i < d; ++i) tr += 2.f * smem[i * d + i];
    out_lap[p] = tr;
  }
}

static size_t param_smem_bytes(int P, int ntg, int rows_max, int RS) {
  const int acc_sz = kPD * RS + RS;
  const int per_warp_al = (acc_sz + 32 * kPZS + 32 * (rows_max + 1) + 1 + 3) & ~3;
  return sizeof(float) * (size_t)(((P + 3) & ~3) + ((ntg + 3) & ~3) + kPNW * per_warp_al);
}

// parametric_fast.cu: production kernels of the GMM model for d in {4, 8, 16, 32}
bool gmm_param_fast_ok(int set_kind, const ResidualArgs& a);
int gmm_param_fast_accumulate(int set_kind, const ResidualArgs& a, int K, cudaStream_t st);
bool quad_param_fast_ok(int set_kind, const ResidualArgs& a);
int quad_param_fast_accumulate(int set_kind, const ResidualArgs& a, cudaStream_t st);

int param_residual_accumulate(int set_kind, int model_kind, const ResidualArgs& a, int n_gaussian, cudaStream_t st) {
  PDEIP_REQUIRE(a.d >= 1 && a.d <= kPD, PDEIP_ERR_UNSUPPORTED, "parametric residual supports 1 <= d <= %d", kPD);
  const int ntg = true_grad_floats(a.tg, a.d);
  const int grid = residual_grid();
  if (model_kind == PDEIP_MODEL_GMM) {
    PDEIP_REQUIRE(n_gaussian >= 1 && n_gaussian <= kPK, PDEIP_ERR_UNSUPPORTED, "1 <= n_gaussian <= %d supported", kPK);
    PDEIP_REQUIRE(set_kind == PDEIP_SET_KFP_0T || set_kind == PDEIP_SET_KFP_BOUNDARY, PDEIP_ERR_UNSUPPORTED,
                  "GMM parametric model supports the kinetic point sets only");
    if (gmm_param_fast_ok(set_kind, a)) return gmm_param_fast_accumulate(set_kind, a, n_gaussian, st);
    const size_t smem = param_smem_bytes(n_gaussian * a.d, ntg, kPK, 64);
    if (set_kind == PDEIP_SET_KFP_0T) {
      auto kern = gmm_param_residual_kernel<PDEIP_SET_KFP_0T>;
      PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kPNW * 32, smem, st>>>(a, n_gaussian);
    } else {
      auto kern = gmm_param_residual_kernel<PDEIP_SET_KFP_BOUNDARY>;
      PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kPNW * 32, smem, st>>>(a, n_gaussian);
    }
  } else if (model_kind == PDEIP_MODEL_QUADRATIC) {
    if (quad_param_fast_ok(set_kind, a)) return quad_param_fast_accumulate(set_kind, a, st);
    const size_t smem = param_smem_bytes(a.d * a.d + a.d, ntg, kPD, 32);
#define LAUNCH_Q(SET)                                                                                   \
  do {                                                                                                  \
    auto kern = quad_param_residual_kernel<SET>;                                                        \
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    kern<<<grid, kPNW * 32, smem, st>>>(a);                                                             \
  } while (0)
    if (set_kind == PDEIP_SET_KFP_0T) LAUNCH_Q(PDEIP_SET_KFP_0T);
    else if (set_kind == PDEIP_SET_KFP_BOUNDARY) LAUNCH_Q(PDEIP_SET_KFP_BOUNDARY);
    else if (set_kind == PDEIP_SET_KMV_PAIRS) LAUNCH_Q(PDEIP_SET_KMV_PAIRS);
    else PDEIP_REQUIRE(false, PDEIP_ERR_UNSUPPORTED, "quadratic parametric model: unsupported point set %d", set_kind);
#undef LAUNCH_Q
  } else {
    PDEIP_REQUIRE(false, PDEIP_ERR_INVALID_ARG, "unknown model kind %d", model_kind);
  }
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

int param_eval(int model_kind, const float* params, int d, int n_gaussian, const float* x, const float* v,
               float* out_value, float* out_grad, float* out_vHv, float* out_lap, int64_t n, cudaStream_t st) {
  PDEIP_REQUIRE(d >= 1 && d <= kPD, PDEIP_ERR_UNSUPPORTED, "parametric model_eval supports 1 <= d <= %d", kPD);
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (model_kind == PDEIP_MODEL_GMM) {
    PDEIP_REQUIRE(n_gaussian >= 1 && n_gaussian <= kPK, PDEIP_ERR_UNSUPPORTED, "1 <= n_gaussian <= %d supported", kPK);
    gmm_param_eval_kernel<<<grid, 128, sizeof(float) * n_gaussian * d, st>>>(params, n_gaussian, d, x, v, out_value,
                                                                            out_grad, out_vHv, out_lap, n);
  } else if (model_kind == PDEIP_MODEL_QUADRATIC) {
    quad_param_eval_kernel<<<grid, 128, sizeof(float) * (d * d + d), st>>>(params, d, x, v, out_value, out_grad,
                                                                          out_vHv, out_lap, n);
  } else {
    PDEIP_REQUIRE(false, PDEIP_ERR_INVALID_ARG, "unknown model kind %d", model_kind);
  }
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

}  // namespace pdeip
