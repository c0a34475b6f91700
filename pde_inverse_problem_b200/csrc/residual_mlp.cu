// residual_mlp.cu — K3/K4 (fp32 parity path): self-consistency residual of the potential MLP,
// loss terms and parameter gradient in one persistent kernel.
//
// Replaces value_and_grad_fn of methods/consistency_instances/kinetic_fokker_planck.py:11-69 and
// methods/consistency_instances/fokker_planck.py:33-63 (jax.value_and_grad over jvp-of-grad /
// jacfwd-of-grad).  Per point set:
//   KFP 0T        l = |g|^2 - 2 D_v^2 V + 2 gamma D_v V        (kinetic_fokker_planck.py:40-45)
//   KFP boundary  l = coef D_v V, coef = +-2/T                 (:34-39,48-50)
//   FP 0T         l = |g|^2 - 2 sum_i D_{e_i}^2 V              (fokker_planck.py:50-51)
//   FP boundary   l = coef V                                   (:48-49,53)
// plus sum|g_true|^2 (constant in theta) and "loss ground truth" sum|g_true - g|^2 on the 0T sets.
#include "mlp_thread.cuh"
#include "residual_common.cuh"

namespace pdeip {

template <int H>
__device__ __forceinline__ int flat_to_acc(const MlpShape<H>& sh, int idx) {
  constexpr int RS = MlpShape<H>::RS;
  for (int l = 0; l <= sh.LH; ++l) {
    const int w0 = sh.w_off(l), no = sh.n_out(l), nw = sh.n_in(l) * no;
    if (idx < w0 + nw) {
      const int rel = idx - w0;
      const int i = rel / no, j = rel - i * no;
      return sh.acc_w_off(l) + j * RS + i;
    }
    if (idx < w0 + nw + no) return sh.acc_b_off(l) + (idx - w0 - nw);
  }
  return 0;
}

template <int H, int LHMAX, int SET, int NW>
__global__ void __launch_bounds__(NW * 32, 1) mlp_residual_kernel(const ResidualArgs a) {
  extern __shared__ __align__(16) float smem[];
  const MlpShape<H> sh{a.d, a.layers};
  const int d = a.d;
  const int P = sh.num_params();
  const int P4 = (P + 3) & ~3;
  float* sp = smem;
  float* tp = sp + P4;
  int ntg = 0;
  if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = d * d;
  else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * d;
  float* warp_base = tp + ((ntg + 3) & ~3);
  const int acc_sz = (sh.acc_size() + 3) & ~3;
  constexpr int tileA_sz = 32 * (H + 1);
  constexpr int tileZ_sz = 32 * kTileZStride;
  const int per_warp = acc_sz + tileA_sz + tileZ_sz;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpScratch<H> ws;
  ws.acc = warp_base + warp * per_warp;
  ws.tileA = ws.acc + acc_sz;
  ws.tileZ = ws.tileA + tileA_sz;

  for (int i = threadIdx.x; i < P; i += blockDim.x) sp[i] = a.params[i];
  for (int i = threadIdx.x; i < ntg; i += blockDim.x) tp[i] = a.tg.params[i];
  for (int i = lane; i < acc_sz; i += 32) ws.acc[i] = 0.f;
  __syncthreads();

  float sums[PDEIP_NUM_SUMS];
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;

  PointState<H, LHMAX> st;
  MlpThread<H, LHMAX> net(sh, sp, st);
  constexpr bool kinetic = (SET == PDEIP_SET_KFP_0T || SET == PDEIP_SET_KFP_BOUNDARY);
  // row width of the point set; PDEIP_DRIFT_IN_POINTS appends grad V_true (d floats) to every point
  const int dim = (kinetic ? 2 * d : d) + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? d : 0);
  const int gt_off = kinetic ? 2 * d : d;
  const int64_t tile_pts = NW * 32;
  const int64_t n_tiles = (a.n_points + tile_pts - 1) / tile_pts;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * tile_pts + threadIdx.x;
    const bool valid = p < a.n_points;
    const float wt = valid ? a.weight : 0.f;
    float v[kDMax];
    float kmv_u[kDMax];
    float kmv_kappa = 0.f;
    if (SET == PDEIP_SET_KMV_PAIRS) {
      // pair index p -> (i, j, t): Delta = x[j,t] - x[i,t], v = v[j,t], u = G[j,t], kappa = 2 c[j,t]  (:23,:46-48)
      const int64_t n = a.kmv_n;
      const int nt = a.kmv_nt;
      const int64_t pp = valid ? p : 0;
      const int t = (int)(pp % nt);
      const int64_t ij = pp / nt;
      const int64_t j = ij % n, i = ij / n;
      const float* rj = a.points + (j * nt + t) * 2 * d;
      const float* ri = (a.ref ? a.ref : a.points) + (i * nt + t) * 2 * d;
      for (int c = 0; c < d; ++c) {
        st.x[c] = valid ? rj[c] - ri[c] : 0.f;
        v[c] = valid ? rj[d + c] : 0.f;
        kmv_u[c] = (valid && a.G) ? a.G[(j * nt + t) * d + c] : 0.f;
      }
      kmv_kappa = valid ? 2.f * a.c[j * nt + t] * wt : 0.f;
    } else {
      for (int i = 0; i < d; ++i) {
        st.x[i] = valid ? a.points[elem_index(a.layout, p, i, a.n_points, dim)] : 0.f;
        if (kinetic) v[i] = valid ? a.points[elem_index(a.layout, p, d + i, a.n_points, dim)] : 0.f;
      }
    }
    net.reset_adjoints();
    net.primal_forward();
    float u1[kOut], u2[kOut];

    if (SET == PDEIP_SET_KFP_0T || SET == PDEIP_SET_FP_0T) {
      float g[kDMax], gt[kDMax];
      net.input_gradient(g);
      if (a.tg.kind == PDEIP_DRIFT_IN_POINTS) {
        for (int i = 0; i < d; ++i)
          gt[i] = valid ? a.points[elem_index(a.layout, p, gt_off + i, a.n_points, dim)] : 0.f;
      } else {
        true_grad_thread(a.tg, tp, d, st.x, gt);
      }
      float g2 = 0.f, gt2 = 0.f, gd2 = 0.f;
      for (int i = 0; i < d; ++i) {
        g2 = fmaf(g[i], g[i], g2);
        gt2 = fmaf(gt[i], gt[i], gt2);
        const float df = gt[i] - g[i];
        gd2 = fmaf(df, df, gd2);
      }
      float D1 = 0.f, D2 = 0.f, lval;
      if (SET == PDEIP_SET_KFP_0T) {
        const float gamma = a.coef;
        net.template direction_forward<true>(v, u1, u2, D1, D2);
        net.template direction_reverse<true>(ws, v, -2.f * wt, 2.f * gamma * wt, u1, u2);
        lval = g2 - 2.f * D2 + 2.f * gamma * D1 + gt2;
      } else {
        float wdir[kDMax];
        for (int i = 0; i < d; ++i) wdir[i] = 0.f;
        float lap = 0.f;
        for (int k = 0; k < d; ++k) {  // exact Laplacian: d tangent streams e_k (fokker_planck.py:35-37)
          wdir[k] = 1.f;
          float d1k, d2k;
          net.template direction_forward<true>(wdir, u1, u2, d1k, d2k);
          net.template direction_reverse<true>(ws, wdir, -2.f * wt, 0.f, u1, u2);
          lap += d2k;
          wdir[k] = 0.f;
        }
        D2 = lap;
        lval = g2 - 2.f * lap + gt2;
      }
      // |g|^2 through the stop-gradient direction w = g (beta = 2)
      float d1g, d2g;
      net.template direction_forward<false>(g, u1, u2, d1g, d2g);
      net.template direction_reverse<false>(ws, g, 0.f, 2.f * wt, u1, u2);
      net.primal_reverse(ws, 0.f);
      sums[PDEIP_SUM_G2] += wt * g2;
      sums[PDEIP_SUM_D2] += wt * D2;
      sums[PDEIP_SUM_D1] += wt * D1;
      sums[PDEIP_SUM_GTRUE2] += wt * gt2;
      sums[PDEIP_SUM_GT] += wt * gd2;
      sums[PDEIP_SUM_LOSS] += wt * lval;
    } else if (SET == PDEIP_SET_KFP_BOUNDARY) {
      float D1, D2;
      net.template direction_forward<false>(v, u1, u2, D1, D2);
      net.template direction_reverse<false>(ws, v, 0.f, a.coef * wt, u1, u2);
      net.primal_reverse(ws, 0.f);
      sums[PDEIP_SUM_BOUNDARY] += wt * a.coef * D1;
      sums[PDEIP_SUM_LOSS] += wt * a.coef * D1;
    } else if (SET == PDEIP_SET_KMV_PAIRS) {
      const float V = net.value();
      float D1, D2;
      net.template direction_forward<true>(v, u1, u2, D1, D2);
      net.template direction_reverse<true>(ws, v, -2.f * wt, 0.f, u1, u2);
      if (a.G) {  // d|G_j|^2 through the constant direction u = G_j
        float d1g, d2g;
        net.template direction_forward<false>(kmv_u, u1, u2, d1g, d2g);
        net.template direction_reverse<false>(ws, kmv_u, 0.f, 2.f * wt, u1, u2);
      }
      net.primal_reverse(ws, kmv_kappa);
      sums[PDEIP_SUM_D2] += wt * D2;
      sums[PDEIP_SUM_D1] += 0.5f * kmv_kappa * V;
      sums[PDEIP_SUM_LOSS] += -2.f * wt * D2 + kmv_kappa * V;
    } else {  // FP boundary
      const float V = net.value();
      net.primal_reverse(ws, a.coef * wt);
      sums[PDEIP_SUM_BOUNDARY] += wt * a.coef * V;
      sums[PDEIP_SUM_LOSS] += wt * a.coef * V;
    }
  }

  // ---- CTA reduction into this CTA's partial vector ------------------------------------------
  __syncthreads();
  float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
  for (int idx = threadIdx.x; idx < P; idx += blockDim.x) {
    const int ai = flat_to_acc<H>(sh, idx);
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += warp_base[w * per_warp + ai];
    part[idx] += s;
  }
  __syncthreads();  // everyone is done with the accumulators; reuse warp 0's tileZ for the sums
  float* red = warp_base + acc_sz + tileA_sz;  // warp 0 tileZ
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
    const float s = warp_sum(sums[k]);
    if (lane == 0) red[warp * PDEIP_NUM_SUMS + k] = s;
  }
  __syncthreads();
  if (threadIdx.x < PDEIP_NUM_SUMS) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red[w * PDEIP_NUM_SUMS + threadIdx.x];
    part[P + threadIdx.x] += s;
  }
}

template <int H, int LHMAX, int NW>
static int launch_mlp_residual(int set_kind, const ResidualArgs& a, cudaStream_t st) {
  const MlpShape<H> sh{a.d, a.layers};
  const int P4 = (sh.num_params() + 3) & ~3;
  const int ntg4 = (true_grad_floats(a.tg, a.d) + 3) & ~3;
  const int acc_sz = (sh.acc_size() + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)P4 + ntg4 + (size_t)NW * (acc_sz + 32 * (H + 1) + 32 * kTileZStride));
  PDEIP_REQUIRE(smem <= 227 * 1024, PDEIP_ERR_UNSUPPORTED, "residual kernel needs %zu B of shared memory", smem);
  const int grid = residual_grid();
#define LAUNCH_SET(SET)                                                                              \
  do {                                                                                               \
    auto kern = mlp_residual_kernel<H, LHMAX, SET, NW>;                                              \
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, NW * 32, smem, st>>>(a);                                                            \
  } while (0)
  switch (set_kind) {
    case PDEIP_SET_KFP_0T: LAUNCH_SET(PDEIP_SET_KFP_0T); break;
    case PDEIP_SET_KFP_BOUNDARY: LAUNCH_SET(PDEIP_SET_KFP_BOUNDARY); break;
    case PDEIP_SET_FP_0T: LAUNCH_SET(PDEIP_SET_FP_0T); break;
    case PDEIP_SET_FP_BOUNDARY: LAUNCH_SET(PDEIP_SET_FP_BOUNDARY); break;
    case PDEIP_SET_KMV_PAIRS: LAUNCH_SET(PDEIP_SET_KMV_PAIRS); break;
    default: PDEIP_REQUIRE(false, PDEIP_ERR_INVALID_ARG, "unknown point-set kind %d", set_kind);
  }
#undef LAUNCH_SET
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

int mlp_residual_accumulate_fp32(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st) {
  PDEIP_REQUIRE(hidden == 32, PDEIP_ERR_UNSUPPORTED,
                "fp32 MLP residual is built for hidden_dim == 32 (pad smaller widths with zeros); got %d", hidden);
  PDEIP_REQUIRE(a.layers >= 1 && a.layers <= 8, PDEIP_ERR_UNSUPPORTED, "1 <= layers <= 8 supported (got %d)",
                a.layers);
  PDEIP_REQUIRE(a.d >= 1 && a.d <= kDMax, PDEIP_ERR_UNSUPPORTED, "1 <= d <= %d supported (got %d)", kDMax, a.d);
  if (a.layers <= 2) return launch_mlp_residual<32, 2, 8>(set_kind, a, st);
  if (a.layers <= 4) return launch_mlp_residual<32, 4, 4>(set_kind, a, st);
  return launch_mlp_residual<32, 8, 4>(set_kind, a, st);  // configurations/neural_network/MLP.yaml: layers = 8
}

// ------------------------------------------------------------------------------------------------
// model evaluation (value / gradient / v'Hv / Laplacian), one thread per point, no reverse pass
// ------------------------------------------------------------------------------------------------
template <int H, int LHMAX>
__global__ void __launch_bounds__(128) mlp_eval_kernel(const float* __restrict__ params, int d, int layers,
                                                       const float* __restrict__ x, const float* __restrict__ v,
                                                       float* out_value, float* out_grad, float* out_vHv,
                                                       float* out_lap, int64_t n) {
  extern __shared__ __align__(16) float smem[];
  const MlpShape<H> sh{d, layers};
  const int P = sh.num_params();
  for (int i = threadIdx.x; i < P; i += blockDim.x) smem[i] = params[i];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  PointState<H, LHMAX> st;
  MlpThread<H, LHMAX> net(sh, smem, st);
  for (int i = 0; i < d; ++i) st.x[i] = x[p * d + i];
  net.primal_forward();
  if (out_value) out_value[p] = net.value();
  if (out_grad) {
    float g[kDMax];
    net.input_gradient(g);
    for (int i = 0; i < d; ++i) out_grad[p * d + i] = g[i];
  }
  float u1[kOut], u2[kOut];
  if (out_vHv && v) {
    float w[kDMax], D1, D2;
    for (int i = 0; i < d; ++i) w[i] = v[p * d + i];
    net.template direction_forward<true>(w, u1, u2, D1, D2);
    out_vHv[p] = D2;
  }
  if (out_lap) {
    float w[kDMax], lap = 0.f;
    for (int i = 0; i < d; ++i) w[i] = 0.f;
    for (int k = 0; k < d; ++k) {
      w[k] = 1.f;
      float D1, D2;
      net.template direction_forward<true>(w, u1, u2, D1, D2);
      lap += D2;
      w[k] = 0.f;
    }
    out_lap[p] = lap;
  }
}

int mlp_eval_fp32(const float* params, int d, int hidden, int layers, const float* x, const float* v,
                  float* out_value, float* out_grad, float* out_vHv, float* out_lap, int64_t n,
                  cudaStream_t st) {
  PDEIP_REQUIRE(hidden == 32, PDEIP_ERR_UNSUPPORTED, "model_eval is built for hidden_dim == 32 (got %d)", hidden);
  PDEIP_REQUIRE(layers >= 1 && layers <= 8, PDEIP_ERR_UNSUPPORTED, "1 <= layers <= 8 supported (got %d)", layers);
  PDEIP_REQUIRE(d >= 1 && d <= kDMax, PDEIP_ERR_UNSUPPORTED, "1 <= d <= %d supported (got %d)", kDMax, d);
  const MlpShape<32> sh{d, layers};
  const size_t smem = sizeof(float) * (size_t)sh.num_params();
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (layers <= 2) {
    mlp_eval_kernel<32, 2><<<grid, 128, smem, st>>>(params, d, layers, x, v, out_value, out_grad, out_vHv, out_lap, n);
  } else if (layers <= 4) {
    mlp_eval_kernel<32, 4><<<grid, 128, smem, st>>>(params, d, layers, x, v, out_value, out_grad, out_vHv, out_lap, n);
  } else {
    mlp_eval_kernel<32, 8><<<grid, 128, smem, st>>>(params, d, layers, x, v, out_value, out_grad, out_vHv, out_lap, n);
  }
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

}  // namespace pdeip
