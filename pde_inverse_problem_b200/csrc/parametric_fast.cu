// parametric_fast.cu — K5, production kernels: closed-form residuals of the parametric GMM and quadratic models for
// d in {4, 8, 16, 32}.
//
// Replaces jax.value_and_grad of the KFP loss (methods/consistency_instances/kinetic_fokker_planck.py:33-61) for the
// model of example_problems/kinetic_fokker_planck_example_GMM.py:214-234 — the model the reference's launch scripts
// train by default (configurations/config.yaml:40, scripts/run_KGMM.sh).  Same closed forms as parametric.cu
// (checked against autodiff in oracle/taylor.py); that file stays the generic path (any d <= 32).
//
// One thread = one point, everything that is indexed by a coordinate lives in registers (compile-time d, packed
// fp32x2 arithmetic), the centres sit in shared memory (rows padded to a multiple of the tile with far-away centres,
// whose weight is exactly 0).  Two passes over the centres instead of per-thread arrays of length K:
//   pass A  online softmax in log2 units over tiles of TK centres (branch-free, one rescale per tile, MUFU ex2) with the
//           running sums  se = sum e_k,  sum e_k mu_k,  sum e_k c_k,  sum e_k c_k^2   (c_k = r_k . v, r_k = y - mu_k)
//           -> g = E[r], D_v V = E[c], D_v^2 V = |v|^2 - Var_w(c);   E[r_k . g] = |g|^2 needs no pass of its own
//   pass B  per centre: r_k, w_k, c_k, r_k . g recomputed, the three coefficients of SURVEY §9.5, and the point's
//           contribution  t = sg_k g + sv_k v + sr_k r_k  to d l / d mu_k  (r_k instead of y: no separate sum of sr_k).
// The sum of t over the 32 points of a warp is a TRANSPOSE-REDUCE by shuffles (DP - 1 + log2(32 / DP) exchanges
// instead of 5 DP for DP butterflies): afterwards lane L holds the warp total of component L / (32 / DP), and one lane
// per component adds it to the warp's accumulator [K][DP] in shared memory (conflict-free, no atomics).  Per-CTA
// partials go to the workspace of the residual protocol (residual_common.cuh): bit-reproducible.
#include <stdlib.h>

#include "common.cuh"
#include "residual_common.cuh"

namespace pdeip {

namespace pfast {

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// t[0..N) per lane -> returns in t[0] the sum over the 32 lanes of component (lane / (32 / N))
template <int N>
__device__ __forceinline__ float transpose_reduce(float (&t)[N], int lane) {
  int o = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? t[i] : t[i + n / 2];
      const float keep = upper ? t[i + n / 2] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  for (; o >= 1; o >>= 1) t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
  return t[0];
}

template <int DP>
struct Tile {
  static constexpr int TK = DP >= 32 ? 4 : 8;            // centres per softmax tile (register pressure at d = 32)
  static constexpr int THREADS = DP >= 16 ? 256 : 512;    // one CTA per SM (one workspace partial per CTA); 128-register cap at 512
};

template <int DP, int SET>
__global__ void __launch_bounds__(Tile<DP>::THREADS, 1) gmm_param_fast_kernel(const ResidualArgs a, int K, int k_pad) {
  static_assert(DP % 4 == 0 && DP <= 32, "d in {4, 8, 16, 32}");
  constexpr int TK = Tile<DP>::TK, NT = Tile<DP>::THREADS, NW = NT / 32;
  extern __shared__ __align__(16) float smem[];
  float* mus_s = smem;                                    // [k_pad][DP]
  float* tp = mus_s + k_pad * DP;                         // true-gradient parameters (LINEAR / GMM kinds)
  int ntg = 0;
  if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = DP * DP;
  else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * DP;
  float* acc_base = tp + ((ntg + 3) & ~3);                // [NW][k_pad][DP] gradient accumulators, then [NW][8] sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* acc_s = acc_base + warp * (k_pad * DP);
  for (int idx = tid; idx < k_pad * DP; idx += NT) mus_s[idx] = (idx / DP) < K ? a.params[idx] : 1.0e18f;
  for (int i = tid; i < ntg; i += NT) tp[i] = a.tg.params[i];
  for (int idx = tid; idx < NW * k_pad * DP; idx += NT) acc_base[idx] = 0.f;
  __syncthreads();

  const int dimw = 2 * DP + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? DP : 0);
  const int64_t cstride = comp_stride(a.layout, a.n_points);
  constexpr float c2 = -0.5f * 1.4426950408889634f;       // logits in log2 units (sigma = 1: GMM.py:232)
  const float2 neg1 = make_float2(-1.f, -1.f);
  constexpr int SH = 32 / DP;                             // lanes per component after the transpose-reduce
  const int comp = lane / SH;
  const bool writer = (lane % SH) == 0;

  float sums[PDEIP_NUM_SUMS];
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;

  const int64_t n_tiles = (a.n_points + NT - 1) / NT;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * NT + tid;
    const bool valid = p < a.n_points;
    const float wt = valid ? a.weight : 0.f;
    const float* pb = a.points + point_base(a.layout, valid ? p : 0, dimw);
    float2 y[DP / 2], v[DP / 2];
#pragma unroll
    for (int i = 0; i < DP / 2; ++i) {
      y[i] = make_float2(__ldg(pb + (2 * i) * cstride), __ldg(pb + (2 * i + 1) * cstride));
      v[i] = make_float2(__ldg(pb + (DP + 2 * i) * cstride), __ldg(pb + (DP + 2 * i + 1) * cstride));
    }
    // ---- pass A: online softmax with the moments of c ----------------------------------------------------------
    float m = -INFINITY, se = 0.f, sc = 0.f, sc2 = 0.f;
    float2 am[DP / 2];
#pragma unroll
    for (int i = 0; i < DP / 2; ++i) am[i] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int k0 = 0; k0 < k_pad; k0 += TK) {
      float al[TK], cv[TK];
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        const float4* mu4 = reinterpret_cast<const float4*>(mus_s + (k0 + kk) * DP);
        float2 s2 = make_float2(0.f, 0.f), cc = make_float2(0.f, 0.f);
#pragma unroll
        for (int i4 = 0; i4 < DP / 4; ++i4) {
          const float4 t = mu4[i4];
          const float2 r0 = __ffma2_rn(make_float2(t.x, t.y), neg1, y[2 * i4]);
          const float2 r1 = __ffma2_rn(make_float2(t.z, t.w), neg1, y[2 * i4 + 1]);
          s2 = __ffma2_rn(r0, r0, s2);
          s2 = __ffma2_rn(r1, r1, s2);
          cc = __ffma2_rn(r0, v[2 * i4], cc);
          cc = __ffma2_rn(r1, v[2 * i4 + 1], cc);
        }
        al[kk] = c2 * (s2.x + s2.y);
        cv[kk] = cc.x + cc.y;
      }
      float tm = al[0];
#pragma unroll
      for (int kk = 1; kk < TK; ++kk) tm = fmaxf(tm, al[kk]);
      const float mn = fmaxf(m, tm);
      const float scl = ex2(m - mn);  // first tile: ex2(-inf) = 0 and the sums are 0 anyway
      m = mn;
      se *= scl; sc *= scl; sc2 *= scl;
      const float2 scl2 = make_float2(scl, scl);
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) am[i] = __fmul2_rn(am[i], scl2);
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        const float e = ex2(al[kk] - m);  // padding centres: al = -1e36 -> 0
        const float ec = e * cv[kk];
        se += e;
        sc += ec;
        sc2 = fmaf(ec, cv[kk], sc2);
        const float2 e2 = make_float2(e, e);
        const float4* mu4 = reinterpret_cast<const float4*>(mus_s + (k0 + kk) * DP);
#pragma unroll
        for (int i4 = 0; i4 < DP / 4; ++i4) {
          const float4 t = mu4[i4];
          am[2 * i4] = __ffma2_rn(e2, make_float2(t.x, t.y), am[2 * i4]);
          am[2 * i4 + 1] = __ffma2_rn(e2, make_float2(t.z, t.w), am[2 * i4 + 1]);
        }
      }
    }
    const float inv = 1.f / se;
    const float Ec = sc * inv, Ec2 = sc2 * inv;
    float2 g[DP / 2];
    float2 g2v = make_float2(0.f, 0.f), v2v = make_float2(0.f, 0.f);
    const float2 ninv2 = make_float2(-inv, -inv);
#pragma unroll
    for (int i = 0; i < DP / 2; ++i) {
      g[i] = __ffma2_rn(am[i], ninv2, y[i]);  // g = E[r] = y - sum w_k mu_k
      g2v = __ffma2_rn(g[i], g[i], g2v);
      v2v = __ffma2_rn(v[i], v[i], v2v);
    }
    const float g2 = g2v.x + g2v.y;           // also E[r_k . g]
    const float D1 = Ec, D2 = (v2v.x + v2v.y) - (Ec2 - Ec * Ec);

    float alpha, beta, cg;
    if (SET == PDEIP_SET_KFP_0T) {
      const float gamma = a.coef;
      alpha = -2.f * wt; beta = 2.f * gamma * wt; cg = wt;
      float gt2 = 0.f, gd2 = 0.f;
      if (a.tg.kind == PDEIP_DRIFT_IN_POINTS) {
#pragma unroll
        for (int i = 0; i < DP / 2; ++i) {
          const float ga = __ldg(pb + (2 * DP + 2 * i) * cstride), gb = __ldg(pb + (2 * DP + 2 * i + 1) * cstride);
          gt2 = fmaf(ga, ga, fmaf(gb, gb, gt2));
          const float da = ga - g[i].x, db = gb - g[i].y;
          gd2 = fmaf(da, da, fmaf(db, db, gd2));
        }
      } else {
        float yy[DP], gt[DP];
#pragma unroll
        for (int i = 0; i < DP / 2; ++i) { yy[2 * i] = y[i].x; yy[2 * i + 1] = y[i].y; }
        true_grad_thread(a.tg, tp, DP, yy, gt);
#pragma unroll
        for (int i = 0; i < DP / 2; ++i) {
          gt2 = fmaf(gt[2 * i], gt[2 * i], fmaf(gt[2 * i + 1], gt[2 * i + 1], gt2));
          const float da = gt[2 * i] - g[i].x, db = gt[2 * i + 1] - g[i].y;
          gd2 = fmaf(da, da, fmaf(db, db, gd2));
        }
      }
      sums[PDEIP_SUM_G2] += wt * g2;
      sums[PDEIP_SUM_D2] += wt * D2;
      sums[PDEIP_SUM_D1] += wt * D1;
      sums[PDEIP_SUM_GTRUE2] += wt * gt2;
      sums[PDEIP_SUM_GT] += wt * gd2;
      sums[PDEIP_SUM_LOSS] += wt * (g2 - 2.f * D2 + 2.f * gamma * D1 + gt2);
    } else {
      alpha = 0.f; beta = a.coef * wt; cg = 0.f;
      sums[PDEIP_SUM_BOUNDARY] += wt * a.coef * D1;
      sums[PDEIP_SUM_LOSS] += wt * a.coef * D1;
    }
    // ---- pass B: d l / d mu_k, summed over the warp's 32 points by a transpose-reduce ------------------------------
    const float beta_g = 2.f * cg;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      const float4* mu4 = reinterpret_cast<const float4*>(mus_s + k * DP);
      float2 r[DP / 2];
      float2 s2 = make_float2(0.f, 0.f), cc = make_float2(0.f, 0.f), cgv = make_float2(0.f, 0.f);
#pragma unroll
      for (int i4 = 0; i4 < DP / 4; ++i4) {
        const float4 t = mu4[i4];
        r[2 * i4] = __ffma2_rn(make_float2(t.x, t.y), neg1, y[2 * i4]);
        r[2 * i4 + 1] = __ffma2_rn(make_float2(t.z, t.w), neg1, y[2 * i4 + 1]);
        s2 = __ffma2_rn(r[2 * i4], r[2 * i4], s2);
        s2 = __ffma2_rn(r[2 * i4 + 1], r[2 * i4 + 1], s2);
        cc = __ffma2_rn(r[2 * i4], v[2 * i4], cc);
        cc = __ffma2_rn(r[2 * i4 + 1], v[2 * i4 + 1], cc);
        cgv = __ffma2_rn(r[2 * i4], g[2 * i4], cgv);
        cgv = __ffma2_rn(r[2 * i4 + 1], g[2 * i4 + 1], cgv);
      }
      const float w = ex2(c2 * (s2.x + s2.y) - m) * inv;
      const float c = cc.x + cc.y, cgk = cgv.x + cgv.y;
      const float sg = -beta_g * w;
      const float sv = -beta * w + alpha * (2.f * w * c - 2.f * Ec * w);
      const float sr = beta_g * w * (cgk - g2) + beta * w * (c - Ec) + alpha * (-w * (c * c - Ec2) + 2.f * Ec * w * (c - Ec));
      const float2 sg2 = make_float2(sg, sg), sv2 = make_float2(sv, sv), sr2 = make_float2(sr, sr);
      float t[DP];
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        const float2 ti = __ffma2_rn(sg2, g[i], __ffma2_rn(sv2, v[i], __fmul2_rn(sr2, r[i])));
        t[2 * i] = ti.x;
        t[2 * i + 1] = ti.y;
      }
      const float tot = transpose_reduce<DP>(t, lane);
      if (writer) acc_s[k * DP + comp] += tot;
    }
  }
  // ---- CTA reduction into this CTA's partial (fixed order) -----------------------------------------------------------
  __syncthreads();
  float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
  const int P = K * DP;
  for (int idx = tid; idx < P; idx += NT) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += acc_base[w * (k_pad * DP) + idx];
    part[idx] += s;
  }
  __syncthreads();
  float* red = acc_base;  // accumulators consumed: reuse
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
    const float s = warp_sum(sums[k]);
    if (lane == 0) red[warp * PDEIP_NUM_SUMS + k] = s;
  }
  __syncthreads();
  if (tid < PDEIP_NUM_SUMS) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red[w * PDEIP_NUM_SUMS + tid];
    part[P + tid] += s;
  }
}

template <int DP>
static int launch(int set_kind, const ResidualArgs& a, int K, cudaStream_t st) {
  constexpr int TK = Tile<DP>::TK, NT = Tile<DP>::THREADS, NW = NT / 32;
  const int k_pad = (K + TK - 1) / TK * TK;
  const int ntg = true_grad_floats(a.tg, a.d);
  const size_t smem = sizeof(float) * ((size_t)k_pad * DP + ((ntg + 3) & ~3) + (size_t)NW * k_pad * DP + NW * PDEIP_NUM_SUMS);
  PDEIP_REQUIRE(smem <= 227 * 1024, PDEIP_ERR_UNSUPPORTED, "parametric GMM kernel needs %zu B of shared memory", smem);
  const int grid = residual_grid();
  if (set_kind == PDEIP_SET_KFP_0T) {
    auto kern = gmm_param_fast_kernel<DP, PDEIP_SET_KFP_0T>;
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(a, K, k_pad);
  } else {
    auto kern = gmm_param_fast_kernel<DP, PDEIP_SET_KFP_BOUNDARY>;
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(a, K, k_pad);
  }
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Quadratic parametric model  V(y) = y.(y W + b)  (example_problems/kinetic_fokker_planck_example_OU.py:209-220,
// kinetic_mckean_vlasov_example_quadratic.py:205-216): g = (W + W^T) y + b, D_v V = g.v, D_v^2 V = v^T (W + W^T) v.
//   dW[i][j] += a1_i y_j + a2_i g_j + a3_i v_j,  db_i += a1_i,   a1 = 2 c_g g + beta v,  a2 = 2 c_g y,  a3 = 2 alpha v + beta y
// (same coefficients as parametric.cu).  One thread = one point, Ws = W + W^T in shared memory, the
// d x d gradient is reduced over the warp ROW BY ROW with the same transpose-reduce as the GMM kernel.
// ------------------------------------------------------------------------------------------------------------
template <int DP, int SET>
__global__ void __launch_bounds__(Tile<DP>::THREADS, 1) quad_param_fast_kernel(const ResidualArgs a) {
  constexpr int NT = Tile<DP>::THREADS, NW = NT / 32;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                  // [DP][DP] = W + W^T
  float* bs = Ws + DP * DP;          // [DP]
  float* tp = bs + DP;               // true-gradient parameters (LINEAR / GMM kinds)
  int ntg = 0;
  if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = DP * DP;
  else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * DP;
  float* acc_base = tp + ((ntg + 3) & ~3);  // [NW][DP * DP + DP]
  constexpr int ACC = DP * DP + DP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* acc_s = acc_base + warp * ACC;
  for (int idx = tid; idx < DP * DP; idx += NT) {
    const int i = idx / DP, j = idx - i * DP;
    Ws[idx] = a.params[i * DP + j] + a.params[j * DP + i];
  }
  for (int i = tid; i < DP; i += NT) bs[i] = a.params[DP * DP + i];
  for (int i = tid; i < ntg; i += NT) tp[i] = a.tg.params[i];
  for (int idx = tid; idx < NW * ACC; idx += NT) acc_base[idx] = 0.f;
  __syncthreads();

  const int dimw = 2 * DP + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? DP : 0);
  const int64_t cstride = comp_stride(a.layout, a.n_points);
  constexpr int SH = 32 / DP;
  const int comp = lane / SH;
  const bool writer = (lane % SH) == 0;
  float sums[PDEIP_NUM_SUMS];
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;

  const int64_t n_tiles = (a.n_points + NT - 1) / NT;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * NT + tid;
    const bool valid = p < a.n_points;
    const float wt = valid ? a.weight : 0.f;
    const float* pb = a.points + point_base(a.layout, valid ? p : 0, dimw);
    float y[DP], v[DP], g[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      y[i] = __ldg(pb + i * cstride);
      v[i] = __ldg(pb + (DP + i) * cstride);
    }
    float D1 = 0.f, D2 = 0.f, g2 = 0.f;
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      const float4* row = reinterpret_cast<const float4*>(Ws + i * DP);
      float s = bs[i], h = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < DP / 4; ++j4) {
        const float4 t = row[j4];
        s = fmaf(t.x, y[4 * j4], fmaf(t.y, y[4 * j4 + 1], fmaf(t.z, y[4 * j4 + 2], fmaf(t.w, y[4 * j4 + 3], s))));
        h = fmaf(t.x, v[4 * j4], fmaf(t.y, v[4 * j4 + 1], fmaf(t.z, v[4 * j4 + 2], fmaf(t.w, v[4 * j4 + 3], h))));
      }
      g[i] = s;
      D1 = fmaf(s, v[i], D1);
      D2 = fmaf(h, v[i], D2);
      g2 = fmaf(s, s, g2);
    }
    float alpha, beta, cg;
    if (SET == PDEIP_SET_KFP_0T) {
      const float gamma = a.coef;
      alpha = -2.f * wt; beta = 2.f * gamma * wt; cg = wt;
      float gt[DP];
      if (a.tg.kind == PDEIP_DRIFT_IN_POINTS) {
#pragma unroll
        for (int i = 0; i < DP; ++i) gt[i] = __ldg(pb + (2 * DP + i) * cstride);
      } else {
        true_grad_thread(a.tg, tp, DP, y, gt);
      }
      float gt2 = 0.f, gd2 = 0.f;
#pragma unroll
      for (int i = 0; i < DP; ++i) {
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
    } else {
      alpha = 0.f; beta = a.coef * wt; cg = 0.f;
      sums[PDEIP_SUM_BOUNDARY] += wt * a.coef * D1;
      sums[PDEIP_SUM_LOSS] += wt * a.coef * D1;
    }
    const float c2g = 2.f * cg, a2c = 2.f * alpha;
    float a1[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) a1[i] = fmaf(c2g, g[i], beta * v[i]);
#pragma unroll
    for (int i = 0; i < DP; ++i) {  // row i of dW: a1_i y + a2_i g + a3_i v
      const float a2i = c2g * y[i], a3i = fmaf(a2c, v[i], beta * y[i]);
      float t[DP];
#pragma unroll
      for (int j = 0; j < DP; ++j) t[j] = fmaf(a1[i], y[j], fmaf(a2i, g[j], a3i * v[j]));
      const float tot = transpose_reduce<DP>(t, lane);
      if (writer) acc_s[i * DP + comp] += tot;
    }
    {
      const float tot = transpose_reduce<DP>(a1, lane);  // db
      if (writer) acc_s[DP * DP + comp] += tot;
    }
  }
  __syncthreads();
  float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
  for (int idx = tid; idx < ACC; idx += NT) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += acc_base[w * ACC + idx];
    part[idx] += s;  // flat layout of the quadratic model: W [d][d] row-major, then b [d]
  }
  __syncthreads();
  float* red = acc_base;
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
    const float s = warp_sum(sums[k]);
    if (lane == 0) red[warp * PDEIP_NUM_SUMS + k] = s;
  }
  __syncthreads();
  if (tid < PDEIP_NUM_SUMS) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red[w * PDEIP_NUM_SUMS + tid];
    part[ACC + tid] += s;
  }
}

template <int DP>
static int launch_quad(int set_kind, const ResidualArgs& a, cudaStream_t st) {
  constexpr int NT = Tile<DP>::THREADS, NW = NT / 32;
  const int ntg = true_grad_floats(a.tg, a.d);
  const size_t smem = sizeof(float) * ((size_t)DP * DP + DP + ((ntg + 3) & ~3) + (size_t)NW * (DP * DP + DP) + NW * PDEIP_NUM_SUMS);
  PDEIP_REQUIRE(smem <= 227 * 1024, PDEIP_ERR_UNSUPPORTED, "parametric quadratic kernel needs %zu B of shared memory", smem);
  const int grid = residual_grid();
  if (set_kind == PDEIP_SET_KFP_0T) {
    auto kern = quad_param_fast_kernel<DP, PDEIP_SET_KFP_0T>;
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(a);
  } else {
    auto kern = quad_param_fast_kernel<DP, PDEIP_SET_KFP_BOUNDARY>;
    PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, st>>>(a);
  }
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

}  // namespace pfast

// true if (d, set) is served by the production kernel; the caller falls back to parametric.cu otherwise
bool gmm_param_fast_ok(int set_kind, const ResidualArgs& a) {
  return (a.d == 4 || a.d == 8 || a.d == 16 || a.d == 32) &&
         (set_kind == PDEIP_SET_KFP_0T || set_kind == PDEIP_SET_KFP_BOUNDARY) && getenv("PDEIP_NO_FAST_PARAMETRIC") == nullptr;
}

int gmm_param_fast_accumulate(int set_kind, const ResidualArgs& a, int K, cudaStream_t st) {
  switch (a.d) {
    case 4: return pfast::launch<4>(set_kind, a, K, st);
    case 8: return pfast::launch<8>(set_kind, a, K, st);
    case 16: return pfast::launch<16>(set_kind, a, K, st);
    default: return pfast::launch<32>(set_kind, a, K, st);
  }
}

}  // namespace pdeip

namespace pdeip {

bool quad_param_fast_ok(int set_kind, const ResidualArgs& a) { return gmm_param_fast_ok(set_kind, a); }

int quad_param_fast_accumulate(int set_kind, const ResidualArgs& a, cudaStream_t st) {
  switch (a.d) {
    case 4: return pfast::launch_quad<4>(set_kind, a, st);
    case 8: return pfast::launch_quad<8>(set_kind, a, st);
    case 16: return pfast::launch_quad<16>(set_kind, a, st);
    default: return pfast::launch_quad<32>(set_kind, a, st);
  }
}

}  // namespace pdeip
