// mlp_thread.cuh — per-thread (one point per thread) evaluation of the potential MLP, its forward
// Taylor streams and the reverse pass through them, in fp32 on the CUDA cores.
//
// Model: core/model.py:51-62 — Dense d -> H -> ... -> H -> 40, tanh, V = sum u^2.
// What jax.grad / jax.jvp(jax.grad) / jax.jacfwd(jax.grad) / jax.value_and_grad compute for the
// residuals (kinetic_fokker_planck.py:17-23,60-61; fokker_planck.py:35-37,58-59) is evaluated here in
// closed form (checked against autodiff in oracle/taylor.py):
//
//   primal        a_0 = x, z_l = a_l W_l + b_l, a_{l+1} = tanh z_l, u = z_L
//   input grad    za_L = 2u, aa_l = za_l W_l^T, za_{l-1} = aa_l (1 - a_l^2), grad V = aa_0
//   Taylor stream along w (orders 1,2):  a1_0 = w, a2_0 = 0, z1_l = a1_l W_l, z2_l = a2_l W_l,
//                 a1_{l+1} = s1 z1_l, a2_{l+1} = s1 z2_l + s2 z1_l^2, s1 = 1 - t^2, s2 = -2 t s1
//                 D_w V = 2 u.u1,  D_w^2 V = 2 (u1.u1 + u.u2)
//   reverse       for l(x) = alpha D_w^2 V + beta D_w V + kappa V  (one reverse pass through the
//                 forward streams only; |grad V|^2 enters through the stop-gradient direction w = grad V)
//
// Weights are read from shared memory (warp-uniform addresses -> broadcast); the per-point vectors that
// are indexed dynamically live in thread-local memory; the GEMV accumulators are registers.
// The parameter gradient dW_l = sum_points a_l^T zbar_l is a batch-reduced outer product: each warp
// transposes its 32 points through two shared-memory tiles and lane r accumulates row r of dW_l.
#pragma once

#include "common.cuh"

namespace pdeip {

constexpr int kOut = 40;   // core/model.py:43
constexpr int kDMax = 32;  // largest input dimension of the per-thread path
constexpr int kTileZStride = 44;

template <int H>
struct MlpShape {
  int d;
  int LH;  // number of hidden layers (cfg.neural_network.layers); Dense layers = LH + 1
  __host__ __device__ int n_in(int l) const { return l == 0 ? d : H; }
  __host__ __device__ int n_out(int l) const { return l == LH ? kOut : H; }
  __host__ __device__ int w_off(int l) const { return l == 0 ? 0 : d * H + H + (l - 1) * (H * H + H); }
  __host__ __device__ int b_off(int l) const { return w_off(l) + n_in(l) * n_out(l); }
  __host__ __device__ int num_params() const { return b_off(LH) + kOut; }
  // per-warp accumulator layout: weights [l][j][RS] then biases
  static constexpr int RS = H > 32 ? 64 : 32;
  __host__ __device__ int acc_w_off(int l) const { return l * H * RS; }
  __host__ __device__ int acc_b_off(int l) const { return (LH * H + kOut) * RS + l * H; }
  __host__ __device__ int acc_size() const { return (LH * H + kOut) * (RS + 1); }
};

template <int N>
__device__ __forceinline__ float dot_row(const float (&z)[N], const float* __restrict__ row) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j4 = 0; j4 < N / 4; ++j4) {
    const float4 w = reinterpret_cast<const float4*>(row)[j4];
    s0 = fmaf(z[4 * j4 + 0], w.x, s0);
    s1 = fmaf(z[4 * j4 + 1], w.y, s1);
    s2 = fmaf(z[4 * j4 + 2], w.z, s2);
    s3 = fmaf(z[4 * j4 + 3], w.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

template <int N>
__device__ __forceinline__ void dot_row2(const float (&za)[N], const float (&zb)[N],
                                         const float* __restrict__ row, float& oa, float& ob) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
  for (int j4 = 0; j4 < N / 4; ++j4) {
    const float4 w = reinterpret_cast<const float4*>(row)[j4];
    a0 = fmaf(za[4 * j4 + 0], w.x, a0);
    a1 = fmaf(za[4 * j4 + 1], w.y, a1);
    a2 = fmaf(za[4 * j4 + 2], w.z, a2);
    a3 = fmaf(za[4 * j4 + 3], w.w, a3);
    b0 = fmaf(zb[4 * j4 + 0], w.x, b0);
    b1 = fmaf(zb[4 * j4 + 1], w.y, b1);
    b2 = fmaf(zb[4 * j4 + 2], w.z, b2);
    b3 = fmaf(zb[4 * j4 + 3], w.w, b3);
  }
  oa = (a0 + a1) + (a2 + a3);
  ob = (b0 + b1) + (b2 + b3);
}

// acc[j] += sum_i in[i] * W[i][j]    (W: smem row-major [n_in][N])
template <int N>
__device__ __forceinline__ void gemv1(const float* __restrict__ W, int n_in, const float* in, float (&acc)[N]) {
#pragma unroll 2
  for (int i = 0; i < n_in; ++i) {
    const float a = in[i];
    const float4* row = reinterpret_cast<const float4*>(W + i * N);
#pragma unroll
    for (int j4 = 0; j4 < N / 4; ++j4) {
      const float4 w = row[j4];
      acc[4 * j4 + 0] = fmaf(a, w.x, acc[4 * j4 + 0]);
      acc[4 * j4 + 1] = fmaf(a, w.y, acc[4 * j4 + 1]);
      acc[4 * j4 + 2] = fmaf(a, w.z, acc[4 * j4 + 2]);
      acc[4 * j4 + 3] = fmaf(a, w.w, acc[4 * j4 + 3]);
    }
  }
}

template <int N>
__device__ __forceinline__ void gemv2(const float* __restrict__ W, int n_in, const float* ina, const float* inb,
                                      float (&acca)[N], float (&accb)[N]) {
#pragma unroll 2
  for (int i = 0; i < n_in; ++i) {
    const float a = ina[i], b = inb[i];
    const float4* row = reinterpret_cast<const float4*>(W + i * N);
#pragma unroll
    for (int j4 = 0; j4 < N / 4; ++j4) {
      const float4 w = row[j4];
      acca[4 * j4 + 0] = fmaf(a, w.x, acca[4 * j4 + 0]);
      acca[4 * j4 + 1] = fmaf(a, w.y, acca[4 * j4 + 1]);
      acca[4 * j4 + 2] = fmaf(a, w.z, acca[4 * j4 + 2]);
      acca[4 * j4 + 3] = fmaf(a, w.w, acca[4 * j4 + 3]);
      accb[4 * j4 + 0] = fmaf(b, w.x, accb[4 * j4 + 0]);
      accb[4 * j4 + 1] = fmaf(b, w.y, accb[4 * j4 + 1]);
      accb[4 * j4 + 2] = fmaf(b, w.z, accb[4 * j4 + 2]);
      accb[4 * j4 + 3] = fmaf(b, w.w, accb[4 * j4 + 3]);
    }
  }
}

// Per-warp scratch for the batch-reduced outer products.
template <int H>
struct WarpScratch {
  float* acc;    // [acc_size]  per-warp partial parameter gradient
  float* tileA;  // [32][H+1]   layer inputs of the warp's 32 points
  float* tileZ;  // [32][44]    layer-output adjoints of the warp's 32 points
};

// dW_l[r][:] += sum_{p in warp} a_p[r] * zb_p[:]   (and db_l += sum_p zb_p when WITH_BIAS)
template <int H, int N, bool WITH_BIAS, typename AFn>
__device__ __forceinline__ void outer_acc(const MlpShape<H>& sh, const WarpScratch<H>& ws, int l, int n_in,
                                          AFn a_fn, const float (&zb)[N]) {
  constexpr int RS = MlpShape<H>::RS;
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (int i = 0; i < n_in; ++i) ws.tileA[lane * (H + 1) + i] = a_fn(i);
#pragma unroll
  for (int j4 = 0; j4 < N / 4; ++j4)
    reinterpret_cast<float4*>(ws.tileZ + lane * kTileZStride)[j4] =
        make_float4(zb[4 * j4], zb[4 * j4 + 1], zb[4 * j4 + 2], zb[4 * j4 + 3]);
  __syncwarp();
  for (int r = lane; r < n_in; r += 32) {
    float acc[N];
#pragma unroll
    for (int j = 0; j < N; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (int p = 0; p < 32; ++p) {
      const float a = ws.tileA[p * (H + 1) + r];
      const float4* zr = reinterpret_cast<const float4*>(ws.tileZ + p * kTileZStride);
#pragma unroll
      for (int j4 = 0; j4 < N / 4; ++j4) {
        const float4 z = zr[j4];
        acc[4 * j4 + 0] = fmaf(a, z.x, acc[4 * j4 + 0]);
        acc[4 * j4 + 1] = fmaf(a, z.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(a, z.z, acc[4 * j4 + 2]);
        acc[4 * j4 + 3] = fmaf(a, z.w, acc[4 * j4 + 3]);
      }
    }
    float* dst = ws.acc + sh.acc_w_off(l) + r;
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j * RS] += acc[j];
  }
  if (WITH_BIAS) {
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
#pragma unroll 8
      for (int p = 0; p < 32; ++p) s += ws.tileZ[p * kTileZStride + j];
      ws.acc[sh.acc_b_off(l) + j] += s;
    }
  }
}

// One point's worth of state.  Everything indexed dynamically lives in local memory.
template <int H, int LHMAX>
struct PointState {
  float x[kDMax];
  float t[LHMAX][H];     // primal activations a_1..a_LH (tanh outputs)
  float tbar[LHMAX][H];  // adjoint contributions of the tangent streams to a_1..a_LH
  float z1[LHMAX][H];    // order-1 pre-activations of the current direction
  float z2[LHMAX][H];    // order-2 pre-activations of the current direction
  float a1[H], a2[H];    // inputs of the layer being processed (current direction)
  float zbuf1[H], zbuf2[H];
  float u[kOut];
  float ubar[kOut];
};

template <int H, int LHMAX>
struct MlpThread {
  const MlpShape<H>& sh;
  const float* sp;  // parameters in shared memory (flat layout of pdeip.h)
  PointState<H, LHMAX>& s;

  __device__ __forceinline__ MlpThread(const MlpShape<H>& shape, const float* smem_params,
                                       PointState<H, LHMAX>& st)
      : sh(shape), sp(smem_params), s(st) {}

  // ---- primal: t[l], u ----------------------------------------------------------------------
  __device__ __forceinline__ void primal_forward() {
    const int LH = sh.LH;
    {
      float acc[H];
      const float* b = sp + sh.b_off(0);
#pragma unroll
      for (int j = 0; j < H; ++j) acc[j] = b[j];
      gemv1<H>(sp + sh.w_off(0), sh.d, s.x, acc);
#pragma unroll
      for (int j = 0; j < H; ++j) s.t[0][j] = tanhf(acc[j]);
    }
    for (int l = 1; l < LH; ++l) {
      float acc[H];
      const float* b = sp + sh.b_off(l);
#pragma unroll
      for (int j = 0; j < H; ++j) acc[j] = b[j];
      gemv1<H>(sp + sh.w_off(l), H, s.t[l - 1], acc);
#pragma unroll
      for (int j = 0; j < H; ++j) s.t[l][j] = tanhf(acc[j]);
    }
    {
      float acc[kOut];
      const float* b = sp + sh.b_off(LH);
#pragma unroll
      for (int j = 0; j < kOut; ++j) acc[j] = b[j];
      gemv1<kOut>(sp + sh.w_off(LH), H, s.t[LH - 1], acc);
#pragma unroll
      for (int j = 0; j < kOut; ++j) s.u[j] = acc[j];
    }
  }

  __device__ __forceinline__ float value() const {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < kOut; ++j) v = fmaf(s.u[j], s.u[j], v);
    return v;
  }

  // ---- input gradient g = grad_x V  (written to g[0..d)) -------------------------------------
  __device__ __forceinline__ void input_gradient(float* g) {
    const int LH = sh.LH;
    {
      float zb[kOut];
#pragma unroll
      for (int j = 0; j < kOut; ++j) zb[j] = 2.f * s.u[j];
      const float* W = sp + sh.w_off(LH);
      for (int i = 0; i < H; ++i) {
        const float aa = dot_row<kOut>(zb, W + i * kOut);
        const float tt = s.t[LH - 1][i];
        s.zbuf1[i] = aa * (1.f - tt * tt);
      }
    }
    for (int l = LH - 1; l >= 1; --l) {
      float zb[H];
#pragma unroll
      for (int j = 0; j < H; ++j) zb[j] = s.zbuf1[j];
      const float* W = sp + sh.w_off(l);
      for (int i = 0; i < H; ++i) {
        const float aa = dot_row<H>(zb, W + i * H);
        const float tt = s.t[l - 1][i];
        s.zbuf1[i] = aa * (1.f - tt * tt);
      }
    }
    {
      float zb[H];
#pragma unroll
      for (int j = 0; j < H; ++j) zb[j] = s.zbuf1[j];
      const float* W = sp + sh.w_off(0);
      for (int i = 0; i < sh.d; ++i) g[i] = dot_row<H>(zb, W + i * H);
    }
  }

  // ---- forward Taylor stream along w; leaves z1/z2/a1/a2 for the reverse; returns u1,u2 -------
  template <bool ORDER2>
  __device__ __forceinline__ void direction_forward(const float* w, float (&u1)[kOut], float (&u2)[kOut],
                                                    float& D1, float& D2) {
    const int LH = sh.LH;
    {
      float acc1[H];
#pragma unroll
      for (int j = 0; j < H; ++j) acc1[j] = 0.f;
      gemv1<H>(sp + sh.w_off(0), sh.d, w, acc1);
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const float tt = s.t[0][j];
        const float s1 = 1.f - tt * tt;
        s.z1[0][j] = acc1[j];
        s.a1[j] = s1 * acc1[j];
        if (ORDER2) {
          s.z2[0][j] = 0.f;
          s.a2[j] = (-2.f * tt * s1) * acc1[j] * acc1[j];
        }
      }
    }
    for (int l = 1; l < LH; ++l) {
      float acc1[H], acc2[H];
#pragma unroll
      for (int j = 0; j < H; ++j) { acc1[j] = 0.f; acc2[j] = 0.f; }
      if (ORDER2) gemv2<H>(sp + sh.w_off(l), H, s.a1, s.a2, acc1, acc2);
      else gemv1<H>(sp + sh.w_off(l), H, s.a1, acc1);
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const float tt = s.t[l][j];
        const float s1 = 1.f - tt * tt;
        s.z1[l][j] = acc1[j];
        s.a1[j] = s1 * acc1[j];
        if (ORDER2) {
          s.z2[l][j] = acc2[j];
          s.a2[j] = s1 * acc2[j] + (-2.f * tt * s1) * acc1[j] * acc1[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kOut; ++j) { u1[j] = 0.f; u2[j] = 0.f; }
    if (ORDER2) gemv2<kOut>(sp + sh.w_off(LH), H, s.a1, s.a2, u1, u2);
    else gemv1<kOut>(sp + sh.w_off(LH), H, s.a1, u1);
    float d1 = 0.f, d2a = 0.f, d2b = 0.f;
#pragma unroll
    for (int j = 0; j < kOut; ++j) {
      d1 = fmaf(s.u[j], u1[j], d1);
      if (ORDER2) {
        d2a = fmaf(u1[j], u1[j], d2a);
        d2b = fmaf(s.u[j], u2[j], d2b);
      }
    }
    D1 = 2.f * d1;
    D2 = 2.f * (d2a + d2b);
  }

  // ---- reverse of the stream just run forward, for l = alpha D_w^2 V + beta D_w V ---------------
  // (u1,u2 are consumed; w is the direction itself, needed for dW_0)
  template <bool ORDER2>
  __device__ __forceinline__ void direction_reverse(const WarpScratch<H>& ws, const float* w, float alpha,
                                                    float beta, float (&u1)[kOut], float (&u2)[kOut]) {
    const int LH = sh.LH;
    // seeds
#pragma unroll
    for (int j = 0; j < kOut; ++j) {
      const float uj = s.u[j];
      s.ubar[j] += (ORDER2 ? 2.f * alpha * u2[j] : 0.f) + 2.f * beta * u1[j];
      const float zb1 = 4.f * alpha * u1[j] + 2.f * beta * uj;
      u2[j] = 2.f * alpha * uj;  // zb2
      u1[j] = zb1;
    }
    // last Dense layer
    {
      outer_acc<H, kOut, false>(sh, ws, LH, H, [&](int i) { return s.a1[i]; }, u1);
      if (ORDER2) outer_acc<H, kOut, false>(sh, ws, LH, H, [&](int i) { return s.a2[i]; }, u2);
      const float* W = sp + sh.w_off(LH);
      for (int i = 0; i < H; ++i) {
        float ab1, ab2 = 0.f;
        if (ORDER2) dot_row2<kOut>(u1, u2, W + i * kOut, ab1, ab2);
        else ab1 = dot_row<kOut>(u1, W + i * kOut);
        tanh_reverse<ORDER2>(LH - 1, i, ab1, ab2);
      }
    }
    for (int l = LH - 1; l >= 1; --l) {
      float zb1[H], zb2[H];
#pragma unroll
      for (int j = 0; j < H; ++j) { zb1[j] = s.zbuf1[j]; zb2[j] = ORDER2 ? s.zbuf2[j] : 0.f; }
      // inputs of layer l are a1_l = s1 z1_{l-1}, a2_l = s1 z2_{l-1} + s2 z1_{l-1}^2
      outer_acc<H, H, false>(sh, ws, l, H, [&](int i) {
        const float tt = s.t[l - 1][i];
        return (1.f - tt * tt) * s.z1[l - 1][i];
      }, zb1);
      if (ORDER2)
        outer_acc<H, H, false>(sh, ws, l, H, [&](int i) {
          const float tt = s.t[l - 1][i];
          const float s1 = 1.f - tt * tt;
          const float z1p = s.z1[l - 1][i];
          return s1 * s.z2[l - 1][i] + (-2.f * tt * s1) * z1p * z1p;
        }, zb2);
      const float* W = sp + sh.w_off(l);
      for (int i = 0; i < H; ++i) {
        float ab1, ab2 = 0.f;
        if (ORDER2) dot_row2<H>(zb1, zb2, W + i * H, ab1, ab2);
        else ab1 = dot_row<H>(zb1, W + i * H);
        tanh_reverse<ORDER2>(l - 1, i, ab1, ab2);
      }
    }
    {
      float zb1[H];
#pragma unroll
      for (int j = 0; j < H; ++j) zb1[j] = s.zbuf1[j];
      outer_acc<H, H, false>(sh, ws, 0, sh.d, [&](int i) { return w[i]; }, zb1);  // a2_0 = 0
    }
  }

  // through tanh of hidden layer `lt` (0-based), unit i: accumulate tbar, produce next zbar1/zbar2
  template <bool ORDER2>
  __device__ __forceinline__ void tanh_reverse(int lt, int i, float ab1, float ab2) {
    const float tt = s.t[lt][i];
    const float s1 = 1.f - tt * tt;
    const float z1p = s.z1[lt][i];
    float tb = ab1 * z1p * (-2.f * tt);
    float nz1 = ab1 * s1;
    if (ORDER2) {
      const float z2p = s.z2[lt][i];
      const float s2 = -2.f * tt * s1;
      tb += ab2 * (z2p * (-2.f * tt) + z1p * z1p * (6.f * tt * tt - 2.f));
      nz1 += ab2 * 2.f * s2 * z1p;
      s.zbuf2[i] = ab2 * s1;
    }
    s.tbar[lt][i] += tb;
    s.zbuf1[i] = nz1;
  }

  // ---- primal reverse: zbar_L = ubar + 2 kappa u, plus the tangent streams' tbar ------------------
  __device__ __forceinline__ void primal_reverse(const WarpScratch<H>& ws, float kappa) {
    const int LH = sh.LH;
    {
      float zb[kOut];
#pragma unroll
      for (int j = 0; j < kOut; ++j) zb[j] = s.ubar[j] + 2.f * kappa * s.u[j];
      outer_acc<H, kOut, true>(sh, ws, LH, H, [&](int i) { return s.t[LH - 1][i]; }, zb);
      const float* W = sp + sh.w_off(LH);
      for (int i = 0; i < H; ++i) {
        const float ab = dot_row<kOut>(zb, W + i * kOut) + s.tbar[LH - 1][i];
        const float tt = s.t[LH - 1][i];
        s.zbuf1[i] = ab * (1.f - tt * tt);
      }
    }
    for (int l = LH - 1; l >= 1; --l) {
      float zb[H];
#pragma unroll
      for (int j = 0; j < H; ++j) zb[j] = s.zbuf1[j];
      outer_acc<H, H, true>(sh, ws, l, H, [&](int i) { return s.t[l - 1][i]; }, zb);
      const float* W = sp + sh.w_off(l);
      for (int i = 0; i < H; ++i) {
        const float ab = dot_row<H>(zb, W + i * H) + s.tbar[l - 1][i];
        const float tt = s.t[l - 1][i];
        s.zbuf1[i] = ab * (1.f - tt * tt);
      }
    }
    {
      float zb[H];
#pragma unroll
      for (int j = 0; j < H; ++j) zb[j] = s.zbuf1[j];
      outer_acc<H, H, true>(sh, ws, 0, sh.d, [&](int i) { return s.x[i]; }, zb);
    }
  }

  __device__ __forceinline__ void reset_adjoints() {
#pragma unroll
    for (int j = 0; j < kOut; ++j) s.ubar[j] = 0.f;
    for (int l = 0; l < sh.LH; ++l)
#pragma unroll
      for (int j = 0; j < H; ++j) s.tbar[l][j] = 0.f;
  }
};

}  // namespace pdeip
