// residual_tensor.cu — K3/K4 on the 5th-generation tensor cores (tcgen05 + TMEM), KFP 0T point set.
//
// Same mathematics as residual_mlp.cu (reference: kinetic_fokker_planck.py:11-69), organised as a
// chain of 128-point x {32,48}-unit GEMMs:  one CTA = 128 threads = one 128-point tile at a time, thread p owns
// point p (TMEM lane p) in every epilogue, one elected thread issues the MMAs.
//
//   operands   bf16 in shared memory, no-swizzle core-matrix layout (umma.cuh); every tile is written once and
//              used both K-major (layer GEMMs: points x units) and through the transposed view (batch-reduced
//              dW GEMMs: units x points).  Weights are split W = W_hi + W_lo (two bf16 MMAs per k-step), which
//              removes the systematic weight-rounding error; activations are rounded to bf16 once.
//   accumulators  fp32 in TMEM; epilogues (tanh, s1 = 1-t^2, s2, Taylor-stream and adjoint updates) in fp32.
//   phases     S0 x,v -> z0, z1_0 | S1 t1,a1,a2 -> z1, z1_1, z2_1 | S2 t2,a1,a2 -> u,u1,u2 | S3-S5 input-gradient
//              chain -> g | S6-S8 stop-gradient stream along g -> ug | S9 seeds -> abar (layer 2) + dW2 + db2 |
//              S10 tanh-reverse -> abar (layer 1) + dW1 + db1 | S11 tanh-reverse -> dW0 + db0 | S12 read-out.
//   dW         D[(stream,unit_in)][unit_out] = sum_points ACT^T Zbar_s with the 4 streams stacked along M; only
//              the band of stream s is meaningful in region s, and that band is exactly the TMEM lane quadrant
//              of warp s, so each warp reads back 32 lanes x N columns per tile and keeps the running sums in
//              registers.  db_l comes from a constant-one unit in the x|v|g tile (persistent TMEM regions).
//   parity     rtol 1e-2 (BASELINE.json, bf16 GEMM path); measured against the oracle in tests/test_gpu_tensor.py.
#include "mlp_thread.cuh"
#include "residual_common.cuh"
#include "umma.cuh"

namespace pdeip {

int mlp_residual_accumulate_fp32(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st);

namespace tc {

using namespace umma;

constexpr int H = 32;
constexpr int OP = 48;  // output width 40 padded to a multiple of 16

// TMEM column map (512 columns allocated)
constexpr uint32_t C_T1 = 0, C_T2 = 32, C_DB2 = 64, C_DB1 = 112;   // persistent
constexpr uint32_t C_Z10 = 144, C_Z11 = 176, C_Z21 = 208;            // S0/S1 transients
constexpr uint32_t C_U = 144, C_U1 = 192, C_U2 = 240, C_UG = 288;    // S2..S9
constexpr uint32_t C_AA2 = 336, C_AA1 = 368, C_G = 400, C_ZG0 = 432, C_ZG1 = 464;
constexpr uint32_t C_DW = 144;   // dW regions of the reverse phases
constexpr uint32_t C_AB = 336;   // abar_s regions: C_AB + 32 s

template <int DP>
struct Smem {
  static constexpr uint32_t RG_T0 = DP / 8 * 128, RG_T1 = 512, RG_T2 = 512;
  static constexpr uint32_t SZ_T0 = 4 * RG_T0, SZ_T1 = 4 * RG_T1, SZ_T2 = 6 * RG_T2;
  static constexpr uint32_t RG_X = 4 * DP / 8 * 128, SZ_X = 16 * RG_X;
  static constexpr uint32_t RG_A = 2048, SZ_A = 16 * RG_A;
  static constexpr uint32_t RG_Z = 4 * OP / 8 * 128, SZ_Z = 16 * RG_Z;
  static constexpr uint32_t O_T0H = 0, O_T0L = O_T0H + SZ_T0, O_T1H = O_T0L + SZ_T0, O_T1L = O_T1H + SZ_T1,
                            O_T2H = O_T1L + SZ_T1, O_T2L = O_T2H + SZ_T2, O_BIAS = O_T2L + SZ_T2;
  static constexpr uint32_t O_X = O_BIAS + 512, O_A1 = O_X + SZ_X, O_A2 = O_A1 + SZ_A, O_Z = O_A2 + SZ_A;
  static constexpr uint32_t O_STASH = O_Z + SZ_Z, SZ_STASH = 5 * 32 * 128 * 4;
  static constexpr uint32_t O_TRUE = O_STASH + SZ_STASH, SZ_TRUE = 4096;
  static constexpr uint32_t O_MISC = O_TRUE + SZ_TRUE, TOTAL = O_MISC + 64;
};

// stash arrays (fp32 [unit][point])
enum { ST_Z10 = 0, ST_ZG0 = 1, ST_Z11 = 2, ST_Z21 = 3, ST_ZG1 = 4 };

struct Ctx {
  uint32_t tbase;      // TMEM base (lane 0)
  uint32_t lane_addr;  // TMEM address of this thread's lane quadrant
  uint32_t mbar;       // shared address of the mbarrier
  uint32_t parity;
  int* status;
  bool ok;
};

// epilogue done -> make operand writes visible to the tensor core, order TMEM reads, sync the CTA
__device__ __forceinline__ void phase_sync() {
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
}

__device__ __forceinline__ void phase_wait(Ctx& c) {
  if (c.ok && !mbar_wait(c.mbar, c.parity)) {
    c.ok = false;
    if (threadIdx.x == 0) atomicExch(c.status, 1);
  }
  c.parity ^= 1u;
  fence_after_sync();
}

// two MMAs per k-step: hi and lo halves of the split weights (forward: K-major B tile [out][in])
__device__ __forceinline__ void gemm_fwd(uint32_t d, uint32_t a_tile, uint32_t a_rg, int a_c0, uint32_t w_hi,
                                         uint32_t w_lo, uint32_t w_rg, int K, int N) {
  gemm_kk(d, a_tile, a_rg, a_c0, w_hi, w_rg, 0, K, N, 0);
  gemm_kk(d, a_tile, a_rg, a_c0, w_lo, w_rg, 0, K, N, 1);
}
// backward: abar = zbar W^T, the same weight tile through the transposed view (operand rows = in units)
__device__ __forceinline__ void gemm_bwd(uint32_t d, uint32_t a_tile, uint32_t a_rg, int a_c0, uint32_t w_hi,
                                         uint32_t w_lo, uint32_t w_rg, int K, int N) {
  gemm_km(d, a_tile, a_rg, a_c0, w_hi, w_rg, 0, 0, K, N, 0);
  gemm_km(d, a_tile, a_rg, a_c0, w_lo, w_rg, 0, 0, K, N, 1);
}

template <int DP>
__global__ void __launch_bounds__(128, 1) mlp_residual_tc_kernel(const ResidualArgs a, int* status) {
  using S = Smem<DP>;
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int d = a.d;
  const MlpShape<H> sh{d, 2};
  const int P = sh.num_params();
  float* bias_s = reinterpret_cast<float*>(sm + S::O_BIAS);   // b0[32] b1[32] b2[48]
  float* stash = reinterpret_cast<float*>(sm + S::O_STASH);
  float* tp = reinterpret_cast<float*>(sm + S::O_TRUE);
  uint64_t* mbar_p = reinterpret_cast<uint64_t*>(sm + S::O_MISC);
  uint32_t* tmem_p = reinterpret_cast<uint32_t*>(sm + S::O_MISC + 16);

  const int64_t n_tiles = (a.n_points + 127) / 128;
  if ((int64_t)blockIdx.x >= n_tiles) return;  // nothing to do for this CTA (uniform)

  // ---- one-time set-up: TMEM, mbarrier, split weights in core-matrix layout, constants ------------------
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_p), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(mbar_p), 1);
    fence_mbar_init();
  }
  {
    const float* W0 = a.params + sh.w_off(0);
    const float* W1 = a.params + sh.w_off(1);
    const float* W2 = a.params + sh.w_off(2);
    auto put = [&](uint32_t o_hi, uint32_t o_lo, uint32_t rg, int r, int c, float w) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(w);
      const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
      const uint32_t off = chunk_off(r, c >> 3, rg) + (uint32_t)(c & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16*>(sm + o_hi + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(sm + o_lo + off) = lo;
    };
    for (int idx = tid; idx < 32 * DP; idx += 128) {  // T0[r = out][c = in] = W0[c][r]
      const int r = idx / DP, c = idx % DP;
      put(S::O_T0H, S::O_T0L, S::RG_T0, r, c, c < d ? W0[c * H + r] : 0.f);
    }
    for (int idx = tid; idx < 32 * 32; idx += 128) {
      const int r = idx / 32, c = idx % 32;
      put(S::O_T1H, S::O_T1L, S::RG_T1, r, c, W1[c * H + r]);
    }
    for (int idx = tid; idx < OP * 32; idx += 128) {  // T2[r = out (48)][c = in] = W2[c][r], rows >= 40 zero
      const int r = idx / 32, c = idx % 32;
      put(S::O_T2H, S::O_T2L, S::RG_T2, r, c, r < kOut ? W2[c * kOut + r] : 0.f);
    }
    for (int j = tid; j < 32; j += 128) {
      bias_s[j] = a.params[sh.b_off(0) + j];
      bias_s[32 + j] = a.params[sh.b_off(1) + j];
    }
    for (int j = tid; j < OP; j += 128) bias_s[64 + j] = j < kOut ? a.params[sh.b_off(2) + j] : 0.f;
    int ntg = 0;
    if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = d * d;
    else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * d;
    for (int i = tid; i < ntg; i += 128) tp[i] = a.tg.params[i];
    // x|v|g tile: zero everything once, then the constant-one unit (column 3*DP) of every point
    for (uint32_t o = tid * 16; o < S::SZ_X; o += 128 * 16) *reinterpret_cast<uint4*>(sm + S::O_X + o) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store_chunk(sm + S::O_X, chunk_off(tid, 3 * DP / 8, S::RG_X), ones);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  Ctx c;
  c.tbase = *tmem_p;
  c.lane_addr = c.tbase + ((uint32_t)(warp * 32) << 16);
  c.mbar = smem_u32(mbar_p);
  c.parity = 0;
  c.status = status;
  c.ok = true;
  const uint32_t LA = c.lane_addr, TB = c.tbase;
  const uint32_t xs = smem_u32(sm + S::O_X), a1s = smem_u32(sm + S::O_A1), a2s = smem_u32(sm + S::O_A2),
                 zs = smem_u32(sm + S::O_Z);
  const uint32_t t0h = smem_u32(sm + S::O_T0H), t0l = smem_u32(sm + S::O_T0L), t1h = smem_u32(sm + S::O_T1H),
                 t1l = smem_u32(sm + S::O_T1L), t2h = smem_u32(sm + S::O_T2H), t2l = smem_u32(sm + S::O_T2L);
  uint8_t* X = sm + S::O_X;
  uint8_t* A1 = sm + S::O_A1;
  uint8_t* A2 = sm + S::O_A2;
  uint8_t* Z = sm + S::O_Z;
  auto ST = [&](int arr, int unit) -> float& { return stash[(arr * 32 + unit) * 128 + tid]; };

  // running sums kept in registers across the CTA's tiles
  float acc2[40], acc1[32], acc0[32];
#pragma unroll
  for (int j = 0; j < 40; ++j) acc2[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { acc1[j] = 0.f; acc0[j] = 0.f; }
  float sums[PDEIP_NUM_SUMS];
#pragma unroll
  for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;
  const float gamma = a.coef;
  uint32_t not_first = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t p = tile * 128 + tid;
    const bool valid = p < a.n_points;
    const float wt = valid ? a.weight : 0.f;
    const float alpha = -2.f * wt, beta = 2.f * gamma * wt, beta_g = 2.f * wt;
    float x[DP];
    // ---- S0: x, v bands ---------------------------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < DP / 8; ++cg) {
      float xv[8], vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int u = cg * 8 + i;
        const bool in = valid && u < d;
        xv[i] = in ? a.points[elem_index(a.layout, p, u, a.n_points, 2 * d)] : 0.f;
        vv[i] = in ? a.points[elem_index(a.layout, p, d + u, a.n_points, 2 * d)] : 0.f;
        x[u] = xv[i];
      }
      store_chunk(X, chunk_off(tid, cg, S::RG_X), xv);
      store_chunk(X, chunk_off(tid, DP / 8 + cg, S::RG_X), vv);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_fwd(TB + C_T1, xs, S::RG_X, 0, t0h, t0l, S::RG_T0, DP, 32);
      gemm_fwd(TB + C_Z10, xs, S::RG_X, DP, t0h, t0l, S::RG_T0, DP, 32);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S1: t1, a1_1, a2_1 ---------------------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float z0[8], z1[8], t[8], q1[8], q2[8];
      tmem_ld8x2(LA + C_T1 + 8 * cg, LA + C_Z10 + 8 * cg, z0, z1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t[i] = tanhf(z0[i] + bias_s[cg * 8 + i]);
        const float s1 = 1.f - t[i] * t[i];
        q1[i] = s1 * z1[i];
        q2[i] = (-2.f * t[i] * s1) * z1[i] * z1[i];
        ST(ST_Z10, cg * 8 + i) = z1[i];
      }
      store_chunk(A1, chunk_off(tid, cg, S::RG_A), t);
      store_chunk(A1, chunk_off(tid, 4 + cg, S::RG_A), q1);
      store_chunk(A1, chunk_off(tid, 8 + cg, S::RG_A), q2);
      tmem_st8(LA + C_T1 + 8 * cg, t);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_fwd(TB + C_T2, a1s, S::RG_A, 0, t1h, t1l, S::RG_T1, 32, 32);
      gemm_fwd(TB + C_Z11, a1s, S::RG_A, 32, t1h, t1l, S::RG_T1, 32, 32);
      gemm_fwd(TB + C_Z21, a1s, S::RG_A, 64, t1h, t1l, S::RG_T1, 32, 32);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S2: t2, a1_2, a2_2 ---------------------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float z0[8], z1[8], z2[8], t[8], q1[8], q2[8];
      tmem_ld8x3(LA + C_T2 + 8 * cg, LA + C_Z11 + 8 * cg, LA + C_Z21 + 8 * cg, z0, z1, z2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t[i] = tanhf(z0[i] + bias_s[32 + cg * 8 + i]);
        const float s1 = 1.f - t[i] * t[i];
        q1[i] = s1 * z1[i];
        q2[i] = s1 * z2[i] + (-2.f * t[i] * s1) * z1[i] * z1[i];
        ST(ST_Z11, cg * 8 + i) = z1[i];
        ST(ST_Z21, cg * 8 + i) = z2[i];
      }
      store_chunk(A2, chunk_off(tid, cg, S::RG_A), t);
      store_chunk(A2, chunk_off(tid, 4 + cg, S::RG_A), q1);
      store_chunk(A2, chunk_off(tid, 8 + cg, S::RG_A), q2);
      tmem_st8(LA + C_T2 + 8 * cg, t);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_fwd(TB + C_U, a2s, S::RG_A, 0, t2h, t2l, S::RG_T2, 32, OP);
      gemm_fwd(TB + C_U1, a2s, S::RG_A, 32, t2h, t2l, S::RG_T2, 32, OP);
      gemm_fwd(TB + C_U2, a2s, S::RG_A, 64, t2h, t2l, S::RG_T2, 32, OP);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S3: za_2 = 2u  ->  aa_2 = za_2 W2^T -----------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < OP / 8; ++cg) {
      float u[8];
      tmem_ld8(LA + C_U + 8 * cg, u);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = 2.f * (u[i] + bias_s[64 + cg * 8 + i]);
      store_chunk(Z, chunk_off(tid, cg, S::RG_Z), u);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_bwd(TB + C_AA2, zs, S::RG_Z, 0, t2h, t2l, S::RG_T2, OP, 32);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S4: za_1 = aa_2 (1 - t2^2) -> aa_1 ------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float aa[8], t[8];
      tmem_ld8x2(LA + C_AA2 + 8 * cg, LA + C_T2 + 8 * cg, aa, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) aa[i] *= (1.f - t[i] * t[i]);
      store_chunk(Z, chunk_off(tid, OP / 8 + cg, S::RG_Z), aa);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_bwd(TB + C_AA1, zs, S::RG_Z, OP, t1h, t1l, S::RG_T1, 32, 32);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S5: za_0 = aa_1 (1 - t1^2) -> g ---------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float aa[8], t[8];
      tmem_ld8x2(LA + C_AA1 + 8 * cg, LA + C_T1 + 8 * cg, aa, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) aa[i] *= (1.f - t[i] * t[i]);
      store_chunk(Z, chunk_off(tid, 2 * OP / 8 + cg, S::RG_Z), aa);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_bwd(TB + C_G, zs, S::RG_Z, 2 * OP, t0h, t0l, S::RG_T0, 32, DP);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S6: g band, |g|^2, true gradient -------------------------------------------------------------
    float g2 = 0.f, gt2 = 0.f, gd2 = 0.f;
    {
      float gq[DP], gt[DP];
#pragma unroll
      for (int cg = 0; cg < DP / 8; ++cg) {
        float gv[8];
        tmem_ld8(LA + C_G + 8 * cg, gv);
#pragma unroll
        for (int i = 0; i < 8; ++i) gq[cg * 8 + i] = gv[i];
        store_chunk(X, chunk_off(tid, 2 * DP / 8 + cg, S::RG_X), gv);
      }
      true_grad_thread(a.tg, tp, d, x, gt);
      for (int i = 0; i < d; ++i) {
        g2 = fmaf(gq[i], gq[i], g2);
        gt2 = fmaf(gt[i], gt[i], gt2);
        const float df = gt[i] - gq[i];
        gd2 = fmaf(df, df, gd2);
      }
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_fwd(TB + C_ZG0, xs, S::RG_X, 2 * DP, t0h, t0l, S::RG_T0, DP, 32);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S7: ag_1 = (1 - t1^2) zg_0 --------------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float zg[8], t[8];
      tmem_ld8x2(LA + C_ZG0 + 8 * cg, LA + C_T1 + 8 * cg, zg, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ST(ST_ZG0, cg * 8 + i) = zg[i];
        zg[i] *= (1.f - t[i] * t[i]);
      }
      store_chunk(A1, chunk_off(tid, 12 + cg, S::RG_A), zg);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_fwd(TB + C_ZG1, a1s, S::RG_A, 96, t1h, t1l, S::RG_T1, 32, 32);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S8: ag_2 = (1 - t2^2) zg_1 --------------------------------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float zg[8], t[8];
      tmem_ld8x2(LA + C_ZG1 + 8 * cg, LA + C_T2 + 8 * cg, zg, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ST(ST_ZG1, cg * 8 + i) = zg[i];
        zg[i] *= (1.f - t[i] * t[i]);
      }
      store_chunk(A2, chunk_off(tid, 12 + cg, S::RG_A), zg);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_fwd(TB + C_UG, a2s, S::RG_A, 96, t2h, t2l, S::RG_T2, 32, OP);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S9: loss terms and seeds ------------------------------------------------------------
    {
      float d1 = 0.f, d2a = 0.f, d2b = 0.f;
#pragma unroll
      for (int cg = 0; cg < OP / 8; ++cg) {
        float u[8], u1[8], u2[8], ug[8], s0[8], s1v[8], s2v[8], sg[8];
        tmem_ld8x4(LA + C_U + 8 * cg, LA + C_U1 + 8 * cg, LA + C_U2 + 8 * cg, LA + C_UG + 8 * cg, u, u1, u2, ug);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float uu = u[i] + bias_s[64 + cg * 8 + i];
          d1 = fmaf(uu, u1[i], d1);
          d2a = fmaf(u1[i], u1[i], d2a);
          d2b = fmaf(uu, u2[i], d2b);
          s0[i] = 2.f * alpha * u2[i] + 2.f * beta * u1[i] + 2.f * beta_g * ug[i];
          s1v[i] = 4.f * alpha * u1[i] + 2.f * beta * uu;
          s2v[i] = 2.f * alpha * uu;
          sg[i] = 2.f * beta_g * uu;
        }
        store_chunk(Z, chunk_off(tid, cg, S::RG_Z), s0);
        store_chunk(Z, chunk_off(tid, OP / 8 + cg, S::RG_Z), s1v);
        store_chunk(Z, chunk_off(tid, 2 * OP / 8 + cg, S::RG_Z), s2v);
        store_chunk(Z, chunk_off(tid, 3 * OP / 8 + cg, S::RG_Z), sg);
      }
      const float D1 = 2.f * d1, D2 = 2.f * (d2a + d2b);
      sums[PDEIP_SUM_G2] += wt * g2;
      sums[PDEIP_SUM_D2] += wt * D2;
      sums[PDEIP_SUM_D1] += wt * D1;
      sums[PDEIP_SUM_GTRUE2] += wt * gt2;
      sums[PDEIP_SUM_GT] += wt * gd2;
      sums[PDEIP_SUM_LOSS] += wt * (g2 - 2.f * D2 + 2.f * gamma * D1 + gt2);
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        gemm_bwd(TB + C_AB + 32 * s, zs, S::RG_Z, OP * s, t2h, t2l, S::RG_T2, OP, 32);
        gemm_mm(TB + C_DW + OP * s, a2s, S::RG_A, 0, zs, S::RG_Z, OP * s, 128, OP, 0);
      }
      gemm_mm(TB + C_DB2, xs, S::RG_X, 0, zs, S::RG_Z, 0, 128, OP, not_first);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S10: through tanh of hidden layer 2; read dW2 band ----------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float ab[8], ab1[8], ab2[8], abg[8], t[8];
      tmem_ld8x4(LA + C_AB + 8 * cg, LA + C_AB + 32 + 8 * cg, LA + C_AB + 64 + 8 * cg, LA + C_AB + 96 + 8 * cg, ab, ab1,
                 ab2, abg);
      tmem_ld8(LA + C_T2 + 8 * cg, t);
      float o0[8], o1[8], o2[8], og[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int un = cg * 8 + i;
        const float tt = t[i], s1 = 1.f - tt * tt, s2 = -2.f * tt * s1;
        const float z1 = ST(ST_Z11, un), z2 = ST(ST_Z21, un), zg = ST(ST_ZG1, un);
        const float tb = ab[i] + ab1[i] * z1 * (-2.f * tt) + ab2[i] * (z2 * (-2.f * tt) + z1 * z1 * (6.f * tt * tt - 2.f)) +
                         abg[i] * zg * (-2.f * tt);
        o0[i] = tb * s1;
        o1[i] = ab1[i] * s1 + ab2[i] * 2.f * s2 * z1;
        o2[i] = ab2[i] * s1;
        og[i] = abg[i] * s1;
      }
      store_chunk(Z, chunk_off(tid, cg, S::RG_Z), o0);
      store_chunk(Z, chunk_off(tid, OP / 8 + cg, S::RG_Z), o1);
      store_chunk(Z, chunk_off(tid, 2 * OP / 8 + cg, S::RG_Z), o2);
      store_chunk(Z, chunk_off(tid, 3 * OP / 8 + cg, S::RG_Z), og);
    }
#pragma unroll
    for (int cg = 0; cg < 5; ++cg) {  // warp s = stream s: rows of dW2 owned by this lane
      float w[8];
      tmem_ld8(LA + C_DW + OP * warp + 8 * cg, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc2[cg * 8 + i] += w[i];
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        gemm_bwd(TB + C_AB + 32 * s, zs, S::RG_Z, OP * s, t1h, t1l, S::RG_T1, 32, 32);
        gemm_mm(TB + C_DW + 32 * s, a1s, S::RG_A, 0, zs, S::RG_Z, OP * s, 128, 32, 0);
      }
      gemm_mm(TB + C_DB1, xs, S::RG_X, 0, zs, S::RG_Z, 0, 128, 32, not_first);
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S11: through tanh of hidden layer 1; read dW1 band ----------------------------------------------
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float ab[8], ab1[8], ab2[8], abg[8], t[8];
      tmem_ld8x4(LA + C_AB + 8 * cg, LA + C_AB + 32 + 8 * cg, LA + C_AB + 64 + 8 * cg, LA + C_AB + 96 + 8 * cg, ab, ab1,
                 ab2, abg);
      tmem_ld8(LA + C_T1 + 8 * cg, t);
      float o0[8], o1[8], og[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int un = cg * 8 + i;
        const float tt = t[i], s1 = 1.f - tt * tt, s2 = -2.f * tt * s1;
        const float z1 = ST(ST_Z10, un), zg = ST(ST_ZG0, un);  // z2_0 = 0
        const float tb = ab[i] + ab1[i] * z1 * (-2.f * tt) + ab2[i] * (z1 * z1 * (6.f * tt * tt - 2.f)) +
                         abg[i] * zg * (-2.f * tt);
        o0[i] = tb * s1;
        o1[i] = ab1[i] * s1 + ab2[i] * 2.f * s2 * z1;
        og[i] = abg[i] * s1;
      }
      store_chunk(Z, chunk_off(tid, cg, S::RG_Z), o0);
      store_chunk(Z, chunk_off(tid, OP / 8 + cg, S::RG_Z), o1);
      store_chunk(Z, chunk_off(tid, 2 * OP / 8 + cg, S::RG_Z), og);
    }
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      float w[8];
      tmem_ld8(LA + C_DW + 32 * warp + 8 * cg, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc1[cg * 8 + i] += w[i];
    }
    phase_sync();
    if (tid == 0) {
      fence_after_sync();
      gemm_mm(TB + C_DW, xs, S::RG_X, 0, zs, S::RG_Z, 0, 128, 32, 0);           // rows x      <-> zbar  (primal)
      gemm_mm(TB + C_DW + 32, xs, S::RG_X, 0, zs, S::RG_Z, OP, 128, 32, 0);     // rows v      <-> zbar1
      gemm_mm(TB + C_DW + 64, xs, S::RG_X, 0, zs, S::RG_Z, 2 * OP, 128, 32, 0); // rows g      <-> zbar_g
      commit(c.mbar);
    }
    phase_wait(c);
    // ---- S12: read dW0 rows (lane = unit of the x|v|g|one tile) --------------------------------------------
    if (warp * 32 < 4 * DP) {  // warp-uniform
      const int band = tid / DP;  // 0: x rows, 1: v rows, 2: g rows, 3: the constant-one row (db0) and padding
      const int sel = (band == 1) ? 1 : (band == 2 ? 2 : 0);
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        float w0[8], w1[8], w2[8];
        tmem_ld8x3(LA + C_DW + 8 * cg, LA + C_DW + 32 + 8 * cg, LA + C_DW + 64 + 8 * cg, w0, w1, w2);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc0[cg * 8 + i] += (sel == 0) ? w0[i] : (sel == 1 ? w1[i] : w2[i]);
      }
    }
    not_first = 1;
    fence_before_sync();  // TMEM reads of this tile are ordered before the next tile's MMAs by the next phase_sync
  }

  // ---- write-out: per-CTA partial gradient and sums -------------------------------------------------------
  __syncthreads();
  // four contributor slices (stream / band / warp): every (slice, index) has exactly one writer, and the slices
  // are summed in a fixed order, so the result is bit-reproducible.
  float* red = stash;  // the stash is dead now (4 * (P + 8) floats <= 80 KB)
  const int PS = P + PDEIP_NUM_SUMS;
  for (int i = tid; i < 4 * PS; i += 128) red[i] = 0.f;
  __syncthreads();
  {
    const int lane = tid & 31;
    float* mine = red + warp * PS;  // dW2 / dW1: warp s holds stream s's contribution to row (lane) of the layer
#pragma unroll
    for (int j = 0; j < kOut; ++j) mine[sh.w_off(2) + lane * kOut + j] = acc2[j];
#pragma unroll
    for (int j = 0; j < 32; ++j) mine[sh.w_off(1) + lane * H + j] = acc1[j];
    // dW0 / db0: lane = unit of the x|v|g|one tile; slice = band
    if (tid < 4 * DP) {
      const int band = tid / DP, i = tid % DP;
      if (band < 3 && i < d) {
#pragma unroll
        for (int j = 0; j < 32; ++j) red[band * PS + sh.w_off(0) + i * H + j] = acc0[j];
      } else if (tid == 3 * DP) {
#pragma unroll
        for (int j = 0; j < 32; ++j) red[sh.b_off(0) + j] = acc0[j];
      }
    }
    // db2 / db1 from the persistent regions (row of the constant-one unit)
    if (warp == (3 * DP) / 32) {  // warp-uniform: the warp that owns lane 3*DP
#pragma unroll
      for (int cg = 0; cg < 5; ++cg) {
        float w[8];
        tmem_ld8(LA + C_DB2 + 8 * cg, w);
        if (tid == 3 * DP)
          for (int i = 0; i < 8; ++i) red[sh.b_off(2) + cg * 8 + i] = w[i];
      }
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        float w[8];
        tmem_ld8(LA + C_DB1 + 8 * cg, w);
        if (tid == 3 * DP)
          for (int i = 0; i < 8; ++i) red[sh.b_off(1) + cg * 8 + i] = w[i];
      }
    }
#pragma unroll
    for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
      const float sk = warp_sum(sums[k]);
      if ((tid & 31) == 0) mine[P + k] = sk;
    }
  }
  fence_before_sync();
  __syncthreads();
  float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
  if (c.ok)
    for (int i = tid; i < PS; i += 128) part[i] += (red[i] + red[PS + i]) + (red[2 * PS + i] + red[3 * PS + i]);
  if (warp == 0) tmem_dealloc(c.tbase, 512);
}

}  // namespace tc

static int* tensor_status_word() {
  static int* w = nullptr;
  if (!w) {
    if (cudaMalloc(&w, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(w, 0, sizeof(int));
  }
  return w;
}

#ifdef PDEIP_HAVE_TENSOR_PATH
int mlp_residual_accumulate_tensor(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st) {
  // boundary sets (two launches over n points vs n*S points of the 0T set) stay on the fp32 kernel
  if (set_kind != PDEIP_SET_KFP_0T) return mlp_residual_accumulate_fp32(set_kind, a, hidden, st);
  PDEIP_REQUIRE(hidden == 32 && a.layers == 2, PDEIP_ERR_UNSUPPORTED,
                "tensor path is built for hidden_dim == 32, layers == 2 (got %d, %d)", hidden, a.layers);
  PDEIP_REQUIRE(a.d >= 1 && a.d <= 16, PDEIP_ERR_UNSUPPORTED, "tensor path supports 1 <= d <= 16 (got %d)", a.d);
  PDEIP_REQUIRE(true_grad_floats(a.tg, a.d) <= 1024, PDEIP_ERR_UNSUPPORTED,
                "tensor path: true-gradient parameters exceed 4 KB");
  int* status = tensor_status_word();
  PDEIP_REQUIRE(status != nullptr, PDEIP_ERR_CUDA, "cannot allocate the tensor-path status word");
  using S = tc::Smem<16>;
  auto kern = tc::mlp_residual_tc_kernel<16>;
  PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL));
  kern<<<residual_grid(), 128, S::TOTAL, st>>>(a, status);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}
#endif

int tensor_path_status(cudaStream_t st, int* out) {
  int* status = tensor_status_word();
  if (!status) return PDEIP_ERR_CUDA;
  if (cudaMemcpyAsync(out, status, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return PDEIP_ERR_CUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) return PDEIP_ERR_CUDA;
  return PDEIP_OK;
}

}  // namespace pdeip

// 0 = every tcgen05 phase completed; 1 = a bounded mbarrier wait timed out (results invalid).  Synchronises.
extern "C" int pdeip_tensor_path_status(void* stream, int* out_status) {
  PDEIP_REQUIRE(out_status != nullptr, PDEIP_ERR_INVALID_ARG, "out_status is NULL");
  return pdeip::tensor_path_status((cudaStream_t)stream, out_status);
}
