// residual_tensor.cu — K3/K4 on the 5th-generation tensor cores (tcgen05 + TMEM), KFP 0T point set, version 2.
//
// Mathematics: SURVEY.md §9 (reference: methods/consistency_instances/kinetic_fokker_planck.py:11-69), restated
// phase by phase in tests/tensor_v2_model.py (float64 twin, checked against oracle/taylor.py to 1e-16).
// Per point (x, v):  l = |g|^2 - 2 D_v^2 V + 2 gamma D_v V,  g = grad_x V,  V = |MLP(x)|^2,  and dl/dtheta.
//
// Organisation.  One CTA per SM, persistent over 128-point tiles; NS tiles ("slots") are in flight per CTA so
// that one slot's epilogue hides the other's GEMM hand-off latency (NS = 2 for d <= 8, 1 otherwise).
//   * 16 epilogue warps: thread = (point row == TMEM lane, 8-unit chunk); 1 MMA warp: one elected lane issues
//     every tcgen05.mma of both slots.
//   * operands: bf16 in shared memory, no-swizzle core-matrix layout (umma.cuh); every tile is written once and
//     used K-major (layer GEMMs, points x units) and through the transposed view (batch-reduced dW GEMMs).
//     Weights and the input x are split hi + lo (two bf16 terms), activations are rounded to bf16 once.
//   * accumulators: fp32 in TMEM.  Values an epilogue needs again later are "parked" in TMEM as packed bf16 pairs
//     (s1 = 1 - t^2 of both hidden layers, the partial seed s0p, the pz terms).
//   * streams.  Forward Taylor streams along v: a (primal), a1, a2~ = -2 a2.  Input gradient chain za_l / aa_l.
//     Stop-gradient stream along g~ = 2 g (order 1).  The adjoints of the a2 and g streams are proportional to the
//     input-gradient chain (zbar2 = -2 za, zbar_g = 2 za), so only TWO adjoint streams need GEMMs: the primal
//     (zbar0) and the order-1 stream (zbar1); the latter rides along with the input-gradient chain (P3..P5).
//   * dW_l = t^T zbar0 + a1^T zbar1 + (a2~ + ag~)^T za accumulates over ALL tiles of the CTA in persistent TMEM
//     columns: each term is an M = 128 GEMM over the 128 points whose A operand is the transposed view of the
//     activation tile started at the band of that stream ("band-shifted accumulate"): rows 0..31 of D are the
//     wanted 32 x N block, rows >= 32 are never read.  db_l = sum_p zbar0 is kept in registers.
//   phases  P0 z0,z1_0 | P1 z1,z1_1,z2~_1 | P2 u,u1,u2~ | P3 aa2, ab1_2 (+dW2: a1) | P4 aa1, ab1_1 (+dW1: a1) |
//           P5 g (+dW0: v) | P6 zg~_0 | P7 zg~_1 | P8 ug~ | P9 ab_2, aa2 (+dW2: t, c) | P10 ab_1, aa1 (+dW1: t, c) |
//           P11 (dW0: x_hi, x_lo, g~).   E_k = epilogue between P_{k-1} and P_k.
//   parity  rtol 1e-2 (BASELINE.json, bf16 GEMM path); measured against the oracle in tests/test_gpu_tensor.py.
#include "mlp_thread.cuh"
#include "residual_common.cuh"
#include "umma.cuh"

namespace pdeip {

int mlp_residual_accumulate_fp32(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st);

namespace tc {

using namespace umma;

constexpr int H = 32;
constexpr int OP = 48;  // output width 40 padded to a multiple of 16

// ---- TMEM column map: [0, 112) persistent dW accumulators, then NS slots of 200 columns -----------------------
constexpr uint32_t C_DW2 = 0, C_DW1 = 48, C_DW0 = 80, C_SLOT0 = 112, SLOT_COLS = 200;
// slot-relative
constexpr uint32_t C_S1P1 = 0, C_S1P2 = 16;  // parked s1 of hidden layers 1 / 2 (packed bf16 pairs, 16 columns)
constexpr uint32_t C_R = 32;                  // 144 transient columns
constexpr uint32_t C_Z0 = C_R + 0, C_Z10 = C_R + 32;                       // P0
constexpr uint32_t C_Z1 = C_R + 64, C_Z11 = C_R + 96, C_Z21 = C_R + 0;      // P1
constexpr uint32_t C_U = C_R + 0, C_U1 = C_R + 48, C_U2 = C_R + 96;         // P2
constexpr uint32_t C_S0P = 176;                                            // parked s0p (24 columns)
constexpr uint32_t C_AA2 = C_R + 0, C_AB12 = C_R + 32, C_PZ2 = C_R + 128;   // P3 / E4
constexpr uint32_t C_AA1 = C_R + 64, C_AB11 = C_R + 96, C_PZ1 = C_R + 0;    // P4 / E5
constexpr uint32_t C_G = C_R + 16;                                         // P5
constexpr uint32_t C_ZG0 = C_R + 64, C_ZG1 = C_R + 96, C_UG = C_R + 48;     // P6, P7, P8
constexpr uint32_t C_AB2 = C_R + 16, C_AA2R = C_R + 96;                     // P9
constexpr uint32_t C_AB1 = C_R + 48, C_AA1R = C_R + 96;                     // P10

// ---- shared memory ---------------------------------------------------------------------------------------------
// Z tile chunk map (24 chunks of 8 columns)
constexpr int ZC_ZA2 = 0, ZC_ZA1 = 6, ZC_ZA0 = 10, ZC_TA = 14, ZC_TB = 20;
// A tile chunk map (12 chunks): t | a1 (later the g-stream operand ag~) | a2~ (later c = a2~ + ag~)
constexpr int AC_T = 0, AC_A1 = 4, AC_C = 8;

template <int DP, int NS>
struct Cfg {
  static constexpr int XC = DP / 8;                  // chunks per input band
  static constexpr int NX = 4 * XC;                  // x_hi | x_lo | v | g~
  static constexpr int KX = 2 * DP;                  // K of the x GEMM ([x_hi | x_lo] against [W0; W0])
  static constexpr int KV = DP < 16 ? 16 : DP;       // K of the v / g~ GEMMs, N of the g GEMM
  static constexpr int XC_HI = 0, XC_LO = XC, XC_V = 2 * XC, XC_G = 3 * XC;
  static constexpr uint32_t RG_X = NX * 128, RG_A = 12 * 128, RG_Z = 24 * 128;
  static constexpr uint32_t SZ_X = 16 * RG_X, SZ_A = 16 * RG_A, SZ_Z = 16 * RG_Z;
  static constexpr uint32_t O_X = 0, O_A1 = O_X + SZ_X, O_A2 = O_A1 + SZ_A, O_Z = O_A2 + SZ_A, SLOT = O_Z + SZ_Z;
  // weights (rows = output units, columns = input units), hi and lo halves
  static constexpr uint32_t RG_T0X = KX / 8 * 128, SZ_T0X = 4 * RG_T0X;
  static constexpr uint32_t RG_T0V = KV / 8 * 128, SZ_T0V = 4 * RG_T0V;
  static constexpr uint32_t RG_T1 = 512, SZ_T1 = 4 * RG_T1, RG_T2 = 512, SZ_T2 = 6 * RG_T2;
  static constexpr uint32_t O_T0XH = NS * SLOT, O_T0XL = O_T0XH + SZ_T0X, O_T0VH = O_T0XL + SZ_T0X,
                            O_T0VL = O_T0VH + SZ_T0V, O_T1H = O_T0VL + SZ_T0V, O_T1L = O_T1H + SZ_T1,
                            O_T2H = O_T1L + SZ_T1, O_T2L = O_T2H + SZ_T2, O_BIAS = O_T2L + SZ_T2;
  static constexpr uint32_t TP_BYTES = NS == 2 ? 2048 : 4096;
  static constexpr uint32_t O_TRUE = O_BIAS + 512, O_MISC = O_TRUE + TP_BYTES, TOTAL = O_MISC + 128;
  // prefetched input items per thread: x chunks, v chunks, (stored true gradient chunks), dealt round-robin to sub
  static constexpr int NI = (3 * XC + 3) / 4;
};

constexpr int kEpiThreads = 512;            // 16 epilogue warps: thread = (point row, 8-unit chunk)
constexpr int kThreads = kEpiThreads + 32;  // + one MMA-issuing warp

#ifdef PDEIP_TC_PROBE
#define TC_PROBE(id, unit0, arr)                                                              \
  do {                                                                                        \
    if (probe_on) {                                                                           \
      for (int _i = 0; _i < 8; ++_i) probe[((id) * 128 + row) * 48 + (unit0) + _i] = (arr)[_i]; \
    }                                                                                         \
  } while (0)
#else
#define TC_PROBE(id, unit0, arr) do { } while (0)
#endif

// epilogue side: operands written -> visible to the tensor core; TMEM accesses ordered; signal the MMA warp
__device__ __forceinline__ void epi_arrive(int slot) {
  fence_async_smem();
  fence_before_sync();
  asm volatile("bar.arrive %0, %1;" ::"r"(1 + slot), "n"(kThreads) : "memory");
}
// MMA warp: wait until every epilogue thread has arrived
__device__ __forceinline__ void mma_wait_operands(int slot) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(kThreads) : "memory");
  fence_after_sync();
}

// ---- descriptors as (lo, hi) 32-bit halves; advancing = one add on lo (addresses stay below 256 KB) ----------
struct Desc {
  uint32_t lo, hi;
};
__device__ __forceinline__ Desc mk_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  Desc d;
  d.lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  d.hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);  // descriptor version 1 (bit 46)
  return d;
}
__device__ __forceinline__ void mma(uint32_t d_tmem, Desc a, uint32_t a_adv, Desc b, uint32_t b_adv, uint32_t idesc,
                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a.lo + (a_adv >> 4)), "r"(a.hi), "r"(b.lo + (b_adv >> 4)), "r"(b.hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// NG independent forward GEMMs D_g[128 x N] = A_g(K-major, K columns) * W^T(K-major tile [N][K]), hi + lo halves of
// the weights, issued round-robin over g.  a_off[g] / d[g]: byte offset of the A band / TMEM column of GEMM g.
template <int K, int N, int NG>
__device__ __forceinline__ void mm_fwd(const uint32_t (&d)[NG], Desc a, const uint32_t (&a_off)[NG], Desc w_hi, Desc w_lo) {
  constexpr uint32_t idesc = make_idesc(N, 0, 0);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g) mma(d[g], a, a_off[g] + k * 16, w_hi, k * 16, idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g) mma(d[g], a, a_off[g] + k * 16, w_lo, k * 16, idesc, 1u);
}
// NG backward GEMMs D_g[128 x N] = A_g(K-major, K columns) * W (transposed view of the weight tile, rows = K)
template <int K, int N, int NG>
__device__ __forceinline__ void mm_bwd(const uint32_t (&d)[NG], Desc a, const uint32_t (&a_off)[NG], Desc w_hi_m,
                                       Desc w_lo_m, uint32_t w_rg) {
  constexpr uint32_t idesc = make_idesc(N, 0, 1);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g) mma(d[g], a, a_off[g] + k * 16, w_hi_m, (k >> 3) * w_rg, idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g) mma(d[g], a, a_off[g] + k * 16, w_lo_m, (k >> 3) * w_rg, idesc, 1u);
}
// band-shifted batch-reduced outer product: D[m][n] += sum_points A[p][a_col0 + m] * B[p][b_col0 + n]
// (both operands transposed views; a_off / b_off = byte offsets of the first chunk of the band)
template <int N>
__device__ __forceinline__ void mm_outer(uint32_t d, Desc a_m, uint32_t a_off, uint32_t a_rg, Desc b_m, uint32_t b_off,
                                         uint32_t b_rg, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(N, 1, 1);
#pragma unroll
  for (int k = 0; k < 128; k += 16)
    mma(d, a_m, a_off + (k >> 3) * a_rg, b_m, b_off + (k >> 3) * b_rg, idesc, k > 0 ? 1u : accumulate);
}

// MUFU tanh (max relative error ~2^-11, below the bf16 rounding applied to every activation operand)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed bf16 helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xffff0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xffff0000u);
}
__device__ __forceinline__ void load_chunk(const uint8_t* tile, uint32_t off, float (&v)[8]) {
  unpack8(*reinterpret_cast<const uint4*>(tile + off), v);
}
// park 8 floats as 4 TMEM columns of packed bf16 pairs / read them back
__device__ __forceinline__ void tmem_park8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};\n\t"
      "tcgen05.wait::st.sync.aligned;\n" ::"r"(taddr),
      "r"(pack2(v[0], v[1])), "r"(pack2(v[2], v[3])), "r"(pack2(v[4], v[5])), "r"(pack2(v[6], v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_ld4_raw(uint32_t taddr, uint4& q) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n\t"
               : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8_raw(uint32_t taddr, float (&v)[8]) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr)
               : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int DP, int NS>
__global__ void __launch_bounds__(kThreads, 1) mlp_residual_tc_kernel(const ResidualArgs a, int* status) {
  using S = Cfg<DP, NS>;
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_mma_warp = warp == kEpiThreads / 32;
  const int q = warp & 3;         // TMEM lane quadrant of this warp
  const int sub = warp >> 2;      // which 8-unit chunk(s) this thread owns (epilogue warps: 0..3)
  const int row = q * 32 + lane;  // point within the tile == TMEM lane == operand tile row
  const int d = a.d;
  const MlpShape<H> sh{d, 2};
  const int P = sh.num_params();
  float* bias_s = reinterpret_cast<float*>(sm + S::O_BIAS);   // b0[32] b1[32] b2[48]
  float* tp = reinterpret_cast<float*>(sm + S::O_TRUE);
  uint64_t* mbar_p = reinterpret_cast<uint64_t*>(sm + S::O_MISC);  // [slot]
  uint32_t* tmem_p = reinterpret_cast<uint32_t*>(sm + S::O_MISC + 64);

  const int64_t n_tiles = (a.n_points + 127) / 128;
  if ((int64_t)blockIdx.x * NS >= n_tiles) return;  // nothing to do for this CTA (uniform)

  // ---- one-time set-up: TMEM, mbarriers, zeroed operand tiles, split weights in core-matrix layout ------------
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_p), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(mbar_p + s), 1);
    fence_mbar_init();
  }
  // every operand byte is finite from the start: zero-weight columns multiply whatever the neighbouring band holds
  for (uint32_t o = tid * 16; o < S::O_BIAS; o += kThreads * 16) *reinterpret_cast<uint4*>(sm + o) = make_uint4(0, 0, 0, 0);
  __syncthreads();
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
    for (int idx = tid; idx < 32 * S::KX; idx += kThreads) {  // T0X[r = out][c] = W0[c mod DP][r]  ([W0; W0])
      const int r = idx / S::KX, c = idx % S::KX, ci = c % DP;
      put(S::O_T0XH, S::O_T0XL, S::RG_T0X, r, c, ci < d ? W0[ci * H + r] : 0.f);
    }
    for (int idx = tid; idx < 32 * S::KV; idx += kThreads) {  // T0V[r = out][c = in] = W0[c][r], columns >= d zero
      const int r = idx / S::KV, c = idx % S::KV;
      put(S::O_T0VH, S::O_T0VL, S::RG_T0V, r, c, c < d ? W0[c * H + r] : 0.f);
    }
    for (int idx = tid; idx < 32 * 32; idx += kThreads) {
      const int r = idx / 32, c = idx % 32;
      put(S::O_T1H, S::O_T1L, S::RG_T1, r, c, W1[c * H + r]);
    }
    for (int idx = tid; idx < OP * 32; idx += kThreads) {  // T2[r = out (48)][c = in] = W2[c][r], rows >= 40 zero
      const int r = idx / 32, c = idx % 32;
      put(S::O_T2H, S::O_T2L, S::RG_T2, r, c, r < kOut ? W2[c * kOut + r] : 0.f);
    }
    for (int j = tid; j < 32; j += kThreads) {
      bias_s[j] = a.params[sh.b_off(0) + j];
      bias_s[32 + j] = a.params[sh.b_off(1) + j];
    }
    for (int j = tid; j < OP; j += kThreads) bias_s[64 + j] = j < kOut ? a.params[sh.b_off(2) + j] : 0.f;
    int ntg = 0;
    if (a.tg.kind == PDEIP_DRIFT_LINEAR) ntg = d * d;
    else if (a.tg.kind == PDEIP_DRIFT_GMM) ntg = a.tg.n_gaussian * d;
    for (int i = tid; i < ntg; i += kThreads) tp[i] = a.tg.params[i];
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t TB = *tmem_p;
  const int64_t tile_stride = (int64_t)gridDim.x * NS;

  // ==========================================================================================================
  // MMA warp: one lane issues every tcgen05.mma, phase by phase, slot by slot
  // ==========================================================================================================
  if (is_mma_warp) {
    // weights: K-major views (LBO = 128: next 8 columns, SBO = row-group bytes) and transposed views
    const Desc T0XHK = mk_desc(smem_u32(sm + S::O_T0XH), 128, S::RG_T0X), T0XLK = mk_desc(smem_u32(sm + S::O_T0XL), 128, S::RG_T0X);
    const Desc T0VHK = mk_desc(smem_u32(sm + S::O_T0VH), 128, S::RG_T0V), T0VLK = mk_desc(smem_u32(sm + S::O_T0VL), 128, S::RG_T0V);
    const Desc T1HK = mk_desc(smem_u32(sm + S::O_T1H), 128, S::RG_T1), T1LK = mk_desc(smem_u32(sm + S::O_T1L), 128, S::RG_T1);
    const Desc T2HK = mk_desc(smem_u32(sm + S::O_T2H), 128, S::RG_T2), T2LK = mk_desc(smem_u32(sm + S::O_T2L), 128, S::RG_T2);
    const Desc T0VHM = mk_desc(smem_u32(sm + S::O_T0VH), S::RG_T0V, 128), T0VLM = mk_desc(smem_u32(sm + S::O_T0VL), S::RG_T0V, 128);
    const Desc T1HM = mk_desc(smem_u32(sm + S::O_T1H), S::RG_T1, 128), T1LM = mk_desc(smem_u32(sm + S::O_T1L), S::RG_T1, 128);
    const Desc T2HM = mk_desc(smem_u32(sm + S::O_T2H), S::RG_T2, 128), T2LM = mk_desc(smem_u32(sm + S::O_T2L), S::RG_T2, 128);
    constexpr uint32_t CH = 128;  // bytes per chunk (8 operand columns)
    uint32_t dw_started = 0;      // becomes 1 after the first tile's P3..P5 background GEMMs
#pragma unroll 1
    for (int64_t base = (int64_t)blockIdx.x * NS; base < n_tiles; base += tile_stride) {
#pragma unroll 1
      for (int ph = 0; ph < 12; ++ph) {
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
          const uint32_t sb = smem_u32(sm) + (uint32_t)s * S::SLOT;
          const uint32_t TS = TB + C_SLOT0 + (uint32_t)s * SLOT_COLS;
          const uint32_t mb = smem_u32(mbar_p + s);
          const Desc XK = mk_desc(sb + S::O_X, 128, S::RG_X), A1K = mk_desc(sb + S::O_A1, 128, S::RG_A),
                     A2K = mk_desc(sb + S::O_A2, 128, S::RG_A), ZK = mk_desc(sb + S::O_Z, 128, S::RG_Z);
          const Desc XM = mk_desc(sb + S::O_X, S::RG_X, 128), A1M = mk_desc(sb + S::O_A1, S::RG_A, 128),
                     A2M = mk_desc(sb + S::O_A2, S::RG_A, 128), ZM = mk_desc(sb + S::O_Z, S::RG_Z, 128);
          const uint32_t acc_first = (dw_started | (uint32_t)s) ? 1u : 0u;
          mma_wait_operands(s);
          if (elect_one()) {
            switch (ph) {
              case 0: {  // z0 = [x_hi | x_lo] [W0; W0],  z1_0 = v W0
                mm_fwd<S::KX, 32, 1>({TS + C_Z0}, XK, {S::XC_HI * CH}, T0XHK, T0XLK);
                mm_fwd<S::KV, 32, 1>({TS + C_Z10}, XK, {S::XC_V * CH}, T0VHK, T0VLK);
                commit(mb);
              } break;
              case 1: {  // z1, z1_1, z2~_1
                mm_fwd<32, 32, 3>({TS + C_Z1, TS + C_Z11, TS + C_Z21}, A1K, {AC_T * CH, AC_A1 * CH, AC_C * CH}, T1HK, T1LK);
                commit(mb);
              } break;
              case 2: {  // u, u1, u2~
                mm_fwd<32, OP, 3>({TS + C_U, TS + C_U1, TS + C_U2}, A2K, {AC_T * CH, AC_A1 * CH, AC_C * CH}, T2HK, T2LK);
                commit(mb);
              } break;
              case 3: {  // aa2 = za2 W2^T, ab1_2 = s1v W2^T;  dW2 += a1_2^T s1v
                mm_bwd<OP, 32, 2>({TS + C_AA2, TS + C_AB12}, ZK, {ZC_ZA2 * CH, ZC_TA * CH}, T2HM, T2LM, S::RG_T2);
                commit(mb);
                mm_outer<OP>(TB + C_DW2, A2M, AC_A1 * CH, S::RG_A, ZM, ZC_TA * CH, S::RG_Z, acc_first);
              } break;
              case 4: {  // aa1 = za1 W1^T, ab1_1 = zbar1' W1^T;  dW1 += a1_1^T zbar1'
                mm_bwd<32, 32, 2>({TS + C_AA1, TS + C_AB11}, ZK, {ZC_ZA1 * CH, ZC_TB * CH}, T1HM, T1LM, S::RG_T1);
                commit(mb);
                mm_outer<32>(TB + C_DW1, A1M, AC_A1 * CH, S::RG_A, ZM, ZC_TB * CH, S::RG_Z, acc_first);
              } break;
              case 5: {  // g = za0 W0^T;  dW0 += v^T zbar1''
                mm_bwd<32, S::KV, 1>({TS + C_G}, ZK, {ZC_ZA0 * CH}, T0VHM, T0VLM, S::RG_T0V);
                commit(mb);
                mm_outer<32>(TB + C_DW0, XM, S::XC_V * CH, S::RG_X, ZM, ZC_TA * CH, S::RG_Z, acc_first);
              } break;
              case 6: {  // zg~_0 = g~ W0
                mm_fwd<S::KV, 32, 1>({TS + C_ZG0}, XK, {S::XC_G * CH}, T0VHK, T0VLK);
                commit(mb);
              } break;
              case 7: {  // zg~_1 = ag~_1 W1  (operand in the a1 band)
                mm_fwd<32, 32, 1>({TS + C_ZG1}, A1K, {AC_A1 * CH}, T1HK, T1LK);
                commit(mb);
              } break;
              case 8: {  // ug~ = ag~_2 W2
                mm_fwd<32, OP, 1>({TS + C_UG}, A2K, {AC_A1 * CH}, T2HK, T2LK);
                commit(mb);
              } break;
              case 9: {  // ab_2 = s0 W2^T, aa2 again;  dW2 += t2^T s0 + c_2^T za2
                mm_bwd<OP, 32, 2>({TS + C_AB2, TS + C_AA2R}, ZK, {ZC_TA * CH, ZC_ZA2 * CH}, T2HM, T2LM, S::RG_T2);
                commit(mb);
                mm_outer<OP>(TB + C_DW2, A2M, AC_T * CH, S::RG_A, ZM, ZC_TA * CH, S::RG_Z, 1u);
                mm_outer<OP>(TB + C_DW2, A2M, AC_C * CH, S::RG_A, ZM, ZC_ZA2 * CH, S::RG_Z, 1u);
              } break;
              case 10: {  // ab_1 = zbar0' W1^T, aa1 again;  dW1 += t1^T zbar0' + c_1^T za1
                mm_bwd<32, 32, 2>({TS + C_AB1, TS + C_AA1R}, ZK, {ZC_TB * CH, ZC_ZA1 * CH}, T1HM, T1LM, S::RG_T1);
                commit(mb);
                mm_outer<32>(TB + C_DW1, A1M, AC_T * CH, S::RG_A, ZM, ZC_TB * CH, S::RG_Z, 1u);
                mm_outer<32>(TB + C_DW1, A1M, AC_C * CH, S::RG_A, ZM, ZC_ZA1 * CH, S::RG_Z, 1u);
              } break;
              default: {  // dW0 += x_hi^T zbar0'' + x_lo^T zbar0'' + g~^T za0
                mm_outer<32>(TB + C_DW0, XM, S::XC_HI * CH, S::RG_X, ZM, ZC_TA * CH, S::RG_Z, 1u);
                mm_outer<32>(TB + C_DW0, XM, S::XC_LO * CH, S::RG_X, ZM, ZC_TA * CH, S::RG_Z, 1u);
                mm_outer<32>(TB + C_DW0, XM, S::XC_G * CH, S::RG_X, ZM, ZC_ZA0 * CH, S::RG_Z, 1u);
                commit(mb);  // the next tile's E0 overwrites the x | v bands
              } break;
            }
          }
          __syncwarp();
        }
      }
      dw_started = 1u;
    }
  } else {
    // ========================================================================================================
    // epilogue warps
    // ========================================================================================================
    int* const status_w = status;
    bool ok = true;
    uint32_t par[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) par[s] = 0;
#ifdef PDEIP_TC_PROBE
    float* probe = reinterpret_cast<float*>(status) + 64;
#endif
    // bias-gradient sums over this thread's row (chunk `sub` of the hidden layers; chunks sub, sub + 4 of the output)
    float db0a[8], db1a[8], db2a[8], db2b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { db0a[i] = 0.f; db1a[i] = 0.f; db2a[i] = 0.f; db2b[i] = 0.f; }
    float sum_g2 = 0.f, sum_gt2 = 0.f, sum_gd2 = 0.f, sum_d1 = 0.f, sum_d2 = 0.f;
    const float gamma = a.coef;
    const int dimw = 2 * d + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? d : 0);  // floats per point
    // prefetched input chunks: item j = sub + 4 i;  j in [0, XC): x chunk j;  [XC, 2 XC): v;  [2 XC, 3 XC): stored gt
    float xin[NS][S::NI][8];
    auto load_inputs = [&](int s, int64_t t) {
#pragma unroll
      for (int i = 0; i < S::NI; ++i) {
        const int j = sub + 4 * i;
        const int band = j / S::XC, cg = j % S::XC;
        const int64_t pp = t * 128 + row;
        const bool ok_p = j < 3 * S::XC && t < n_tiles && pp < a.n_points && (band < 2 || a.tg.kind == PDEIP_DRIFT_IN_POINTS);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int u = cg * 8 + e;
          xin[s][i][e] = (ok_p && u < d) ? __ldg(a.points + elem_index(a.layout, pp, band * d + u, a.n_points, dimw)) : 0.f;
        }
      }
    };
#pragma unroll
    for (int s = 0; s < NS; ++s) load_inputs(s, (int64_t)blockIdx.x * NS + s);

    bool first = true;
#pragma unroll 1
    for (int64_t base = (int64_t)blockIdx.x * NS; base < n_tiles; base += tile_stride) {
#pragma unroll 1
      for (int ph = 0; ph < 12; ++ph) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          uint8_t* const X = sm + (uint32_t)s * S::SLOT + S::O_X;
          uint8_t* const A1 = sm + (uint32_t)s * S::SLOT + S::O_A1;
          uint8_t* const A2 = sm + (uint32_t)s * S::SLOT + S::O_A2;
          uint8_t* const Z = sm + (uint32_t)s * S::SLOT + S::O_Z;
          const uint32_t LA = TB + ((uint32_t)(q * 32) << 16) + C_SLOT0 + (uint32_t)s * SLOT_COLS;
          const int64_t tile = base + s;
          const int64_t p = tile * 128 + row;
          const bool valid = tile < n_tiles && p < a.n_points;
          const float mk = valid ? 1.f : 0.f;
#ifdef PDEIP_TC_PROBE
          const bool probe_on = blockIdx.x == 0 && tile == 0;
#endif
          // wait for the GEMMs of the previous phase of this slot (P11 of the previous tile before E0)
          if (!(first && ph == 0)) {
            if (ok && !mbar_wait(smem_u32(mbar_p + s), par[s])) {
              ok = false;
              atomicExch(status_w, 1);
            }
            par[s] ^= 1u;
            fence_after_sync();
          }
          switch (ph) {
            case 0: {  // E0: x (hi + lo) and v bands of this tile; prefetch registers are refilled in E1
#pragma unroll
              for (int i = 0; i < S::NI; ++i) {
                const int j = sub + 4 * i;
                const int band = j / S::XC, cg = j % S::XC;
                if (band == 0) {
                  float hi[8], lo[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    hi[e] = __bfloat162float(__float2bfloat16_rn(xin[s][i][e]));
                    lo[e] = xin[s][i][e] - hi[e];
                  }
                  store_chunk(X, chunk_off(row, S::XC_HI + cg, S::RG_X), hi);
                  store_chunk(X, chunk_off(row, S::XC_LO + cg, S::RG_X), lo);
                } else if (band == 1) {
                  store_chunk(X, chunk_off(row, S::XC_V + cg, S::RG_X), xin[s][i]);
                }
              }
            } break;
            case 1: {  // E1: t1, a1_1, a2~_1 = 4 t a1 z1
              float z0[8], z1[8], t[8], s1[8], q1[8], q2[8];
              tmem_ld8x2(LA + C_Z0 + 8 * sub, LA + C_Z10 + 8 * sub, z0, z1);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                t[i] = tanh_fast(z0[i] + bias_s[sub * 8 + i]);
                s1[i] = 1.f - t[i] * t[i];
                q1[i] = s1[i] * z1[i];
                q2[i] = 4.f * t[i] * q1[i] * z1[i];
              }
              store_chunk(A1, chunk_off(row, AC_T + sub, S::RG_A), t);
              store_chunk(A1, chunk_off(row, AC_A1 + sub, S::RG_A), q1);
              store_chunk(A1, chunk_off(row, AC_C + sub, S::RG_A), q2);
              tmem_park8(LA + C_S1P1 + 4 * sub, s1);
              TC_PROBE(0, sub * 8, t);
              TC_PROBE(1, sub * 8, q1);
              TC_PROBE(2, sub * 8, q2);
            } break;
            case 2: {  // E2: t2, a1_2, a2~_2 = s1 z2~ + 4 t a1 z1
              float z0[8], z1[8], z2[8], t[8], s1[8], q1[8], q2[8];
              tmem_ld8x3(LA + C_Z1 + 8 * sub, LA + C_Z11 + 8 * sub, LA + C_Z21 + 8 * sub, z0, z1, z2);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                t[i] = tanh_fast(z0[i] + bias_s[32 + sub * 8 + i]);
                s1[i] = 1.f - t[i] * t[i];
                q1[i] = s1[i] * z1[i];
                q2[i] = fmaf(s1[i], z2[i], 4.f * t[i] * q1[i] * z1[i]);
              }
              store_chunk(A2, chunk_off(row, AC_T + sub, S::RG_A), t);
              store_chunk(A2, chunk_off(row, AC_A1 + sub, S::RG_A), q1);
              store_chunk(A2, chunk_off(row, AC_C + sub, S::RG_A), q2);
              tmem_park8(LA + C_S1P2 + 4 * sub, s1);
              TC_PROBE(3, sub * 8, t);
              TC_PROBE(4, sub * 8, q1);
              TC_PROBE(5, sub * 8, q2);
            } break;
            case 3: {  // E3: za2 = 2u, s1v, s0p, D_v V, D_v^2 V
              float d1 = 0.f, d2a = 0.f, d2b = 0.f;
#pragma unroll 1
              for (int cg = sub; cg < OP / 8; cg += 4) {
                float u[8], u1[8], u2[8], za[8], sv[8], sp[8];
                tmem_ld8x3(LA + C_U + 8 * cg, LA + C_U1 + 8 * cg, LA + C_U2 + 8 * cg, u, u1, u2);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float uu = u[i] + bias_s[64 + cg * 8 + i];
                  d1 = fmaf(uu, u1[i], d1);
                  d2a = fmaf(u1[i], u1[i], d2a);
                  d2b = fmaf(uu, u2[i], d2b);
                  za[i] = 2.f * mk * uu;
                  sv[i] = mk * fmaf(-8.f, u1[i], 4.f * gamma * uu);
                  sp[i] = mk * fmaf(2.f, u2[i], 4.f * gamma * u1[i]);
                }
                store_chunk(Z, chunk_off(row, ZC_ZA2 + cg, S::RG_Z), za);
                store_chunk(Z, chunk_off(row, ZC_TA + cg, S::RG_Z), sv);
                tmem_park8(LA + C_S0P + 4 * cg, sp);
                TC_PROBE(6, cg * 8, za);
                TC_PROBE(7, cg * 8, sv);
                TC_PROBE(8, cg * 8, sp);
              }
              // this thread's share of D_v V = 2 u.u1 and D_v^2 V = 2 (u1.u1 + u.u2), u2 = -u2~ / 2
              sum_d1 += mk * 2.f * d1;
              sum_d2 += mk * (2.f * d2a - d2b);
            } break;
            case 4:    // E4: za1, zbar1', pz2  (hidden layer 2)
            case 5: {  // E5: za0, zbar1'', pz1 (hidden layer 1)
              const bool l2 = ph == 4;
              const uint8_t* At = l2 ? A2 : A1;
              float aa[8], ab1[8], s1[8], t[8], a1[8], za[8], zb[8], pz[8];
              uint4 s1q;
              tmem_ld8_raw(LA + (l2 ? C_AA2 : C_AA1) + 8 * sub, aa);
              tmem_ld8_raw(LA + (l2 ? C_AB12 : C_AB11) + 8 * sub, ab1);
              tmem_ld4_raw(LA + (l2 ? C_S1P2 : C_S1P1) + 4 * sub, s1q);
              load_chunk(At, chunk_off(row, AC_T + sub, S::RG_A), t);
              load_chunk(At, chunk_off(row, AC_A1 + sub, S::RG_A), a1);
              tmem_wait_ld();
              unpack8(s1q, s1);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                za[i] = aa[i] * s1[i];
                const float qq = aa[i] * a1[i];
                zb[i] = fmaf(s1[i], ab1[i], 8.f * t[i] * qq);
                pz[i] = fmaf(-2.f * t[i] * ab1[i], a1[i], 4.f * qq * a1[i]);
              }
              store_chunk(Z, chunk_off(row, (l2 ? ZC_ZA1 : ZC_ZA0) + sub, S::RG_Z), za);
              store_chunk(Z, chunk_off(row, (l2 ? ZC_TB : ZC_TA) + sub, S::RG_Z), zb);
              tmem_park8(LA + (l2 ? C_PZ2 : C_PZ1) + 4 * sub, pz);
              TC_PROBE(l2 ? 9 : 12, sub * 8, za);
              TC_PROBE(l2 ? 10 : 13, sub * 8, zb);
              TC_PROBE(l2 ? 11 : 14, sub * 8, pz);
            } break;
            case 6: {  // E6: g~ = 2 g band; |g|^2, |g_true|^2, |g_true - g|^2
              // item j of this thread: band 2 (stored true gradient) chunks own the sums when the gradient is stored;
              // otherwise the x-chunk owners do (they hold x for the inline true gradient)
#pragma unroll
              for (int i = 0; i < S::NI; ++i) {
                const int j = sub + 4 * i;
                const int band = j / S::XC, cg = j % S::XC;
                if (j < 3 * S::XC && band == 0) {  // writes the g~ chunk
                  float gv[8], g2v[8];
                  tmem_ld8(LA + C_G + 8 * cg, gv);
#pragma unroll
                  for (int e = 0; e < 8; ++e) g2v[e] = 2.f * gv[e];
                  store_chunk(X, chunk_off(row, S::XC_G + cg, S::RG_X), g2v);
                  TC_PROBE(15, cg * 8, gv);
                  if (a.tg.kind != PDEIP_DRIFT_IN_POINTS) {
                    float gt[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) gt[e] = 0.f;
                    if (a.tg.kind == PDEIP_DRIFT_LINEAR || a.tg.kind == PDEIP_DRIFT_GMM) {
                      float x[DP];
#pragma unroll
                      for (int u = 0; u < DP; ++u)
                        x[u] = (valid && u < d) ? __ldg(a.points + elem_index(a.layout, p, u, a.n_points, dimw)) : 0.f;
                      if (a.tg.kind == PDEIP_DRIFT_LINEAR) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                          const int u = cg * 8 + e;
                          if (u < d) {
                            float acc = 0.f;
#pragma unroll
                            for (int k = 0; k < DP; ++k)
                              if (k < d) acc = fmaf(tp[u * d + k], x[k], acc);
                            gt[e] = acc;
                          }
                        }
                      } else {
                        float m = -INFINITY;
                        for (int k = 0; k < a.tg.n_gaussian; ++k) {
                          float s2 = 0.f;
#pragma unroll
                          for (int u = 0; u < DP; ++u)
                            if (u < d) {
                              const float r = x[u] - tp[k * d + u];
                              s2 = fmaf(r, r, s2);
                            }
                          m = fmaxf(m, -0.5f * a.tg.inv_sigma2 * s2);
                        }
                        float se = 0.f, wm[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) wm[e] = 0.f;
                        for (int k = 0; k < a.tg.n_gaussian; ++k) {
                          float s2 = 0.f;
#pragma unroll
                          for (int u = 0; u < DP; ++u)
                            if (u < d) {
                              const float r = x[u] - tp[k * d + u];
                              s2 = fmaf(r, r, s2);
                            }
                          const float ek = __expf(-0.5f * a.tg.inv_sigma2 * s2 - m);
                          se += ek;
#pragma unroll
                          for (int e = 0; e < 8; ++e)
                            if (cg * 8 + e < d) wm[e] = fmaf(ek, tp[k * d + cg * 8 + e], wm[e]);
                        }
                        const float inv = 1.f / se;
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                          if (cg * 8 + e < d) gt[e] = (x[cg * 8 + e] - wm[e] * inv) * a.tg.inv_sigma2;
                      }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                      sum_g2 = fmaf(mk * gv[e], gv[e], sum_g2);
                      sum_gt2 = fmaf(mk * gt[e], gt[e], sum_gt2);
                      sum_gd2 = fmaf(mk * (gt[e] - gv[e]), gt[e] - gv[e], sum_gd2);
                    }
                  }
                } else if (j < 3 * S::XC && band == 2 && a.tg.kind == PDEIP_DRIFT_IN_POINTS) {
                  float gv[8];
                  tmem_ld8(LA + C_G + 8 * cg, gv);
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float gt = xin[s][i][e];
                    sum_g2 = fmaf(mk * gv[e], gv[e], sum_g2);
                    sum_gt2 = fmaf(mk * gt, gt, sum_gt2);
                    sum_gd2 = fmaf(mk * (gt - gv[e]), gt - gv[e], sum_gd2);
                  }
                }
              }
            } break;
            case 7:    // E7: ag~_1 = s1_1 zg~_0 -> a1 band of A1;  c_1 = a2~_1 + ag~_1
            case 8: {  // E8: ag~_2 = s1_2 zg~_1 -> a1 band of A2;  c_2
              const bool l1 = ph == 7;
              uint8_t* At = l1 ? A1 : A2;
              float zg[8], s1[8], a2[8], ag[8], c[8];
              uint4 s1q;
              tmem_ld8_raw(LA + (l1 ? C_ZG0 : C_ZG1) + 8 * sub, zg);
              tmem_ld4_raw(LA + (l1 ? C_S1P1 : C_S1P2) + 4 * sub, s1q);
              load_chunk(At, chunk_off(row, AC_C + sub, S::RG_A), a2);
              tmem_wait_ld();
              unpack8(s1q, s1);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                ag[i] = s1[i] * zg[i];
                c[i] = a2[i] + ag[i];
              }
              store_chunk(At, chunk_off(row, AC_A1 + sub, S::RG_A), ag);
              store_chunk(At, chunk_off(row, AC_C + sub, S::RG_A), c);
              TC_PROBE(l1 ? 16 : 17, sub * 8, c);
            } break;
            case 9: {  // E9: s0 = s0p + 2 ug~;  db2 += s0
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int cg = sub + 4 * h;
                if (cg < OP / 8) {
                  float ug[8], sp[8], s0[8];
                  uint4 spq;
                  tmem_ld8_raw(LA + C_UG + 8 * cg, ug);
                  tmem_ld4_raw(LA + C_S0P + 4 * cg, spq);
                  tmem_wait_ld();
                  unpack8(spq, sp);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    s0[i] = fmaf(2.f, ug[i], sp[i]);
                    if (h == 0) db2a[i] += s0[i];
                    else db2b[i] += s0[i];
                  }
                  store_chunk(Z, chunk_off(row, ZC_TA + cg, S::RG_Z), s0);
                  TC_PROBE(18, cg * 8, s0);
                }
              }
            } break;
            case 10:    // E10: zbar0' = s1_2 ab_2 + pz2 - 2 t2 aa2 c_2;  db1
            default: {  // E11: zbar0'' = s1_1 ab_1 + pz1 - 2 t1 aa1 c_1;  db0
              const bool l2 = ph == 10;
              const uint8_t* At = l2 ? A2 : A1;
              float ab[8], aa[8], s1[8], pz[8], t[8], c[8], zb[8];
              uint4 s1q, pzq;
              tmem_ld8_raw(LA + (l2 ? C_AB2 : C_AB1) + 8 * sub, ab);
              tmem_ld8_raw(LA + (l2 ? C_AA2R : C_AA1R) + 8 * sub, aa);
              tmem_ld4_raw(LA + (l2 ? C_S1P2 : C_S1P1) + 4 * sub, s1q);
              tmem_ld4_raw(LA + (l2 ? C_PZ2 : C_PZ1) + 4 * sub, pzq);
              load_chunk(At, chunk_off(row, AC_T + sub, S::RG_A), t);
              load_chunk(At, chunk_off(row, AC_C + sub, S::RG_A), c);
              tmem_wait_ld();
              unpack8(s1q, s1);
              unpack8(pzq, pz);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                zb[i] = fmaf(s1[i], ab[i], fmaf(-2.f * t[i] * aa[i], c[i], pz[i]));
                if (l2) db1a[i] += zb[i];
                else db0a[i] += zb[i];
              }
              store_chunk(Z, chunk_off(row, (l2 ? ZC_TB : ZC_TA) + sub, S::RG_Z), zb);
              TC_PROBE(l2 ? 19 : 20, sub * 8, zb);
              // prefetch the next tile of this slot while its last GEMMs run
              if (!l2) load_inputs(s, tile + tile_stride);
            } break;
          }
          epi_arrive(s);
        }
      }
      first = false;
    }
    // ---- drain: wait for P11 of the last tile of every slot (covers every MMA issued before it) --------------
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (ok && !mbar_wait(smem_u32(mbar_p + s), par[s])) {
        ok = false;
        atomicExch(status_w, 1);
      }
      par[s] ^= 1u;
    }
    fence_after_sync();
    asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");

    // ---- write-out: this CTA's partial (one writer per index, fixed order -> bit-reproducible) ------------------
    float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
    float* red = reinterpret_cast<float*>(sm);  // [4 quadrants][128] bias sums + [16 warps][8] loss sums
    const float wt = a.weight;
    if (ok && q == 0) {  // TMEM lanes 0..31 = input unit of the dW blocks
      const uint32_t L0 = TB;  // quadrant 0
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cg = sub + 4 * h;
        if (cg < kOut / 8) {
          float w[8];
          tmem_ld8(L0 + C_DW2 + 8 * cg, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) part[sh.w_off(2) + lane * kOut + cg * 8 + i] += wt * w[i];
        }
      }
      {
        float w[8];
        tmem_ld8(L0 + C_DW1 + 8 * sub, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) part[sh.w_off(1) + lane * H + sub * 8 + i] += wt * w[i];
      }
      {
        float w[8];
        tmem_ld8(L0 + C_DW0 + 8 * sub, w);
        if (lane < d) {
#pragma unroll
          for (int i = 0; i < 8; ++i) part[sh.w_off(0) + lane * H + sub * 8 + i] += wt * w[i];
        }
      }
    }
    // bias gradients: sum over the 32 rows of the warp, then over the 4 quadrants through shared memory
    // red layout: [q][0..31] db0, [32..63] db1, [64..111] db2
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float r0 = warp_sum(db0a[i]), r1 = warp_sum(db1a[i]), r2 = warp_sum(db2a[i]), r3 = warp_sum(db2b[i]);
      if (lane == 0) {
        red[q * 128 + sub * 8 + i] = r0;
        red[q * 128 + 32 + sub * 8 + i] = r1;
        red[q * 128 + 64 + sub * 8 + i] = r2;
        if (sub + 4 < OP / 8) red[q * 128 + 64 + (sub + 4) * 8 + i] = r3;
      }
    }
    {
      const float sums5[5] = {sum_g2, sum_gt2, sum_gd2, sum_d1, sum_d2};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float sk = warp_sum(sums5[k]);
        if (lane == 0) red[512 + warp * 8 + k] = sk;
      }
    }
    asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
    if (ok) {
      if (tid < 112) {
        const float v = (red[tid] + red[128 + tid]) + (red[256 + tid] + red[384 + tid]);
        if (tid < 32) part[sh.b_off(0) + tid] += wt * v;
        else if (tid < 64) part[sh.b_off(1) + tid - 32] += wt * v;
        else if (tid - 64 < kOut) part[sh.b_off(2) + tid - 64] += wt * v;
      }
      if (tid == 128) {
        float t5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < kEpiThreads / 32; ++w)
          for (int k = 0; k < 5; ++k) t5[k] += red[512 + w * 8 + k];
        const float g2 = wt * t5[0], gt2 = wt * t5[1], gd2 = wt * t5[2], D1 = wt * t5[3], D2 = wt * t5[4];
        part[P + PDEIP_SUM_G2] += g2;
        part[P + PDEIP_SUM_GTRUE2] += gt2;
        part[P + PDEIP_SUM_GT] += gd2;
        part[P + PDEIP_SUM_D1] += D1;
        part[P + PDEIP_SUM_D2] += D2;
        part[P + PDEIP_SUM_LOSS] += g2 + gt2 - 2.f * D2 + 2.f * gamma * D1;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(TB, 512);
}

}  // namespace tc

static int* tensor_status_word() {
  static int* w = nullptr;
  if (!w) {
    const size_t bytes = 4096 + 24 * 128 * 48 * sizeof(float);
    if (cudaMalloc(&w, bytes) != cudaSuccess) return nullptr;
    cudaMemset(w, 0, bytes);
  }
  return w;
}

#ifdef PDEIP_HAVE_TENSOR_PATH
template <int DP, int NS>
static int launch_tc(const ResidualArgs& a, int* status, cudaStream_t st) {
  using S = tc::Cfg<DP, NS>;
  auto kern = tc::mlp_residual_tc_kernel<DP, NS>;
  PDEIP_REQUIRE(true_grad_floats(a.tg, a.d) * sizeof(float) <= S::TP_BYTES, PDEIP_ERR_UNSUPPORTED,
                "tensor path: true-gradient parameters exceed %u bytes", (unsigned)S::TP_BYTES);
  PDEIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL));
  kern<<<residual_grid(), tc::kThreads, S::TOTAL, st>>>(a, status);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

int mlp_residual_accumulate_tensor(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st) {
  // boundary sets (two launches over n points vs n*S points of the 0T set) stay on the fp32 kernel
  if (set_kind != PDEIP_SET_KFP_0T) return mlp_residual_accumulate_fp32(set_kind, a, hidden, st);
  PDEIP_REQUIRE(hidden == 32 && a.layers == 2, PDEIP_ERR_UNSUPPORTED,
                "tensor path is built for hidden_dim == 32, layers == 2 (got %d, %d)", hidden, a.layers);
  PDEIP_REQUIRE(a.d >= 1 && a.d <= 32, PDEIP_ERR_UNSUPPORTED, "tensor path supports 1 <= d <= 32 (got %d)", a.d);
  int* status = tensor_status_word();
  PDEIP_REQUIRE(status != nullptr, PDEIP_ERR_CUDA, "cannot allocate the tensor-path status word");
  if (a.d <= 8) return launch_tc<8, 2>(a, status, st);
  if (a.d <= 16) return launch_tc<16, 1>(a, status, st);
  return launch_tc<32, 1>(a, status, st);
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

// debug (PDEIP_TC_PROBE builds): per-phase epilogue values of tile 0, [probe id][row 128][unit 48] floats
extern "C" int pdeip_debug_tensor_probe(float* out, int n_floats) {
  int* status = pdeip::tensor_status_word();
  if (!status) return PDEIP_ERR_CUDA;
  if (n_floats > 24 * 128 * 48) return PDEIP_ERR_INVALID_ARG;
  if (cudaDeviceSynchronize() != cudaSuccess) return PDEIP_ERR_CUDA;
  if (cudaMemcpy(out, reinterpret_cast<float*>(status) + 64, sizeof(float) * (size_t)n_floats, cudaMemcpyDeviceToHost) !=
      cudaSuccess)
    return PDEIP_ERR_CUDA;
  return PDEIP_OK;
}
