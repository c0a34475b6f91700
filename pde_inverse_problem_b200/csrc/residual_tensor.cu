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
constexpr uint32_t C_GX = 432, C_GMS = 496;  // S3..S6: GMM true-gradient partials (acc[sub][DP], (m,se)[sub])
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
  uint32_t mb_fast;    // mbarrier: the GEMMs the next epilogue depends on have completed
  uint32_t mb_dw;      // mbarrier: the batch-reduced dW GEMMs of the phase have completed
  uint32_t par_fast, par_dw;
  int* status;
  bool ok;
};

#ifdef PDEIP_TC_TRACE
#define TC_TRACE(slot, ph) do { if (trace_on) trace[(ph) * 4 + (slot)] = clock64(); } while (0)
#else
#define TC_TRACE(slot, ph) do { } while (0)
#endif

constexpr int kEpiThreads = 512;            // 16 epilogue warps: thread = (point row, 8-unit chunk)
constexpr int kThreads = kEpiThreads + 32;  // + one MMA-issuing warp

// epilogue side: operands written -> visible to the tensor core; TMEM reads ordered; signal the MMA warp
__device__ __forceinline__ void epi_arrive() {
  fence_async_smem();
  fence_before_sync();
  asm volatile("bar.arrive 1, %0;" ::"n"(kThreads) : "memory");
}
// MMA warp: wait until every epilogue thread has arrived
__device__ __forceinline__ void mma_wait_operands() {
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
  fence_after_sync();
}

__device__ __forceinline__ void wait_fast(Ctx& c) {
  if (c.ok && !mbar_wait(c.mb_fast, c.par_fast)) {
    c.ok = false;
    atomicExch(c.status, 1);
  }
  c.par_fast ^= 1u;
  fence_after_sync();
}
__device__ __forceinline__ void wait_dw(Ctx& c) {
  if (c.ok && !mbar_wait(c.mb_dw, c.par_dw)) {
    c.ok = false;
    atomicExch(c.status, 1);
  }
  c.par_dw ^= 1u;
  fence_after_sync();
}

// ---- MMA-warp side: descriptors are built once; advancing a descriptor by `bytes` is one 32-bit add on the
// start-address field (addresses stay below 256 KB, so there is no carry out of the field) ---------------
__device__ __forceinline__ uint64_t adv(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// D[128 x N] (+)= A(K-major, K cols from a_k) * B(K-major weight tile [N][K]); hi + lo halves of the weights
template <int K, int N>
__device__ __forceinline__ void mm_fwd(uint32_t d, uint64_t a_k, uint64_t w_hi_k, uint64_t w_lo_k) {
  constexpr uint32_t idesc = make_idesc(N, 0, 0);
#pragma unroll
  for (int k = 0; k < K; k += 16) mma_bf16(d, adv(a_k, k * 16), adv(w_hi_k, k * 16), idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < K; k += 16) mma_bf16(d, adv(a_k, k * 16), adv(w_lo_k, k * 16), idesc, 1u);
}
// NG independent GEMMs (different D, different A column offsets, same weights) issued round-robin so that
// consecutive MMAs never accumulate into the same TMEM region (dependent accumulation costs a full pipeline latency)
template <int K, int N, int NG>
__device__ __forceinline__ void mm_fwd_multi(uint32_t d0, uint32_t d_stride, uint64_t a_k, uint32_t a_stride_bytes,
                                             uint64_t w_hi_k, uint64_t w_lo_k) {
  constexpr uint32_t idesc = make_idesc(N, 0, 0);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g)
      mma_bf16(d0 + g * d_stride, adv(a_k, g * a_stride_bytes + k * 16), adv(w_hi_k, k * 16), idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g)
      mma_bf16(d0 + g * d_stride, adv(a_k, g * a_stride_bytes + k * 16), adv(w_lo_k, k * 16), idesc, 1u);
}
template <int K, int N, int NG>
__device__ __forceinline__ void mm_bwd_multi(uint32_t d0, uint32_t d_stride, uint64_t a_k, uint32_t a_stride_bytes,
                                             uint64_t w_hi_m, uint64_t w_lo_m, uint32_t w_rg) {
  constexpr uint32_t idesc = make_idesc(N, 0, 1);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g)
      mma_bf16(d0 + g * d_stride, adv(a_k, g * a_stride_bytes + k * 16), adv(w_hi_m, (k >> 3) * w_rg), idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < K; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g)
      mma_bf16(d0 + g * d_stride, adv(a_k, g * a_stride_bytes + k * 16), adv(w_lo_m, (k >> 3) * w_rg), idesc, 1u);
}
// NG batch-reduced outer products sharing the A operand, round-robin over the B operands / D regions
template <int N, int NG>
__device__ __forceinline__ void mm_outer_multi(uint32_t d0, uint32_t d_stride, uint64_t a_m, uint32_t a_rg, uint64_t b_m,
                                               uint32_t b_stride_bytes, uint32_t b_rg) {
  constexpr uint32_t idesc = make_idesc(N, 1, 1);
#pragma unroll
  for (int k = 0; k < 128; k += 16)
#pragma unroll
    for (int g = 0; g < NG; ++g)
      mma_bf16(d0 + g * d_stride, adv(a_m, (k >> 3) * a_rg), adv(b_m, g * b_stride_bytes + (k >> 3) * b_rg), idesc,
               k > 0 ? 1u : 0u);
}

// D[128 x N] = A(K-major) * W^T (transposed view of the weight tile: operand rows = tile columns)
template <int K, int N>
__device__ __forceinline__ void mm_bwd(uint32_t d, uint64_t a_k, uint64_t w_hi_m, uint64_t w_lo_m, uint32_t w_rg) {
  constexpr uint32_t idesc = make_idesc(N, 0, 1);
#pragma unroll
  for (int k = 0; k < K; k += 16) mma_bf16(d, adv(a_k, k * 16), adv(w_hi_m, (k >> 3) * w_rg), idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < K; k += 16) mma_bf16(d, adv(a_k, k * 16), adv(w_lo_m, (k >> 3) * w_rg), idesc, 1u);
}
// D[128 x N] (+)= A^T B over the 128 points (both transposed views)
template <int N>
__device__ __forceinline__ void mm_outer(uint32_t d, uint64_t a_m, uint32_t a_rg, uint64_t b_m, uint32_t b_rg,
                                         uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(N, 1, 1);
#pragma unroll
  for (int k = 0; k < 128; k += 16)
    mma_bf16(d, adv(a_m, (k >> 3) * a_rg), adv(b_m, (k >> 3) * b_rg), idesc, k > 0 ? 1u : accumulate);
}

// MUFU tanh (max relative error ~2^-11, below the bf16 rounding applied to every activation operand)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


template <int DP>
__global__ void __launch_bounds__(kThreads, 1) mlp_residual_tc_kernel(const ResidualArgs a, int* status) {
  using S = Smem<DP>;
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_mma_warp = warp == kEpiThreads / 32;
  const int q = warp & 3;     // TMEM lane quadrant of this warp
  const int sub = warp >> 2;  // which 8-unit chunk(s) this thread owns (epilogue warps: 0..3)
  const int row = q * 32 + lane;  // point within the tile == TMEM lane == operand tile row
  const int d = a.d;
  const MlpShape<H> sh{d, 2};
  const int P = sh.num_params();
  float* bias_s = reinterpret_cast<float*>(sm + S::O_BIAS);   // b0[32] b1[32] b2[48]
  float* stash = reinterpret_cast<float*>(sm + S::O_STASH);
  float* tp = reinterpret_cast<float*>(sm + S::O_TRUE);
  uint64_t* mbar_p = reinterpret_cast<uint64_t*>(sm + S::O_MISC);  // [0] fast, [1] dw
  uint32_t* tmem_p = reinterpret_cast<uint32_t*>(sm + S::O_MISC + 16);

  const int64_t n_tiles = (a.n_points + 127) / 128;
  if ((int64_t)blockIdx.x >= n_tiles) return;  // nothing to do for this CTA (uniform)

  // ---- one-time set-up: TMEM, mbarriers, split weights in core-matrix layout, constants ------------------
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_p), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(mbar_p), 1);
    mbar_init(smem_u32(mbar_p + 1), 1);
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
    for (int idx = tid; idx < 32 * DP; idx += kThreads) {  // T0[r = out][c = in] = W0[c][r]
      const int r = idx / DP, c = idx % DP;
      put(S::O_T0H, S::O_T0L, S::RG_T0, r, c, c < d ? W0[c * H + r] : 0.f);
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
    // x|v|g tile: zero everything once, then the constant-one unit (column 3*DP) of every point
    for (uint32_t o = tid * 16; o < S::SZ_X; o += kThreads * 16)
      *reinterpret_cast<uint4*>(sm + S::O_X + o) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  if (!is_mma_warp && sub == 0) {
    float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store_chunk(sm + S::O_X, chunk_off(row, 3 * DP / 8, S::RG_X), ones);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const uint32_t TB = *tmem_p;
  const uint32_t mb_fast = smem_u32(mbar_p), mb_dw = smem_u32(mbar_p + 1);

  // ==========================================================================================================
  // MMA warp: one lane issues every tcgen05.mma of the tile in a fixed order, phase by phase
  // ==========================================================================================================
  if (is_mma_warp) {
    const uint32_t xs = smem_u32(sm + S::O_X), a1s = smem_u32(sm + S::O_A1), a2s = smem_u32(sm + S::O_A2),
                   zs = smem_u32(sm + S::O_Z);
    // K-major views (rows = points / weight outputs): LBO = 128 (next 8 columns), SBO = row-group bytes
    const uint64_t XK = make_desc(xs, 128, S::RG_X), A1K = make_desc(a1s, 128, S::RG_A),
                   A2K = make_desc(a2s, 128, S::RG_A), ZK = make_desc(zs, 128, S::RG_Z);
    const uint64_t T0HK = make_desc(smem_u32(sm + S::O_T0H), 128, S::RG_T0),
                   T0LK = make_desc(smem_u32(sm + S::O_T0L), 128, S::RG_T0),
                   T1HK = make_desc(smem_u32(sm + S::O_T1H), 128, S::RG_T1),
                   T1LK = make_desc(smem_u32(sm + S::O_T1L), 128, S::RG_T1),
                   T2HK = make_desc(smem_u32(sm + S::O_T2H), 128, S::RG_T2),
                   T2LK = make_desc(smem_u32(sm + S::O_T2L), 128, S::RG_T2);
    // transposed views (operand rows = tile columns): LBO = row-group bytes (next 8 tile rows), SBO = 128
    const uint64_t XM = make_desc(xs, S::RG_X, 128), A1M = make_desc(a1s, S::RG_A, 128),
                   A2M = make_desc(a2s, S::RG_A, 128), ZM = make_desc(zs, S::RG_Z, 128);
    const uint64_t T0HM = make_desc(smem_u32(sm + S::O_T0H), S::RG_T0, 128),
                   T0LM = make_desc(smem_u32(sm + S::O_T0L), S::RG_T0, 128),
                   T1HM = make_desc(smem_u32(sm + S::O_T1H), S::RG_T1, 128),
                   T1LM = make_desc(smem_u32(sm + S::O_T1L), S::RG_T1, 128),
                   T2HM = make_desc(smem_u32(sm + S::O_T2H), S::RG_T2, 128),
                   T2LM = make_desc(smem_u32(sm + S::O_T2L), S::RG_T2, 128);
    constexpr uint32_t CB = 16;  // bytes per operand column inside a row of core matrices: 8 columns = 128 B
    uint32_t not_first = 0;
#pragma unroll 1
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#ifdef PDEIP_TC_TRACE
      long long* trace = reinterpret_cast<long long*>(status) + 8;
      const bool trace_on = blockIdx.x == 0 && lane == 0 && tile == 100 * (int64_t)gridDim.x;
      int ph = 0;
#endif
      // S0: z0 = x W0, z1_0 = v W0
#ifdef PDEIP_TC_TRACE
      ph = 0;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_fwd<DP, 32>(TB + C_T1, XK, T0HK, T0LK);          // C_T1 and C_Z10 are not adjacent: two calls
        mm_fwd<DP, 32>(TB + C_Z10, adv(XK, DP * CB), T0HK, T0LK);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S1: z1, z1_1, z2_1
#ifdef PDEIP_TC_TRACE
      ph = 1;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_fwd<32, 32>(TB + C_T2, A1K, T1HK, T1LK);
        mm_fwd_multi<32, 32, 2>(TB + C_Z11, 32, adv(A1K, 32 * CB), 32 * CB, T1HK, T1LK);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S2: u, u1, u2
#ifdef PDEIP_TC_TRACE
      ph = 2;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_fwd_multi<32, OP, 3>(TB + C_U, OP, A2K, 32 * CB, T2HK, T2LK);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S3: aa_2 = za_2 W2^T
#ifdef PDEIP_TC_TRACE
      ph = 3;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_bwd<OP, 32>(TB + C_AA2, ZK, T2HM, T2LM, S::RG_T2);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S4: aa_1 = za_1 W1^T
#ifdef PDEIP_TC_TRACE
      ph = 4;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_bwd<32, 32>(TB + C_AA1, adv(ZK, OP * CB), T1HM, T1LM, S::RG_T1);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S5: g = za_0 W0^T
#ifdef PDEIP_TC_TRACE
      ph = 5;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_bwd<32, DP>(TB + C_G, adv(ZK, 2 * OP * CB), T0HM, T0LM, S::RG_T0);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S6: zg_0 = g W0
#ifdef PDEIP_TC_TRACE
      ph = 6;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_fwd<DP, 32>(TB + C_ZG0, adv(XK, 2 * DP * CB), T0HK, T0LK);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S7: zg_1 = ag_1 W1
#ifdef PDEIP_TC_TRACE
      ph = 7;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_fwd<32, 32>(TB + C_ZG1, adv(A1K, 96 * CB), T1HK, T1LK);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S8: ug = ag_2 W2
#ifdef PDEIP_TC_TRACE
      ph = 8;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_fwd<32, OP>(TB + C_UG, adv(A2K, 96 * CB), T2HK, T2LK);
        commit(mb_fast);
        TC_TRACE(2, ph);
      }
      // S9: abar_s = Ubar_s W2^T (fast);  dW2_s = ACT2^T Ubar_s, db2 (background)
#ifdef PDEIP_TC_TRACE
      ph = 9;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 4; ++s) mm_bwd<OP, 32>(TB + C_AB + 32 * s, adv(ZK, OP * s * CB), T2HM, T2LM, S::RG_T2);
        commit(mb_fast);
        TC_TRACE(2, ph);
#pragma unroll
        for (int s = 0; s < 4; ++s) mm_outer<OP>(TB + C_DW + OP * s, A2M, S::RG_A, adv(ZM, OP * s * CB), S::RG_Z, 0);
        mm_outer<OP>(TB + C_DB2, XM, S::RG_X, ZM, S::RG_Z, not_first);
        commit(mb_dw);
      }
      // S10: abar'_s = zbar_s W1^T (fast);  dW1_s = ACT1^T zbar_s, db1 (background)
#ifdef PDEIP_TC_TRACE
      ph = 10;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 4; ++s) mm_bwd<32, 32>(TB + C_AB + 32 * s, adv(ZK, OP * s * CB), T1HM, T1LM, S::RG_T1);
        commit(mb_fast);
        TC_TRACE(2, ph);
#pragma unroll
        for (int s = 0; s < 4; ++s) mm_outer<32>(TB + C_DW + 32 * s, A1M, S::RG_A, adv(ZM, OP * s * CB), S::RG_Z, 0);
        mm_outer<32>(TB + C_DB1, XM, S::RG_X, ZM, S::RG_Z, not_first);
        commit(mb_dw);
      }
      // S11: dW0 regions: rows x <-> zbar', rows v <-> zbar1', rows g <-> zbar_g'
#ifdef PDEIP_TC_TRACE
      ph = 11;
#endif
      mma_wait_operands();
      TC_TRACE(1, ph);
      if (elect_one()) {
        mm_outer_multi<32, 3>(TB + C_DW, 32, XM, S::RG_X, ZM, OP * CB, S::RG_Z);
        commit(mb_dw);
      }
      not_first = 1;
    }
  } else {
    // ========================================================================================================
    // epilogue warps
    // ========================================================================================================
    Ctx c;
    c.tbase = TB;
    c.lane_addr = TB + ((uint32_t)(q * 32) << 16);
    c.mb_fast = mb_fast;
    c.mb_dw = mb_dw;
    c.par_fast = 0;
    c.par_dw = 0;
    c.status = status;
    c.ok = true;
    const uint32_t LA = c.lane_addr;
    uint8_t* X = sm + S::O_X;
    uint8_t* A1 = sm + S::O_A1;
    uint8_t* A2 = sm + S::O_A2;
    uint8_t* Z = sm + S::O_Z;
    auto ST = [&](int arr, int unit) -> float& { return stash[(arr * 32 + unit) * 128 + row]; };

    // running sums kept in registers across the CTA's tiles (this thread's chunk of the rows its lane owns)
    float acc2[16], acc1[8], acc0[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc2[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc1[j] = 0.f; acc0[j] = 0.f; }
    float sums[PDEIP_NUM_SUMS];
#pragma unroll
    for (int k = 0; k < PDEIP_NUM_SUMS; ++k) sums[k] = 0.f;
    const float gamma = a.coef;
    const int cg4 = sub;  // the chunk of a 32-unit tile owned by this thread
    // this thread's x|v input chunk (sub-warps with sub < 2*DP/8 own one), prefetched one tile ahead
    const int dimw = 2 * d + (a.tg.kind == PDEIP_DRIFT_IN_POINTS ? d : 0);  // floats per point
    const bool has_in = sub < 2 * (DP / 8);
    const int in_band = sub / (DP / 8), in_cg = sub % (DP / 8);
    float xin[8];
    auto load_inputs = [&](int64_t t) {
      const int64_t pp = t * 128 + row;
      const bool ok_p = has_in && t < n_tiles && pp < a.n_points;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int u = in_cg * 8 + i;
        xin[i] = (ok_p && u < d) ? __ldg(a.points + elem_index(a.layout, pp, in_band * d + u, a.n_points, dimw)) : 0.f;
      }
    };
    load_inputs(blockIdx.x);
#ifdef PDEIP_TC_TRACE
    if (blockIdx.x == 0 && tid == 0) (reinterpret_cast<long long*>(status) + 8)[62] = clock64();
#endif

#pragma unroll 1
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t p = tile * 128 + row;
      const bool valid = p < a.n_points;
      const float wt = valid ? a.weight : 0.f;
      const float alpha = -2.f * wt, beta = 2.f * gamma * wt, beta_g = 2.f * wt;
#ifdef PDEIP_TC_TRACE
      long long* trace = reinterpret_cast<long long*>(status) + 8;
      const bool trace_on = blockIdx.x == 0 && tid == 0 && tile == 100 * (int64_t)gridDim.x;
      if (trace_on) trace[60] = clock64();
      int eph = 0;
#endif
      // ---- S0: x, v bands (one prefetched chunk per sub-warp; DP = 16: x0 x1 v0 v1) -----------------------
      static_assert(2 * (DP / 8) <= 4, "one input chunk per sub-warp");
      if (has_in) store_chunk(X, chunk_off(row, in_band * (DP / 8) + in_cg, S::RG_X), xin);
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S1: t1, a1_1, a2_1 -------------------------------------------------------------------------
      {
        float z0[8], z1[8], t[8], q1[8], q2[8];
        tmem_ld8x2(LA + C_T1 + 8 * cg4, LA + C_Z10 + 8 * cg4, z0, z1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i] = tanh_fast(z0[i] + bias_s[cg4 * 8 + i]);
          const float s1 = 1.f - t[i] * t[i];
          q1[i] = s1 * z1[i];
          q2[i] = (-2.f * t[i] * s1) * z1[i] * z1[i];
          ST(ST_Z10, cg4 * 8 + i) = z1[i];
        }
        store_chunk(A1, chunk_off(row, cg4, S::RG_A), t);
        store_chunk(A1, chunk_off(row, 4 + cg4, S::RG_A), q1);
        store_chunk(A1, chunk_off(row, 8 + cg4, S::RG_A), q2);
        tmem_st8(LA + C_T1 + 8 * cg4, t);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S2: t2, a1_2, a2_2 -------------------------------------------------------------------------
      {
        float z0[8], z1[8], z2[8], t[8], q1[8], q2[8];
        tmem_ld8x3(LA + C_T2 + 8 * cg4, LA + C_Z11 + 8 * cg4, LA + C_Z21 + 8 * cg4, z0, z1, z2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t[i] = tanh_fast(z0[i] + bias_s[32 + cg4 * 8 + i]);
          const float s1 = 1.f - t[i] * t[i];
          q1[i] = s1 * z1[i];
          q2[i] = s1 * z2[i] + (-2.f * t[i] * s1) * z1[i] * z1[i];
          ST(ST_Z11, cg4 * 8 + i) = z1[i];
          ST(ST_Z21, cg4 * 8 + i) = z2[i];
        }
        store_chunk(A2, chunk_off(row, cg4, S::RG_A), t);
        store_chunk(A2, chunk_off(row, 4 + cg4, S::RG_A), q1);
        store_chunk(A2, chunk_off(row, 8 + cg4, S::RG_A), q2);
        tmem_st8(LA + C_T2 + 8 * cg4, t);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S3: za_2 = 2u  ->  aa_2 = za_2 W2^T ---------------------------------------------------------
#pragma unroll 1
      for (int cg = sub; cg < OP / 8; cg += 4) {
        float u[8];
        tmem_ld8(LA + C_U + 8 * cg, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = 2.f * (u[i] + bias_s[64 + cg * 8 + i]);
        store_chunk(Z, chunk_off(row, cg, S::RG_Z), u);
      }
      if (a.tg.kind == PDEIP_DRIFT_GMM) {
        // true gradient of the GMM potential (core/potential.py:32-37), centres k = sub, sub+4, ... handled by this
        // sub-warp with an online softmax; the partial (m, se, sum e mu) goes to free TMEM columns of this lane and is
        // combined by sub-warp 0 in S6
        float x[DP];
#pragma unroll
        for (int u = 0; u < DP; ++u)
          x[u] = (valid && u < d) ? __ldg(a.points + elem_index(a.layout, p, u, a.n_points, dimw)) : 0.f;
        float m = -INFINITY, se = 0.f, accg[DP];
#pragma unroll
        for (int u = 0; u < DP; ++u) accg[u] = 0.f;
        for (int k = sub; k < a.tg.n_gaussian; k += 4) {
          const float* mu = tp + k * d;
          float s2 = 0.f;
#pragma unroll
          for (int u = 0; u < DP; ++u)
            if (u < d) {
              const float r = x[u] - mu[u];
              s2 = fmaf(r, r, s2);
            }
          const float ak = -0.5f * a.tg.inv_sigma2 * s2;
          if (ak > m) {
            const float sc = __expf(m - ak);
            se *= sc;
#pragma unroll
            for (int u = 0; u < DP; ++u) accg[u] *= sc;
            m = ak;
          }
          const float e = __expf(ak - m);
          se += e;
#pragma unroll
          for (int u = 0; u < DP; ++u)
            if (u < d) accg[u] = fmaf(e, mu[u], accg[u]);
        }
#pragma unroll
        for (int cg = 0; cg < DP / 8; ++cg) {
          float w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = accg[cg * 8 + i];
          tmem_st8(LA + C_GX + DP * sub + 8 * cg, w8);
        }
        tmem_st2(LA + C_GMS + 2 * sub, m, se);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S4: za_1 = aa_2 (1 - t2^2) -> aa_1 ----------------------------------------------------------
      {
        float aa[8], t[8];
        tmem_ld8x2(LA + C_AA2 + 8 * cg4, LA + C_T2 + 8 * cg4, aa, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) aa[i] *= (1.f - t[i] * t[i]);
        store_chunk(Z, chunk_off(row, OP / 8 + cg4, S::RG_Z), aa);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S5: za_0 = aa_1 (1 - t1^2) -> g -------------------------------------------------------------
      {
        float aa[8], t[8];
        tmem_ld8x2(LA + C_AA1 + 8 * cg4, LA + C_T1 + 8 * cg4, aa, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) aa[i] *= (1.f - t[i] * t[i]);
        store_chunk(Z, chunk_off(row, 2 * OP / 8 + cg4, S::RG_Z), aa);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S6: g band; |g|^2 and the true gradient (sub-warp 0 sees the whole point) ------------------------
#pragma unroll 1
      for (int cg = sub; cg < DP / 8; cg += 4) {
        float gv[8];
        tmem_ld8(LA + C_G + 8 * cg, gv);
        store_chunk(X, chunk_off(row, 2 * DP / 8 + cg, S::RG_X), gv);
      }
      {
        // every sub-warp sees x and g; rows i = sub, sub+4, ... of the true gradient are handled here (all loss terms
        // are sums over i, so no exchange is needed); the GMM softmax partials of S3 are combined first
        float x[DP], gq[DP];
#pragma unroll
        for (int cg = 0; cg < DP / 8; ++cg) {
          float gv[8];
          tmem_ld8(LA + C_G + 8 * cg, gv);
#pragma unroll
          for (int i = 0; i < 8; ++i) gq[cg * 8 + i] = gv[i];
        }
        if (a.tg.kind == PDEIP_DRIFT_LINEAR || a.tg.kind == PDEIP_DRIFT_GMM) {
#pragma unroll
          for (int u = 0; u < DP; ++u)
            x[u] = (valid && u < d) ? __ldg(a.points + elem_index(a.layout, p, u, a.n_points, dimw)) : 0.f;
        }
        float g2 = 0.f, gt2 = 0.f, gd2 = 0.f;
        if (a.tg.kind == PDEIP_DRIFT_GMM) {
          float ms[8];
          tmem_ld8(LA + C_GMS, ms);  // (m, se) of the four sub-warps
          float m = fmaxf(fmaxf(ms[0], ms[2]), fmaxf(ms[4], ms[6]));
          float sc[4], se = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sc[j] = (ms[2 * j + 1] > 0.f) ? __expf(ms[2 * j] - m) : 0.f;  // a sub-warp without centres has se = 0
            se = fmaf(sc[j], ms[2 * j + 1], se);
          }
          const float inv = 1.f / se;
#pragma unroll
          for (int cg = 0; cg < DP / 8; ++cg) {
            float p0[8], p1[8], p2[8], p3[8];
            tmem_ld8x4(LA + C_GX + 8 * cg, LA + C_GX + DP + 8 * cg, LA + C_GX + 2 * DP + 8 * cg,
                       LA + C_GX + 3 * DP + 8 * cg, p0, p1, p2, p3);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int u = cg * 8 + i;
              if ((u & 3) == sub && u < d) {
                const float wmu = (sc[0] * p0[i] + sc[1] * p1[i]) + (sc[2] * p2[i] + sc[3] * p3[i]);
                const float gti = (x[u] - wmu * inv) * a.tg.inv_sigma2;
                g2 = fmaf(gq[u], gq[u], g2);
                gt2 = fmaf(gti, gti, gt2);
                gd2 = fmaf(gti - gq[u], gti - gq[u], gd2);
              }
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < DP; ++u) {
            if ((u & 3) == sub && u < d) {
              float gti = 0.f;
              if (a.tg.kind == PDEIP_DRIFT_IN_POINTS) {  // stored next to the point by the integrator
                gti = valid ? __ldg(a.points + elem_index(a.layout, p, 2 * d + u, a.n_points, dimw)) : 0.f;
              } else if (a.tg.kind == PDEIP_DRIFT_LINEAR) {
#pragma unroll
                for (int k = 0; k < DP; ++k)
                  if (k < d) gti = fmaf(tp[u * d + k], x[k], gti);
              }
              g2 = fmaf(gq[u], gq[u], g2);
              gt2 = fmaf(gti, gti, gt2);
              gd2 = fmaf(gti - gq[u], gti - gq[u], gd2);
            }
          }
        }
        sums[PDEIP_SUM_G2] += wt * g2;
        sums[PDEIP_SUM_GTRUE2] += wt * gt2;
        sums[PDEIP_SUM_GT] += wt * gd2;
        sums[PDEIP_SUM_LOSS] += wt * (g2 + gt2);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S7: ag_1 = (1 - t1^2) zg_0 ------------------------------------------------------------------
      {
        float zg[8], t[8];
        tmem_ld8x2(LA + C_ZG0 + 8 * cg4, LA + C_T1 + 8 * cg4, zg, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ST(ST_ZG0, cg4 * 8 + i) = zg[i];
          zg[i] *= (1.f - t[i] * t[i]);
        }
        store_chunk(A1, chunk_off(row, 12 + cg4, S::RG_A), zg);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S8: ag_2 = (1 - t2^2) zg_1 ------------------------------------------------------------------
      {
        float zg[8], t[8];
        tmem_ld8x2(LA + C_ZG1 + 8 * cg4, LA + C_T2 + 8 * cg4, zg, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ST(ST_ZG1, cg4 * 8 + i) = zg[i];
          zg[i] *= (1.f - t[i] * t[i]);
        }
        store_chunk(A2, chunk_off(row, 12 + cg4, S::RG_A), zg);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S9: loss terms and seeds ----------------------------------------------------------
      {
        float d1 = 0.f, d2a = 0.f, d2b = 0.f;
#pragma unroll 1
        for (int cg = sub; cg < OP / 8; cg += 4) {
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
          store_chunk(Z, chunk_off(row, cg, S::RG_Z), s0);
          store_chunk(Z, chunk_off(row, OP / 8 + cg, S::RG_Z), s1v);
          store_chunk(Z, chunk_off(row, 2 * OP / 8 + cg, S::RG_Z), s2v);
          store_chunk(Z, chunk_off(row, 3 * OP / 8 + cg, S::RG_Z), sg);
        }
        // this thread's share (its chunks) of D_v V = 2 u.u1 and D_v^2 V = 2 (u1.u1 + u.u2); all terms are linear
        const float D1 = 2.f * d1, D2 = 2.f * (d2a + d2b);
        sums[PDEIP_SUM_D2] += wt * D2;
        sums[PDEIP_SUM_D1] += wt * D1;
        sums[PDEIP_SUM_LOSS] += wt * (-2.f * D2 + 2.f * gamma * D1);
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S10: through tanh of hidden layer 2 (while the dW2 GEMMs run), then store zbar, read dW2 ----------
      {
        float ab[8], ab1[8], ab2[8], abg[8], t[8];
        tmem_ld8x4(LA + C_AB + 8 * cg4, LA + C_AB + 32 + 8 * cg4, LA + C_AB + 64 + 8 * cg4, LA + C_AB + 96 + 8 * cg4,
                   ab, ab1, ab2, abg);
        tmem_ld8(LA + C_T2 + 8 * cg4, t);
        float o0[8], o1[8], o2[8], og[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int un = cg4 * 8 + i;
          const float tt = t[i], s1 = 1.f - tt * tt, s2 = -2.f * tt * s1;
          const float z1 = ST(ST_Z11, un), z2 = ST(ST_Z21, un), zg = ST(ST_ZG1, un);
          const float tb = ab[i] + ab1[i] * z1 * (-2.f * tt) +
                           ab2[i] * (z2 * (-2.f * tt) + z1 * z1 * (6.f * tt * tt - 2.f)) + abg[i] * zg * (-2.f * tt);
          o0[i] = tb * s1;
          o1[i] = ab1[i] * s1 + ab2[i] * 2.f * s2 * z1;
          o2[i] = ab2[i] * s1;
          og[i] = abg[i] * s1;
        }
        wait_dw(c);  // dW2 / db2 GEMMs done: the Ubar slots may be overwritten, the dW2 regions read
        store_chunk(Z, chunk_off(row, cg4, S::RG_Z), o0);
        store_chunk(Z, chunk_off(row, OP / 8 + cg4, S::RG_Z), o1);
        store_chunk(Z, chunk_off(row, 2 * OP / 8 + cg4, S::RG_Z), o2);
        store_chunk(Z, chunk_off(row, 3 * OP / 8 + cg4, S::RG_Z), og);
      }
      {  // quadrant q = stream q: this lane owns row `lane` of stream q's dW2 contribution; chunks sub, (4)
        float w[8];
        tmem_ld8(LA + C_DW + OP * q + 8 * sub, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc2[i] += w[i];
        if (sub == 0) {
          tmem_ld8(LA + C_DW + OP * q + 32, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc2[8 + i] += w[i];
        }
      }
      epi_arrive();
      TC_TRACE(0, eph);
      wait_fast(c);
      TC_TRACE(3, eph);
#ifdef PDEIP_TC_TRACE
      ++eph;
#endif
      // ---- S11: through tanh of hidden layer 1 (while the dW1 GEMMs run), then store zbar', read dW1 ---------
      {
        float ab[8], ab1[8], ab2[8], abg[8], t[8];
        tmem_ld8x4(LA + C_AB + 8 * cg4, LA + C_AB + 32 + 8 * cg4, LA + C_AB + 64 + 8 * cg4, LA + C_AB + 96 + 8 * cg4,
                   ab, ab1, ab2, abg);
        tmem_ld8(LA + C_T1 + 8 * cg4, t);
        float o0[8], o1[8], og[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int un = cg4 * 8 + i;
          const float tt = t[i], s1 = 1.f - tt * tt, s2 = -2.f * tt * s1;
          const float z1 = ST(ST_Z10, un), zg = ST(ST_ZG0, un);  // z2_0 = 0
          const float tb = ab[i] + ab1[i] * z1 * (-2.f * tt) + ab2[i] * (z1 * z1 * (6.f * tt * tt - 2.f)) +
                           abg[i] * zg * (-2.f * tt);
          o0[i] = tb * s1;
          o1[i] = ab1[i] * s1 + ab2[i] * 2.f * s2 * z1;
          og[i] = abg[i] * s1;
        }
        wait_dw(c);  // dW1 / db1 GEMMs done
        store_chunk(Z, chunk_off(row, cg4, S::RG_Z), o0);
        store_chunk(Z, chunk_off(row, OP / 8 + cg4, S::RG_Z), o1);
        store_chunk(Z, chunk_off(row, 2 * OP / 8 + cg4, S::RG_Z), og);
      }
      {
        float w[8];
        tmem_ld8(LA + C_DW + 32 * q + 8 * sub, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc1[i] += w[i];
      }
      epi_arrive();
      load_inputs(tile + gridDim.x);  // prefetch the next tile's x|v chunk while the dW0 GEMMs run
      wait_dw(c);  // dW0 GEMMs done (they read the x|v|g tile, which the next tile's S0 overwrites)
      // ---- S12: read dW0 rows (TMEM lane = unit of the x|v|g|one tile), chunk `sub` of the 32 outputs --------
      if (q * 32 < 4 * DP) {  // warp-uniform
        const int band = row / DP;  // 0: x rows, 1: v rows, 2: g rows, 3: the constant-one row (db0) and padding
        const int sel = (band == 1) ? 1 : (band == 2 ? 2 : 0);
        float w0[8], w1[8], w2[8];
        tmem_ld8x3(LA + C_DW + 8 * sub, LA + C_DW + 32 + 8 * sub, LA + C_DW + 64 + 8 * sub, w0, w1, w2);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc0[i] += (sel == 0) ? w0[i] : (sel == 1 ? w1[i] : w2[i]);
      }
#ifdef PDEIP_TC_TRACE
      if (trace_on) trace[61] = clock64();
#endif
    }

#ifdef PDEIP_TC_TRACE
    if (blockIdx.x == 0 && tid == 0) (reinterpret_cast<long long*>(status) + 8)[63] = clock64();
#endif
    // ---- write-out: contributor slices (one writer per (slice, index)), summed in a fixed order below ----------
    asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");  // every epilogue thread is done with the stash
    float* red = stash;  // 4 * P + 16 * 8 floats << 80 KB
    float* wsum = red + 4 * P;
    for (int i = tid; i < 4 * P; i += kEpiThreads) red[i] = 0.f;
    asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
    {
      float* mine = red + q * P;
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[sh.w_off(2) + lane * kOut + sub * 8 + i] = acc2[i];
      if (sub == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mine[sh.w_off(2) + lane * kOut + 32 + i] = acc2[8 + i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[sh.w_off(1) + lane * H + sub * 8 + i] = acc1[i];
      if (q * 32 < 4 * DP) {
        const int band = row / DP, iu = row % DP;
        if (band < 3 && iu < d) {
#pragma unroll
          for (int i = 0; i < 8; ++i) red[band * P + sh.w_off(0) + iu * H + sub * 8 + i] = acc0[i];
        } else if (row == 3 * DP) {
#pragma unroll
          for (int i = 0; i < 8; ++i) red[sh.b_off(0) + sub * 8 + i] = acc0[i];
        }
      }
      // db2 / db1 from the persistent regions: row of the constant-one unit (lane 3*DP), sub-warp 0 of its quadrant
      if (q == (3 * DP) / 32 && sub == 0) {  // warp-uniform
#pragma unroll
        for (int cg = 0; cg < 5; ++cg) {
          float w[8];
          tmem_ld8(LA + C_DB2 + 8 * cg, w);
          if (row == 3 * DP)
            for (int i = 0; i < 8; ++i) red[sh.b_off(2) + cg * 8 + i] = w[i];
        }
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
          float w[8];
          tmem_ld8(LA + C_DB1 + 8 * cg, w);
          if (row == 3 * DP)
            for (int i = 0; i < 8; ++i) red[sh.b_off(1) + cg * 8 + i] = w[i];
        }
      }
#pragma unroll
      for (int k = 0; k < PDEIP_NUM_SUMS; ++k) {
        const float sk = warp_sum(sums[k]);
        if (lane == 0) wsum[warp * PDEIP_NUM_SUMS + k] = sk;
      }
    }
    fence_before_sync();
    asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
    float* part = a.ws + (int64_t)blockIdx.x * a.pstride;
    if (c.ok) {
      for (int i = tid; i < P; i += kEpiThreads) part[i] += (red[i] + red[P + i]) + (red[2 * P + i] + red[3 * P + i]);
      if (tid < PDEIP_NUM_SUMS) {
        float sacc = 0.f;
        for (int w = 0; w < kEpiThreads / 32; ++w) sacc += wsum[w * PDEIP_NUM_SUMS + tid];
        part[P + tid] += sacc;
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(TB, 512);
}

}  // namespace tc

static int* tensor_status_word() {
  static int* w = nullptr;
  if (!w) {
    if (cudaMalloc(&w, 4096) != cudaSuccess) return nullptr;
    cudaMemset(w, 0, 4096);
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
  kern<<<residual_grid(), tc::kThreads, S::TOTAL, st>>>(a, status);
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

// debug: copies the phase trace (PDEIP_TC_TRACE builds) to the host: out[64] clock stamps
extern "C" int pdeip_debug_tensor_trace(long long* out, int n) {
  int* status = pdeip::tensor_status_word();
  if (!status) return PDEIP_ERR_CUDA;
  if (cudaMemcpy(out, reinterpret_cast<long long*>(status) + 8, sizeof(long long) * n, cudaMemcpyDeviceToHost) != cudaSuccess)
    return PDEIP_ERR_CUDA;
  return PDEIP_OK;
}
