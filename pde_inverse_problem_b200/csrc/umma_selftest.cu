// umma_selftest.cu — single-tile tcgen05 GEMMs through the helpers of umma.cuh, for unit tests of the
// descriptor / core-matrix-layout conventions (K-major, MN-major "transposed view", TMEM ld/st).
#include "common.cuh"
#include "umma.cuh"

namespace pdeip {

// mode 0: D[m][n] = sum_k A[m][k] * B[n][k]      A [128][K], B [N][K]        (both K-major)
// mode 1: D[m][n] = sum_k A[m][k] * B[k][n]      A [128][K], B [K][N]        (B transposed view)
// mode 2: D[m][n] = sum_r A[r][m] * B[r][n]      A [K][128], B [K][N]        (both transposed views)
// mode 3: TMEM st/ld round trip: D[m][n] = A[m][n] (K = N)
// mode 4: as mode 2 with the M = 64 instruction shape; TMEM is pre-filled with -777 so the dump shows which lanes are written
// mode 5: as mode 0 with the A operand read from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): thread = row = TMEM
//         lane stores its row as packed bf16 pairs, column c = (A[row][2c], A[row][2c+1]), at columns [64, 64 + K/2)
__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, const float* __restrict__ A,
                                                            const float* __restrict__ B, float* __restrict__ D,
                                                            int K, int N, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool a_t = mode == 2 || mode == 4;
  const int a_rows = a_t ? K : 128, a_cols = a_t ? 128 : K;
  const int b_rows = (mode == 0 || mode == 5) ? N : K, b_cols = (mode == 0 || mode == 5) ? K : N;
  const uint32_t a_rg = (uint32_t)(a_cols / 8) * 128u, b_rg = (uint32_t)(b_cols / 8) * 128u;
  uint8_t* a_tile = sm;
  uint8_t* b_tile = sm + (size_t)(a_rows / 8) * a_rg;
  if (warp == 0) {
    umma::tmem_alloc(umma::smem_u32(&tmem_base_s), 128);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(umma::smem_u32(&mbar), 1);
    umma::fence_mbar_init();
  }
  if (mode != 3) {
    for (int idx = tid; idx < a_rows * (a_cols / 8); idx += 128) {
      const int r = idx / (a_cols / 8), cg = idx % (a_cols / 8);
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = A[r * a_cols + cg * 8 + i];
      umma::store_chunk(a_tile, umma::chunk_off(r, cg, a_rg), v);
    }
    for (int idx = tid; idx < b_rows * (b_cols / 8); idx += 128) {
      const int r = idx / (b_cols / 8), cg = idx % (b_cols / 8);
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = B[r * b_cols + cg * 8 + i];
      umma::store_chunk(b_tile, umma::chunk_off(r, cg, b_rg), v);
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  bool ok = true;
  if (mode == 5) {  // A operand into TMEM: packed bf16 pairs of this thread's row
    for (int c = 0; c < K / 2; c += 4) {
      uint32_t w[4];
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 pr = __floats2bfloat162_rn(A[tid * K + 2 * (c + i)], A[tid * K + 2 * (c + i) + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&pr);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lane_addr + 64 + c), "r"(w[0]),
                   "r"(w[1]), "r"(w[2]), "r"(w[3])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (mode == 4) {  // sentinel fill (every warp its own lane quadrant), ordered before the MMA by the barrier below
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = -777.f;
    for (int c = 0; c < N; c += 8) umma::tmem_st8(lane_addr + c, v);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (mode == 3) {
    for (int c = 0; c < N; c += 8) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = A[tid * N + c + i];
      umma::tmem_st8(lane_addr + c, v);
    }
  } else {
    if (tid == 0) {
      const uint32_t at = umma::smem_u32(a_tile), bt = umma::smem_u32(b_tile);
      if (mode == 0) umma::gemm_kk(tbase, at, a_rg, 0, bt, b_rg, 0, K, N, 0);
      else if (mode == 1) umma::gemm_km(tbase, at, a_rg, 0, bt, b_rg, 0, 0, K, N, 0);
      else if (mode == 2) umma::gemm_mm(tbase, at, a_rg, 0, bt, b_rg, 0, K, N, 0);
      else if (mode == 5) {
        const uint32_t idesc = umma::make_idesc(N, 0, 0);
        for (int k = 0; k < K; k += 16) {
          const uint64_t bd = umma::make_desc(bt + (uint32_t)(k >> 3) * 128u, 128u, b_rg);
          umma::mma_bf16_ts(tbase, tbase + 64 + (uint32_t)(k >> 1), bd, idesc, (k > 0) ? 1u : 0u);
        }
      } else {
        const uint32_t idesc = (umma::make_idesc(N, 1, 1) & ~(0x1Fu << 24)) | ((uint32_t)(64 >> 4) << 24);  // M = 64
        for (int k = 0; k < K; k += 16) {
          const uint64_t ad = umma::make_desc(at + (uint32_t)(k >> 3) * a_rg, a_rg, 128u);
          const uint64_t bd = umma::make_desc(bt + (uint32_t)(k >> 3) * b_rg, b_rg, 128u);
          umma::mma_bf16(tbase, ad, bd, idesc, (k > 0) ? 1u : 0u);
        }
      }
      umma::commit(umma::smem_u32(&mbar));
    }
    ok = umma::mbar_wait(umma::smem_u32(&mbar), 0);
    umma::fence_after_sync();
  }
  if (ok) {
    for (int c = 0; c < N; c += 8) {
      float v[8];
      umma::tmem_ld8(lane_addr + c, v);
      for (int i = 0; i < 8; ++i) D[tid * N + c + i] = v[i];
    }
  } else if (tid == 0) {
    *status = 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 128);
}

}  // namespace pdeip

using namespace pdeip;

extern "C" int pdeip_debug_umma(int mode, const float* A, const float* B, float* D, int K, int N, int* status,
                                void* stream) {
  PDEIP_REQUIRE(mode >= 0 && mode <= 5 && A && D && status, PDEIP_ERR_INVALID_ARG, "bad arguments");
  PDEIP_REQUIRE(K % 16 == 0 && K >= 16 && K <= 128 && N % 16 == 0 && N >= 16 && N <= 64, PDEIP_ERR_INVALID_ARG,
                "K must be a multiple of 16 in [16,128], N a multiple of 16 in [16,64]");
  const size_t smem = (size_t)128 * 128 * 2 + (size_t)128 * 64 * 2 + 1024;
  PDEIP_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, A, B, D, K, N, status);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

// ---- TMEM read-bandwidth probe: every warp streams `reps` x (4 x tcgen05.ld.32x32b.x16) over its lane quadrant ----
namespace pdeip {
__global__ void tmem_bw_kernel(int reps, long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    umma::tmem_alloc(umma::smem_u32(&tmem_base_s), 512);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t la = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(la + (uint32_t)(((r * 4 + j) * 16) & 511))
          : "memory");
      acc ^= v[0] ^ v[15];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = (long long)reps * 4 * 16 * 128 * (blockDim.x / 32);  // bytes read by the CTA: x16 columns x 32 lanes x 4 B per warp load
  }
  if (acc == 0x12345678u) out[2] = acc;
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base_s, 512);
}
}  // namespace pdeip

extern "C" int pdeip_debug_tmem_bw(int n_warps, int reps, long long* out_host) {
  long long* d = nullptr;
  if (cudaMalloc(&d, 32) != cudaSuccess) return PDEIP_ERR_CUDA;
  cudaMemset(d, 0, 32);
  pdeip::tmem_bw_kernel<<<1, 32 * n_warps>>>(reps, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(d); return PDEIP_ERR_CUDA; }
  cudaMemcpy(out_host, d, 32, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return PDEIP_OK;
}

// ---- tcgen05.mma cost probe: `reps` MMAs of shape M x N x 16 (bf16, fp32 accumulate) issued back to back by one thread,
// cycles from the first issue to the completion of the commit.  variant bit 0: A operand from TMEM instead of shared
// memory; bit 1: M = 64 instead of 128; bit 2: rotate over 4 accumulators (independent MMAs) instead of one chain;
// bit 3: B through the MN-major (transposed) view; bit 4: rotate over 4 A tiles; bit 5: rotate over 4 B tiles;
// bit 6: pairs (A0 B0)(A0 B1)(A1 B0)(A1 B1): consecutive MMAs share their A operand. ----
namespace pdeip {
constexpr int kCostATile = 128 * 16 * 2, kCostBTile = 256 * 16 * 2;
__global__ void __launch_bounds__(128) umma_cost_kernel(int N, int reps, int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    umma::tmem_alloc(umma::smem_u32(&tmem_base_s), 512);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(umma::smem_u32(&mbar), 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < 4 * (kCostATile + kCostBTile) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0, 0x3f803f80u);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  if (warp == 0 && umma::elect_one()) {
    const bool ts = variant & 1, m64 = variant & 2, rot = variant & 4, bmn = variant & 8;
    const bool rota = variant & 16, rotb = variant & 32, pairs = variant & 64;
    const uint32_t at = umma::smem_u32(sm), bt = at + 4 * kCostATile;
    uint32_t idesc = umma::make_idesc(N, 0, bmn ? 1 : 0);
    if (m64) idesc = (idesc & ~(0x1Fu << 24)) | ((uint32_t)(64 >> 4) << 24);
    const uint32_t ncol = rot ? (uint32_t)N : 0u;
    const uint32_t d0 = tbase, d1 = tbase + ncol, d2 = tbase + 2 * ncol, d3 = tbase + 3 * ncol;
    uint32_t ta[4];
    uint64_t ad[4], bd[4];
    for (int j = 0; j < 4; ++j) {
      const int ja = pairs ? (j >> 1) : (rota ? j : 0), jb = pairs ? (j & 1) : (rotb ? j : 0);
      ta[j] = tbase + 480 + 8 * ja;
      ad[j] = umma::make_desc(at + ja * kCostATile, 128u, 256u);  // [128][16] K-major: 2 chunks per row group
      bd[j] = bmn ? umma::make_desc(bt + jb * kCostBTile, (uint32_t)(N / 8) * 128u, 128u) : umma::make_desc(bt + jb * kCostBTile, 128u, 256u);
    }
    const long long t0 = clock64();
    if (ts) {
#pragma unroll 1
      for (int r = 0; r < reps; r += 8) {
        umma::mma_bf16_ts(d0, ta[0], bd[0], idesc, 1u); umma::mma_bf16_ts(d1, ta[1], bd[1], idesc, 1u);
        umma::mma_bf16_ts(d2, ta[2], bd[2], idesc, 1u); umma::mma_bf16_ts(d3, ta[3], bd[3], idesc, 1u);
        umma::mma_bf16_ts(d0, ta[0], bd[0], idesc, 1u); umma::mma_bf16_ts(d1, ta[1], bd[1], idesc, 1u);
        umma::mma_bf16_ts(d2, ta[2], bd[2], idesc, 1u); umma::mma_bf16_ts(d3, ta[3], bd[3], idesc, 1u);
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < reps; r += 8) {
        umma::mma_bf16(d0, ad[0], bd[0], idesc, 1u); umma::mma_bf16(d1, ad[1], bd[1], idesc, 1u);
        umma::mma_bf16(d2, ad[2], bd[2], idesc, 1u); umma::mma_bf16(d3, ad[3], bd[3], idesc, 1u);
        umma::mma_bf16(d0, ad[0], bd[0], idesc, 1u); umma::mma_bf16(d1, ad[1], bd[1], idesc, 1u);
        umma::mma_bf16(d2, ad[2], bd[2], idesc, 1u); umma::mma_bf16(d3, ad[3], bd[3], idesc, 1u);
      }
    }
    const long long t1 = clock64();
    umma::commit(umma::smem_u32(&mbar));
    umma::mbar_wait(umma::smem_u32(&mbar), 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;  // issue time
    out[1] = t2 - t0;  // until all complete
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}
}  // namespace pdeip

extern "C" int pdeip_debug_umma_cost(int N, int reps, int variant, long long* out_host) {
  long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return PDEIP_ERR_CUDA;
  cudaMemset(d, 0, 16);
  const int smem = 4 * (pdeip::kCostATile + pdeip::kCostBTile) + 1024;
  cudaFuncSetAttribute(pdeip::umma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pdeip::umma_cost_kernel<<<1, 128, smem>>>(N, reps, variant, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(d); return PDEIP_ERR_CUDA; }
  cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return PDEIP_OK;
}

// ---- hand-off probe: warps 0..7 wait for an mbarrier that warp 8's tcgen05.commit arrives on after `reps` MMAs
// (128 x N x 16, A from TMEM if ts).  wait_kind 0: mbarrier.try_wait spin (may suspend the thread), 1: mbarrier.test_wait
// spin (never suspends), 2: try_wait with a suspend-time hint of `hint_ns`.  out[0] = mean cycles first issue -> commit
// issued, out[1 + w] = mean cycles commit issued -> warp w past its wait (16 rounds, round 0 discarded). ----
namespace pdeip {
__global__ void __launch_bounds__(288) umma_handoff_kernel(int N, int reps, int ts, int wait_kind, uint32_t hint_ns, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t_commit[16], t_first[16], t_wake[16][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    umma::tmem_alloc(umma::smem_u32(&tmem_base_s), 512);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(umma::smem_u32(&mbar), 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < (kCostATile + kCostBTile) / 16; i += 288) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3f803f80u, 0, 0, 0);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_base_s, mb = umma::smem_u32(&mbar);
  const uint32_t idesc = umma::make_idesc(N, 0, 0);
  const uint64_t ad = umma::make_desc(umma::smem_u32(sm), 128u, 256u), bd = umma::make_desc(umma::smem_u32(sm) + kCostATile, 128u, 256u);
  for (int it = 0; it < 16; ++it) {
    const uint32_t par = it & 1;
    asm volatile("bar.sync 1, 288;" ::: "memory");  // everybody starts the round together (the epilogue's arrive)
    if (warp == 8) {
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
          if (ts) umma::mma_bf16_ts(tbase + (r & 1) * 64, tbase + 480, bd, idesc, 1u);
          else umma::mma_bf16(tbase + (r & 1) * 64, ad, bd, idesc, 1u);
        }
        umma::commit(mb);
        t_commit[it] = clock64();
        t_first[it] = t0;
      }
      __syncwarp();
    } else {
      if (wait_kind == 0) {
        while (!umma::mbar_try_wait(mb, par)) {}
      } else if (wait_kind == 1) {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb), "r"(par) : "memory");
      } else {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb), "r"(par), "r"(hint_ns) : "memory");
      }
      umma::fence_after_sync();
      if (lane == 0) t_wake[it][warp] = clock64();
    }
  }
  __syncthreads();
  if (tid == 0) {
    long long a = 0;
    for (int it = 1; it < 16; ++it) a += t_commit[it] - t_first[it];
    out[0] = a / 15;
    for (int w = 0; w < 8; ++w) {
      long long b = 0;
      for (int it = 1; it < 16; ++it) b += t_wake[it][w] - t_commit[it];
      out[1 + w] = b / 15;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}
}  // namespace pdeip

extern "C" int pdeip_debug_umma_handoff(int N, int reps, int ts, int wait_kind, unsigned hint_ns, long long* out_host) {
  long long* d = nullptr;
  if (cudaMalloc(&d, 9 * 8) != cudaSuccess) return PDEIP_ERR_CUDA;
  cudaMemset(d, 0, 72);
  const int smem = pdeip::kCostATile + pdeip::kCostBTile + 1024;
  pdeip::umma_handoff_kernel<<<1, 288, smem>>>(N, reps, ts, wait_kind, hint_ns, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(d); return PDEIP_ERR_CUDA; }
  cudaMemcpy(out_host, d, 72, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return PDEIP_OK;
}
