// umma.cuh — thin inline-PTX layer over the sm_100a tensor-core path used by the residual kernel:
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), tcgen05.ld/st, tcgen05.commit + mbarrier, TMEM alloc.
//
// Operand tiles live in shared memory in the *no-swizzle core-matrix layout*: a tile of R rows x C columns
// of bf16 is stored as 8x8 core matrices (8 rows x 16 bytes, 128 contiguous bytes each)
//     byte(r, c) = (r % 8) * 16 + (c % 8) * 2 + (c / 8) * 128 + (r / 8) * row_group_bytes ,
//     row_group_bytes = (C / 8) * 128.
// Read through a K-major descriptor (LBO = 128, SBO = row_group_bytes) the tile is the operand [rows = M|N,
// cols = K]; read through an MN-major descriptor (SBO = 128, LBO = row_group_bytes) the *same bytes* are the
// transposed operand [M|N = cols, K = rows].  The residual kernel uses this for every tile: activations are
// A operands of the layer GEMMs (points x units, K-major) and A/B operands of the batch-reduced dW GEMMs
// (units x points, MN-major) without a second copy; a weight tile W^T serves the forward GEMM K-major and the
// backward GEMM MN-major.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace pdeip {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- shared-memory matrix descriptor (SWIZZLE_NONE, version 1) -------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // [16,30) leading-dimension byte offset >> 4 (K direction)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;       // [32,46) stride byte offset >> 4 (M/N direction)
  d |= (uint64_t)1 << 46;                                 // [46,48) descriptor version = 1 (Blackwell)
  return d;                                               // base_offset = 0, lbo_mode = 0, layout = SWIZZLE_NONE
}

// ---- instruction descriptor: kind::f16, A = B = bf16, D = fp32, M = 128 -----------------------------------
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                        // c_format = F32
         | (1u << 7)                      // a_format = BF16
         | (1u << 10)                     // b_format = BF16
         | ((uint32_t)a_mn_major << 15)   // a_major: 0 = K, 1 = MN
         | ((uint32_t)b_mn_major << 16)   // b_major
         | ((uint32_t)(n >> 3) << 17)     // n_dim = N >> 3
         | ((uint32_t)(128 >> 4) << 24);  // m_dim = M >> 4
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand from TENSOR MEMORY (K-major by construction: lane = row, 32-bit column c = the bf16 pair (k = 2c, 2c + 1)),
// B from shared memory.  a_tmem addresses the first of the 8 columns of this K = 16 step.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// exactly one lane of a converged warp (the compiler then issues tcgen05.mma / commit without a per-lane
// "waterfall" loop, which a plain `lane == 0` test provokes)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

__device__ __forceinline__ void commit(uint32_t mbar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr)
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint32_t saddr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(saddr), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: returns false if the phase did not complete within ~`budget` polls (a wrong descriptor must
// not hang the GPU box; the caller records the failure and bails out).
__device__ __forceinline__ bool mbar_wait(uint32_t saddr, uint32_t parity, uint32_t budget = (1u << 24)) {
#pragma unroll 1  // (the compiler unrolled this 64 x at every wait site: ~3 KB of code each)
  for (uint32_t i = 0; i < budget; ++i)
    if (mbar_try_wait(saddr, parity)) return true;
  return false;
}

// ---- TMEM ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_saddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 8 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31); includes the wait.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
      : "r"(taddr)
      : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// two / three / four 8-column groups with a single wait (the loads overlap each other)
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, float (&a)[8], float (&b)[8]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
  }
}

__device__ __forceinline__ void tmem_ld8x3(uint32_t ta, uint32_t tb, uint32_t tc, float (&a)[8], float (&b)[8],
                                           float (&c)[8]) {
  uint32_t r[24];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
      : "r"(ta), "r"(tb), "r"(tc)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
    c[i] = __uint_as_float(r[16 + i]);
  }
}

__device__ __forceinline__ void tmem_ld8x4(uint32_t ta, uint32_t tb, uint32_t tc, uint32_t td, float (&a)[8],
                                           float (&b)[8], float (&c)[8], float (&d)[8]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%32];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%33];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%34];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%24,%25,%26,%27,%28,%29,%30,%31}, [%35];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(ta), "r"(tb), "r"(tc), "r"(td)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
    c[i] = __uint_as_float(r[16 + i]);
    d[i] = __uint_as_float(r[24 + i]);
  }
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n\t"
      "tcgen05.wait::st.sync.aligned;\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}

__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};\n\t"
      "tcgen05.wait::st.sync.aligned;\n" ::"r"(taddr),
      "r"(__float_as_uint(a)), "r"(__float_as_uint(b))
      : "memory");
}

// ---- core-matrix tile addressing ---------------------------------------------------------------------------
// byte offset of the 16-byte chunk holding columns [8*cg, 8*cg+8) of row r
__device__ __forceinline__ uint32_t chunk_off(int r, int cg, uint32_t row_group_bytes) {
  return (uint32_t)(r & 7) * 16u + (uint32_t)cg * 128u + (uint32_t)(r >> 3) * row_group_bytes;
}

// pack 8 floats to bf16 (round-to-nearest-even) and store them as one 16-byte chunk
__device__ __forceinline__ void store_chunk(uint8_t* tile, uint32_t off, const float (&v)[8]) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 p3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 q;
  q.x = *reinterpret_cast<uint32_t*>(&p0);
  q.y = *reinterpret_cast<uint32_t*>(&p1);
  q.z = *reinterpret_cast<uint32_t*>(&p2);
  q.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(tile + off) = q;
}

// GEMM helpers (issued by ONE thread).  K-major A [128 x K] at columns a_c0.. of tile A, times operand B.
//   kmajor_b: B tile rows = N, cols = K (e.g. W^T stored [out][in]) at column b_c0
//   mnmajor_b: B tile rows = K, cols = N (same bytes, transposed view)
// All K are multiples of 16 (two core matrices per MMA along K).
__device__ __forceinline__ void gemm_kk(uint32_t d_tmem, uint32_t a_tile, uint32_t a_rg, int a_c0, uint32_t b_tile,
                                        uint32_t b_rg, int b_c0, int K, int N, uint32_t accumulate) {
  const uint32_t idesc = make_idesc(N, 0, 0);
  for (int k = 0; k < K; k += 16) {
    const uint64_t ad = make_desc(a_tile + (uint32_t)((a_c0 + k) >> 3) * 128u, 128u, a_rg);
    const uint64_t bd = make_desc(b_tile + (uint32_t)((b_c0 + k) >> 3) * 128u, 128u, b_rg);
    mma_bf16(d_tmem, ad, bd, idesc, (k > 0) ? 1u : accumulate);
  }
}

// A K-major [128 x K] (cols a_c0..), B^T view of tile B: operand rows N = B-tile columns b_c0.., K = B-tile rows
// b_r0.. (multiples of 8)
__device__ __forceinline__ void gemm_km(uint32_t d_tmem, uint32_t a_tile, uint32_t a_rg, int a_c0, uint32_t b_tile,
                                        uint32_t b_rg, int b_c0, int b_r0, int K, int N, uint32_t accumulate) {
  const uint32_t idesc = make_idesc(N, 0, 1);
  for (int k = 0; k < K; k += 16) {
    const uint64_t ad = make_desc(a_tile + (uint32_t)((a_c0 + k) >> 3) * 128u, 128u, a_rg);
    const uint64_t bd = make_desc(b_tile + (uint32_t)(b_c0 >> 3) * 128u + (uint32_t)((b_r0 + k) >> 3) * b_rg, b_rg, 128u);
    mma_bf16(d_tmem, ad, bd, idesc, (k > 0) ? 1u : accumulate);
  }
}

// both operands transposed views (batch-reduced outer product): D[m][n] = sum_r A[r][a_c0+m] * B[r][b_c0+n],
// r = tile rows (points) r0 .. r0+K
__device__ __forceinline__ void gemm_mm(uint32_t d_tmem, uint32_t a_tile, uint32_t a_rg, int a_c0, uint32_t b_tile,
                                        uint32_t b_rg, int b_c0, int K, int N, uint32_t accumulate) {
  const uint32_t idesc = make_idesc(N, 1, 1);
  for (int k = 0; k < K; k += 16) {
    const uint64_t ad = make_desc(a_tile + (uint32_t)(a_c0 >> 3) * 128u + (uint32_t)(k >> 3) * a_rg, a_rg, 128u);
    const uint64_t bd = make_desc(b_tile + (uint32_t)(b_c0 >> 3) * 128u + (uint32_t)(k >> 3) * b_rg, b_rg, 128u);
    mma_bf16(d_tmem, ad, bd, idesc, (k > 0) ? 1u : accumulate);
  }
}

}  // namespace umma
}  // namespace pdeip
