// philox.cuh — Philox4x32-10 (Salmon et al., SC'11) evaluated in registers, plus Box-Muller.
//
// Replaces the reference's threefry2x32 stream (jax.random.split / normal / uniform,
// utils/sampling_utils.py:8,14,27,32).  The counter layout is restated in oracle/philox.py:
//   c0,c1 = global particle id (lo,hi), c2 = step index or a TAG, c3 = block index, key = seed.
#pragma once

#include <stdint.h>

namespace pdeip {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kTagTau0 = 0xFFFFFFFFu;
constexpr uint32_t kTagInit = 0xFFFFFFFEu;

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                       uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
}

// The ten round keys (k0 + r W0, k1 + r W1) of one seed, computed once on the host and handed to the kernels in
// their argument block: the constant bank feeds them straight into the round's LOP3, so a round costs two
// IMAD.WIDE and two LOP3.
struct PhiloxRoundKeys {
  uint32_t k[20];
};
inline PhiloxRoundKeys philox_round_keys(uint64_t seed) {
  PhiloxRoundKeys rk;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    rk.k[2 * r] = k0;
    rk.k[2 * r + 1] = k1;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return rk;
}

#ifdef __CUDACC__
__device__ __forceinline__ void mul_hilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
}
// same function as philox4x32_10 (bit-identical output)
__device__ __forceinline__ void philox4x32_10_rk(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                 const PhiloxRoundKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mul_hilo(kPhiloxM0, c0, hi0, lo0);
    mul_hilo(kPhiloxM1, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ rk.k[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
}
#endif

// u1 in (0,1], u2 in [0,1) from two uint32 words -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float u1 = ((float)(xa >> 8) + 1.0f) * 5.9604644775390625e-08f;  // 2^-24
  const float ang = fmaf((float)(xb >> 8), 3.7450702829239756e-07f, -3.14159265358979323846f);  // pi (2 u2 - 1), u2 = (xb >> 8) 2^-24
  float l2, r;  // r = sqrt(-2 ln u1) on two MUFU ops (lg2, sqrt), no slow-path branch
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2 * -1.3862943611198906f));
  float s, c;
  __sincosf(ang, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// four normals of block j at (particle, step)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t particle, uint32_t step,
                                               uint32_t block, float (&out)[4]) {
  uint32_t c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32), c2 = step, c3 = block;
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  box_muller(c0, c1, out[0], out[1]);
  box_muller(c2, c3, out[2], out[3]);
}

__device__ __forceinline__ void philox_normal4_rk(const PhiloxRoundKeys& rk, uint64_t particle, uint32_t step,
                                                  uint32_t block, float (&out)[4]) {
  uint32_t c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32), c2 = step, c3 = block;
  philox4x32_10_rk(c0, c1, c2, c3, rk);
  box_muller(c0, c1, out[0], out[1]);
  box_muller(c2, c3, out[2], out[3]);
}

__device__ __forceinline__ float philox_uniform01(uint64_t seed, uint64_t particle, uint32_t tag) {
  uint32_t c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32), c2 = tag, c3 = 0u;
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (float)(c0 >> 8) * 5.9604644775390625e-08f;
}

}  // namespace pdeip
