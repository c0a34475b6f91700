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

// u1 in (0,1], u2 in [0,1) from two uint32 words -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float u1 = ((float)(xa >> 8) + 1.0f) * 5.9604644775390625e-08f;  // 2^-24
  const float u2 = (float)(xb >> 8) * 5.9604644775390625e-08f;
  float l2, r;  // r = sqrt(-2 ln u1) on two MUFU ops (lg2, sqrt), no slow-path branch
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2 * -1.3862943611198906f));
  float s, c;
  __sincosf(3.14159265358979323846f * (2.0f * u2 - 1.0f), &s, &c);
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

__device__ __forceinline__ float philox_uniform01(uint64_t seed, uint64_t particle, uint32_t tag) {
  uint32_t c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32), c2 = tag, c3 = 0u;
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (float)(c0 >> 8) * 5.9604644775390625e-08f;
}

}  // namespace pdeip
