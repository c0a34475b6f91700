// common.cuh — error plumbing and small device helpers shared by all pdeip kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pdeip.h"

namespace pdeip {

void set_error(const char* fmt, ...);
int sm_count();

#define PDEIP_REQUIRE(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::pdeip::set_error(__VA_ARGS__);            \
      return (code);                              \
    }                                             \
  } while (0)

#define PDEIP_CUDA_OK(expr)                                                               \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::pdeip::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return PDEIP_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define PDEIP_LAUNCH_OK()                                                                  \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::pdeip::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),       \
                         __FILE__, __LINE__);                                              \
      return PDEIP_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// the same address as base + c * stride: hoists the layout switch out of per-component loops
__device__ __forceinline__ int64_t point_base(int layout, int64_t n, int dim) {
  if (layout == PDEIP_LAYOUT_BLOCK128) return (n >> 7) * ((int64_t)dim * 128) + (n & 127);
  return layout == PDEIP_LAYOUT_AOS ? n * dim : n;
}
__device__ __forceinline__ int64_t comp_stride(int layout, int64_t n_total) {
  if (layout == PDEIP_LAYOUT_BLOCK128) return 128;
  return layout == PDEIP_LAYOUT_AOS ? 1 : n_total;
}

// element (point n, component c) of a [n][dim] (AOS), [dim][n] (SOA) or [n/128][dim][128] (BLOCK128) array
__device__ __forceinline__ int64_t elem_index(int layout, int64_t n, int c, int64_t n_total, int dim) {
  if (layout == PDEIP_LAYOUT_BLOCK128) return ((n >> 7) * dim + c) * 128 + (n & 127);
  return layout == PDEIP_LAYOUT_AOS ? n * dim + c : (int64_t)c * n_total + n;
}

}  // namespace pdeip
