// Shared helpers for the sesa_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SESA_OK 0
#define SESA_ERR_ARG 1
#define SESA_ERR_CUDA 2
#define SESA_ERR_UNSUPPORTED 3

void sesa_set_error(const char* fmt, ...);

#define SESA_CHECK_ARG(cond, ...)                \
  do {                                           \
    if (!(cond)) {                               \
      sesa_set_error(__VA_ARGS__);               \
      return SESA_ERR_ARG;                       \
    }                                            \
  } while (0)

#define SESA_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      sesa_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                     __FILE__, __LINE__);                                        \
      return SESA_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

#define SESA_LAUNCH_CHECK()                                                      \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      sesa_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                        \
      return SESA_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// index into a length-n signal extended by reflection (torch 'reflect' pad: no edge repeat)
__device__ __forceinline__ int64_t reflect_index(int64_t j, int64_t n) {
  if (j < 0) j = -j;
  if (j >= n) j = 2 * (n - 1) - j;
  return j;
}
