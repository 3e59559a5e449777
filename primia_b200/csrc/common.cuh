// Shared helpers for the primia_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/primia_b200.h"

extern char g_pm_err[512];
int pm_set_err(const char* file, int line, const char* msg);

#define PM_CHECK_ARG(cond)                                             \
  do {                                                                 \
    if (!(cond)) {                                                     \
      snprintf(g_pm_err, sizeof(g_pm_err), "%s:%d: bad argument: %s", __FILE__, __LINE__, #cond); \
      return PM_EINVAL;                                                \
    }                                                                  \
  } while (0)

#define PM_CUDA(expr)                                                  \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) {                                           \
      snprintf(g_pm_err, sizeof(g_pm_err), "%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return PM_ECUDA;                                                 \
    }                                                                  \
  } while (0)

#define PM_LAUNCH_OK()                                                 \
  do {                                                                 \
    cudaError_t _e = cudaGetLastError();                               \
    if (_e != cudaSuccess) {                                           \
      snprintf(g_pm_err, sizeof(g_pm_err), "%s:%d: launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return PM_ECUDA;                                                 \
    }                                                                  \
    return PM_OK;                                                      \
  } while (0)

static inline cudaStream_t S(pm_stream_t s) { return (cudaStream_t)s; }

static inline int pm_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// grid sized as a multiple of the SM count for grid-stride kernels
static inline int pm_grid(size_t work_items, int threads, int per_thread = 1, int max_waves = 8) {
  size_t blocks = (work_items + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
  size_t cap = (size_t)pm_num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
