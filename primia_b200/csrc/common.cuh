// Shared helpers for the primia_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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

// ---- programmatic dependent launch (PDL).  A kernel launched through pm_launch() may be scheduled while its predecessor in
// the stream is still draining: its CTAs run their prologue (barrier init, TMEM allocation, descriptor prefetch) and then block
// in pm_pdl_sync() until the predecessor grid has completed and flushed.  Rules kept by every kernel that uses it:
//   * pm_pdl_sync() is executed by every thread, unconditionally, before the first global-memory access (read OR write);
//   * nothing before it touches global memory.
// Transitivity (kernel C after B after A also sees A's results) holds because B cannot complete before its own wait returned.
// Launched without the attribute (PRIMIA_PDL=0, or a plain <<<>>> launch) both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pm_pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
static inline bool pm_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PRIMIA_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t pm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
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
