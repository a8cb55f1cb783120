// Shared helpers for the cdseg_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define CDSEG_API extern "C" __attribute__((visibility("default")))

#ifdef CDSEG_NO_LDG   // experiment (profiles/r02_two_stream_race.md): no read-only-path (ld.global.nc) loads anywhere
#define __ldg __ldcg
#endif

// status codes returned by every C-ABI entry point (0 == ok, >0 == cudaError_t)
#define CDSEG_OK 0
#define CDSEG_EINVAL (-1)      // bad argument (size / alignment / unsupported shape)
#define CDSEG_ENOSPC (-2)      // workspace too small

#define CDSEG_LAUNCH_CHECK()                         \
  do {                                               \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) return (int)_e;           \
  } while (0)

static inline int cdseg_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

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

__device__ __forceinline__ float gelu_erf(float x) {          // nn.GELU() default (exact erf form)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// kernel launch counter (bench.py's "gpu_launches" claim is read from here)
extern unsigned long long g_cdseg_launches;
#define CDSEG_COUNT_LAUNCH(n) (g_cdseg_launches += (n))
