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

// 256-bit read-only global load (sm_100: LDG.256).  The A gather has every lane of a warp on a different row, so each warp-level load
// costs one L1 wavefront per lane whatever its width: 32-byte loads halve the wavefronts per k-chunk, which is what bounds the deep
// levels' launches (profiles/r02_trace_gemm_deep.txt: 1 200 - 1 800 cycles per k-iteration with four producer warps per SM).
__device__ __forceinline__ void ldg256(const float* ptr, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(ptr));
}

// one lane of a converged warp (elect.sync): keeps the surrounding control flow -- and with it the tcgen05 operands -- warp-uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// ---- programmatic dependent launch (PDL) ----
// Kernels launched through cdseg_launch_pdl carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may be scheduled as soon
// as every CTA of the preceding kernel of the stream has executed pdl_trigger() (or exited), and they must not touch global memory
// before pdl_wait(), which returns once the preceding grid has completed and its writes are visible.  What runs before pdl_wait() --
// barrier initialisation, TMEM allocation, descriptor prefetch -- overlaps the predecessor's tail, and the launch latency itself
// (3.3 -> 1.2 us per dependent launch, profiles/r02_launch_gap.txt) is hidden.  Launched without the attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Where the trigger sits.  At the top of a kernel the dependents' CTAs become resident as early as possible -- and then hold shared memory /
// TMEM at their pdl_wait() while the other stream could have used them (measured: 2.5 % slower forward, profiles/r02_pdl.txt).  "Late":
// when a CTA starts its last piece of work, so the dependents only overlap the tail and the launch latency.
#ifdef CDSEG_PDL_EARLY
#define PDL_TRIGGER_EARLY() pdl_trigger()
#define PDL_TRIGGER_LATE() do {} while (0)
#else
#define PDL_TRIGGER_EARLY() do {} while (0)
#define PDL_TRIGGER_LATE() pdl_trigger()
#endif

extern int g_cdseg_pdl;                     // serialize.cu; env CDSEG_PDL=0 disables the attribute

template <typename... KArgs, typename... Args>
static inline cudaError_t cdseg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_cdseg_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
