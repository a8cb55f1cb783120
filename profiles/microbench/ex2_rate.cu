#include <cuda_fp16.h>
#include <cstdio>
__global__ void k32(float* out, float x, int iters) {
  float a = x + threadIdx.x * 1e-3f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void k16(float* out, float x, int iters) {
  unsigned a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(a + b + c + d);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k32<<<148 * 8, 256>>>(out, 0.5f, iters); else k16<<<148 * 8, 256>>>(out, 0.5f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = 148.0 * 8 * 256 * iters * 4.0;
      if (rep) printf("%s: %.3f ms, %.2f T instr/s (thread-level), %.2f T exp/s\n", which ? "ex2.f16x2" : "ex2.f32", ms, ops / ms / 1e9, ops * (which ? 2 : 1) / ms / 1e9);
    }
  }
  return 0;
}
