// Launch-to-launch cost of a chain of dependent kernels on one stream: plain launches, programmatic dependent launch
// (griddepcontrol.launch_dependents at the top + griddepcontrol.wait before the first global access), and a CUDA graph of the chain.
// Each kernel does a fixed amount of "work" (spin for W ns on 148 CTAs) after a prologue of P ns that touches no global memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap launch_gap.cu && ./launch_gap
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool PDL>
__global__ void link_kernel(const float* in, float* out, int prologue_ns, int work_ns) {
  if (PDL) asm volatile("griddepcontrol.launch_dependents;");
  unsigned long long t0 = gtime();
  while (gtime() - t0 < (unsigned long long)prologue_ns) {}
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = in[blockIdx.x * blockDim.x + threadIdx.x];
  t0 = gtime();
  while (gtime() - t0 < (unsigned long long)work_ns) {}
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + 1.0f;
}

template <bool PDL>
static void launch(cudaStream_t st, const float* in, float* out, int grid, int p, int w) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
  cudaLaunchKernelEx(&cfg, link_kernel<PDL>, in, out, p, w);
}

int main() {
  const int N = 400, GRID = 148;
  float *a, *b; cudaMalloc(&a, GRID * 128 * 4); cudaMalloc(&b, GRID * 128 * 4); cudaMemset(a, 0, GRID * 128 * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int work : {0, 5000, 20000}) for (int pro : {0, 1500}) {
    float ms[3];
    for (int variant = 0; variant < 3; ++variant) {
      cudaGraphExec_t ge = nullptr;
      if (variant == 2) {
        cudaGraph_t g; cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < N; ++i) launch<false>(st, i & 1 ? b : a, i & 1 ? a : b, GRID, pro, work);
        cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
      }
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        if (variant == 2) cudaGraphLaunch(ge, st);
        else for (int i = 0; i < N; ++i) { if (variant) launch<true>(st, i & 1 ? b : a, i & 1 ? a : b, GRID, pro, work); else launch<false>(st, i & 1 ? b : a, i & 1 ? a : b, GRID, pro, work); }
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[variant], e0, e1);
      }
    }
    printf("work %5d ns prologue %4d ns: per link plain %.2f us, PDL %.2f us, graph %.2f us  (kernel body alone %.2f us)\n", work, pro,
           1e3 * ms[0] / N, 1e3 * ms[1] / N, 1e3 * ms[2] / N, (work + pro) / 1e3);
  }
  float h; cudaMemcpy(&h, a, 4, cudaMemcpyDeviceToHost);
  printf("check %g err %s\n", h, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
