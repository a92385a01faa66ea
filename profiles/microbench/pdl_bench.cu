// Kernel-to-kernel hand-off inside a CUDA graph with and without programmatic dependent launch (PDL):
// a chain of NK kernels, each `work` dependent FMAs on `ctas` CTAs; reports microseconds per kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_bench pdl_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void __launch_bounds__(256) link(float* buf, int work, size_t smem_touch) {
  extern __shared__ float sm[];
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  float v = buf[blockIdx.x * 256 + threadIdx.x];
  for (int i = 0; i < work; ++i) v = fmaf(v, 1.0001f, 0.5f);
  if (smem_touch) sm[threadIdx.x] = v;
  buf[blockIdx.x * 256 + threadIdx.x] = v;
}

template <bool PDL>
static void launch(float* buf, int ctas, int work, size_t smem, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = PDL ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, link<PDL>, buf, work, smem);
  if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); exit(1); }
}

template <bool PDL>
static float run(float* buf, int ctas, int work, size_t smem, int NK) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaFuncSetAttribute(link<PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < NK; ++k) launch<PDL>(buf, ctas, work, smem, s);
  cudaStreamEndCapture(s, &g);
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); exit(1); }
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(b, s);
  cudaStreamSynchronize(s);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  e = cudaGetLastError();
  if (e != cudaSuccess) { printf("run: %s\n", cudaGetErrorString(e)); exit(1); }
  return ms * 1e3f / (10 * NK);
}

int main() {
  float* buf;
  cudaMalloc(&buf, 1024 * 256 * sizeof(float));
  cudaMemset(buf, 0, 1024 * 256 * sizeof(float));
  const int NK = 128;
  printf("%6s %8s %8s | %10s %10s %8s\n", "ctas", "work", "smem", "plain us", "pdl us", "saved");
  for (int ctas : {60, 148, 296})
    for (int work : {0, 2000, 10000, 40000})
      for (size_t smem : {(size_t)0, (size_t)200 * 1024}) {
        const float p = run<false>(buf, ctas, work, smem, NK), q = run<true>(buf, ctas, work, smem, NK);
        printf("%6d %8d %8zu | %10.2f %10.2f %8.2f\n", ctas, work, smem, p, q, p - q);
      }
  return 0;
}
