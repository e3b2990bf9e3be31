// How much does a software grid barrier cost against a kernel boundary?  (1/8-resolution layers of the backbone: 44 launches of
// ~7 us of CTA lifetime each, ~12 us apart in the replayed graph.)
//   A: persistent kernel, one CTA per SM, `iters` rounds of [__syncthreads; thread 0: fence, atomicAdd, spin on the counter; __syncthreads]
//   B: `iters` back-to-back launches of an (almost) empty kernel with the same shape and 200 KB of dynamic shared memory, captured
//      in a CUDA graph with programmatic dependent launch edges
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_barrier_bench grid_barrier_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(192, 1) k_persistent(unsigned* counter, int iters, float* sink) {
  extern __shared__ float sm[];
  const unsigned n = gridDim.x;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    sm[threadIdx.x] = acc + it;                    // a token amount of work
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned target = (unsigned)(it + 1) * n;
      while (*(volatile unsigned*)counter < target) { }
      __threadfence();
    }
    __syncthreads();
    acc += sm[(threadIdx.x + 1) % 192];
  }
  if (acc == 12345.f) sink[0] = acc;
}

__global__ void __launch_bounds__(192, 1) k_step(float* sink, int it) {
  extern __shared__ float sm[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  sm[threadIdx.x] = (float)it;
  __syncthreads();
  if (sm[(threadIdx.x + 1) % 192] == 12345.f) sink[0] = 1.f;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* counter; float* sink;
  cudaMalloc(&counter, 4); cudaMalloc(&sink, 4);
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(k_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1000;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemsetAsync(counter, 0, 4, st);
    cudaEventRecord(e0, st);
    k_persistent<<<sms, 192, smem, st>>>(counter, iters, sink);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("A persistent kernel, %d CTAs: %.3f us per grid barrier round (%s)\n", sms, 1000.f * ms / iters, cudaGetErrorString(cudaGetLastError()));
  }
  // B: graph of launches with PDL edges
  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int it = 0; it < iters; ++it) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = sms; cfg.blockDim = 192; cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl;
      cudaLaunchKernelEx(&cfg, k_step, sink, it);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0, st);
      cudaGraphLaunch(ge, st);
      cudaEventRecord(e1, st);
      cudaStreamSynchronize(st);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("B graph of %d kernels (%s): %.3f us per kernel boundary (%s)\n", iters, pdl ? "PDL edges" : "plain edges", 1000.f * ms / iters, cudaGetErrorString(cudaGetLastError()));
    }
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  return 0;
}
