// Probe of the tcgen05.mma (kind::f16, fp32 accumulate in TMEM) ARITHMETIC on sm_100a: how the 16 products of one
// MMA and the accumulator are aligned, truncated and rounded.  The split-fp16 convolution kernels rely on fp32-class
// accumulation; the stage reports (profiles/r02_stage_*_d192.txt) show a systematic negative bias in every tcgen05
// stage, which this probe characterises (tools/ubench/mma_round_probe.py drives it and fits a model).
//
// One CTA.  D[128 x N] = sum over steps s of A_s[128 x 16] * B_s[N x 16]^T, step 0 overwrites (accumulate = 0), every
// later step accumulates; each step is ONE tcgen05.mma followed by a commit + wait, so the order of accumulation is
// exactly the order of the steps.  Every row m is an independent experiment (its own A values).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -o libmma_round_probe.so mma_round_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

// A: [nsteps][128][16] halfs, B: [nsteps][N][16] halfs (row-major, K innermost); D: [128][N] floats.  N = 16..256, % 16 == 0
__global__ void __launch_bounds__(128, 1) k_probe(const __half* A, const __half* B, int nsteps, int N, float* D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sA = smem;                 // no-swizzle K-major: [K half][128 rows][8 halfs]
  uint8_t* sB = smem + 4096;          // [K half][N rows][8 halfs]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t idesc = make_idesc_f16(128, N);
  uint32_t par = 0;
  for (int s = 0; s < nsteps; ++s) {
    for (int i = threadIdx.x; i < 128 * 16; i += 128) {
      const int m = i / 16, k = i % 16;
      reinterpret_cast<__half*>(sA)[(k / 8) * 128 * 8 + m * 8 + k % 8] = A[((size_t)s * 128 + m) * 16 + k];
    }
    for (int i = threadIdx.x; i < N * 16; i += 128) {
      const int n = i / 16, k = i % 16;
      reinterpret_cast<__half*>(sB)[(k / 8) * N * 8 + n * 8 + k % 8] = B[((size_t)s * N + n) * 16 + k];
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      umma_f16(tm, make_smem_desc(smem_u32(sA), 128 * 16, 128), make_smem_desc(smem_u32(sB), (uint32_t)N * 16, 128), idesc, s ? 1u : 0u);
      umma_commit(&bar);
      mbar_wait(&bar, par);
    }
    par ^= 1;
    __syncthreads();
  }
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_16(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

extern "C" __attribute__((visibility("default")))
int mma_round_probe(const void* hA, const void* hB, int nsteps, int N, float* hD) {
  if (N < 16 || N > 256 || N % 16 || nsteps < 1) return -1;
  __half *dA = nullptr, *dB = nullptr; float* dD = nullptr;
  const size_t na = (size_t)nsteps * 128 * 16 * 2, nb = (size_t)nsteps * N * 16 * 2, nd = (size_t)128 * N * 4;
  if (cudaMalloc(&dA, na) != cudaSuccess || cudaMalloc(&dB, nb) != cudaSuccess || cudaMalloc(&dD, nd) != cudaSuccess) return -2;
  cudaMemcpy(dA, hA, na, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, nb, cudaMemcpyHostToDevice);
  k_probe<<<1, 128, 4096 + N * 32 + 256>>>(dA, dB, nsteps, N, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) cudaMemcpy(hD, dD, nd, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  if (e != cudaSuccess) { fprintf(stderr, "mma_round_probe: %s\n", cudaGetErrorString(e)); return -3; }
  return 0;
}
