// Micro-benchmark of tcgen05.ld (TMEM -> registers) throughput on sm_100a: cycles per 32x32b.xN load as a
// function of N (columns per instruction) and of the number of warps draining TMEM concurrently.
// Answers "how fast can the epilogue warps of k_conv_tc drain a 192-column accumulator slot".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

__device__ __forceinline__ void ld_x16(uint32_t ta, float& sink) {
  uint32_t r[16];
  tmem_ld_16(ta, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) sink += __uint_as_float(r[i]);
}

__device__ __forceinline__ void ld_x32(uint32_t ta, float& sink) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(ta) : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) sink += __uint_as_float(r[i]);
}

// mode 0: x16 loads, wait after each; mode 1: x32 loads; mode 2: three x16 loads in flight, then one wait
__global__ void __launch_bounds__(512, 1) k_bench(int nwarps, int iters, int mode, long long* out, float* sinkp) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  float sink = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (uint32_t)((i * 48 + (warp >> 2) * 96) & 255);
      if (mode == 0) { ld_x16(lane_base + col, sink); ld_x16(lane_base + col + 16, sink); ld_x16(lane_base + col + 32, sink); }
      else if (mode == 1) { ld_x32(lane_base + col, sink); ld_x16(lane_base + col + 32, sink); }
      else {
        uint32_t a[16], b[16], c[16];
        tmem_ld_16(lane_base + col, a); tmem_ld_16(lane_base + col + 16, b); tmem_ld_16(lane_base + col + 32, c);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) sink += __uint_as_float(a[k]) + __uint_as_float(b[k]) + __uint_as_float(c[k]);
      }
    }
    t1 = clock64();
  }
  if (warp < nwarps && (threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (sink == 123.456f) *sinkp = sink;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d; float* s;
  cudaMalloc(&d, sms * 16 * sizeof(long long));
  cudaMalloc(&s, 4);
  const int iters = 2000;
  printf("# each iteration drains 48 columns x 32 lanes x 4 B = 6 KB per warp\n");
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 2, 4, 8, 16}) {
      cudaMemset(d, 0, sms * 16 * sizeof(long long));
      k_bench<<<sms, 512>>>(nw, iters, mode, d, s);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d nw %d: %s\n", mode, nw, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(sms * 16);
      cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (auto v : h) mx = std::max(mx, v);
      const double cyc = (double)mx / iters;
      printf("mode %d warps %2d : %7.1f cycles per 48-column drain per warp  -> %7.1f B/cycle/SM\n", mode, nw, cyc,
             nw * 6144.0 / cyc);
    }
  return 0;
}
