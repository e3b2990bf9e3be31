// Micro-benchmark of tcgen05.mma (kind::f16, M=128, cta_group::1) issue behaviour on sm_100a:
// cycles per MMA as a function of N, the number of independent accumulators the issue loop rotates
// through, and the alignment of the A-operand start address in the no-swizzle K-major layout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {   // K-major, 128B swizzle, SBO = 1024
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: no-swizzle layout (LBO = 32 KB apart chunks, SBO = 128); mode 1: 128B swizzle.
// The issue loop is fully unrolled over NACC accumulators x 3 A windows with compile-time offsets, so the
// single issuing thread executes only the MMA and a few uniform adds per iteration.
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) k_bench(int shift16, int iters, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
   const bool leader = elect_one();
   if (leader) {
    const uint32_t sa = smem_u32(smem), sb = sa + 96 * 1024;
    const uint32_t idesc = make_idesc_f16(128, N);
    uint64_t da0, db0;
    if (mode == 0) { da0 = make_smem_desc(sa, 32 * 1024, 128); db0 = make_smem_desc(sb, 8 * 1024, 128); }
    else if (mode == 2) { da0 = make_smem_desc(sa, 2080, 128); db0 = make_smem_desc(sb, 3072, 128); }   // the conv kernels' strides
    else { da0 = make_desc_sw128(sa); db0 = make_desc_sw128(sb); }
    const uint64_t da1 = da0 + (uint32_t)shift16, da2 = da0 + (uint32_t)(2 * shift16);
    for (int a = 0; a < NACC; ++a) umma_f16(tm + a * N, da0, db0, idesc, 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 3 * NACC) {
#pragma unroll
      for (int a = 0; a < NACC; ++a) umma_acc(tm + a * N, da0, db0, idesc);
#pragma unroll
      for (int a = 0; a < NACC; ++a) umma_acc(tm + a * N, da1, db0, idesc);
#pragma unroll
      for (int a = 0; a < NACC; ++a) umma_acc(tm + a * N, da2, db0, idesc);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
   }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int NACC>
static void run(int grid, int shift16, int mode, const char* note, long long* d) {
  const int iters = 3 * NACC * 200;
  cudaFuncSetAttribute(k_bench<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_bench<N, NACC><<<grid, 128, 170 * 1024>>>(shift16, iters, mode, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d nacc=%d: %s\n", N, NACC, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0, mi = 1ll << 60, iss = 0;
  for (int i = 0; i < grid; ++i) { mx = std::max(mx, h[2 * i + 1]); mi = std::min(mi, h[2 * i + 1]); iss = std::max(iss, h[2 * i]); }
  printf("grid=%3d mode=%d N=%3d nacc=%d shift16=%3d : %7.1f cyc/MMA (min CTA %7.1f, issue %6.1f)  floor %5.1f  %s\n",
         grid, mode, N, NACC, shift16, (double)mx / iters, (double)mi / iters, (double)iss / iters, N / 2.0, note);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * 2 * sizeof(long long));
  printf("# dependent chain (1 accumulator) vs rotating accumulators, aligned A windows, no-swizzle layout\n");
  run<32, 1>(sms, 0, 0, "", d); run<32, 2>(sms, 0, 0, "", d); run<32, 4>(sms, 0, 0, "", d); run<32, 8>(sms, 0, 0, "", d); run<32, 16>(sms, 0, 0, "", d);
  run<64, 1>(sms, 0, 0, "", d); run<64, 2>(sms, 0, 0, "", d); run<64, 4>(sms, 0, 0, "", d); run<64, 8>(sms, 0, 0, "", d);
  run<128, 1>(sms, 0, 0, "", d); run<128, 2>(sms, 0, 0, "", d); run<128, 4>(sms, 0, 0, "", d);
  run<256, 1>(sms, 0, 0, "", d); run<256, 2>(sms, 0, 0, "", d);
  printf("# one CTA only\n");
  run<32, 1>(1, 0, 0, "", d); run<32, 8>(1, 0, 0, "", d); run<64, 8>(1, 0, 0, "", d); run<256, 2>(1, 0, 0, "", d);
  printf("# A window shifted per tap: 16 B (misaligned core matrices), 130 px rows, 128 B (aligned)\n");
  run<32, 8>(sms, 1, 0, "16 B", d); run<32, 8>(sms, 130, 0, "130 px", d); run<32, 8>(sms, 8, 0, "128 B", d);
  run<64, 8>(sms, 1, 0, "16 B", d); run<64, 8>(sms, 130, 0, "130 px", d); run<64, 8>(sms, 8, 0, "128 B", d);
  run<128, 4>(sms, 1, 0, "16 B", d); run<256, 2>(sms, 1, 0, "16 B", d);
  printf("# N = 48 / 96 / 192 (the conv kernels' shapes), aligned and 16 B-shifted windows, benchmark and kernel strides\n");
  run<48, 8>(sms, 0, 0, "", d); run<96, 4>(sms, 0, 0, "", d); run<192, 2>(sms, 0, 0, "", d);
  run<48, 8>(sms, 1, 0, "16 B", d); run<96, 4>(sms, 1, 0, "16 B", d); run<192, 2>(sms, 1, 0, "16 B", d);
  run<96, 4>(sms, 1, 2, "16 B, LBO 2080/3072", d); run<192, 2>(sms, 1, 2, "16 B, LBO 2080/3072", d); run<96, 1>(sms, 1, 2, "one accumulator", d);
  printf("# 128B-swizzle layout for comparison\n");
  run<32, 1>(sms, 0, 1, "", d); run<32, 8>(sms, 0, 1, "", d); run<64, 8>(sms, 0, 1, "", d); run<128, 4>(sms, 0, 1, "", d); run<256, 2>(sms, 0, 1, "", d);
  return 0;
}
