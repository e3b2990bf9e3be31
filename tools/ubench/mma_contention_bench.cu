// What slows tcgen05.mma inside the conv kernels?  In isolation an M=128, N=96, K=16 MMA costs 56 cycles
// (mma_bench.cu); inside k_conv_stream / k_resblock_tc the issuer measures 72-92.  This benchmark adds the
// kernels' other activities one by one next to the same MMA stream:
//   bit 0: a tcgen05.commit after every 9 MMAs (as the per-entry x_empty commits)
//   bit 1: four warps draining TMEM with tcgen05.ld in a loop (the epilogue)
//   bit 2: one warp streaming 8 KB bulk copies global -> shared into a ring (the producer)
//   bit 3: four warps writing shared memory with st.shared.v4 (the y-row emission)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_contention_bench mma_contention_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}

template <int N>
__global__ void __launch_bounds__(320, 1) k_bench(int mode, int iters, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, cbar[4], lbar[8];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&lbar[i], 1);
    stop = 0;
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    // ---- MMA issuer: groups of 9 MMAs over 5 accumulators, A windows shifted like the 3 kernel columns ----
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
      const uint32_t idesc = make_idesc_f16(128, N);
      const uint64_t da = make_smem_desc(sa, 2080, 128), db = make_smem_desc(sb, 3072, 128);
      umma_f16(tm, da, db, idesc, 0u);
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t0 = clock64();
      int g = 0;
      for (int i = 0; i < iters; i += 9, ++g) {
        const uint32_t dcol = tm + (g % 5) * 96;
        const uint64_t a = da + (uint64_t)((g & 7) * 520), b = db + (uint64_t)((g & 3) * 1152);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          umma_acc(dcol, a + kx, b + kx * 384, idesc);
          umma_acc(dcol, a + kx, b + kx * 384 + 96, idesc);
          umma_acc(dcol, a + kx + 260, b + kx * 384, idesc);
        }
        if (mode & 1) umma_commit(&cbar[g & 3]);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 1);
      const long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
      stop = 1;
    }
  } else if (warp >= 2 && warp < 6) {
    if (mode & 2) {                              // TMEM drain loop
      const uint32_t la = tm + ((uint32_t)((warp & 3) * 32) << 16);
      float sink = 0.f;
      int k = 0;
      while (!stop) {
        uint32_t r[16];
        tmem_ld_16(la + (k % 5) * 96, r); tmem_ld_wait();
        sink += __uint_as_float(r[0]);
        tmem_ld_16(la + (k % 5) * 96 + 16, r); tmem_ld_wait();
        sink += __uint_as_float(r[1]);
        tmem_ld_16(la + (k % 5) * 96 + 32, r); tmem_ld_wait();
        sink += __uint_as_float(r[2]);
        ++k;
      }
      if (sink == 123.f) out[0] = 0;
    }
  } else if (warp == 6) {
    if ((mode & 4) && lane == 0) {               // bulk-copy stream into a ring behind the operands
      int k = 0;
      while (!stop) {
        const int s = k & 7;
        if (k >= 8) mbar_wait(&lbar[s], ((k >> 3) - 1) & 1);
        mbar_expect_tx(&lbar[s], 8320);
        for (int q = 0; q < 4; ++q)
          bulk_load(smem + 128 * 1024 + s * 8320 + q * 2080, gsrc + (size_t)((k * 4 + q) & 1023) * 2080, 2080, &lbar[s]);
        ++k;
      }
    }
  } else if (warp >= 7) {
    if (mode & 8) {                              // shared-memory store stream
      uint32_t a = smem_u32(smem) + 196 * 1024 + (threadIdx.x - 224) * 16;
      while (!stop) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0x3c003c00u) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + 1536), "r"(0x3c003c00u) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  uint8_t* g;
  cudaMalloc(&d, sms * 2 * sizeof(long long));
  cudaMalloc(&g, 1024 * 2080 + 4096);
  cudaMemset(g, 0x3c, 1024 * 2080 + 4096);
  cudaFuncSetAttribute(k_bench<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 9 * 400;
  for (int mode : {0, 1, 2, 4, 8, 3, 5, 7, 15}) {
    k_bench<96><<<sms, 320, 216 * 1024>>>(mode, iters, g, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(sms * 2);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = std::max(mx, h[2 * i + 1]);
    printf("mode %2d [%s%s%s%s] : %6.1f cycles per N=96 MMA\n", mode, (mode & 1) ? "commit " : "", (mode & 2) ? "tmem-ld " : "",
           (mode & 4) ? "bulk-copy " : "", (mode & 8) ? "st.shared" : "", (double)mx / iters);
  }
  return 0;
}
