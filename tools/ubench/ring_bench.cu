// Micro-benchmark of the MMA-issuer <-> epilogue handshake on sm_100a: one thread issues K tcgen05.mma per
// iteration into TMEM slot (it % S), commits to full[slot]; EW epilogue warps wait full[slot], optionally
// tcgen05.ld the slot, and arrive on empty[slot].  Reports cycles per iteration.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

template <int S, int K, int N, int EW, int LD, int MODE>
__global__ void __launch_bounds__(128 + EW * 32, 1) k_ring(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[S], empty[S];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], EW); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  constexpr int COLS = 512 / S;
  if (warp == 1) {
    const bool leader = elect_one();
    if (leader) {
      const uint32_t sa = smem_u32(smem), sb = sa + 96 * 1024;
      const uint32_t idesc = make_idesc_f16(128, N);
      const uint64_t da = make_smem_desc(sa, 32 * 1024, 128), db = make_smem_desc(sb, 8 * 1024, 128);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int s = it % S;
        if (MODE != 1) { mbar_wait(&empty[s], ((it / S) & 1) ^ 1); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < K; ++k) umma_f16(tm + s * COLS, da + (uint32_t)k, db, idesc, k > 0 ? 1u : 0u);
        if (MODE != 2) umma_commit(&full[s]);
      }
      out[blockIdx.x * 2] = clock64() - t0;
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    float sum = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters && MODE == 0; ++it) {
      const int s = it % S;
      mbar_wait(&full[s], (it / S) & 1);
      tc_fence_after();
      if (LD) {
        float a[16], b[16];
        tmem_ld_2x16(tm + ((uint32_t)(wq * 32) << 16) + s * COLS, tm + ((uint32_t)(wq * 32) << 16) + s * COLS + 16, a, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += a[i] + b[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (warp == 4 && lane == 0) out[blockIdx.x * 2 + 1] = clock64() - t0;
    if (sum == 123.456f) out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int S, int K, int N, int EW, int LD, int MODE = 0>
static void run(long long* d) {
  const int iters = 2000, grid = 148;
  cudaFuncSetAttribute(k_ring<S, K, N, EW, LD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_ring<S, K, N, EW, LD, MODE><<<grid, 128 + EW * 32, 140 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("S=%d K=%d: %s\n", S, K, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0, ex = 0;
  for (int i = 0; i < grid; ++i) { mx = std::max(mx, h[2 * i]); ex = std::max(ex, h[2 * i + 1]); }
  printf("mode=%d slots=%d  MMAs/iter=%2d N=%3d  epi warps=%d ld=%d : issuer %7.1f cyc/iter, epilogue %7.1f cyc/iter   (MMA floor %5.0f)\n", MODE, S, K, N, EW,
         LD, (double)mx / iters, (double)ex / iters, K * (N / 2.0));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  printf("# handshake only (no MMAs)\n");
  run<8, 0, 64, 8, 0>(d); run<8, 0, 64, 8, 1>(d); run<2, 0, 64, 8, 1>(d); run<8, 0, 64, 4, 1>(d); run<8, 0, 64, 1, 0>(d);
  printf("# with MMAs\n");
  run<8, 2, 64, 8, 1>(d); run<8, 6, 64, 8, 1>(d); run<8, 18, 64, 8, 1>(d); run<2, 6, 64, 8, 1>(d); run<2, 18, 64, 8, 1>(d);
  run<2, 6, 192, 8, 1>(d); run<2, 3, 192, 8, 1>(d); run<4, 6, 128, 8, 1>(d);
  printf("# issuer alone: mode 1 = no wait on empty (commit only), mode 2 = no commit (wait only, always passes after first lap? no epilogue -> only first S pass)\n");
  run<8, 0, 64, 8, 0, 1>(d); run<8, 6, 64, 8, 0, 1>(d); run<8, 18, 64, 8, 0, 1>(d); run<8, 6, 32, 8, 0, 1>(d);
  return 0;
}
