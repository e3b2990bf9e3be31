// Does it matter WHO issues tcgen05.mma and in which order accumulators alternate?  (k_resblock_tc: conv_a and conv_b jobs of 18
// MMAs each, one issuing warp per convolution since round 2.)  All modes run the same 2 x iters N = 96 MMAs on one SM:
//   0: one thread, jobs of 18 MMAs alternating between two accumulators (a-job, b-job, a-job ...)
//   1: one thread, accumulators alternating after EVERY MMA
//   2: two threads (different warps), each its own accumulator, jobs of 18, free running
//   3: as 2, a commit + mbarrier wait on the own job every 18 MMAs in each thread (job-level hand-over latency)
//   4: as 0 with the commit + wait after every job (the single issuer with its bubble)
//   5: one thread, taps of three MMAs A_hi x W_hi, A_hi x W_lo, A_lo x W_hi (the merged accumulator), plain
//   7: as 5 with the third MMA on the same A window (is it the 16 KB hop between the hi and lo planes?)
//   6: as 5 with the A-operand collector: A_hi kept by the first MMA (.collector::a::fill), reused by the second (::lastuse)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_two_issuer_bench mma_two_issuer_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}

// A-operand collector: the first MMA keeps its A tile, the second (same A descriptor, other weights) reuses it
__device__ __forceinline__ void umma_acc_keep(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_acc_reuse(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_bench(int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2], jbar[2];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&jbar[0], 1); mbar_init(&jbar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
  const uint32_t idesc = make_idesc_f16(128, 96);
  const uint64_t da = make_smem_desc(sa, 2080, 128), db = make_smem_desc(sb, 3072, 128);
  const bool two = mode == 2 || mode == 3;
  if ((warp == 1 || (two && warp == 2)) && elect_one()) {
    const int me = warp - 1;
    const long long t0 = clock64();
    uint32_t jpar = 0;
    if (mode >= 8) {
      // order study of the 18 MMAs of a merged job: tap t = (chunk, kernel column), products hh = A_hi x W_hi, hl = A_hi x W_lo, lh = A_lo x W_hi
      for (int i = 0; i < 2 * iters; i += 18) {
        if (mode == 8) {                 // hh, lh, hl per tap: consecutive MMAs share the weights, not the A window
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const uint64_t a_hi = da + (uint64_t)(t % 3), a_lo = a_hi + 1040, w_hi = db + (uint64_t)(t * 192), w_lo = w_hi + 96;
            umma_acc(tm, a_hi, w_hi, idesc); umma_acc(tm, a_lo, w_hi, idesc); umma_acc(tm, a_hi, w_lo, idesc);
          }
        } else if (mode == 9) {          // all hh, then all hl, then all lh: neighbours share nothing
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int t = 0; t < 6; ++t) {
              const uint64_t a_hi = da + (uint64_t)(t % 3), a_lo = a_hi + 1040, w_hi = db + (uint64_t)(t * 192), w_lo = w_hi + 96;
              umma_acc(tm, q == 2 ? a_lo : a_hi, q == 1 ? w_lo : w_hi, idesc);
            }
        } else if (mode == 10) {         // as 5 (hh, hl, lh) but the three products go to three different accumulators
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const uint64_t a_hi = da + (uint64_t)(t % 3), a_lo = a_hi + 1040, w_hi = db + (uint64_t)(t * 192), w_lo = w_hi + 96;
            umma_acc(tm, a_hi, w_hi, idesc); umma_acc(tm + 96, a_hi, w_lo, idesc); umma_acc(tm + 192, a_lo, w_hi, idesc);
          }
        } else {                         // 11: hh and hl as ONE N = 192 MMA (the split layout), lh N = 96: cycles per 2 MMAs x 1.5
          const uint32_t idesc2 = make_idesc_f16(128, 192);
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const uint64_t a_hi = da + (uint64_t)(t % 3), a_lo = a_hi + 1040, w_hi = db + (uint64_t)(t * 384);
            umma_acc(tm, a_hi, w_hi, idesc2); umma_acc(tm + 192, a_lo, w_hi, idesc); 
          }
        }
      }
    } else if (mode >= 5) {
      for (int i = 0; i < 2 * iters; i += 18) {
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const uint64_t a_hi = da + (uint64_t)(t % 3), a_lo = a_hi + (mode == 7 ? 0 : 1040), w_hi = db + (uint64_t)(t * 192), w_lo = w_hi + 96;
          if (mode == 6) { umma_acc_keep(tm, a_hi, w_hi, idesc); umma_acc_reuse(tm, a_hi, w_lo, idesc); }
          else { umma_acc(tm, a_hi, w_hi, idesc); umma_acc(tm, a_hi, w_lo, idesc); }
          umma_acc(tm, a_lo, w_hi, idesc);
        }
      }
    } else if (!two) {
      for (int i = 0; i < 2 * iters; i += 18) {
        for (int k = 0; k < 18; ++k) {
          const int which = mode == 1 ? (k & 1) : ((i / 18) & 1);
          umma_acc(tm + which * 96, da + (uint64_t)(which * 1040 + (k % 3)), db + (uint64_t)(which * 1152 + (k % 6) * 96), idesc);
        }
        if (mode == 4) { umma_commit(&jbar[0]); mbar_wait(&jbar[0], jpar); jpar ^= 1; }
      }
    } else {
      for (int i = 0; i < iters; i += 18) {
        for (int k = 0; k < 18; ++k)
          umma_acc(tm + me * 96, da + (uint64_t)(me * 1040 + (k % 3)), db + (uint64_t)(me * 1152 + (k % 6) * 96), idesc);
        if (mode == 3) { umma_commit(&jbar[me]); mbar_wait(&jbar[me], jpar); jpar ^= 1; }
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar[me]);
    mbar_wait(&bar[me], 0);
    const long long t2 = clock64();
    out[blockIdx.x * 4 + me * 2] = t1 - t0;
    out[blockIdx.x * 4 + me * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * 4 * sizeof(long long));
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 18 * 200;
  const char* names[] = {"one thread, jobs of 18 alternate", "one thread, alternate every MMA", "two threads, free running",
                         "two threads, commit+wait per job", "one thread, commit+wait per job", "merged taps, plain",
                         "merged taps, A collector fill/lastuse", "merged taps, plain, A_lo = A_hi window", "taps as hh, lh, hl", "all hh, all hl, all lh",
                         "hh, hl, lh into 3 accumulators", "split layout: N=192 + N=96 per tap (x1.5)"};
  for (int mode = 0; mode < 12; ++mode) {
    cudaMemset(d, 0, sms * 4 * sizeof(long long));
    k_bench<<<sms, 128, 180 * 1024>>>(mode, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(sms * 4);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, mi = 0;
    for (int i = 0; i < sms; ++i) { mx = std::max(mx, std::max(h[4 * i + 1], h[4 * i + 3])); mi = std::max(mi, std::max(h[4 * i], h[4 * i + 2])); }
    printf("mode %d [%-34s] : %6.1f cycles per N=96 MMA to completion, %6.1f to the last issue\n", mode, names[mode], (double)mx / (2 * iters), (double)mi / (2 * iters));
  }
  return 0;
}
