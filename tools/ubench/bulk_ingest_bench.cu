// How fast can one SM pull L2-resident data into shared memory with cp.async.bulk, as a function of the copy size and of how
// many SMs pull at once?  (k_conv_stream at 1/8 resolution: 74 KB of weights + 16 ring entries x 4 copies of 2080 B per CTA.)
// Each CTA streams `total` bytes from its own window of an L2-resident buffer into a ring of 8 slots; one thread issues, waits
// only when the ring is full.  Reports bytes per cycle per SM (max time over CTAs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_ingest_bench bulk_ingest_bench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include "../../hobot_stereonet_b200/csrc/tc_ptx.cuh"

using namespace snb::ptx;

__global__ void __launch_bounds__(128, 1) k_ingest(const uint8_t* src, size_t window, uint32_t copy_bytes, uint32_t per_slot, int nslot_fills, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (size_t)blockIdx.x * window;
    const uint32_t slot_bytes = copy_bytes * per_slot;
    size_t off = 0;
    const long long t0 = clock64();
    for (int f = 0; f < nslot_fills; ++f) {
      const int s = f & 7;
      if (f >= 8) mbar_wait(&full[s], ((f >> 3) - 1) & 1);        // the previous fill of this slot has landed
      mbar_expect_tx(&full[s], slot_bytes);
      for (uint32_t q = 0; q < per_slot; ++q) {
        bulk_load(smem + (size_t)s * slot_bytes + (size_t)q * copy_bytes, base + off, copy_bytes, &full[s]);
        off += copy_bytes;
        if (off + copy_bytes > window) off = 0;
      }
    }
    for (int f = nslot_fills > 8 ? nslot_fills - 8 : 0; f < nslot_fills; ++f) mbar_wait(&full[f & 7], (f >> 3) & 1);
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t window = 512 * 1024;                      // per CTA: 148 x 512 KB = 74 MB, L2-resident after the first pass
  uint8_t* src; long long* d;
  cudaMalloc(&src, window * sms + 65536); cudaMemset(src, 1, window * sms + 65536);
  cudaMalloc(&d, sms * sizeof(long long));
  cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Case { uint32_t copy, per_slot; };
  const Case cases[] = {{2080, 4}, {2080, 8}, {4160, 2}, {8320, 1}, {16640, 1}, {18432, 1}};
  for (int grid : {1, 37, 74, sms}) {
    for (const Case& c : cases) {
      const uint32_t slot = c.copy * c.per_slot;
      const int fills = (int)((4u << 20) / slot);        // 4 MB per CTA
      for (int rep = 0; rep < 2; ++rep) {
        k_ingest<<<grid, 128, 8 * (size_t)slot + 256>>>(src, window, c.copy, c.per_slot, fills, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      const long long mx = *std::max_element(h.begin(), h.end());
      printf("CTAs %3d  copy %5u B x %u per slot (8 slots in flight): %6.1f B/cycle/SM  (%.2f TB/s aggregate at 1.9 GHz)\n", grid, c.copy, c.per_slot,
             (double)fills * slot / mx, (double)fills * slot / mx * grid * 1.9e9 / 1e12);
    }
  }
  return 0;
}
