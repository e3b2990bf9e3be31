// A chain of same-shape streaming convolutions as a LAYER PIPELINE in one launch (SNB_PREC_TC_F16X2).  OPT-IN EXPERIMENT
// (SNB_FLAG_PIPE / SNB_PIPE=1): correct, parity-tested, but slower than one k_conv_stream launch per convolution.
//
// The 1/8-resolution backbone (layer2: 15 identity BasicBlocks = 30 convolutions 64 -> 64 at 68 x 120 px x 2 views at
// config 2) carries ~4 us of tensor work per convolution and costs ~9.5 us per launch inside the CUDA graph.  Here the
// unit of parallelism is the LAYER, not the row chunk:
//   * CTA (l, n, strip, cc) owns convolution l of the chain for view n, one 128-pixel strip and one 32-channel output
//     slice, for ALL rows; its weights are staged once and stay in shared memory; no halo waste (k_conv_stream needs two
//     extra input rows per row chunk: 2x at the two-row chunks that fill 148 SMs);
//   * rows flow from layer to layer through L2: the epilogue warps store output row r and bump a shared counter
//     (red.release.cta), a publisher warp turns it into `done[cta] = r + 1` (st.release.gpu, off the epilogue's path);
//     the producer warp of layer l + 1 polls the counters of the CTAs that write the rows / columns it is about to
//     stream (all channel slices, strips s-1..s+1), fences once (acq_rel.gpu + proxy.async) and bulk-copies the row;
//   * CTAs are numbered layer-major, so when fewer than all CTAs are resident the resident ones are the EARLIEST layers,
//     which never wait for later ones: the pipeline drains front to back and cannot deadlock.  Every layer writes its own
//     output tensor (a consumer may still read row r of layer l - 2 as its residual while layer l runs).
// MEASURED (config 2, B200, profiles/r01_pipe_timeline.txt): layer 0 alone streams a row every 1.15 us (84 us for the 68
// rows), but every hop from layer l to l + 1 costs ~8.5 us before the first row comes out and the per-row rate decays to
// ~2.1 us deep in the chain, so the 30 layers take 434 us against ~295 us as 30 launches (553 vs 608 pairs/s).  Per-job
// trace of one hop: accumulator complete -> row published 1.3 us (drain, stores, release), published -> seen by the
// consumer 0.8 us (poll round trip + fence), seen -> MMAs issued 2.2-3.5 us (bulk copy of 33 KB from L2 + 24 MMAs), on top
// of the two row-times a 3x3 convolution is inherently behind its input.  Removing every fence changes the total by < 8 %:
// the cost is the L2 round trips themselves.  A row hand-over through L2 is as slow as a kernel boundary at this size; the
// version of this idea that could pay keeps consecutive layers in one thread-block cluster and hands rows over through
// distributed shared memory (remote mbarrier arrive instead of flag + poll), which is the lever left for these layers.
// (k_conv_chain.cu of earlier commits - all CTAs on the same layer, neighbour synchronisation between layers - was also
// slower than separate launches, 12.2 vs 10.5 us per convolution, and is replaced by this file.)
// Waits are bounded (trap after ~2 s) like every mbarrier wait in this library.  The per-layer machinery is
// k_conv_stream's SPLIT configuration: warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue, 6 MMAs (N = 192 main|corr,
// N = 96 corr) per 16-channel chunk and input row, two 192-column TMEM slots.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "stream_common.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int CP_THREADS = 224;                 // warp 0 producer, 1 MMA issuer, 2-5 epilogue, 6 publisher
constexpr int CP_EPI_WARPS = 4;
constexpr int CP_SLOTS = 2;

__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_cta_shared(int* p, int v) {
  asm volatile("red.release.cta.shared::cta.add.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// layers[l0 .. l0 + nl) run in this launch; `done` is indexed by the chain-global CTA number
// ((l * N + n) * strips + s) * ccs + cc and must be zero when the first launch of the chain starts.
template <bool PROF>
__global__ void __launch_bounds__(CP_THREADS, 1) k_conv_pipe(const CsParams p, const CsLayer* __restrict__ layers, int l0,
                                                             int* __restrict__ done) {
  constexpr int NCOL = 96, SLOT_STRIDE = 192;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[32];
  __shared__ uint64_t bars[48];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_stored;                                 // epilogue warps x rows stored so far (4 per finished row)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_w = smem;                                     // [k16][kx][K half][2*NCOL rows][8 halfs]
  uint8_t* s_x = smem + p.w_bytes;                         // ring of [plane][chunk][XW px][8 halfs]
  uint64_t* w_full = bars;
  uint64_t* x_full = bars + 2;
  uint64_t* x_empty = x_full + p.nxs;                      // nxs <= 16
  uint64_t* s_full = bars + 36;
  uint64_t* s_empty = s_full + CP_SLOTS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx -> (layer, view, strip, channel slice), layer slowest
  int u = blockIdx.x;
  const int cc = u % p.ccs; u /= p.ccs;
  const int strip = u % p.strips; u /= p.strips;
  const int n = u % p.N;
  const int lrel = u / p.N, l = l0 + lrel;
  const int my_flag = (((l * p.N + n) * p.strips + strip) * p.ccs) + cc;
  const CsLayer& L = layers[l];
  const int x0 = strip * 128;
  const int H = p.H, NJ = H + 2;

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    s_stored = 0;
    for (int i = 0; i < p.nxs; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < CP_SLOTS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], CP_EPI_WARPS); }
    fence_barrier_init();
    mbar_expect_tx(w_full, p.w_bytes);                     // weights are constants of the pass: staged before the dependency wait
    bulk_load(s_w, L.w + (size_t)cc * (p.w_bytes / 2), p.w_bytes, w_full);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  if (threadIdx.x >= 64 && threadIdx.x < 96) s_bias[threadIdx.x - 64] = L.bias[cc * 32 + (threadIdx.x - 64)];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;
  const int d = p.dil;                                      // 1 (one comb): checked by the host
  long long t_start = 0, tw0 = 0, tw1 = 0;
  if (PROF) { t_start = clock64(); if (threadIdx.x == 0) p.prof[blockIdx.x * 24 + 22] = globaltimer_ns(); }
#define CP_WAIT(acc, bar, par) do { if (PROF) { const long long _c = clock64(); mbar_wait(bar, par); acc += clock64() - _c; } else mbar_wait(bar, par); } while (0)

  if (warp == 0) {
    // ================================ producer ================================
    // lane i < nflags owns one counter of the previous layer: (strip s - 1 + i / ccs, slice i % ccs)
    const int s_lo = strip > 0 ? strip - 1 : 0, s_hi = strip + 1 < p.strips ? strip + 1 : p.strips - 1;
    const int nflags = lrel > 0 ? (s_hi - s_lo + 1) * p.ccs : 0;
    const int* flag = done + ((((l - 1) * p.N + n) * p.strips + s_lo) * p.ccs) + lane;   // contiguous over (strip, slice)
    int seen = 0;
    const __half* in = static_cast<const __half*>(L.in.p);
    const int e = lane >> 2, q4 = lane & 3;                // entry of the job, (plane, chunk) of the entry
    const bool act = e < p.nk16;
    uint32_t it = 0;
    for (int j = 0; j < NJ; ++j) {
      // input row j - 1 of this job exists once every writer has finished j rows (rows -1 and H are the zero border)
      const int need = j < H ? j : H;
      const bool poll = __any_sync(0xffffffffu, lane < nflags && seen < need);
      if (lane < nflags && seen < need) {
        const long long t0 = clock64();
        while ((seen = ld_relaxed(flag)) < need) {          // relaxed polls, ONE acquire fence once the row is there
          if (clock64() - t0 > 4000000000ll) __trap();
        }
        fence_acq_rel_gpu();
        if (PROF) tw0 += clock64() - t0;
      }
      __syncwarp();
      if (poll) fence_proxy_async_all();                   // order the bulk copies (async proxy) after the acquires
      const int row = min(j - 1, H + p.in_pad - 1);
      if (act) {
        const uint32_t ge = it + (uint32_t)e, slot = ge % (uint32_t)p.nxs, par = ((ge / (uint32_t)p.nxs) & 1) ^ 1;
        if (q4 == 0) {
          CP_WAIT(tw1, &x_empty[slot], par);
          mbar_expect_tx(&x_full[slot], 4 * p.sub_bytes);
        }
        __syncwarp(0xfu << (e * 4));
        const __half* src = in + (size_t)n * L.in.ss + (size_t)(q4 >> 1) * L.in.lo +          // q4 = plane*2 + chunk
                            (size_t)(e * 2 + (q4 & 1)) * L.in.slice + ((ptrdiff_t)row * L.in.ws + (x0 - d)) * 8;
        bulk_load(s_x + (size_t)slot * p.slot_bytes + (size_t)q4 * p.sub_bytes, src, p.sub_bytes, &x_full[slot]);
      }
      __syncwarp();
      it += (uint32_t)p.nk16;
    }
    if (PROF) {
      long long m0 = tw0;                                   // the slowest lane's counter wait
      for (int o = 16; o; o >>= 1) m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, o));
      if (lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[0] = clock64() - t_start; q[1] = m0; q[2] = tw1; q[3] = l; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, NCOL), idesc2 = make_idesc_f16(128, 2 * NCOL);
    const uint32_t b_lbo = 2 * NCOL * 16;                   // bytes between the two K halves of a weight block
    const uint32_t wblk = 3 * 2 * 2 * NCOL * 16;            // weight bytes of one 16-channel chunk: [kx][K half][2*NCOL][8]
    const uint64_t dil16 = (uint64_t)d;
    const uint64_t a_desc0 = make_smem_desc(smem_u32(s_x), p.sub_bytes, 128);
    const uint64_t a_lo_off = (uint64_t)(2 * p.sub_bytes >> 4), a_slot16 = (uint64_t)(p.slot_bytes >> 4);
    const uint64_t w_desc0 = make_smem_desc(smem_u32(s_w), b_lbo, 128);
    const uint64_t wblk16 = (uint64_t)(wblk >> 4), wkx16 = (uint64_t)(2 * b_lbo >> 4);
    uint32_t slot = 0, xpar = 0, ts = 0, spar = 1;
    CP_WAIT(tw0, w_full, 0);
    for (int j = 0; j < NJ; ++j) {
      CP_WAIT(tw1, &s_empty[ts], spar);
      const uint32_t dcol = tmem_base + ts * SLOT_STRIDE;
      uint32_t acc = 0;
      uint64_t w_hi = w_desc0;
      for (int k16 = 0; k16 < p.nk16; ++k16, w_hi += wblk16) {
        CP_WAIT(tw0, &x_full[slot], xpar);
        tc_fence_after();
        const uint64_t a_hi = a_desc0 + (uint64_t)slot * a_slot16, a_lo = a_hi + a_lo_off;
        if (leader) {
          cs_issue_split(dcol, a_hi, a_lo, w_hi, dil16, wkx16, idesc, idesc2, acc);
          umma_commit(&x_empty[slot]);
        }
        __syncwarp();
        acc = 1;
        if (++slot == (uint32_t)p.nxs) { slot = 0; xpar ^= 1; }
      }
      if (leader) umma_commit(&s_full[ts]);
      __syncwarp();
      if (++ts == CP_SLOTS) { ts = 0; spar ^= 1; }
    }
    if (PROF && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[8] = clock64() - t_start; q[9] = tw0; q[10] = tw1; }
  } else if (warp < 6) {
    // ================================ epilogue ================================
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int opx = x0 + m;
    const bool col_ok = opx < p.W;
    uint32_t ts = 0, fpar = 0;
    const __half* res = L.has_res ? static_cast<const __half*>(L.res.p) : nullptr;
    __half* out = static_cast<__half*>(L.out.p);
    float a0[32], a1[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a0[c] = a1[c] = 0.f;
    const size_t o_base = (size_t)n * L.out.ss + (size_t)(cc * 4) * L.out.slice + (size_t)opx * 8;
    const size_t r_base = (size_t)n * L.res.ss + (size_t)(cc * 4) * L.res.slice + (size_t)opx * 8;
    for (int j = 0; j < NJ; ++j) {
      const int row = j - 2;                                // the output row this job completes
      const bool ok = col_ok && j >= 2;
      uint4 rh[4], rl[4];
      if (ok && res) {
        // the residual may have been written by THIS launch two layers ago: read it through L2 (ld.global.cg), never
        // L1 / .nc.  It is complete: row `row` of layer l - 1 (consumed by the previous job) needed rows <= row + 1 of it.
        const __half* rp = res + r_base + (size_t)row * L.res.ws * 8;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          rh[cb] = __ldcg(reinterpret_cast<const uint4*>(rp + (size_t)cb * L.res.slice));
          rl[cb] = __ldcg(reinterpret_cast<const uint4*>(rp + (size_t)cb * L.res.slice + L.res.lo));
        }
      }
      CP_WAIT(tw0, &s_full[ts], fpar);
      tc_fence_after();
      const uint32_t ts_cur = ts;
      if (++ts == CP_SLOTS) { ts = 0; fpar ^= 1; }
      // drain first (the finished row goes to f, the partial rows roll over), hand the slot back, then emit
      float f[32];
      cs_drain_split(lane_addr + ts_cur * SLOT_STRIDE, a0, a1, f, s_bias);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[ts_cur]);
      if (ok) {
        __half* op = out + o_base + (size_t)row * L.out.ws * 8;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          float g[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] = f[cb * 8 + q];
          if (res) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&rh[cb]);
            const __half2* l2 = reinterpret_cast<const __half2*>(&rl[cb]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = __half22float2(h2[q]), b = __half22float2(l2[q]);
              g[2 * q] += a.x + b.x; g[2 * q + 1] += a.y + b.y;
            }
          }
          if (L.relu) {
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = fmaxf(g[q], 0.f);
          }
          uint4 oh, ol;
          cs_split8(g, oh, ol);
          *reinterpret_cast<uint4*>(op + (size_t)cb * L.out.slice) = oh;
          *reinterpret_cast<uint4*>(op + (size_t)cb * L.out.slice + L.out.lo) = ol;
        }
      }
      if (j >= 2) {
        // row `row` of this warp's 32 pixels is stored: tell the publisher warp (release at CTA scope; it releases at gpu
        // scope, cumulatively, so the fence latency is not on this warp's path to the next drain)
        __syncwarp();
        if (lane == 0) red_release_cta_shared(&s_stored, 1);
      }
    }
    if (PROF && warp == 2 && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[16] = clock64() - t_start; q[17] = tw0; }
  } else if (warp == 6) {
    // ================================ publisher ================================
    // done[cta] = number of finished rows, as soon as all four epilogue warps have stored them
    if (lane == 0) {
      int pub = 0;
      long long g0 = 0;
      while (pub < H) {
        const long long t0 = clock64();
        int rows;
        while ((rows = ld_acquire_cta_shared(&s_stored) / CP_EPI_WARPS) <= pub) {
          if (clock64() - t0 > 4000000000ll) __trap();
        }
        pub = rows;
        st_release(done + my_flag, pub);
        if (PROF && g0 == 0) g0 = globaltimer_ns();
      }
      if (PROF) { long long* q = p.prof + blockIdx.x * 24; q[20] = g0; q[21] = globaltimer_ns(); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
// Plans a pipeline for tensors of `in`'s geometry with cin = cout = ch (multiple of 32), 3x3, stride 1, dilation 1.
cudaError_t conv_pipe_plan(CsPlan* plan, const Tens& in, int ch, int num_sms) {
  if (ch % 32 || ch / 16 > 8 || in.planes != 2 || in.pad < 1 || in.d != 1 || in.cb * 8 != ch) return cudaErrorInvalidValue;
  cudaError_t e = conv_stream_plan(plan, in, ch, ch, 1, 1, num_sms);
  if (e != cudaSuccess) return e;
  CsParams& p = plan->p;
  if (!p.split || p.nxs < p.nk16 + 1 || p.ccs * 3 > 32) return cudaErrorInvalidValue;
  p.nslots = CP_SLOTS; p.tmem_cols = 512; p.ostride = 1;
  return cudaSuccess;
}

// Layers per launch for N samples: all CTAs of a launch should be co-resident (one per SM).  0: not worth it (fewer
// than two layers fit), the caller launches layer by layer.
int conv_pipe_layers_per_launch(const CsPlan& plan, int N, int nlayers) {
  const int upl = plan.p.ccs * N * plan.p.strips;
  const int lmax = plan.num_sms / upl;
  if (lmax < 2) return 0;
  const int nseg = cdiv(nlayers, lmax);
  return cdiv(nlayers, nseg);
}

cudaError_t launch_conv_pipe(const CsPlan& plan, int N, const CsLayer* d_layers, int nlayers, int* d_done, cudaStream_t st) {
  static bool attr_done[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 31]) {
    cudaFuncSetAttribute(k_conv_pipe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(k_conv_pipe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    attr_done[dev & 31] = true;
  }
  CsParams p = plan.p;
  p.N = N;
  const int upl = p.ccs * N * p.strips;
  const int lpl = conv_pipe_layers_per_launch(plan, N, nlayers);
  if (lpl < 2) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(d_done, 0, (size_t)nlayers * upl * sizeof(int), st);
  if (e != cudaSuccess) return e;
  static const int prof = getenv("SNB_TC_PROF") ? atoi(getenv("SNB_TC_PROF")) : 0;
  static long long* d_prof = nullptr;
  if (prof && !d_prof) cudaMalloc(&d_prof, 256 * 24 * sizeof(long long));
  for (int l0 = 0; l0 < nlayers; l0 += lpl) {
    const int nl = std::min(lpl, nlayers - l0), grid = nl * upl;
    if (!prof) {
      e = launch_k(k_conv_pipe<false>, grid, CP_THREADS, plan.smem, st, p, d_layers, l0, d_done);
      if (e != cudaSuccess) return e;
      continue;
    }
    // diagnostics only: per-role cycle counters, synchronous read-back
    p.prof = d_prof;
    cudaMemsetAsync(d_prof, 0, 256 * 24 * sizeof(long long), st);
    e = launch_k(k_conv_pipe<true>, grid, CP_THREADS, plan.smem, st, p, d_layers, l0, d_done);
    cudaStreamSynchronize(st);
    std::vector<long long> h(grid * 24);
    cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx[24] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 24; ++k) mx[k] = std::max(mx[k], h[b * 24 + k]);
    const long long* first = &h[0]; const long long* last = &h[(size_t)(grid - 1) * 24];
    fprintf(stderr, "[cpprof] H%d W%d C%d N%d layers %d-%d grid %d nxs %d | producer total max %lld flag-wait max %lld x_empty max %lld | "
            "issuer total max %lld wait_x max %lld wait_slot max %lld | epilogue total max %lld wait_full max %lld | "
            "first layer: issuer %lld epilogue %lld; last layer: producer flag-wait %lld issuer %lld epilogue %lld\n",
            p.H, p.W, p.nk16 * 16, N, l0, l0 + nl - 1, grid, p.nxs, mx[0], mx[1], mx[2], mx[8], mx[9], mx[10], mx[16], mx[17],
            first[8], first[16], last[1], last[8], last[16]);
    long long gmin = h[22];
    for (int b = 0; b < grid; ++b) gmin = std::min(gmin, h[(size_t)b * 24 + 22]);
    fprintf(stderr, "[cpprof] timeline (us after the first CTA started; CTA of view 0, strip 0, slice 0): layer start / first row published / last row published\n");
    for (int l = 0; l < nl; l += (l + 1 < 4 || l + 4 >= nl) ? 1 : 4) {
      const long long* q = &h[(size_t)l * upl * 24];
      fprintf(stderr, "[cpprof]   layer %2d: %8.1f %8.1f %8.1f\n", l0 + l, (q[22] - gmin) * 1e-3, (q[20] - gmin) * 1e-3, (q[21] - gmin) * 1e-3);
    }
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace snb
