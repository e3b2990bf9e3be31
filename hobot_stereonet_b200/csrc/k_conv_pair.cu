// One BasicBlock of the 1/8-resolution backbone (conv_a -> ReLU -> conv_b -> + x -> ReLU, C -> C channels, C = 64) as ONE
// launch of thread-block CLUSTERS: the intermediate tensor never leaves the chip (SNB_PREC_TC_F16X2).  OPT-IN EXPERIMENT
// (SNB_FLAG_PAIR / SNB_PAIR=1): correct and parity-tested, not faster than two k_conv_stream launches at one pair per pass.
//
// layer2 is 15 such blocks = 30 convolutions at 68 x 120 px x 2 views (config 2).  As separate k_conv_stream launches each
// convolution costs ~9.5 us in the CUDA graph for ~4 us of tensor work.  Two attempts to chain these layers through L2 (all
// CTAs on one layer with neighbour flags; a layer pipeline with row-granular flags, k_conv_pipe.cu) lost to the launches:
// a hand-over through L2 costs ~5 us.  Here the hand-over goes through DISTRIBUTED SHARED MEMORY instead:
//   cluster     2 * ccs CTAs (ccs = C / 32 output-channel slices): ranks [0, ccs) run conv_a, ranks [ccs, 2 ccs) conv_b, all
//               on the same (view, row chunk).  An a-CTA streams x rows from global memory exactly like k_conv_stream (bulk
//               copies, weights of its slice resident, 6 MMAs per 16-channel chunk into main | corr accumulators); its
//               epilogue stages every finished y row (bias, ReLU, zero outside the image, hi/lo split) in its own shared
//               memory in ring-entry layout and hands it to every b-CTA with cp.async.bulk.shared::cluster.shared::cta:
//               the bytes land in the peer's x-row RING and complete on the peer's `x_full` mbarrier (armed remotely with
//               mbarrier.arrive.expect_tx.relaxed.cluster), async proxy on both ends, no generic-proxy fences on the path.
//               A b-CTA has no bulk-copy producer: its issuer consumes the ring as if the rows had come from global memory;
//               its epilogue adds the residual x and stores the block's output.
//   credits     a ring row (= a staging row) may be overwritten once every b-CTA has consumed it: warp 0 of a b-CTA (idle
//               otherwise) waits for the local `x_empty` commits and arrives on the `cred[row % RING_ROWS]` barrier of the
//               a-CTA that owns the entry; the a-epilogue waits for that barrier before it writes row r >= RING_ROWS.
//   unit        (view, chunk of rpc output rows): conv_b needs y rows i0-1 .. i0+rpc, so conv_a runs rpc + 4 jobs on x rows
//               i0-2 .. i0+rpc+1 and conv_b rpc + 2 jobs, three jobs behind conv_a.
// MEASURED (config 2, B200): 33 clusters of 4 can be resident (cudaOccupancyMaxActiveClusters); with 28 clusters (rpc 5) a
// block takes 15 us in-kernel (role counters: conv_a issues its 9 jobs in 23 k cycles, conv_b waits 13.7 k of its 28 k for
// rows: the two-jobs-plus-hop lag and the 4 halo rows are structural) against 2 x 7 us as two launches.  First version
// (st.shared::cluster stores + fence.proxy.async + release arrives from one thread): 56 us - a remote release-arrive costs
// ~700 cycles, four of them per row sat in the epilogue's critical path.  Inside the CUDA graph the cluster launch then costs
// ~5 us more than a plain launch (1.646 vs 1.573 ms per pass with programmatic dependent launch, 1.733 vs 1.682 without):
// a cluster needs four free SMs in one GPC, so it overlaps nothing of its predecessor's tail.  Left opt-in; what it gives the
// next round is a working DSMEM hand-over (bulk copy to a peer ring + remote credits) to build deeper fusions on.
// Restrictions (host checks, otherwise the block runs as two k_conv_stream launches): dilation 1, one 128-pixel strip
// (W <= 126), C = 64 (weights of one slice + 4 ring rows = 207 KB), input border >= 2.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "stream_common.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int CPR_THREADS = 192;
constexpr int CPR_EPI_WARPS = 4;
constexpr int CPR_SLOTS = 2;                    // TMEM slots of 192 columns (main | corr)
constexpr int CPR_RING_ROWS = 4;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// shared memory of this CTA -> shared memory of a cluster peer, completing `bytes` on the PEER's mbarrier (async proxy on both ends)
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t rdst, const void* src, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(rdst), "r"(smem_u32(src)), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t rbar, uint32_t bytes) {
  // relaxed: the bytes travel with the bulk copy (async proxy); nothing this thread wrote has to be visible to the peer
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(rbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mbarrier wait with cluster-scope acquire (the data behind it was written by another CTA of the cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

template <bool PROF>
__global__ void __launch_bounds__(CPR_THREADS, 1) k_conv_pair(const CsParams p, const CsLayer la, const CsLayer lb) {
  constexpr int NCOL = 96, SLOT_STRIDE = 192;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[32];
  __shared__ uint64_t bars[48];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_w = smem;                                     // [k16][kx][K half][2*NCOL rows][8 halfs]
  uint8_t* s_x = smem + p.w_bytes;                         // ring of [plane][chunk][XW px][8 halfs]: conv_b nxs = RING_ROWS * nk16 entries,
                                                           // conv_a nxs / 2 (x rows), followed by its y staging rows
  uint64_t* w_full = bars;
  uint64_t* x_full = bars + 2;
  uint64_t* x_empty = x_full + p.nxs;                      // nxs <= 16
  uint64_t* s_full = bars + 36;
  uint64_t* s_empty = s_full + CPR_SLOTS;
  uint64_t* cred = bars + 40;                              // [RING_ROWS], a-CTAs only: ring row r may be rewritten

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int csize = 2 * p.ccs;
  const int role = (int)rank / p.ccs, cc = (int)rank % p.ccs;           // role 0: conv_a, 1: conv_b
  const int cl = blockIdx.x / csize;                                     // cluster -> (view, row chunk)
  const int chunk = cl % p.nchunk, n = cl / p.nchunk;
  const CsLayer& L = role ? lb : la;
  const int H = p.H;
  const int i0b = chunk * p.rpc, nrb = min(H, i0b + p.rpc) - i0b;       // conv_b's output rows [i0b, i0b + nrb)
  const int i0 = role ? i0b : i0b - 1, nr = role ? nrb : nrb + 2;       // this CTA's output rows (conv_a: the y rows conv_b reads)
  const int NJ = nr + 2;                                                 // jobs = input rows i0-1 .. i0+nr
  const uint32_t nx = role ? (uint32_t)p.nxs : (uint32_t)p.nxs / 2;    // entries of this CTA's x-row ring
  uint8_t* s_y = s_x + (size_t)nx * p.slot_bytes;                        // conv_a: [RING_ROWS][2 entries] finished y rows, the
                                                                         // image of what lands in the b-CTAs' rings

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < p.nxs; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < CPR_SLOTS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], CPR_EPI_WARPS); }
    for (int i = 0; i < CPR_RING_ROWS; ++i) mbar_init(&cred[i], 2 * p.ccs);   // 2 entries per row x every b-CTA
    fence_barrier_init();
    mbar_expect_tx(w_full, p.w_bytes);                     // weights are constants of the pass: staged before the dependency wait
    bulk_load(s_w, L.w + (size_t)cc * (p.w_bytes / 2), p.w_bytes, w_full);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  if (threadIdx.x >= 64 && threadIdx.x < 96) s_bias[threadIdx.x - 64] = L.bias[cc * 32 + (threadIdx.x - 64)];
  if (!role) {
    // the border pixels of a y row (x = -1 and x >= 128) are never written by the epilogue: zero the staging rows once
    uint4* z = reinterpret_cast<uint4*>(s_y);
    const int nz = (int)((size_t)CPR_RING_ROWS * 2 * p.slot_bytes / 16);
    for (int i = threadIdx.x; i < nz; i += CPR_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                                      // every CTA's barriers exist before a peer signals them
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;
  const int d = 1;
  long long t_start = 0, tw0 = 0, tw1 = 0, tw2 = 0;
  if (PROF) t_start = clock64();
#define CPR_WAIT(acc, stmt) do { if (PROF) { const long long _c = clock64(); stmt; acc += clock64() - _c; } else { stmt; } } while (0)

  if (warp == 0) {
    if (role == 0) {
      // ================================ conv_a: bulk-copy producer (x rows from global memory) ================================
      const __half* in = static_cast<const __half*>(L.in.p);
      const int e = lane >> 2, q4 = lane & 3;              // entry of the job, (plane, chunk) of the entry
      const bool act = e < p.nk16;
      uint32_t it = 0;
      for (int j = 0; j < NJ; ++j) {
        const int row = min(i0 - 1 + j, H + p.in_pad - 1);  // >= -2: inside the zero border (host checks pad >= 2)
        if (act) {
          const uint32_t ge = it + (uint32_t)e, slot = ge % nx, par = ((ge / nx) & 1) ^ 1;
          if (q4 == 0) {
            mbar_wait(&x_empty[slot], par);
            mbar_expect_tx(&x_full[slot], 4 * p.sub_bytes);
          }
          __syncwarp(0xfu << (e * 4));
          const __half* src = in + (size_t)n * L.in.ss + (size_t)(q4 >> 1) * L.in.lo +          // q4 = plane*2 + chunk
                              (size_t)(e * 2 + (q4 & 1)) * L.in.slice + ((ptrdiff_t)row * L.in.ws - d) * 8;
          bulk_load(s_x + (size_t)slot * p.slot_bytes + (size_t)q4 * p.sub_bytes, src, p.sub_bytes, &x_full[slot]);
        }
        __syncwarp();
        it += (uint32_t)p.nk16;
      }
    } else if (lane == 0) {
      // ================================ conv_b: credit return ================================
      // entry ge = (y row jb, chunk k16) was consumed once its MMAs retired (x_empty, committed by the issuer): tell the
      // a-CTA that wrote it (slice k16 / 2) that ring row jb % RING_ROWS lost one reader
      const uint32_t cred0 = smem_u32(cred);
      for (int jb = 0; jb < NJ; ++jb)
        for (int k16 = 0; k16 < p.nk16; ++k16) {
          const uint32_t ge = (uint32_t)(jb * p.nk16 + k16);
          mbar_wait(&x_empty[ge % (uint32_t)p.nxs], (ge / (uint32_t)p.nxs) & 1);
          mbar_arrive_cluster(mapa(cred0 + 8u * (uint32_t)(jb % CPR_RING_ROWS), (uint32_t)(k16 >> 1)));
        }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (both roles) ================================
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
    CPR_WAIT(tw0, mbar_wait(w_full, 0));
    for (int j = 0; j < NJ; ++j) {
      CPR_WAIT(tw1, mbar_wait(&s_empty[ts], spar));
      const uint32_t dcol = tmem_base + ts * SLOT_STRIDE;
      uint32_t acc = 0;
      uint64_t w_hi = w_desc0;
      for (int k16 = 0; k16 < p.nk16; ++k16, w_hi += wblk16) {
        if (role) CPR_WAIT(tw2, mbar_wait_cluster(&x_full[slot], xpar)); else CPR_WAIT(tw2, mbar_wait(&x_full[slot], xpar));
        tc_fence_after();
        const uint64_t a_hi = a_desc0 + (uint64_t)slot * a_slot16, a_lo = a_hi + a_lo_off;
        if (leader) {
          cs_issue_split(dcol, a_hi, a_lo, w_hi, dil16, wkx16, idesc, idesc2, acc);
          umma_commit(&x_empty[slot]);
        }
        __syncwarp();
        acc = 1;
        if (++slot == nx) { slot = 0; xpar ^= 1; }
      }
      if (leader) umma_commit(&s_full[ts]);
      __syncwarp();
      if (++ts == CPR_SLOTS) { ts = 0; spar ^= 1; }
    }
    if (PROF && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[0] = clock64() - t_start; q[1] = tw0; q[2] = tw1; q[3] = tw2; q[4] = role; }
  } else {
    // ================================ epilogue ================================
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const bool col_ok = m < p.W;
    uint32_t ts = 0, fpar = 0;
    const __half* res = role ? static_cast<const __half*>(L.res.p) : nullptr;
    __half* out = static_cast<__half*>(L.out.p);
    float a0[32], a1[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) a0[c] = a1[c] = 0.f;
    const size_t o_base = (size_t)n * L.out.ss + (size_t)(cc * 4) * L.out.slice + (size_t)m * 8;
    const size_t r_base = (size_t)n * L.res.ss + (size_t)(cc * 4) * L.res.slice + (size_t)m * 8;
    const uint32_t ring0 = smem_u32(s_x), xfull0 = smem_u32(x_full);
    for (int j = 0; j < NJ; ++j) {
      const int row = i0 + j - 2;                           // the output row this job completes
      const bool in_img = j >= 2 && row >= 0 && row < H;
      const bool ok = col_ok && in_img;
      uint4 rh[4], rl[4];
      if (role && ok) {                                     // residual prefetch while the job's MMAs finish
        const __half* rp = res + r_base + (size_t)row * L.res.ws * 8;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          rh[cb] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)cb * L.res.slice));
          rl[cb] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)cb * L.res.slice + L.res.lo));
        }
      }
      CPR_WAIT(tw0, mbar_wait(&s_full[ts], fpar));
      tc_fence_after();
      const uint32_t ts_cur = ts;
      if (++ts == CPR_SLOTS) { ts = 0; fpar ^= 1; }
      float f[32];
      cs_drain_split(lane_addr + ts_cur * SLOT_STRIDE, a0, a1, f, s_bias);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[ts_cur]);
      if (role) {
        if (ok) {
          __half* op = out + o_base + (size_t)row * L.out.ws * 8;
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            float g[8];
            const __half2* h2 = reinterpret_cast<const __half2*>(&rh[cb]);
            const __half2* l2 = reinterpret_cast<const __half2*>(&rl[cb]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = __half22float2(h2[q]), b = __half22float2(l2[q]);
              g[2 * q] = fmaxf(f[cb * 8 + 2 * q] + (a.x + b.x), 0.f);
              g[2 * q + 1] = fmaxf(f[cb * 8 + 2 * q + 1] + (a.y + b.y), 0.f);
            }
            uint4 oh, ol;
            cs_split8(g, oh, ol);
            *reinterpret_cast<uint4*>(op + (size_t)cb * L.out.slice) = oh;
            *reinterpret_cast<uint4*>(op + (size_t)cb * L.out.slice + L.out.lo) = ol;
          }
        }
      } else if (j >= 2) {
        // y row jb of this chunk -> ring row jb % RING_ROWS of every b-CTA, entries 2 cc and 2 cc + 1 (this slice's 32
        // channels), pixel m at ring position m + 1.  Rows outside the image and columns >= W are conv_b's zero padding.
        const int jb = j - 2;
        if (jb >= CPR_RING_ROWS) {
          if (lane == 0) CPR_WAIT(tw1, mbar_wait_cluster(&cred[jb % CPR_RING_ROWS], (uint32_t)((jb / CPR_RING_ROWS) - 1) & 1));
          __syncwarp();
        }
        // stage the row locally in ring-entry layout ([entry][plane][half][px][8]), then two bulk copies per b-CTA
        uint8_t* srow = s_y + (size_t)(jb % CPR_RING_ROWS) * 2 * p.slot_bytes;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          float g[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) g[q] = ok ? fmaxf(f[cb * 8 + q], 0.f) : 0.f;
          uint4 oh, ol;
          cs_split8(g, oh, ol);
          uint8_t* e = srow + (size_t)(cb >> 1) * p.slot_bytes + (size_t)(cb & 1) * p.sub_bytes + (size_t)(m + 1) * 16;
          *reinterpret_cast<uint4*>(e) = oh;                               // plane 0 (hi): sub-blocks 0, 1
          *reinterpret_cast<uint4*>(e + 2 * p.sub_bytes) = ol;             // plane 1 (lo): sub-blocks 2, 3
        }
        fence_proxy_async();                                  // the bulk copies read the staged row through the async proxy
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // 2 entries x ccs b-CTAs hand-overs, spread over the first lanes of the four warps (a remote arrive is slow)
        const int hw = (warp & 3) * 32 + lane;
        if (lane < 2 && (warp & 3) < p.ccs) {
          const int t = warp & 3, e2 = lane;
          (void)hw;
          const uint32_t rrow = (uint32_t)(jb % CPR_RING_ROWS) * (uint32_t)p.nk16;
          const uint32_t rring = mapa(ring0, (uint32_t)(p.ccs + t)), rxf = mapa(xfull0, (uint32_t)(p.ccs + t));
          const uint32_t slot = rrow + (uint32_t)(2 * cc + e2);
          mbar_expect_tx_cluster(rxf + 8u * slot, p.slot_bytes);
          bulk_copy_to_peer(rring + slot * p.slot_bytes, srow + (size_t)e2 * p.slot_bytes, p.slot_bytes, rxf + 8u * slot);
        }
      }
    }
  }

  if (PROF && warp == 2 && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[8] = clock64() - t_start; q[9] = tw0; q[10] = tw1; }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // nobody's shared memory disappears while a peer may still touch it
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
cudaError_t conv_pair_plan(CsPlan* plan, const Tens& in, int ch, int num_sms) {
  if (ch != 64 || in.planes != 2 || in.pad < 2 || in.d != 1 || in.cb * 8 != ch || in.w > 126) return cudaErrorInvalidValue;
  cudaError_t e = conv_stream_plan(plan, in, ch, ch, 1, 1, num_sms);
  if (e != cudaSuccess) return e;
  CsParams& p = plan->p;
  if (!p.split || p.strips != 1 || p.nxs < CPR_RING_ROWS * p.nk16 || 2 * p.ccs > 8) return cudaErrorInvalidValue;
  p.nxs = CPR_RING_ROWS * p.nk16;
  p.nslots = CPR_SLOTS; p.tmem_cols = 512; p.ostride = 1;
  plan->smem = 128 + (size_t)p.w_bytes + (size_t)p.nxs * p.slot_bytes;
  return cudaSuccess;
}

cudaError_t launch_conv_pair(const CsPlan& plan, int N, const CsLayer& la, const CsLayer& lb, cudaStream_t st) {
  static bool attr_done[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 31]) {
    cudaFuncSetAttribute(k_conv_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(k_conv_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    attr_done[dev & 31] = true;
  }
  CsParams p = plan.p;
  p.N = N;
  const int csize = 2 * p.ccs;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CPR_THREADS); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // clusters that can be resident at once (one CTA per SM, a cluster inside one GPC): one wave of them, not 1.06
  static int max_clusters[32] = {};
  if (!max_clusters[dev & 31]) {
    cfg.gridDim = dim3((unsigned)(csize * 64));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_conv_pair<false>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = plan.num_sms / csize / 2; }
    max_clusters[dev & 31] = n;
    if (getenv("SNB_TC_PROF")) fprintf(stderr, "[pair] cluster size %d: %d clusters resident at once\n", csize, n);
  }
  static const int env_clusters = getenv("SNB_PAIR_CLUSTERS") ? atoi(getenv("SNB_PAIR_CLUSTERS")) : 0;
  const int cap = env_clusters > 0 ? env_clusters : max_clusters[dev & 31];
  int nchunk = cap / N;
  nchunk = std::max(1, std::min(nchunk, cdiv(p.H, 2)));
  p.rpc = cdiv(p.H, nchunk);
  p.nchunk = cdiv(p.H, p.rpc);
  p.total_units = N * p.nchunk;
  cfg.gridDim = dim3((unsigned)(p.total_units * csize));
  static const int prof = getenv("SNB_TC_PROF") ? atoi(getenv("SNB_TC_PROF")) : 0;
  if (!prof) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv_pair<false>, p, la, lb);
    return e != cudaSuccess ? e : cudaGetLastError();
  }
  // diagnostics only: per-role cycle counters, synchronous read-back, max over the CTAs of each role
  static long long* d_prof = nullptr;
  if (!d_prof) cudaMalloc(&d_prof, 256 * 24 * sizeof(long long));
  p.prof = d_prof;
  cudaMemsetAsync(d_prof, 0, 256 * 24 * sizeof(long long), st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv_pair<true>, p, la, lb);
  cudaStreamSynchronize(st);
  const int grid = p.total_units * csize;
  std::vector<long long> h((size_t)grid * 24);
  cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx[2][24] = {};
  for (int b = 0; b < grid; ++b) { const int r = (b % csize) / p.ccs; for (int k = 0; k < 24; ++k) mx[r][k] = std::max(mx[r][k], h[(size_t)b * 24 + k]); }
  fprintf(stderr, "[cprprof] H%d W%d C%d N%d rpc %d clusters %d | conv_a issuer total %lld wait_w %lld wait_slot %lld wait_x %lld, epilogue total %lld wait_full %lld wait_credit %lld | "
          "conv_b issuer total %lld wait_w %lld wait_slot %lld wait_x %lld, epilogue total %lld wait_full %lld\n", p.H, p.W, p.nk16 * 16, N, p.rpc, p.total_units,
          mx[0][0], mx[0][1], mx[0][2], mx[0][3], mx[0][8], mx[0][9], mx[0][10], mx[1][0], mx[1][1], mx[1][2], mx[1][3], mx[1][8], mx[1][9]);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace snb
