// A CHAIN of same-shape streaming convolutions in ONE persistent launch (SNB_PREC_TC_F16X2).
//
// The 1/8-resolution backbone (layer2: 15 identity BasicBlocks = 30 convolutions 64 -> 64 at 68 x 120 px x 2 views at
// config 2) carries ~4 us of tensor work per convolution and paid ~7 us of launch, prologue, weight staging and
// first-row latency per launch (profiles/r01_final_summary.md).  Here every CTA keeps its unit (channel slice, view,
// row chunk) for the whole chain and the layers are separated only by NEIGHBOUR synchronisation:
//   * after its last store of layer l a CTA publishes `done[unit] = l + 1` (threadfence + release store);
//   * before streaming the input rows of layer l + 1 the producer warp acquires `done >= l + 1` of the (at most
//     ccs x 3 x 3) units that wrote the rows and columns it reads - no grid-wide barrier;
//   * weights are double-buffered in shared memory: layer l + 1's slice is bulk-copied while layer l computes;
//   * TMEM, mbarriers and the x-row ring live across layers.
// All CTAs must be co-resident (grid = units <= #SM, one CTA per SM), which the host checks; otherwise the layers run as
// separate k_conv_stream launches.  Waits are bounded (trap after ~2 s) like every mbarrier wait in this library.
// MEASURED (config 2, B200): correct (same EPE), but 12.2 us per convolution against ~10.5 us for one k_conv_stream launch
// each inside the CUDA graph: store -> threadfence -> release -> acquire poll -> proxy fence -> bulk copy is as long as a
// kernel boundary, and every input row depends on the sibling channel-slice CTA, so nothing overlaps.  Kept opt-in
// (SNB_FLAG_CHAIN) as the record of that experiment; the default path launches the layers separately.
// The per-layer pipeline is k_conv_stream's: warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue, 9 MMAs (N = 96)
// per 16-channel chunk and input row into one of 5 TMEM slots, every drained job emits one finished output row.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "stream_common.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int CH_THREADS = 192;
constexpr int CH_EPI_WARPS = 4;
constexpr int CH_SLOTS = 5;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(CH_THREADS, 1) k_conv_chain(const CsParams p, const CsLayer* __restrict__ layers, int nlayers,
                                                              int* __restrict__ done) {
  constexpr int NCOL = 96, SLOT_STRIDE = 96;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[32];
  __shared__ uint64_t bars[56];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_w = smem;                                     // two weight buffers
  uint8_t* s_x = smem + 2 * (size_t)p.w_bytes;             // ring of [plane][chunk][XW px][8 halfs]
  uint64_t* w_full = bars;                                 // [2]
  uint64_t* w_empty = bars + 2;                            // [2]
  uint64_t* x_full = bars + 4;
  uint64_t* x_empty = x_full + p.nxs;                      // nxs <= 16
  uint64_t* s_full = bars + 40;
  uint64_t* s_empty = s_full + CH_SLOTS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CsUnit un = cs_decode(p, blockIdx.x);              // one unit per CTA for the whole chain
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < p.nxs; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < CH_SLOTS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], CH_EPI_WARPS); }
    fence_barrier_init();
    if (un.nr > 0) {                                       // layer 0's weights: constants, staged before the dependency wait
      mbar_expect_tx(&w_full[0], p.w_bytes);
      bulk_load(s_w, layers[0].w + (size_t)un.cc * (p.w_bytes / 2), p.w_bytes, &w_full[0]);
    }
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;
  const int d = p.dil;
  const uint32_t wblk = 3 * 2 * 2 * NCOL * 16;              // weight bytes of one 16-channel chunk: [kx][K half][2*NCOL][8]

  if (un.nr <= 0 && threadIdx.x == 0) st_release(done + blockIdx.x, nlayers);   // an empty unit never blocks its neighbours
  if (un.nr > 0) {
    if (warp == 0) {
      // ================================ producer ================================
      uint32_t slot = 0, xpar = 1;
      // units whose output this unit reads: every channel slice, row chunks +-1, strips +-1 (same view, same comb)
      const int chunk = un.i0 / p.rpc, strip = un.x0 / 128;
      for (int l = 0; l < nlayers; ++l) {
        const CsLayer& L = layers[l];
        if (lane == 0 && l + 1 < nlayers) {                // next layer's weights into the other buffer
          const int b = (l + 1) & 1;
          mbar_wait(&w_empty[b], (((l + 1) >> 1) & 1) ^ 1);
          mbar_expect_tx(&w_full[b], p.w_bytes);
          bulk_load(s_w + (size_t)b * p.w_bytes, layers[l + 1].w + (size_t)un.cc * (p.w_bytes / 2), p.w_bytes, &w_full[b]);
        }
        if (l > 0) {                                       // neighbour synchronisation: inputs of layer l are outputs of layer l-1
          for (int i = lane; i < p.ccs * 9; i += 32) {
            const int cc = i / 9, dc = (i % 9) / 3 - 1, ds = i % 3 - 1;
            const int nc = chunk + dc, ns = strip + ds;
            if (nc < 0 || nc >= p.nchunk || ns < 0 || ns >= p.strips) continue;
            const int nu = ((((cc * p.N + un.n) * p.D + un.d) * p.strips + ns) * p.dil + un.c) * p.nchunk + nc;
            const long long t0 = clock64();
            while (ld_acquire(done + nu) < l) {
              if (clock64() - t0 > 4000000000ll) __trap();
            }
          }
          __syncwarp();
          fence_proxy_async_all();                         // order the bulk copies (async proxy) after the acquires
        }
        const __half* in = static_cast<const __half*>(L.in.p);
        for (int j = 0; j < un.nr + 2; ++j) {
          const int row = min(un.c + d * (un.i0 - 1 + j), p.H + p.in_pad - 1);
          for (int k16 = 0; k16 < p.nk16; ++k16) {
            if (lane == 0) {
              mbar_wait(&x_empty[slot], xpar);
              mbar_expect_tx(&x_full[slot], 4 * p.sub_bytes);
            }
            __syncwarp();
            if (lane < 4) {                                 // lane = plane*2 + chunk
              const __half* src = in + (size_t)un.n * L.in.ss + (size_t)(lane >> 1) * L.in.lo +
                                  (size_t)(k16 * 2 + (lane & 1)) * L.in.slice + ((ptrdiff_t)row * L.in.ws + (un.x0 - d)) * 8;
              bulk_load(s_x + (size_t)slot * p.slot_bytes + (size_t)lane * p.sub_bytes, src, p.sub_bytes, &x_full[slot]);
            }
            if (++slot == (uint32_t)p.nxs) { slot = 0; xpar ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_f16(128, NCOL);
      const uint32_t b_lbo = 2 * NCOL * 16;
      const uint64_t dil16 = (uint64_t)d;
      const uint64_t a_desc0 = make_smem_desc(smem_u32(s_x), p.sub_bytes, 128);
      const uint64_t a_lo_off = (uint64_t)(2 * p.sub_bytes >> 4), a_slot16 = (uint64_t)(p.slot_bytes >> 4);
      const uint64_t w_desc0 = make_smem_desc(smem_u32(s_w), b_lbo, 128);
      const uint64_t wblk16 = (uint64_t)(wblk >> 4), wkx16 = (uint64_t)(2 * b_lbo >> 4), wbuf16 = (uint64_t)(p.w_bytes >> 4);
      uint32_t slot = 0, xpar = 0, ts = 0, spar = 1;
      for (int l = 0; l < nlayers; ++l) {
        const int b = l & 1;
        mbar_wait(&w_full[b], (l >> 1) & 1);
        const uint64_t w_l = w_desc0 + (uint64_t)b * wbuf16;
        for (int j = 0; j < un.nr + 2; ++j) {
          mbar_wait(&s_empty[ts], spar);
          const uint32_t dcol = tmem_base + ts * SLOT_STRIDE;
          uint32_t acc = 0;
          uint64_t w_hi = w_l;
          for (int k16 = 0; k16 < p.nk16; ++k16, w_hi += wblk16) {
            mbar_wait(&x_full[slot], xpar);
            tc_fence_after();
            const uint64_t a_hi = a_desc0 + (uint64_t)slot * a_slot16, a_lo = a_hi + a_lo_off;
            if (leader) {
              umma_f16(dcol, a_hi, w_hi, idesc, acc);
              umma_f16_acc(dcol, a_hi, w_hi + NCOL, idesc);
              umma_f16_acc(dcol, a_lo, w_hi, idesc);
              umma_f16_acc(dcol, a_hi + dil16, w_hi + wkx16, idesc);
              umma_f16_acc(dcol, a_hi + dil16, w_hi + wkx16 + NCOL, idesc);
              umma_f16_acc(dcol, a_lo + dil16, w_hi + wkx16, idesc);
              umma_f16_acc(dcol, a_hi + 2 * dil16, w_hi + 2 * wkx16, idesc);
              umma_f16_acc(dcol, a_hi + 2 * dil16, w_hi + 2 * wkx16 + NCOL, idesc);
              umma_f16_acc(dcol, a_lo + 2 * dil16, w_hi + 2 * wkx16, idesc);
              umma_commit(&x_empty[slot]);
            }
            __syncwarp();
            acc = 1;
            if (++slot == (uint32_t)p.nxs) { slot = 0; xpar ^= 1; }
          }
          if (leader) umma_commit(&s_full[ts]);
          __syncwarp();
          if (++ts == CH_SLOTS) { ts = 0; spar ^= 1; }
        }
        if (leader) umma_commit(&w_empty[b]);              // this weight buffer may be refilled once the MMAs above retired
        __syncwarp();
      }
    } else {
      // ================================ epilogue ================================
      const int m = (warp & 3) * 32 + lane;
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const int opx = un.x0 + m;
      const bool col_ok = opx < p.W;
      uint32_t ts = 0, fpar = 0;
      for (int l = 0; l < nlayers; ++l) {
        const CsLayer& L = layers[l];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x - 64 < 32) s_bias[threadIdx.x - 64] = L.bias[un.cc * 32 + (threadIdx.x - 64)];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const __half* res = L.has_res ? static_cast<const __half*>(L.res.p) : nullptr;
        __half* out = static_cast<__half*>(L.out.p);
        float a0[32], a1[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a0[c] = a1[c] = 0.f;
        const size_t o_base = (size_t)un.n * L.out.ss + (size_t)(un.cc * 4) * L.out.slice + (size_t)opx * 8;
        const size_t r_base = (size_t)un.n * L.res.ss + (size_t)(un.cc * 4) * L.res.slice + (size_t)opx * 8;
        for (int j = 0; j < un.nr + 2; ++j) {
          const int row = un.c + d * (un.i0 - 2 + j);
          const bool ok = col_ok && j >= 2 && row < p.H;
          uint4 rh[4], rl[4];
          if (ok && res) {
            // the residual was written by THIS kernel two layers ago: read it through L2 (ld.global.cg), never L1/.nc
            const __half* rp = res + r_base + (size_t)row * L.res.ws * 8;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              rh[cb] = __ldcg(reinterpret_cast<const uint4*>(rp + (size_t)cb * L.res.slice));
              rl[cb] = __ldcg(reinterpret_cast<const uint4*>(rp + (size_t)cb * L.res.slice + L.res.lo));
            }
          }
          mbar_wait(&s_full[ts], fpar);
          tc_fence_after();
          const uint32_t ts_cur = ts;
          if (++ts == CH_SLOTS) { ts = 0; fpar ^= 1; }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float v0[16], v1[16], v2[16];
            const uint32_t col = lane_addr + ts_cur * SLOT_STRIDE + hf * 16;
            cs_ld3x16(col, col + 32, col + 64, v0, v1, v2);
            if (hf == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s_empty[ts_cur]);
            }
            if (ok) {
              __half* op = out + o_base + (size_t)row * L.out.ws * 8;
#pragma unroll
              for (int jb = 0; jb < 2; ++jb) {
                const int cb = hf * 2 + jb;
                float f[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = a0[hf * 16 + jb * 8 + q] + v2[jb * 8 + q];
                if (res) {
                  const __half2* h2 = reinterpret_cast<const __half2*>(&rh[cb]);
                  const __half2* l2 = reinterpret_cast<const __half2*>(&rl[cb]);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float2 a = __half22float2(h2[q]), b = __half22float2(l2[q]);
                    f[2 * q] += a.x + b.x; f[2 * q + 1] += a.y + b.y;
                  }
                }
                if (L.relu) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.f);
                }
                uint4 oh, ol;
                cs_split8(f, oh, ol);
                *reinterpret_cast<uint4*>(op + (size_t)cb * L.out.slice) = oh;
                *reinterpret_cast<uint4*>(op + (size_t)cb * L.out.slice + L.out.lo) = ol;
              }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) { a0[hf * 16 + c] = a1[hf * 16 + c] + v1[c]; a1[hf * 16 + c] = v0[c] + s_bias[hf * 16 + c]; }
          }
        }
        // publish: every store of this layer is visible device-wide before the counter moves
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) st_release(done + blockIdx.x, l + 1);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
// Plans a chain for tensors of `in`'s geometry with cin = cout (multiple of 32), 3x3, stride 1.
cudaError_t conv_chain_plan(CsPlan* plan, const Tens& in, int ch, int dil, int num_sms) {
  if (ch % 32 || in.planes != 2 || in.pad < dil || in.d != 1 || in.cb * 8 != ch) return cudaErrorInvalidValue;
  *plan = CsPlan();
  CsParams& p = plan->p;
  p.in = view(in);
  p.D = 1; p.H = in.h; p.W = in.w; p.dil = dil; p.kz = 1; p.nk16 = ch / 16; p.in_pad = in.pad;
  p.nco = 32; p.ccs = ch / 32; p.ncb_out = 4; p.nbias = 32; p.ostride = 1;
  p.XW = 128 + 2 * dil;
  p.sub_bytes = (uint32_t)p.XW * 16;
  p.slot_bytes = 4 * p.sub_bytes;
  p.strips = cdiv(p.W, 128);
  p.w_bytes = (uint32_t)p.nk16 * (3 * 2 * 2 * 96 * 16);
  p.nslots = CH_SLOTS; p.tmem_cols = 512;
  plan->num_sms = num_sms;
  const long avail = 227L * 1024 - 2048 - 128 - 2L * p.w_bytes;
  int nxs = avail > 0 ? (int)(avail / p.slot_bytes) : 0;
  nxs = nxs > 16 ? 16 : nxs;
  if (nxs < p.nk16 + 2) return cudaErrorInvalidValue;
  p.nxs = nxs;
  plan->smem = 128 + 2 * (size_t)p.w_bytes + (size_t)nxs * p.slot_bytes;
  return cudaSuccess;
}

// Unit decomposition for N samples; false when the units do not fit one wave (the caller then launches layer by layer).
bool conv_chain_units(CsPlan* plan, int N) {
  CsParams& p = plan->p;
  p.N = N;
  const int rc_max = cdiv(p.H, p.dil);
  const long columns = (long)p.ccs * N * p.strips * p.dil;
  int nchunk = (int)(plan->num_sms / columns);
  if (nchunk < 1) return false;
  if (nchunk > cdiv(rc_max, 2)) nchunk = cdiv(rc_max, 2);
  p.rpc = cdiv(rc_max, nchunk);
  p.nchunk = cdiv(rc_max, p.rpc);
  p.total_units = (int)(columns * p.nchunk);
  return p.total_units <= plan->num_sms;
}

cudaError_t launch_conv_chain(const CsPlan& plan, const CsLayer* d_layers, int nlayers, int* d_done, cudaStream_t st) {
  static bool attr_done[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 31]) {
    cudaFuncSetAttribute(k_conv_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    attr_done[dev & 31] = true;
  }
  cudaError_t e = cudaMemsetAsync(d_done, 0, (size_t)plan.p.total_units * sizeof(int), st);
  if (e != cudaSuccess) return e;
  // plain launch (no programmatic early start): every CTA of the chain must be resident before any of them spins on a neighbour
  k_conv_chain<<<plan.p.total_units, CH_THREADS, plan.smem, st>>>(plan.p, d_layers, nlayers, d_done);
  return cudaGetLastError();
}

}  // namespace snb
