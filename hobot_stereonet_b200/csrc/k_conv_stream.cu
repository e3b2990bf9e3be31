// Streaming tcgen05 convolution (SNB_PREC_TC_F16X2): stride-1 3x3 / 3x3x3 convolutions with the weights of one
// 32-output-channel slice RESIDENT in shared memory and the input streamed row by row - the single-convolution
// sibling of k_resblock_tc.cu.  Covers SURVEY.md §8a rows M1 (layer2-4, firstconv.1), M3 (head.filter.1-4,
// conv3d_alone) and M5 (refinement conv_out).
//
//   unit          (output-channel slice, sample, depth, 128-pixel strip, comb, row chunk); a CTA walks DOWN the rows
//                 of its unit.  Per input row one "job": for every depth tap and every 16-channel chunk of the input,
//                 9 MMAs (3 kernel columns x {hi*hi, hi*lo, lo*hi}) accumulate into ONE TMEM slot of 3*NCO columns:
//                 M = 128 pixels, N = NCO output channels x 3 kernel rows, K = 16.  Row i of the input therefore
//                 contributes to output rows i-1, i, i+1; the epilogue keeps two partial output rows in registers.
//   NCO           32: C8 split-fp16 output (+ bias, optional residual, optional ReLU).
//                 16: single-output-channel convolutions (N must be a multiple of 16 at M = 128: W_hi | W_lo | W_hi
//                     for the lo plane sit in columns 0 | 1 | 2 of the group, 2 MMAs per tap instead of 3) writing an
//                     fp32 plane (+ bias, optional residual, optional ReLU).
//   stride 2      computed at stride 1, every other row / column stored (firstconv.0/.2, layer2.0): 4x the MMAs of a
//                 true strided kernel but still 2-3x faster than the CUDA-core path these small layers used before.
//   1x1           packed as the centre tap of a 3x3 (shortcut convolutions, lastconv.1).
//   pipeline      warp 0 bulk-copy producer (weights when the channel slice changes, then one ring entry per
//                 (row, depth tap, 16-channel chunk)), warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue.
//                 5 TMEM slots decouple the issuer from the epilogue; no end-of-tile burst: every drained job emits
//                 one finished output row.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"
#include "stream_common.cuh"

namespace snb {

using namespace ptx;

// weights of output slice cc: the first ccs - ccs2 slices belong to the main convolution, the rest to the second head
__device__ __forceinline__ const __half* cs_wslice(const CsParams& p, int cc) {
  const int na = p.ccs - p.ccs2;
  return cc < na ? p.w + (size_t)cc * (p.w_bytes / 2) : p.w2 + (size_t)(cc - na) * (p.w_bytes / 2);
}

constexpr int CS_THREADS = 192;
constexpr int CS_EPI_WARPS = 4;
constexpr int CS_SLOTS = 5;                     // TMEM slots of the large configuration (the co-resident one uses 2)

// SPLIT (NCO = 32, long jobs): two accumulators per job - `main` = hi*hi and `corr` = hi*lo + lo*hi - filled by
// A_hi x [W_hi | W_lo] (N = 192) and A_lo x W_hi (N = 96).  The tensor core adds into TMEM with truncation, a bias that
// grows with the chain length and the accumulator magnitude: keeping the small terms apart cuts it (config-3 shape EPE
// 6.7e-4 -> see DESIGN.md §6) and the N = 192 MMA reads the A tile once for both halves.  Costs TMEM: 2 slots, not 5.
template <int NCO, bool PROF, int MINB, bool SPLIT>
__global__ void __launch_bounds__(CS_THREADS, MINB) k_conv_stream(const CsParams p) {
  constexpr int NCOL = 3 * NCO;                // accumulator columns of one job: [ky][NCO]
  constexpr int SLOT_STRIDE = SPLIT ? 192 : (NCO == 32 ? 96 : 64);
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[32];
  __shared__ uint64_t bars[48];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_w = smem;                                     // [k16][dz][kx][chunk][2*NCOL rows][8 halfs]
  uint8_t* s_x = smem + p.w_bytes;                         // ring of [plane][chunk][XW px][8 halfs]
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + 1;
  uint64_t* x_full = bars + 2;
  uint64_t* x_empty = x_full + p.nxs;                      // nxs <= 16
  uint64_t* s_full = bars + 36;
  uint64_t* s_empty = s_full + CS_SLOTS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1); mbar_init(w_empty, 1);
    for (int i = 0; i < p.nxs; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < CS_SLOTS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], CS_EPI_WARPS); }
    fence_barrier_init();
    // the first unit's weights are constants of the pass: stage them before waiting on the previous kernel, so that
    // under programmatic dependent launch (two of these CTAs fit on one SM) the load overlaps the predecessor's tail
    int u0 = blockIdx.x;
    while (u0 < p.total_units && cs_decode(p, u0).nr <= 0) u0 += gridDim.x;
    if (u0 < p.total_units) {
      mbar_expect_tx(w_full, p.w_bytes);
      bulk_load(s_w, cs_wslice(p, cs_decode(p, u0).cc), p.w_bytes, w_full);
    }
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;
  const int d = p.dil, zpad = p.kz >> 1;
  const uint32_t wblk = 3 * 2 * 2 * NCOL * 16;              // weight bytes of one (k16, dz): [kx][chunk][2*NCOL][8]
  long long t_start = 0, tw0 = 0, tw1 = 0, tw2 = 0;
  if (PROF) t_start = clock64();
#define CS_WAIT(acc, bar, par) do { if (PROF) { const long long _c = clock64(); mbar_wait(bar, par); acc += clock64() - _c; } else mbar_wait(bar, par); } while (0)

  if (warp == 0) {
    // ================================ bulk-copy producer ================================
    const __half* in = static_cast<const __half*>(p.in.p);
    uint32_t it = 0, nw = 0;
    int cur_cc = -1;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const CsUnit un = cs_decode(p, u);
      if (un.nr <= 0) continue;
      if (un.cc != cur_cc) {                                 // (re)load this slice's weights once every MMA reading the old ones retired
        if (lane == 0 && nw > 0) {                           // (the first slice was staged in the prologue)
          CS_WAIT(tw0, w_empty, (nw & 1) ^ 1);
          mbar_expect_tx(w_full, p.w_bytes);
          bulk_load(s_w, cs_wslice(p, un.cc), p.w_bytes, w_full);
        }
        cur_cc = un.cc; ++nw;
      }
      // A ring entry costs this warp ~700 cycles when issued one at a time (mbarrier poll, expect_tx, four bulk copies) -
      // more than the ~450 cycles its MMAs take (profiles/: wait_x was a third of the issuer's time).  So the warp works
      // on G entries at once: lane group g = lane / 4 takes entry e0 + g, its first lane polls that entry's slot and arms
      // the barrier, its four lanes issue the (plane, chunk) copies.
      const int dz0 = un.d - zpad < 0 ? zpad - un.d : 0, dz1 = un.d + p.kz - 1 - zpad >= p.D ? p.D - 1 - un.d + zpad : p.kz - 1;
      const int EJ = (dz1 - dz0 + 1) * p.nk16, E = (un.nr + 2) * EJ;      // entries per job, per unit
      const int G = min(8, max(1, p.nxs / 2));
      const int g = lane >> 2, q4 = lane & 3;
      for (int e0 = 0; e0 < E; e0 += G) {
        const int e = e0 + g;
        const bool act = g < G && e < E;
        if (act) {
          const uint32_t ge = it + (uint32_t)e, slot = ge % (uint32_t)p.nxs, par = ((ge / (uint32_t)p.nxs) & 1) ^ 1;
          const int j = e / EJ, r = e - j * EJ, dz = dz0 + r / p.nk16, k16 = r % p.nk16;
          const int row = min(un.c + d * (un.i0 - 1 + j), p.H + p.in_pad - 1);
          const int zin = un.d + dz - zpad;
          if (q4 == 0) {
            CS_WAIT(tw1, &x_empty[slot], par);
            mbar_expect_tx(&x_full[slot], 4 * p.sub_bytes);
          }
          __syncwarp(0xfu << (g * 4));
          const __half* src = in + (size_t)un.n * p.in.ss + (size_t)(q4 >> 1) * p.in.lo +      // q4 = plane*2 + chunk
                              ((size_t)(k16 * 2 + (q4 & 1)) * p.D + zin) * p.in.slice + ((ptrdiff_t)row * p.in.ws + (un.x0 - d)) * 8;
          bulk_load(s_x + (size_t)slot * p.slot_bytes + (size_t)q4 * p.sub_bytes, src, p.sub_bytes, &x_full[slot]);
        }
        __syncwarp();
      }
      it += (uint32_t)E;
    }
    if (PROF && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[0] = clock64() - t_start; q[1] = tw0; q[2] = tw1; q[3] = it; }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The tensor pipe queues only a few MMAs: uniform-datapath work between two MMA groups is exposed, so the loop
    // keeps ring positions as counters (no division) and builds descriptors with adds from hoisted bases.
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, NCOL), idesc2 = make_idesc_f16(128, 2 * NCOL);
    (void)idesc2;
    const uint32_t b_lbo = 2 * NCOL * 16;                   // bytes between the two K halves of a weight block
    const uint64_t dil16 = (uint64_t)d;
    const uint64_t a_desc0 = make_smem_desc(smem_u32(s_x), p.sub_bytes, 128);     // ring slot 0, hi plane
    const uint64_t a_lo_off = (uint64_t)(2 * p.sub_bytes >> 4), a_slot16 = (uint64_t)(p.slot_bytes >> 4);
    const uint64_t w_desc0 = make_smem_desc(smem_u32(s_w), b_lbo, 128);
    const uint64_t wblk16 = (uint64_t)(wblk >> 4), wkx16 = (uint64_t)(2 * b_lbo >> 4);
    uint32_t slot = 0, xpar = 0, ts = 0, spar = 1, nw = 0;
    int cur_cc = -1;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const CsUnit un = cs_decode(p, u);
      if (un.nr <= 0) continue;
      if (un.cc != cur_cc) { CS_WAIT(tw0, w_full, nw & 1); cur_cc = un.cc; ++nw; }
      const int dz0 = un.d - zpad < 0 ? zpad - un.d : 0, dz1 = un.d + p.kz - 1 - zpad >= p.D ? p.D - 1 - un.d + zpad : p.kz - 1;
      for (int j = 0; j < un.nr + 2; ++j) {
        CS_WAIT(tw1, &s_empty[ts], spar);
        const uint32_t dcol = tmem_base + ts * SLOT_STRIDE;
        uint32_t acc = 0;
        for (int dz = dz0; dz <= dz1; ++dz) {
          uint64_t w_hi = w_desc0 + (uint64_t)dz * wblk16;
          for (int k16 = 0; k16 < p.nk16; ++k16, w_hi += (uint64_t)p.kz * wblk16) {
            CS_WAIT(tw2, &x_full[slot], xpar);
            tc_fence_after();
            const uint64_t a_hi = a_desc0 + (uint64_t)slot * a_slot16, a_lo = a_hi + a_lo_off;
            if (leader) {
              if constexpr (SPLIT) {
                cs_issue_split(dcol, a_hi, a_lo, w_hi, dil16, wkx16, idesc, idesc2, acc);
              } else if constexpr (NCO == 16) {
                // one real output channel: W_hi | W_lo share a 16-column group (columns 0, 1), the second weight block
                // holds W_hi in column 2 for the lo plane: 2 MMAs per tap, hi*hi | hi*lo | lo*hi in columns 0 | 1 | 2
                umma_f16(dcol, a_hi, w_hi, idesc, acc);
                umma_f16_acc(dcol, a_lo, w_hi + NCOL, idesc);
                umma_f16_acc(dcol, a_hi + dil16, w_hi + wkx16, idesc);
                umma_f16_acc(dcol, a_lo + dil16, w_hi + wkx16 + NCOL, idesc);
                umma_f16_acc(dcol, a_hi + 2 * dil16, w_hi + 2 * wkx16, idesc);
                umma_f16_acc(dcol, a_lo + 2 * dil16, w_hi + 2 * wkx16 + NCOL, idesc);
              } else {
                // order hi*hi, lo*hi, hi*lo: two consecutive MMAs of one accumulator on the SAME A window cost 68 cycles each
                // instead of 56 (tools/ubench/mma_two_issuer_bench.cu modes 5 / 8, profiles/r02_ubench_two_issuer.txt)
                umma_f16(dcol, a_hi, w_hi, idesc, acc);
                umma_f16_acc(dcol, a_lo, w_hi, idesc);
                umma_f16_acc(dcol, a_hi, w_hi + NCOL, idesc);
                umma_f16_acc(dcol, a_hi + dil16, w_hi + wkx16, idesc);
                umma_f16_acc(dcol, a_lo + dil16, w_hi + wkx16, idesc);
                umma_f16_acc(dcol, a_hi + dil16, w_hi + wkx16 + NCOL, idesc);
                umma_f16_acc(dcol, a_hi + 2 * dil16, w_hi + 2 * wkx16, idesc);
                umma_f16_acc(dcol, a_lo + 2 * dil16, w_hi + 2 * wkx16, idesc);
                umma_f16_acc(dcol, a_hi + 2 * dil16, w_hi + 2 * wkx16 + NCOL, idesc);
              }
              umma_commit(&x_empty[slot]);
            }
            __syncwarp();
            acc = 1;
            if (++slot == (uint32_t)p.nxs) { slot = 0; xpar ^= 1; }
          }
        }
        if (leader) umma_commit(&s_full[ts]);
        __syncwarp();
        if (++ts == (uint32_t)p.nslots) { ts = 0; spar ^= 1; }
      }
      // the next unit of this CTA needs other weights: tell the producer when the MMAs above have retired
      int un_next = u + gridDim.x;
      while (un_next < p.total_units && cs_decode(p, un_next).nr <= 0) un_next += gridDim.x;
      if (un_next < p.total_units && cs_decode(p, un_next).cc != cur_cc) {
        if (leader) umma_commit(w_empty);
        __syncwarp();
      }
    }
    if (PROF && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[8] = clock64() - t_start; q[9] = tw0; q[10] = tw1; q[11] = tw2; q[12] = 0; }
  } else {
    // ================================ epilogue ================================
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ts = 0, fpar = 0;
    int cur_cc = -1;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const CsUnit un = cs_decode(p, u);
      if (un.nr <= 0) continue;
      // which convolution this slice belongs to (second head: the 1x1 shortcut sharing the launch)
      const int na = p.ccs - p.ccs2;
      const bool h2 = un.cc >= na;
      const int ccl = h2 ? un.cc - na : un.cc;
      const float wsc = h2 ? p.wsc2 : p.wsc;
      const int relu = h2 ? p.relu2 : p.relu, ncb_out = h2 ? p.ncb_out2 : p.ncb_out;
      if (un.cc != cur_cc) {
        // all four epilogue warps use the same 32 biases; a named barrier keeps the refill ordered within the group
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // partial rows live in the accumulator domain (weights x 2^k): so does the bias; finished values are multiplied by 2^-k
        if (threadIdx.x - 64 < 32)
          s_bias[threadIdx.x - 64] = (int)(threadIdx.x - 64) < (h2 ? p.nbias2 : p.nbias) ? (h2 ? p.bias2 : p.bias)[ccl * NCO + (threadIdx.x - 64)] / wsc : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_cc = un.cc;
      }
      const int opx = un.x0 + m;
      const bool col_ok = opx < p.W;
      // MMAs that add non-zero products to one TMEM accumulator of a job: (valid depth taps) x (16-channel chunks) x kernel
      // columns carrying weights, x 3 when hi*hi, hi*lo and lo*hi share the accumulator; feeds the round-toward-zero compensation
      const int ndz = min(un.d + p.kz - 1 - zpad, p.D - 1) - max(un.d - zpad, 0) + 1;
      const float kn = p.rzk * (float)(ndz * p.nk16 * (h2 ? p.taps2 : p.taps) * ((SPLIT || NCO == 16) ? 1 : 3)) *
                       (NCO == 16 ? RZ_KAPPA_1CH_PER_MMA / RZ_KAPPA_PER_MMA : 1.f);
      if constexpr (NCO == 32) {
        const __half* res = h2 ? nullptr : static_cast<const __half*>(p.res.p);
        const TV& ov = h2 ? p.out2 : p.out;
        __half* out = static_cast<__half*>(ov.p);
        float a0[32], a1[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a0[c] = a1[c] = 0.f;
        const size_t o_base = (size_t)un.n * ov.ss + ((size_t)(ccl * 4) * p.D + un.d) * ov.slice + (size_t)(opx / p.ostride) * 8;
        const size_t r_base = (size_t)un.n * p.res.ss + ((size_t)(ccl * 4) * p.D + un.d) * p.res.slice + (size_t)opx * 8;
        for (int j = 0; j < un.nr + 2; ++j) {
          const int row = un.c + d * (un.i0 - 2 + j);       // the (stride-1) output row this job completes
          const bool ok = col_ok && j >= 2 && row < p.H && (p.ostride == 1 || !((row | opx) & 1));
          uint4 rh[4], rl[4];
          if (ok && res) {                                  // residual prefetch while the job's MMAs finish
            const __half* rp = res + r_base + (size_t)row * p.res.ws * 8;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              rh[cb] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)cb * p.D * p.res.slice));
              rl[cb] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)cb * p.D * p.res.slice + p.res.lo));
            }
          }
          CS_WAIT(tw0, &s_full[ts], fpar);
          tc_fence_after();
          const uint32_t ts_cur = ts;
          if (++ts == (uint32_t)p.nslots) { ts = 0; fpar ^= 1; }
          // drain first (the finished row goes to f, the partial rows roll over), hand the slot back, then emit
          float f[32];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float v0[16], v1[16], v2[16];
            const uint32_t col = lane_addr + ts_cur * SLOT_STRIDE + hf * 16;
            cs_ld3x16(col, col + 32, col + 64, v0, v1, v2);
            if constexpr (SPLIT) {                                             // + the corr accumulator
              float c0[16], c1[16], c2[16];
              cs_ld3x16(col + NCOL, col + NCOL + 32, col + NCOL + 64, c0, c1, c2);
#pragma unroll
              for (int c = 0; c < 16; ++c) { v0[c] += c0[c]; v1[c] += c1[c]; v2[c] += c2[c]; }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              f[hf * 16 + c] = a0[hf * 16 + c] + v2[c];
              a0[hf * 16 + c] = a1[hf * 16 + c] + v1[c];
              a1[hf * 16 + c] = v0[c] + s_bias[hf * 16 + c];
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[ts_cur]);
          if (ok) {
            __half* op = out + o_base + (size_t)(row / p.ostride) * ov.ws * 8;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              if (cb >= ncb_out) continue;                   // Cout < 32: the slice's upper channel blocks do not exist
              float g[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) g[q] = rz_comp(f[cb * 8 + q], kn) * wsc;   // truncation loss back, out of the 2^k weight scale
              if (res) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&rh[cb]);
                const __half2* l2 = reinterpret_cast<const __half2*>(&rl[cb]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 a = __half22float2(h2[q]), b = __half22float2(l2[q]);
                  g[2 * q] += a.x + b.x; g[2 * q + 1] += a.y + b.y;
                }
              }
              if (relu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) g[q] = fmaxf(g[q], 0.f);
              }
              uint4 oh, ol;
              cs_split8(g, oh, ol);
              *reinterpret_cast<uint4*>(op + (size_t)cb * p.D * ov.slice) = oh;
              *reinterpret_cast<uint4*>(op + (size_t)cb * p.D * ov.slice + ov.lo) = ol;
            }
          }
        }
      } else {
        // single output channel: columns [ky*16 + 0..2]; fp32 plane out [n][D][H][W], optional residual
        float a0 = 0.f, a1 = 0.f;
        const float bias = s_bias[0];
        // The residual of a row is one cold global load: issued just before the wait on the row's accumulator its latency
        // was the epilogue's whole period (1.8 k cycles per job against ~0.5 k of MMAs at full resolution), so the loads
        // run RES_AHEAD jobs ahead of their use.
        constexpr int RES_AHEAD = 3;
        struct ResRaw { unsigned short h, l; float f, f1, f2, f3; };    // kept as loaded: converting would wait for the load
        ResRaw rq[RES_AHEAD];
        // res_mode 3: the residual is the x2 bilinear upsample (PyTorch, align_corners = False) of the previous stage's
        // disparity [n][H/2][W/2] - the same expression k_refine_head feeds conv_in with; the column weights are per thread
        const int ch = p.H >> 1, cw = p.W >> 1;
        const float sxr = fmaxf((opx + 0.5f) * 0.5f - 0.5f, 0.f);
        const int rx0 = min((int)sxr, cw - 1), rx1 = min(rx0 + 1, cw - 1);
        const float lx1 = sxr - (float)(int)sxr, lx0 = 1.f - lx1;
        auto res_load = [&](int j) -> ResRaw {
          ResRaw q; q.h = 0; q.l = 0; q.f = 0.f; q.f1 = q.f2 = q.f3 = 0.f;
          const int row = un.c + d * (un.i0 - 2 + j);
          if (!(col_ok && j >= 2 && j < un.nr + 2 && row < p.H)) return q;
          if (p.res_mode == 3) {
            const float sy = fmaxf((row + 0.5f) * 0.5f - 0.5f, 0.f);
            const int y0 = (int)sy, y1 = min(y0 + 1, ch - 1);
            const float* rp = p.res_plane + (size_t)un.n * ch * cw;
            q.f = __ldg(rp + (size_t)y0 * cw + rx0); q.f1 = __ldg(rp + (size_t)y0 * cw + rx1);
            q.f2 = __ldg(rp + (size_t)y1 * cw + rx0); q.f3 = __ldg(rp + (size_t)y1 * cw + rx1);
            return q;
          }
          if (p.res_mode == 1) {                            // channel 0 of a C8 split-fp16 tensor
            const unsigned short* rp = reinterpret_cast<const unsigned short*>(p.res.p) + (size_t)un.n * p.res.ss + ((size_t)row * p.res.ws + opx) * 8;
            q.h = __ldg(rp); q.l = __ldg(rp + p.res.lo);
          } else if (p.res_mode == 2) {
            q.f = __ldg(p.res_plane + (((size_t)un.n * p.D + un.d) * p.H + row) * p.W + opx);
          }
          return q;
        };
#pragma unroll
        for (int k = 0; k < RES_AHEAD; ++k) rq[k] = res_load(k);
        for (int j = 0; j < un.nr + 2; ++j) {
          const int row = un.c + d * (un.i0 - 2 + j);
          const bool ok = col_ok && j >= 2 && row < p.H;
          const size_t o = (((size_t)un.n * p.D + un.d) * p.H + (ok ? row : 0)) * p.W + (ok ? opx : 0);
          const ResRaw rr = rq[0];
#pragma unroll
          for (int k = 0; k + 1 < RES_AHEAD; ++k) rq[k] = rq[k + 1];
          rq[RES_AHEAD - 1] = res_load(j + RES_AHEAD);
          CS_WAIT(tw0, &s_full[ts], fpar);
          tc_fence_after();
          float v0[16], v1[16], v2[16];
          const uint32_t col = lane_addr + ts * SLOT_STRIDE;
          cs_ld3x16(col, col + 16, col + 32, v0, v1, v2);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[ts]);
          if (++ts == (uint32_t)p.nslots) { ts = 0; fpar ^= 1; }
          if (ok) {
            float r;
            if (p.res_mode == 3) {
              const float sy = fmaxf((row + 0.5f) * 0.5f - 0.5f, 0.f);
              const float ly1 = sy - (float)(int)sy, ly0 = 1.f - ly1;
              r = ly0 * (lx0 * rr.f + lx1 * rr.f1) + ly1 * (lx0 * rr.f2 + lx1 * rr.f3);
            } else {
              r = rr.f + (__half2float(__ushort_as_half(rr.h)) + __half2float(__ushort_as_half(rr.l)));
            }
            float f = fmaf(rz_comp(a0 + (v2[0] + (v2[1] + v2[2])), kn), wsc, r);   // hi*hi + (hi*lo + lo*hi), truncation loss back, out of the 2^k weight scale
            if (p.relu) f = fmaxf(f, 0.f);
            p.out_plane[o] = f;
            // the last conv_out also emits the model output tensor (stereonet_node.cpp:1033): s32 NCHW, cropped to the valid size
            if (p.io.q && row < p.qH && opx < p.qW) p.io.q[((size_t)un.n * p.qH + row) * p.qW + opx] = __float2int_rn(f * p.qmul);
          }
          a0 = a1 + (v1[0] + (v1[1] + v1[2]));
          a1 = (v0[0] + (v0[1] + v0[2])) + bias;
        }
      }
    }
  }

  if (PROF && warp == 2 && lane == 0) { long long* q = p.prof + blockIdx.x * 24; q[16] = clock64() - t_start; q[17] = tw0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---- host side -----------------------------------------------------------------------------------
// Weight packing: [cc][k16][dz][kx][K half][2*3*NCO rows][8]: rows [W_hi: ky*NCO + co | W_lo: 3*NCO + ky*NCO + co]
// ks = 1: the single tap lands at the centre (ky = kx = 1) of an otherwise zero 3x3; cin is padded up to a multiple of 16.
void cs_pack_weights(const float* W, int cout, int cin, int kz, int ks, int NCO, int wlog2, std::vector<__half>& out) {
  const int ccs = NCO == 32 ? (cout + 31) / 32 : 1, nk16 = (cin + 15) / 16, ncol = 3 * NCO;
  out.assign((size_t)ccs * nk16 * kz * 3 * 2 * 2 * ncol * 8, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int dz = 0; dz < kz; ++dz)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            if (ks == 1 && (ky != 1 || kx != 1)) continue;
            const float v = ldexpf(ks == 1 ? W[((size_t)co * cin + ci) * kz + dz] : W[((((size_t)co * cin + ci) * kz + dz) * 3 + ky) * 3 + kx], wlog2);
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const int cc = NCO == 32 ? co / 32 : 0, cl = NCO == 32 ? co % 32 : co;
            const int k16 = ci / 16, half = (ci % 16) / 8, e = ci % 8;
            const size_t blk = ((((size_t)cc * nk16 + k16) * kz + dz) * 3 + kx) * 2 + half;
            if (NCO == 16) {     // one output channel: block 1 = [W_hi, W_lo, 0...] (x A_hi), block 2 = [0, 0, W_hi, 0...] (x A_lo)
              out[(blk * 2 * ncol + ky * NCO + 0) * 8 + e] = hi;
              out[(blk * 2 * ncol + ky * NCO + 1) * 8 + e] = lo;
              out[(blk * 2 * ncol + ncol + ky * NCO + 2) * 8 + e] = hi;
              continue;
            }
            out[(blk * 2 * ncol + ky * NCO + cl) * 8 + e] = hi;
            out[(blk * 2 * ncol + ncol + ky * NCO + cl) * 8 + e] = lo;
          }
}

// in: split-fp16 C8 tensor with pad >= dil, cin % 16 == 0.  cout % 32 == 0 (C8 output) or cout == 1 (plane output).
cudaError_t conv_stream_plan(CsPlan* plan, const Tens& in, int cin, int cout, int dil, int kz, int num_sms) {
  if (cin % 16 || !(cout % 32 == 0 || cout == 1 || (cout < 32 && cout % 8 == 0)) || in.planes != 2 || in.pad < dil || dil < 1 || dil > 16 || (kz != 1 && kz != 3))
    return cudaErrorInvalidValue;
  *plan = CsPlan();
  CsParams& p = plan->p;
  p.in = view(in);
  p.D = in.d; p.H = in.h; p.W = in.w; p.dil = dil; p.kz = kz; p.nk16 = cin / 16; p.in_pad = in.pad;
  p.taps = 3;
  p.nco = cout == 1 ? 16 : 32;
  p.ccs = cout == 1 ? 1 : (cout + 31) / 32;
  p.ncb_out = cout >= 32 ? 4 : cout / 8;
  p.nbias = cout == 1 ? 1 : (cout >= 32 ? 32 : cout);
  p.XW = 128 + 2 * dil;
  p.sub_bytes = (uint32_t)p.XW * 16;
  p.slot_bytes = 4 * p.sub_bytes;
  p.strips = cdiv(p.W, 128);
  p.w_bytes = (uint32_t)(p.nk16 * kz) * (3 * 2 * 2 * 3 * p.nco * 16);
  plan->num_sms = num_sms;
  // Two CTAs per SM (<= 113 KB, 256 TMEM columns each) when the weights are small: with programmatic dependent launch
  // the next convolution's prologue and weight staging then overlap this one's tail (layer2: 31 back-to-back launches).
  auto ring = [&](long budget) { const long avail = budget - 128 - (long)p.w_bytes; int n = avail > 0 ? (int)(avail / p.slot_bytes) : 0; return n > 16 ? 16 : n; };
  // the ring must hold more than one job's entries or the producer cannot run ahead of the issuer
  const int need = std::max(4, p.nk16 + 2);
  // opt-in (SNB_STREAM_SMALL=1): measured 567 vs 574 pairs/s at config 2 - two TMEM slots instead of five cost more
  // than overlapping the next kernel's prologue gains
  static const bool small_ok = getenv("SNB_STREAM_SMALL") && atoi(getenv("SNB_STREAM_SMALL"));
  int nxs = pdl_enabled() && small_ok ? ring(112L * 1024) : 0;
  p.nslots = 2; p.tmem_cols = 256;
  if (nxs < need) { nxs = ring(227L * 1024 - 2048); p.nslots = CS_SLOTS; p.tmem_cols = 512; }
  if (nxs < std::max(4, p.nk16 + 1)) return cudaErrorInvalidValue;
  // long jobs: main | corr accumulators (better accuracy, faster N = 192 MMAs), two 192-column TMEM slots
  static const bool split_ok = !getenv("SNB_STREAM_SPLIT") || atoi(getenv("SNB_STREAM_SPLIT"));
  p.split = (split_ok && p.nco == 32 && p.nk16 * kz >= 4 && p.tmem_cols == 512) ? 1 : 0;
  if (p.split) p.nslots = 2;
  p.nxs = nxs;
  plan->smem = 128 + (size_t)p.w_bytes + (size_t)nxs * p.slot_bytes;
  return cudaSuccess;
}

template <int NCO, int MINB, bool SPLIT>
static cudaError_t cs_launch_t(CsParams p, int grid, size_t smem, cudaStream_t st) {
  static bool attr_done[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 31]) {
    cudaFuncSetAttribute(k_conv_stream<NCO, false, MINB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(k_conv_stream<NCO, true, MINB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    attr_done[dev & 31] = true;
  }
  static const int prof = getenv("SNB_TC_PROF") ? atoi(getenv("SNB_TC_PROF")) : 0;
  if (!prof) return launch_k(k_conv_stream<NCO, false, MINB, SPLIT>, grid, CS_THREADS, smem, st, p);
  // diagnostics only: per-role cycle counters, synchronous read-back, max over CTAs
  static long long* d_prof = nullptr;
  if (!d_prof) cudaMalloc(&d_prof, 256 * 24 * sizeof(long long));
  p.prof = d_prof;
  cudaMemsetAsync(d_prof, 0, 256 * 24 * sizeof(long long), st);
  cudaError_t e = launch_k(k_conv_stream<NCO, true, MINB, SPLIT>, grid, CS_THREADS, smem, st, p);
  cudaStreamSynchronize(st);
  std::vector<long long> h(grid * 24);
  cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx[24] = {0};
  for (int b = 0; b < grid; ++b) for (int k = 0; k < 24; ++k) mx[k] = std::max(mx[k], h[b * 24 + k]);
  fprintf(stderr, "[csprof] H%d W%d D%d Cin%d nco%d ccs%d kz%d dil%d N%d units %d (rpc %d) grid %d nxs %d slots %d w_bytes %u | producer total %lld wait_w_empty %lld wait_x_empty %lld entries %lld | "
          "issuer total %lld wait_w %lld wait_slot %lld wait_x %lld | epilogue total %lld wait_full %lld\n",
          p.H, p.W, p.D, p.nk16 * 16, p.nco, p.ccs, p.kz, p.dil, p.N, p.total_units, p.rpc, grid, p.nxs, p.nslots, p.w_bytes,
          mx[0], mx[1], mx[2], mx[3], mx[8], mx[9], mx[10], mx[11], mx[16], mx[17]);
  return e;
}

cudaError_t launch_conv_stream(const CsPlan& plan, int N, const void* w, int wlog2, const float* bias, const Tens* out, const Tens* res,
                               float* out_plane, const float* res_plane, int res_c8_ch0, int relu, int ostride, cudaStream_t st,
                               int res_plane_mode, const IoPtrs* io, int qH, int qW, float qmul, const CsHead2* head2) {
  CsParams p = plan.p;
  if (head2) {                                      // the shortcut convolution of the same input rides along as extra output slices
    p.ccs2 = (head2->cout + 31) / 32;
    p.ccs += p.ccs2;
    p.w2 = static_cast<const __half*>(head2->w); p.bias2 = head2->bias; p.out2 = view(*head2->out); p.relu2 = head2->relu;
    p.taps2 = head2->ks == 1 ? 1 : 3; p.wsc2 = ldexpf(1.f, -head2->wlog2);
    p.ncb_out2 = head2->cout >= 32 ? 4 : head2->cout / 8; p.nbias2 = head2->cout >= 32 ? 32 : head2->cout;
  }
  p.ostride = ostride;
  p.N = N; p.w = static_cast<const __half*>(w); p.bias = bias; p.relu = relu;
  p.rzk = rz_unit();
  p.wsc = ldexpf(1.f, -wlog2);
  if (out) p.out = view(*out);
  p.res_mode = 0;
  if (res) { p.res = view(*res); p.res_mode = res_c8_ch0 ? 1 : 0; }
  if (res_plane) { p.res_plane = res_plane; p.res_mode = res_plane_mode; }
  if (io) { p.io = *io; p.qH = qH; p.qW = qW; p.qmul = qmul; }
  p.out_plane = out_plane;
  const int rc_max = cdiv(p.H, p.dil);
  const long columns = (long)p.ccs * N * p.D * p.strips * p.dil;
  int nchunk = (int)(plan.num_sms / columns);
  if (nchunk < 1) nchunk = 1;
  if (nchunk > cdiv(rc_max, 2)) nchunk = cdiv(rc_max, 2);
  p.rpc = cdiv(rc_max, nchunk);
  p.nchunk = cdiv(rc_max, p.rpc);
  p.total_units = (int)(columns * p.nchunk);
  const int grid = p.total_units < plan.num_sms ? p.total_units : plan.num_sms;
  cudaError_t e;
  if (p.split) e = cs_launch_t<32, 1, true>(p, grid, plan.smem, st);
  else if (p.nslots == 2) e = p.nco == 32 ? cs_launch_t<32, 2, false>(p, grid, plan.smem, st) : cs_launch_t<16, 2, false>(p, grid, plan.smem, st);
  else e = p.nco == 32 ? cs_launch_t<32, 1, false>(p, grid, plan.smem, st) : cs_launch_t<16, 1, false>(p, grid, plan.smem, st);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace snb
