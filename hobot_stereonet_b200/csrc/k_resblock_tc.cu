// Fused residual block on the tcgen05 tensor cores (SNB_PREC_TC_F16X2), 32 channels:
//     out = ReLU( conv_b( ReLU( conv_a(x) + ba ) ) + bb + res )          conv_a, conv_b: 3x3, dilation d
// i.e. one edge-aware-refinement "residual_astrous_block" (hbm `_head_edge_aware_refinements_*_residual_astrous_blocks_*`)
// or one 32-channel backbone BasicBlock (`_backbone_layer1_*`), SURVEY.md §8a rows M5 / M1, in ONE pass over x:
// the intermediate activation never goes to HBM, and the block moves 8 B/element instead of 20.
//
//   streaming     a CTA owns a strip of 128-2d output columns and walks DOWN the rows of one comb (rows d apart,
//                 so vertical taps stay inside the comb).  Per row it runs two "jobs" on the tensor core:
//                   a-job(i): x row i  (smem ring, bulk-copied)  x  Wa  -> contribution to y rows i-1, i, i+1
//                   b-job(i): y row i  (smem ring, written by epilogue group A) x Wb -> out rows i-1, i, i+1
//                 GEMM view: M = 128 pixels of the input row, K = 16 channels per MMA, N = 32 channels x 3 kernel rows
//                 stacked with the hi/lo weight halves: A_hi x [W_hi | W_lo] (N = 192 -> main | corr columns) and
//                 A_lo x W_hi (N = 96 -> corr); the kernel column is a 16-byte shift of the A descriptor.  A job is
//                 12 MMAs (2 channel chunks x 3 columns x 2) chained in one 192-column TMEM slot.  (Three N = 96
//                 MMAs into one 96-column accumulator were measured 25 % slower: each re-reads the 4 KB A tile and
//                 the job becomes shared-memory-bandwidth-bound.)
//   halo          only 4 extra rows at the top of a column walk (two per conv) instead of 2 per 6-row tile twice.
//   epilogue      two warp groups of EIGHT warps, one per conv, so the y emission and the output emission overlap; a warp
//                 owns a TMEM lane quarter (32 pixels) and one 16-channel half (four warps per group left a chain of six
//                 TMEM round trips per thread and job: 61 -> 57 us at 544x960).
//                 group A (warps 2-9) drains a-jobs, keeps two partial y rows in fp32 registers; a finished row gets
//                 ReLU, zero outside the image (conv_b's zero padding), hi/lo fp16 split, and goes into the y ring
//                 in the A-operand layout.  Group B (warps 10-17) drains b-jobs; a finished output row gets the
//                 residual, ReLU, the split, and goes to HBM.  Biases are added when a partial row is born.
//   pipeline      warp 0: bulk-copy producer (weights once, then one x row per a-job); warp 1: TMEM owner + issuer of
//                 conv_a's MMAs; warp 18: issuer of conv_b's MMAs (one issuing thread left a ~300-cycle bubble between
//                 jobs: barrier polls + commits against a tensor-pipe queue of a few MMAs).
//                 TMEM: conv_a has one 192-column slot (main | corr; its drain overlaps the b-job in flight); conv_b, whose
//                 drain carries the residual and the HBM stores, has two 96-column slots with hi*hi + hi*lo + lo*hi
//                 merged in one accumulator (three N = 96 MMAs per tap: +10 % tensor time for the b-jobs).
//   shared memory is the binding resource: an M = 128 MMA reads its 4 KB A tile and N x 32 B of weights from shared memory at
//                 128 B/cycle - cost max(N/2, 32 + N/4) cycles, i.e. N = 96 is shared-memory-bound (56 = 7 KB / 128) and
//                 N = 192 tensor-bound (96 vs 80).  A row costs 1920 MMA cycles + 2 x 130 cycles of x / y row writes against
//                 ~2800 measured; more TMEM slots for conv_a make it worse (see launch_resblock_tc).
//                 Persistent over (sample, strip, comb, row-chunk) units.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int RB_THREADS = 608;
constexpr int RB_GROUP_WARPS = 8;                       // per epilogue group: TMEM lane quarter (warp & 3) x 16-channel half
constexpr int RB_ISSUER_B = 2 + 2 * RB_GROUP_WARPS;     // warp 18: issues conv_b's MMAs (warp 1 issues conv_a's)
constexpr int RB_SLOT_COLS = 192;                       // [main | corr] x [half][ky][16 ch]
constexpr int RB_WROWS = 192;                           // packed weight rows per (k16, kx, chunk): [W_hi 96 | W_lo 96]
constexpr uint32_t RB_W_BYTES = 2 * 3 * 2 * RB_WROWS * 16;   // one conv: [k16][kx][chunk][192 rows][8 halfs]
constexpr int RB_YSLOTS = 3;
constexpr int RB_BCOLS = 96;                            // conv_b: one merged accumulator per job

struct RbUnit { int n, x0, c, i0, nr; };

__device__ __forceinline__ RbUnit rb_decode(const RbParams& p, int u) {
  RbUnit r;
  const int chunk = u % p.nchunk; u /= p.nchunk;
  r.c = u % p.dil; u /= p.dil;
  const int strip = u % p.strips;
  r.n = u / p.strips;
  r.x0 = strip * p.OW;
  const int rc = r.c < p.H ? (p.H - r.c + p.dil - 1) / p.dil : 0;     // rows of this comb
  r.i0 = chunk * p.rpc;
  const int i1 = min(rc, r.i0 + p.rpc);
  r.nr = i1 - r.i0;
  return r;
}

// one kernel row of one 16-channel half: main + corr columns
__device__ __forceinline__ void rb_ld_sum(uint32_t col, float (&v)[16]) {
  float m[16], c[16];
  tmem_ld_2x16(col, col + RB_SLOT_COLS / 2, m, c);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = m[i] + c[i];
}

// one kernel row of one 16-channel half of a merged accumulator
__device__ __forceinline__ void rb_ld1(uint32_t col, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld_16(col, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void rb_split8(const float* f, uint4& oh, uint4& ol) {
  __half2* ph = reinterpret_cast<__half2*>(&oh);
  __half2* pl = reinterpret_cast<__half2*>(&ol);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hh = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
    const float2 hf = __half22float2(hh);
    ph[j] = hh;
    pl[j] = __floats2half2_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
  }
}

__device__ __forceinline__ void rb_sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// AM: conv_a like conv_b (hi*hi, hi*lo, lo*hi chained in ONE 96-column accumulator); AS / BS: TMEM slots of conv_a / conv_b
template <bool PROF, bool AM, int AS, int BS>
__global__ void __launch_bounds__(RB_THREADS, 1) k_resblock_tc(const RbParams p) {
  constexpr bool RB_A_MERGED = AM;
  constexpr int RB_ACOLS = AM ? 96 : 192, RB_ASLOTS = AS, RB_BSLOTS = BS;
  constexpr int RB_BBASE = RB_ASLOTS * RB_ACOLS;           // first TMEM column of the conv_b slots
  static_assert(RB_BBASE + RB_BSLOTS * RB_BCOLS <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[64];                 // [ba 32 | bb 32]
  __shared__ uint64_t bars[40];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_wa = smem;
  uint8_t* s_wb = smem + RB_W_BYTES;
  uint8_t* s_x = smem + 2 * RB_W_BYTES;
  uint8_t* s_y = s_x + (size_t)p.nxs * p.slot_bytes;
  uint64_t* w_full = bars;
  uint64_t* x_full = bars + 1;
  uint64_t* x_empty = x_full + p.nxs;          // nxs <= 8
  uint64_t* y_full = bars + 17;
  uint64_t* y_empty = y_full + RB_YSLOTS;
  uint64_t* sa_full = y_empty + RB_YSLOTS;
  uint64_t* sa_empty = sa_full + RB_ASLOTS;
  uint64_t* sb_full = sa_empty + RB_ASLOTS;
  uint64_t* sb_empty = sb_full + RB_BSLOTS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < p.nxs; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < RB_YSLOTS; ++i) { mbar_init(&y_full[i], RB_GROUP_WARPS); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < RB_ASLOTS; ++i) { mbar_init(&sa_full[i], 1); mbar_init(&sa_empty[i], RB_GROUP_WARPS); }
    for (int i = 0; i < RB_BSLOTS; ++i) { mbar_init(&sb_full[i], 1); mbar_init(&sb_empty[i], RB_GROUP_WARPS); }
    fence_barrier_init();
    // weights are constants of the pass: stage them before waiting on the previous kernel
    mbar_expect_tx(w_full, 2 * RB_W_BYTES);
    bulk_load(s_wa, p.wa, RB_W_BYTES, w_full);
    bulk_load(s_wb, p.wb, RB_W_BYTES, w_full);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  // biases in the accumulator domain of their convolution (weights x 2^k, common.cuh weight_scale_log2)
  if (threadIdx.x >= 64 && threadIdx.x < 128) s_bias[threadIdx.x - 64] = threadIdx.x < 96 ? p.ba[threadIdx.x - 64] / p.wsa : p.bb[threadIdx.x - 96] / p.wsb;
  // the y ring's columns >= 128 are read by shifted taps of masked pixels only; keep them finite
  for (int i = threadIdx.x; i < (int)(RB_YSLOTS * p.slot_bytes / 16); i += RB_THREADS)
    reinterpret_cast<uint4*>(s_y)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();          // the next kernel may start its prologue
  pdl_wait();             // everything above touched only shared memory / TMEM / constant weights
  const uint32_t tmem_base = tmem_slot;
  const int d = p.dil;
  long long t_start = 0, tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0;
  if (PROF) t_start = clock64();
#define RB_WAIT(acc, bar, par) do { if (PROF) { const long long _c = clock64(); mbar_wait(bar, par); acc += clock64() - _c; } else mbar_wait(bar, par); } while (0)

  if (warp == 0) {
    // ================================ bulk-copy producer ================================
    const __half* in = static_cast<const __half*>(p.in.p);
    uint32_t itx = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const RbUnit un = rb_decode(p, u);
      if (un.nr <= 0) continue;
      const __half* src0 = in + (size_t)un.n * p.in.ss + (size_t)((lane >> 2) & 1) * p.in.lo + (size_t)(lane & 3) * p.in.slice +
                           (ptrdiff_t)(un.x0 - 2 * d) * 8;
      for (int ja = 0; ja < un.nr + 4; ++ja, ++itx) {
        int row = un.c + d * (un.i0 - 2 + ja);
        row = min(row, p.H + p.in_pad - 1);       // rows past the bottom border only feed y rows that are forced to zero
        const uint32_t slot = itx % p.nxs;
        if (lane == 0) {
          mbar_wait(&x_empty[slot], ((itx / p.nxs) & 1) ^ 1);
          mbar_expect_tx(&x_full[slot], 8 * p.sub_bytes);
        }
        __syncwarp();
        if (lane < 8)
          bulk_load(s_x + (size_t)slot * p.slot_bytes + (size_t)lane * p.sub_bytes, src0 + (ptrdiff_t)row * p.in.ws * 8, p.sub_bytes,
                    &x_full[slot]);
      }
    }
  } else if (warp == 1 || warp == RB_ISSUER_B) {
    // ================================ MMA issuers ================================
    // One warp per convolution.  Between two jobs an issuer polls two mbarriers, fences and commits twice (~300 cycles);
    // the tensor pipe queues only a few MMAs, so with a single issuing thread that gap was a bubble after every job (1280
    // cycles per 960-cycle job, profiles/r02_roleprof.txt).  Two threads issue into the same pipe: one's gap is covered by
    // the other's queued MMAs.  conv_a and conv_b are coupled only through the y ring's barriers, as before.
    // SNB_RB_DEBUG=32 (measurement): warp 1 issues both, in the old interleaved order.
    const bool one_issuer = (p.dbg & 32) != 0;
    const bool swap = (p.dbg & 64) != 0;      // measurement: which SM sub-partition issues which convolution
    const bool do_a = warp == (swap ? RB_ISSUER_B : 1), do_b = one_issuer ? do_a : warp == (swap ? 1 : RB_ISSUER_B);
    const bool leader = elect_one();
    const uint32_t idesc1 = make_idesc_f16(128, RB_SLOT_COLS), idesc2 = make_idesc_f16(128, RB_SLOT_COLS / 2);
    const uint32_t b_lbo = RB_WROWS * 16;
    const uint32_t dil16 = (uint32_t)d;                     // descriptor address units of 16 B = one pixel
    const uint32_t wa_addr = smem_u32(s_wa), wb_addr = smem_u32(s_wb);
    mbar_wait(w_full, 0);
    // The tensor pipe queues only a few MMAs, so uniform-datapath work between two jobs is exposed: ring positions
    // are counters (no division), descriptors are adds on hoisted bases.
    const uint64_t x_desc0 = make_smem_desc(smem_u32(s_x), p.sub_bytes, 128), y_desc0 = make_smem_desc(smem_u32(s_y), p.sub_bytes, 128);
    const uint64_t wa_desc0 = make_smem_desc(wa_addr, b_lbo, 128), wb_desc0 = make_smem_desc(wb_addr, b_lbo, 128);
    const uint64_t slot16 = (uint64_t)(p.slot_bytes >> 4), sub16 = (uint64_t)(p.sub_bytes >> 4), wkx16 = (uint64_t)(2 * b_lbo >> 4);
    uint32_t xs = 0, xpar = 0, ys = 0, ypar = 0, apar = 1, bpar = 1, njobs = 0;

    // measurement only (SNB_RB_DEBUG=16, full-resolution launches): drop the A_lo x W_hi products, i.e. fp16-rounded activations
    const bool no_lo = (p.dbg & 16) != 0;
    // 12 MMAs of one job: A rows at descriptor `a` ([plane][chunk][px][8]), weights at descriptor `w`
    auto issue_job = [&](uint64_t a, uint64_t w, uint32_t dcol) {
#pragma unroll
      for (int k16 = 0; k16 < 2; ++k16) {
        const uint64_t a_hi = a + (uint64_t)(2 * k16) * sub16, a_lo = a_hi + 4 * sub16;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint64_t db = w + (uint64_t)(k16 * 3 + kx) * wkx16;
          const uint64_t sh = (uint64_t)(kx * dil16);
          if (k16 == 0 && kx == 0) umma_f16_zero(dcol, a_hi, db, idesc1);
          else umma_f16_acc(dcol, a_hi + sh, db, idesc1);
          if (!no_lo) umma_f16_acc(dcol + RB_SLOT_COLS / 2, a_lo + sh, db, idesc2);
        }
      }
    };

    // conv_b job: 18 MMAs (N = 96) into one merged accumulator
    auto issue_job_b = [&](uint64_t a, uint64_t w, uint32_t dcol) {
#pragma unroll
      for (int k16 = 0; k16 < 2; ++k16) {
        const uint64_t a_hi = a + (uint64_t)(2 * k16) * sub16, a_lo = a_hi + 4 * sub16;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint64_t w_hi = w + (uint64_t)(k16 * 3 + kx) * wkx16, w_lo = w_hi + (uint64_t)RB_BCOLS;
          const uint64_t sh = (uint64_t)(kx * dil16);
          // order hi*hi, lo*hi, hi*lo: two consecutive MMAs of one accumulator on the SAME A window cost 68 cycles each
          // instead of 56 (tools/ubench/mma_two_issuer_bench.cu modes 5 / 8) - the b-job was 1230 cycles, not 1008
          if (k16 == 0 && kx == 0) umma_f16_zero(dcol, a_hi, w_hi, idesc2);
          else umma_f16_acc(dcol, a_hi + sh, w_hi, idesc2);
          if (!no_lo) umma_f16_acc(dcol, a_lo + sh, w_hi, idesc2);
          umma_f16_acc(dcol, a_hi + sh, w_lo, idesc2);
        }
      }
    };
    uint32_t bs = 0, as = 0;

    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const RbUnit un = rb_decode(p, u);
      if (un.nr <= 0) continue;
      for (int step = 0; step < un.nr + 5; ++step) {
        if (do_a && step < un.nr + 4) {                    // a-job: x row i0 - 2 + step
          RB_WAIT(tw0, &x_full[xs], xpar);
          RB_WAIT(tw1, &sa_empty[as], apar);
          tc_fence_after();
          if (leader) {
            if constexpr (RB_A_MERGED) issue_job_b(x_desc0 + (uint64_t)xs * slot16, wa_desc0, tmem_base + as * RB_ACOLS);
            else issue_job(x_desc0 + (uint64_t)xs * slot16, wa_desc0, tmem_base + as * RB_ACOLS);
            umma_commit(&sa_full[as]);
            umma_commit(&x_empty[xs]);
          }
          __syncwarp();
          if (++xs == (uint32_t)p.nxs) { xs = 0; xpar ^= 1; }
          if (++as == RB_ASLOTS) { as = 0; apar ^= 1; }
          ++njobs;
        }
        if (do_b && step >= 3 && step - 3 < un.nr + 2) {   // b-job: y row i0 - 1 + (step - 3)
          RB_WAIT(tw2, &y_full[ys], ypar);
          RB_WAIT(tw3, &sb_empty[bs], bpar);
          tc_fence_after();
          if (leader) {
            issue_job_b(y_desc0 + (uint64_t)ys * slot16, wb_desc0, tmem_base + RB_BBASE + bs * RB_BCOLS);
            umma_commit(&sb_full[bs]);
            umma_commit(&y_empty[ys]);
          }
          __syncwarp();
          if (++ys == RB_YSLOTS) { ys = 0; ypar ^= 1; }
          if (++bs == RB_BSLOTS) { bs = 0; bpar ^= 1; }
          ++njobs;
        }
      }
    }
    if (PROF && lane == 0) {
      long long* q = p.prof + blockIdx.x * 24;
      if (do_a) { q[0] = clock64() - t_start; q[1] = tw0; q[2] = tw1; q[5] = njobs; }
      if (do_b) { q[6] = clock64() - t_start; q[3] = tw2; q[4] = tw3; q[7] = njobs; }
    }
  } else if (warp < 2 + RB_GROUP_WARPS) {
    // ================================ epilogue group A: a-jobs -> y rows ================================
    // Eight warps: a warp owns a TMEM lane quarter (32 pixels) and one 16-channel half.  With four warps doing both halves
    // the drain of one job was a chain of six TMEM round trips per thread (~1500 cycles per job against 960 of MMAs) on an
    // SM sub-partition holding two resident epilogue warps; half the chain per warp and twice the warps hide it.
    const int m = (warp & 3) * 32 + lane;     // TMEM lane = M row = pixel within the 128-wide window
    const int hf = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + hf * 48;
    const uint32_t sy_addr = smem_u32(s_y) + (uint32_t)m * 16;
    const float kna = p.rzk * (RB_A_MERGED ? 18.f : 6.f);   // MMAs per conv_a accumulator: 2 chunks x 3 kernel columns (x 3 products)
    uint32_t na = 0, ny = 0;
    long long t_emit = 0, t_drain = 0, t_fence = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const RbUnit un = rb_decode(p, u);
      if (un.nr <= 0) continue;
      float a0[16], a1[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) a0[c] = a1[c] = 0.f;
      const int ypx = un.x0 - d + m;          // image column of this thread's y pixel
      const bool col_ok = ypx >= 0 && ypx < p.W;
      for (int ja = 0; ja < un.nr + 4; ++ja, ++na) {
        const uint32_t ts = na % RB_ASLOTS, ys = ny % RB_YSLOTS;             // y rows are emitted from ja = 2 on
        const int row = un.c + d * (un.i0 - 3 + ja);
        const bool ok = col_ok && row >= 0 && row < p.H;
        RB_WAIT(tw0, &sa_full[ts], (na / RB_ASLOTS) & 1);
        tc_fence_after();
        if (ja >= 2) RB_WAIT(tw1, &y_empty[ys], ((ny / RB_YSLOTS) & 1) ^ 1);
        const long long c0 = PROF ? clock64() : 0;
        // drain first (the finished row goes to f, the partial rows roll over), hand the slot back, then emit
        float f[16];
        {
          const uint32_t col = lane_addr + ts * RB_ACOLS;
          float v[16];
          if constexpr (RB_A_MERGED) rb_ld1(col + 32, v); else rb_ld_sum(col + 32, v);     // kernel row 2 completes the oldest partial row
#pragma unroll
          for (int c = 0; c < 16; ++c) f[c] = a0[c] + v[c];
          if constexpr (RB_A_MERGED) rb_ld1(col + 16, v); else rb_ld_sum(col + 16, v);
#pragma unroll
          for (int c = 0; c < 16; ++c) a0[c] = a1[c] + v[c];
          if constexpr (RB_A_MERGED) rb_ld1(col, v); else rb_ld_sum(col, v);
#pragma unroll
          for (int c = 0; c < 16; ++c) a1[c] = v[c] + s_bias[hf * 16 + c];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sa_empty[ts]);
        if (PROF) t_drain += clock64() - c0;
        if (ja >= 2) {
#pragma unroll
          for (int jb = 0; jb < 2; ++jb) {
            float g[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = ok ? fmaxf(rz_comp(f[jb * 8 + q], kna), 0.f) * p.wsa : 0.f;   // truncation loss back (common.cuh), ReLU, out of the weight scale
            uint4 oh, ol;
            rb_split8(g, oh, ol);
            const uint32_t a = sy_addr + ys * p.slot_bytes + (uint32_t)(hf * 2 + jb) * p.sub_bytes;
            rb_sts16(a, oh);
            rb_sts16(a + 4 * p.sub_bytes, ol);
          }
          const long long c1 = PROF ? clock64() : 0;
          fence_proxy_async();
          if (PROF) t_fence += clock64() - c1;
          __syncwarp();
          if (lane == 0) mbar_arrive(&y_full[ys]);
          ++ny;
        }
        if (PROF) t_emit += clock64() - c0;
      }
    }
    if (PROF && warp == 2 && lane == 0) {
      long long* q = p.prof + blockIdx.x * 24;
      q[8] = clock64() - t_start; q[9] = tw0; q[10] = tw1; q[11] = t_emit; q[12] = t_drain; q[13] = t_fence;
    }
  } else if (warp < RB_ISSUER_B) {
    // ================================ epilogue group B: b-jobs -> output rows ================================
    const int m = (warp & 3) * 32 + lane;
    const int hf = (warp - (2 + RB_GROUP_WARPS)) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + RB_BBASE + hf * 48;
    const __half* res = static_cast<const __half*>(p.res.p);
    __half* out = static_cast<__half*>(p.out.p);
    const float knb = p.rzk * ((p.dbg & 16) ? 12.f : 18.f);      // conv_b merged accumulator: 2 chunks x 3 columns x {hi*hi, hi*lo, lo*hi}
    uint32_t nb = 0;
    long long t_emit = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const RbUnit un = rb_decode(p, u);
      if (un.nr <= 0) continue;
      float b0[16], b1[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) b0[c] = b1[c] = 0.f;
      const int opx = un.x0 + m;              // image column of this thread's output pixel
      const bool col_ok = m < p.OW && opx < p.W;
      const size_t r_base = (size_t)un.n * p.res.ss + (size_t)opx * 8 + (size_t)(hf * 2) * p.res.slice;
      const size_t o_base = (size_t)un.n * p.out.ss + (size_t)opx * 8 + (size_t)(hf * 2) * p.out.slice;
      for (int jb_ = 0; jb_ < un.nr + 2; ++jb_, ++nb) {
        const uint32_t ts = nb % RB_BSLOTS;
        const int io = jb_ - 2;
        const int row = un.c + d * (un.i0 + io);
        const bool ok = col_ok && io >= 0 && row < p.H;
        uint4 rh[2], rl[2];
        if (ok) {                                          // residual prefetch: in flight while the job's MMAs finish
          const __half* rp = res + r_base + (size_t)row * p.res.ws * 8;
#pragma unroll
          for (int jb = 0; jb < 2; ++jb) {
            rh[jb] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)jb * p.res.slice));
            rl[jb] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)jb * p.res.slice + p.res.lo));
          }
        }
        RB_WAIT(tw0, &sb_full[ts], (nb / RB_BSLOTS) & 1);
        tc_fence_after();
        const long long c0 = PROF ? clock64() : 0;
        const uint32_t col = lane_addr + ts * RB_BCOLS;
        float v[16];
        rb_ld1(col + 32, v);
        if (ok) {
          __half* op = out + o_base + (size_t)row * p.out.ws * 8;
#pragma unroll
          for (int jb = 0; jb < 2; ++jb) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&rh[jb]);
            const __half2* l2 = reinterpret_cast<const __half2*>(&rl[jb]);
            float f[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 a = __half22float2(h2[j]), b = __half22float2(l2[j]);
              const int c = jb * 8 + 2 * j;
              f[2 * j] = fmaxf(fmaf(rz_comp(b0[c] + v[c], knb), p.wsb, a.x + b.x), 0.f);
              f[2 * j + 1] = fmaxf(fmaf(rz_comp(b0[c + 1] + v[c + 1], knb), p.wsb, a.y + b.y), 0.f);
            }
            uint4 oh, ol;
            rb_split8(f, oh, ol);
            *reinterpret_cast<uint4*>(op + (size_t)jb * p.out.slice) = oh;
            *reinterpret_cast<uint4*>(op + (size_t)jb * p.out.slice + p.out.lo) = ol;
          }
        }
        rb_ld1(col + 16, v);
#pragma unroll
        for (int c = 0; c < 16; ++c) b0[c] = b1[c] + v[c];
        rb_ld1(col, v);
#pragma unroll
        for (int c = 0; c < 16; ++c) b1[c] = v[c] + s_bias[32 + hf * 16 + c];
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb_empty[ts]);
        if (PROF) t_emit += clock64() - c0;
      }
    }
    if (PROF && warp == 2 + RB_GROUP_WARPS && lane == 0) {
      long long* q = p.prof + blockIdx.x * 24;
      q[16] = clock64() - t_start; q[17] = tw0; q[18] = t_emit;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
// in/out/res: split-fp16 C8 tensors with 32 channels, d = 1, same H x W; in.pad >= 2*dil.
cudaError_t resblock_tc_plan(RbPlan* plan, const Tens& in, const Tens& out, const Tens& res, int dil, int num_sms) {
  if (in.planes != 2 || out.planes != 2 || res.planes != 2 || in.cb != 4 || out.cb != 4 || res.cb != 4 || in.d != 1 ||
      in.pad < 2 * dil || dil < 1 || dil > 16 || out.h != in.h || out.w != in.w || res.h != in.h || res.w != in.w)
    return cudaErrorInvalidValue;
  *plan = RbPlan();
  RbParams& p = plan->p;
  p.in = view(in); p.out = view(out); p.res = view(res);
  p.H = in.h; p.W = in.w; p.dil = dil; p.in_pad = in.pad;
  p.OW = 128 - 2 * dil;
  p.XW = 128 + 2 * dil;
  p.sub_bytes = (uint32_t)p.XW * 16;
  p.slot_bytes = 8 * p.sub_bytes;
  p.strips = cdiv(p.W, p.OW);
  plan->num_sms = num_sms;
  const size_t fixed = 128 + 2 * (size_t)RB_W_BYTES + (size_t)RB_YSLOTS * p.slot_bytes;
  int nxs = (int)((227 * 1024 - 2048 - fixed) / p.slot_bytes);       // 2 KB: static shared (biases, barriers) + slack
  nxs = nxs > 6 ? 6 : nxs;
  if (nxs < 3) return cudaErrorInvalidValue;
  p.nxs = nxs;
  plan->smem = fixed + (size_t)nxs * p.slot_bytes;
  return cudaSuccess;
}

cudaError_t launch_resblock_tc(const RbPlan& plan, int N, const void* wa, const void* wb, int wlog2a, int wlog2b, const float* ba,
                               const float* bb, cudaStream_t st) {
  RbParams p = plan.p;
  p.N = N; p.wa = static_cast<const __half*>(wa); p.wb = static_cast<const __half*>(wb); p.ba = ba; p.bb = bb;
  p.rzk = rz_unit();
  p.wsa = ldexpf(1.f, -wlog2a); p.wsb = ldexpf(1.f, -wlog2b);
  { static const int dbg = getenv("SNB_RB_DEBUG") ? atoi(getenv("SNB_RB_DEBUG")) : 0; p.dbg = ((long)p.H * p.W >= 400000 ? dbg : 0) | (dbg & (32 | 64)); }
  // row chunks: about one unit per SM (every unit pays 4 halo rows, but an idle SM pays more), at least 2 rows each
  const int rc_max = cdiv(p.H, p.dil);
  const int columns = N * p.strips * p.dil;                 // independent column walks
  int nchunk = plan.num_sms / columns;                      // floor: never spill a few units into a second wave
  if (nchunk < 1) nchunk = 1;
  if (nchunk > cdiv(rc_max, 2)) nchunk = cdiv(rc_max, 2);
  p.rpc = cdiv(rc_max, nchunk);
  p.nchunk = cdiv(rc_max, p.rpc);
  p.total_units = columns * p.nchunk;
  // Slot configurations (SNB_RB_VARIANT, measurement): 0 = the product; 1 and 2 give conv_a a second TMEM slot (merged
  // accumulator, or conv_b cut to one slot).  Every configuration that lets a-jobs run ahead is SLOWER (67 vs 57 us at
  // 544x960): an N = 96 MMA reads 4 KB of A + 3 KB of B per 56 cycles = all 128 B/cycle of shared memory, so with the tensor
  // pipe never pausing the y-row stores of epilogue group A starve (emit phase 1600 cycles per row in the role profile), y rows
  // come late and conv_b idles.  profiles/r02_resblock_variants.txt.
  using Kern = void (*)(const RbParams);
  static const Kern variants[] = {k_resblock_tc<false, false, 1, 2>, k_resblock_tc<false, true, 2, 2>, k_resblock_tc<false, false, 2, 1>};
  constexpr int NV = sizeof(variants) / sizeof(variants[0]);
  static const int variant = getenv("SNB_RB_VARIANT") ? atoi(getenv("SNB_RB_VARIANT")) % NV : 0;
  static bool attr_done[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 31]) {
    for (int i = 0; i < NV; ++i) cudaFuncSetAttribute(variants[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);   // static shared: 592 B
    cudaFuncSetAttribute(k_resblock_tc<true, false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(k_resblock_tc<true, true, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    attr_done[dev & 31] = true;
  }
  const int grid = p.total_units < plan.num_sms ? p.total_units : plan.num_sms;
  static const int prof = getenv("SNB_RB_PROF") ? atoi(getenv("SNB_RB_PROF")) : getenv("SNB_TC_PROF") ? atoi(getenv("SNB_TC_PROF")) : 0;
  if (!prof) {
    launch_k(variants[variant], grid, RB_THREADS, plan.smem, st, p);
    return cudaGetLastError();
  }
  // diagnostics only: per-role cycle counters, synchronous read-back, max over CTAs
  static long long* d_prof = nullptr;
  if (!d_prof) cudaMalloc(&d_prof, 256 * 24 * sizeof(long long));
  p.prof = d_prof;
  if (prof == 2) {           // the instrumented build, timed from outside like the plain one
    launch_k(variant == 1 ? k_resblock_tc<true, true, 2, 2> : k_resblock_tc<true, false, 1, 2>, grid, RB_THREADS, plan.smem, st, p);
    return cudaGetLastError();
  }
  cudaMemsetAsync(d_prof, 0, 256 * 24 * sizeof(long long), st);
  launch_k(variant == 1 ? k_resblock_tc<true, true, 2, 2> : k_resblock_tc<true, false, 1, 2>, grid, RB_THREADS, plan.smem, st, p);
  cudaStreamSynchronize(st);
  std::vector<long long> h(grid * 24);
  cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx[24] = {0};
  for (int b = 0; b < grid; ++b) for (int k = 0; k < 24; ++k) mx[k] = std::max(mx[k], h[b * 24 + k]);
  fprintf(stderr, "[rbprof] H%d W%d dil%d N%d units %d (rpc %d) grid %d | issuer(a) total %lld wait_x %lld wait_slot(a) %lld jobs %lld issuer(b) total %lld wait_y %lld wait_slot(b) %lld | "
          "groupA total %lld wait_full %lld wait_y_empty %lld drain+emit %lld (drain %lld, fence %lld) | groupB total %lld wait_full %lld drain+emit %lld\n",
          p.H, p.W, p.dil, N, p.total_units, p.rpc, grid, mx[0], mx[1], mx[2], mx[5], mx[6], mx[3], mx[4], mx[8], mx[9], mx[10], mx[11], mx[12], mx[13],
          mx[16], mx[17], mx[18]);
  return cudaGetLastError();
}

}  // namespace snb
