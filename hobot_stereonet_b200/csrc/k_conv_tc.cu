// tcgen05 implicit-GEMM convolution (SNB_PREC_TC_F16X2): stride-1 3x3 (any dilation) and 3x3x3
// convolutions as shifted-window GEMMs on the 5th-gen tensor cores, with fp32-class accuracy.
//
//   GEMM view      M = 128 consecutive pixels of one INPUT row, K = 16 input channels per tcgen05.mma
//                  (kind::f16, fp32 accumulate in TMEM), N = the 32 output channels of the tile stacked over
//                  the three kernel rows ky and over the hi/lo halves of the weights:
//                    MMA1  A_hi x [W_hi(ky0..2) | W_lo(ky0..2)]   N = 192  -> TMEM columns [main | corr]
//                    MMA2  A_lo x  W_hi(ky0..2)                   N =  96  -> accumulates into corr
//                  Input row j therefore yields, in one pass over its pixels, its contribution to output rows
//                  j, j-1 and j-2 of the tile; the kernel column kx is a 16-byte shift of the A start address.
//                  Measured on B200 (tools/ubench): one M=128 MMA costs max(N/2, 32 + N/4) cycles - the 4 KB
//                  A tile is re-read from shared memory by every MMA - so wide N is what fills the tensor
//                  pipe: 96 + 56 cycles per (row, kx, 16 ch) here against 9 x 88 for per-tap N = 64/32 MMAs.
//   operands       activations and weights are split fp16 (x = hi + lo); hi*hi lands in `main`, hi*lo + lo*hi
//                  in `corr`, lo*lo (2^-22) is dropped.
//   accumulation   the tensor core adds into TMEM with truncation, which biases long chains (1.4e-3 px
//                  end-point error when all 27*Cin/16 steps shared one accumulator).  A TMEM chain here is
//                  the 3 kx taps of one 16-channel chunk; the epilogue warps drain every chain with
//                  tcgen05.ld and accumulate in fp32 registers (round to nearest).
//   tile           128 pixels x R output rows x 32 channels.  Rows of a tile are `dil` apart (a comb), so
//                  the vertical taps always hit the R + 2 input rows of the same comb, whatever the dilation.
//   A staging      activations live in HBM with a zero border (common.cuh), so a stage's input rows are whole
//                  in-bounds runs: one cp.async.bulk per (plane, 8-channel chunk, row) lands them in the
//                  no-swizzle K-major core-matrix layout [plane][chunk][row][pixel][8ch].
//   pipeline       warp 0: bulk-copy producer (all lanes issue), warp 1: MMA issuer (warp-uniform control,
//                  one elected lane issues), warp 2: TMEM allocator, warps 4-11: epilogue (two warps per TMEM
//                  lane quadrant, 16 channels each).  smem ring of `nstages` stages (full/empty mbarriers);
//                  TMEM holds two 192-column row slots (slot_full/slot_empty) so draining input row j overlaps
//                  the MMAs of row j+1.  Persistent: grid = min(tiles, #SM).
// Covers SURVEY.md §8a rows M1 (backbone), M3 (3-D aggregation), M5 (refinement blocks).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int TC_THREADS = 384;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_NT = 32;                    // output channels per tile
constexpr int TC_SLOT_COLS = 6 * TC_NT;      // 192: [main ky0..2 | corr ky0..2], ordered [half][ky][16 ch]
constexpr uint32_t TC_W_BYTES = 3 * 2 * TC_SLOT_COLS * 16;   // one stage of weights: [kx][chunk][192 rows][8 halfs]

struct TileCoord { int cc, tx, d, n, ybase; };

__device__ __forceinline__ TileCoord decode_tile(const TcConvParams& p, int t, int R) {
  TileCoord c;
  c.cc = t % p.ccs; t /= p.ccs;
  c.tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  c.d = t % p.D; c.n = t / p.D;
  c.ybase = (ty / p.dil) * (R * p.dil) + ty % p.dil;      // first output row of the comb
  return c;
}

template <int R>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const TcConvParams p) {
  constexpr int BH = R + 2;                  // input rows per stage
  constexpr int NT = TC_NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstages;
  uint64_t* slot_full = bars + 2 * p.nstages;
  uint64_t* slot_empty = slot_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();          // the next kernel may start its prologue
  pdl_wait();             // everything above touched only shared memory / TMEM
  const uint32_t tmem_base = *tmem_slot;
  const int zpad = p.kz >> 1;

  if (warp == 0) {
    // ================================ bulk-copy producer ================================
    const __half* in = static_cast<const __half*>(p.in.p);
    const uint32_t row_bytes = (uint32_t)p.BW * 16;
    const int npc = p.cin8 ? 2 : 4;                                // (plane, chunk) regions per stage
    uint32_t it = 0;
    long long w_empty = 0, t_start = clock64();
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t, R);
      const int x0 = tc.tx * 128, y0 = tc.ybase - p.dil;           // first input row
      // rows at or below H + pad feed masked outputs only: skip them (keeps every copy inside the slice)
      int nrows = (p.H + p.in_pad - 1 - y0) / p.dil + 1;
      nrows = nrows > BH ? BH : nrows;
      for (int k16 = 0; k16 < p.nk16; ++k16) {
        for (int dz = 0; dz < p.kz; ++dz) {
          const int zin = tc.d + dz - zpad;
          if (zin < 0 || zin >= p.D) continue;
          const int slot = it % p.nstages;
          uint8_t* sa = smem + (size_t)slot * p.stage_bytes;
          if (lane == 0) {
            const long long c0 = clock64();
            mbar_wait(&empty[slot], ((it / p.nstages) & 1) ^ 1);
            w_empty += clock64() - c0;
            mbar_expect_tx(&full[slot], (p.dbg & 1) ? 0u : (uint32_t)npc * nrows * row_bytes + p.w_bytes);
          }
          __syncwarp();
          ++it;
          if (p.dbg & 1) continue;
          for (int i = lane; i < npc * BH; i += 32) {
            const int pc = i / BH, j = i - pc * BH;          // pc = plane*2 + chunk (cin8: pc = plane, one chunk)
            if (j >= nrows) continue;
            const int y = y0 + j * p.dil;
            const int plane = p.cin8 ? pc : pc >> 1, cb = p.cin8 ? 0 : k16 * 2 + (pc & 1);
            const __half* src = in + (size_t)tc.n * p.in.ss + (size_t)plane * p.in.lo +
                                ((size_t)cb * p.D + zin) * p.in.slice +
                                ((ptrdiff_t)y * p.in.ws + (x0 - p.dil)) * 8;
            bulk_load(sa + (size_t)pc * p.a_chunk_bytes + (size_t)j * row_bytes, src, row_bytes, &full[slot]);
          }
          if (lane == 0) {
            const __half* wsrc = p.w + (((size_t)tc.cc * p.nk16 + k16) * p.kz + dz) * (size_t)(p.w_bytes / 2);
            bulk_load(sa + (size_t)npc * p.a_chunk_bytes, wsrc, p.w_bytes, &full[slot]);
          }
        }
      }
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 0] = clock64() - t_start; p.prof[blockIdx.x * 16 + 1] = w_empty; p.prof[blockIdx.x * 16 + 2] = it; }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // Control flow is warp-uniform; only the elected lane executes tcgen05.mma / commit, so the compiler
    // emits straight-line UTCHMMA with uniform-register descriptor arithmetic.
    const bool leader = elect_one();
    const uint32_t idesc1 = make_idesc_f16(128, TC_SLOT_COLS), idesc2 = make_idesc_f16(128, TC_SLOT_COLS / 2);
    const uint32_t a_lbo = p.a_chunk_bytes, b_lbo = (uint32_t)(TC_SLOT_COLS * 16);
    const uint32_t row16 = (uint32_t)p.BW, dil16 = (uint32_t)p.dil;          // descriptor address units of 16 B
    uint32_t it = 0, rs = 0;
    long long w_full = 0, w_slot = 0, t_start = clock64();
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int d = (t / (p.ccs * p.tiles_x * p.tiles_y)) % p.D;
      for (int k16 = 0; k16 < p.nk16; ++k16) {
        for (int dz = 0; dz < p.kz; ++dz) {
          const int zin = d + dz - zpad;
          if (zin < 0 || zin >= p.D) continue;
          const int slot = it % p.nstages;
          { const long long c0 = clock64(); mbar_wait(&full[slot], (it / p.nstages) & 1); w_full += clock64() - c0; }
          ++it;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)slot * p.stage_bytes);
          // cin8: the second K half of an MMA is the NEXT tap's pixel (LBO = dil pixels), one chunk per plane
          const uint64_t da_hi = make_smem_desc(sa, p.cin8 ? dil16 * 16 : a_lbo, 128);
          const uint64_t da_lo = make_smem_desc(sa + (p.cin8 ? 1 : 2) * p.a_chunk_bytes, p.cin8 ? dil16 * 16 : a_lbo, 128);
          const uint64_t db = make_smem_desc(sa + (p.cin8 ? 2 : 4) * p.a_chunk_bytes, b_lbo, 128);
          constexpr uint32_t KX_B = 2 * TC_SLOT_COLS;                // 16-byte units between the kx weight blocks
#pragma unroll
          for (int j = 0; j < BH; ++j, ++rs) {
            const uint32_t ts = rs & 1;
            { const long long c0 = clock64(); mbar_wait(&slot_empty[ts], ((rs >> 1) & 1) ^ 1); w_slot += clock64() - c0; }
            tc_fence_after();
            if (leader) {
              const uint32_t dcol = tmem_base + ts * TC_SLOT_COLS;
              const uint64_t a_hi = da_hi + (uint64_t)(j * row16), a_lo = da_lo + (uint64_t)(j * row16);
              if (p.cin8) {                      // taps (kx0, kx1) then (kx2, zero weights)
                umma_f16_zero(dcol, a_hi, db, idesc1);
                umma_f16_acc(dcol + TC_SLOT_COLS / 2, a_lo, db, idesc2);
                umma_f16_acc(dcol, a_hi + 2 * dil16, db + KX_B, idesc1);
                umma_f16_acc(dcol + TC_SLOT_COLS / 2, a_lo + 2 * dil16, db + KX_B, idesc2);
              } else if (!(p.dbg & 2)) {
                umma_f16_zero(dcol, a_hi, db, idesc1);
                umma_f16_acc(dcol + TC_SLOT_COLS / 2, a_lo, db, idesc2);
                umma_f16_acc(dcol, a_hi + dil16, db + KX_B, idesc1);
                umma_f16_acc(dcol + TC_SLOT_COLS / 2, a_lo + dil16, db + KX_B, idesc2);
                umma_f16_acc(dcol, a_hi + 2 * dil16, db + 2 * KX_B, idesc1);
                umma_f16_acc(dcol + TC_SLOT_COLS / 2, a_lo + 2 * dil16, db + 2 * KX_B, idesc2);
              }
              umma_commit(&slot_full[ts]);       // this input row's chain is complete once these MMAs retire
            }
            __syncwarp();
          }
          if (leader) umma_commit(&empty[slot]);  // frees the smem stage once every MMA above has read it
          __syncwarp();
        }
      }
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 4] = clock64() - t_start; p.prof[blockIdx.x * 16 + 5] = w_full; p.prof[blockIdx.x * 16 + 6] = w_slot; p.prof[blockIdx.x * 16 + 7] = rs; }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int wq = warp & 3;                  // TMEM lane quadrant this warp may read
    const int hf = (warp - 4) >> 2;           // which 16 of the tile's 32 channels
    const int m = wq * 32 + lane;             // pixel within the 128-wide row segment
    const __half* res = static_cast<const __half*>(p.res.p);
    __half* out = static_cast<__half*>(p.out.p);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + hf * 48;
    uint32_t rs = 0;
    long long w_slotf = 0, t_fin = 0, t_start = clock64();
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t, R);
      const int x = tc.tx * 128 + m;
      float acc[R][16];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;

      for (int k16 = 0; k16 < p.nk16; ++k16) {
        for (int dz = 0; dz < p.kz; ++dz) {
          const int zin = tc.d + dz - zpad;
          if (zin < 0 || zin >= p.D) continue;
#pragma unroll
          for (int j = 0; j < BH; ++j, ++rs) {
            const uint32_t ts = rs & 1;
            { const long long c0 = clock64(); mbar_wait(&slot_full[ts], (rs >> 1) & 1); w_slotf += clock64() - c0; }
            tc_fence_after();
            const uint32_t col = lane_addr + ts * TC_SLOT_COLS;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int r = j - ky;             // input row j is kernel row ky of output row j - ky
              if (r < 0 || r >= R) continue;
              float vm[16], vc[16];
              tmem_ld_2x16(col + ky * 16, col + TC_SLOT_COLS / 2 + ky * 16, vm, vc);
#pragma unroll
              for (int c = 0; c < 16; ++c) acc[r][c] += vm[c] + vc[c];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[ts]);
          }
        }
      }

      // bias (+ residual) (+ ReLU), split into hi/lo, 16-byte stores (a warp writes 512 contiguous bytes)
      const long long cfin = clock64();
      const int co0 = tc.cc * NT + hf * 16;
      {
        // out of the accumulator domain, in place (keeps the finalisation below as lean as it was): the expected truncation loss
        // of the tensor core's accumulate back (common.cuh rz_comp; 3 - or 2 - MMAs per TMEM chain), then the 2^-k weight scale
        const float kn = p.rzk * (p.cin8 ? 2.f : 3.f), wsc = p.wsc;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[r][c] = rz_comp(acc[r][c], kn) * wsc;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int y = tc.ybase + r * p.dil;
        if (y < p.H && x < p.W && !(p.dbg & 4)) {
#pragma unroll
          for (int jb = 0; jb < 2; ++jb) {
            const int cbo = (co0 >> 3) + jb;
            const size_t o = (size_t)tc.n * p.out.ss + ((size_t)cbo * p.D + tc.d) * p.out.slice + ((size_t)y * p.out.ws + x) * 8;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = acc[r][jb * 8 + j] + __ldg(p.bias + co0 + jb * 8 + j);
            if (res) {
              const size_t ro = (size_t)tc.n * p.res.ss + ((size_t)cbo * p.D + tc.d) * p.res.slice + ((size_t)y * p.res.ws + x) * 8;
              const uint4 rh = __ldg(reinterpret_cast<const uint4*>(res + ro));
              const uint4 rl = __ldg(reinterpret_cast<const uint4*>(res + ro + p.res.lo));
              const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
              const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 a = __half22float2(h2[j]), b = __half22float2(l2[j]);
                f[2 * j] += a.x + b.x; f[2 * j + 1] += a.y + b.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            uint4 oh, ol;
            __half2* ph = reinterpret_cast<__half2*>(&oh);
            __half2* pl = reinterpret_cast<__half2*>(&ol);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __half2 hh = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
              const float2 hfv = __half22float2(hh);
              ph[j] = hh;
              pl[j] = __floats2half2_rn(f[2 * j] - hfv.x, f[2 * j + 1] - hfv.y);
            }
            *reinterpret_cast<uint4*>(out + o) = oh;
            *reinterpret_cast<uint4*>(out + o + p.out.lo) = ol;
          }
        }
      }
      t_fin += clock64() - cfin;
    }
    if (p.prof && warp == 4 && lane == 0) { p.prof[blockIdx.x * 16 + 8] = clock64() - t_start; p.prof[blockIdx.x * 16 + 9] = w_slotf; p.prof[blockIdx.x * 16 + 10] = t_fin; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
static const int SMEM_BUDGET = 227 * 1024 - 1024;

static int tc_layout(TcConvParams& p, int R) {
  p.R = R;
  p.BW = 128 + (p.cin8 ? 3 : 2) * p.dil;       // cin8: the zero-weight half of the last tap pair reads one tap further
  p.BH = R + 2;
  p.a_chunk_bytes = (uint32_t)(p.BH * p.BW * 16);
  p.w_bytes = p.cin8 ? TC_W_BYTES * 2 / 3 : TC_W_BYTES;
  p.stage_bytes = ((p.cin8 ? 2 : 4) * p.a_chunk_bytes + p.w_bytes + 127) / 128 * 128;
  const int fixed = 128 + (2 * 8 + 4) * 8 + 16;
  int ns = (SMEM_BUDGET - fixed) / (int)p.stage_bytes;
  return ns > 4 ? 4 : ns;
}

// Chooses the tile height / pipeline depth for one convolution.  in: split-fp16 tensor with pad >= dil.
cudaError_t tc_conv_plan(TcConvPlan* plan, const Tens& in, const Tens& out, int cin, int cout, int dil, int kz, int num_sms) {
  const bool cin8 = cin <= 8 && in.cb == 1 && kz == 1;
  if ((!cin8 && cin % 16) || cout % TC_NT || in.planes != 2 || out.planes != 2 || in.pad < dil) return cudaErrorInvalidValue;
  *plan = TcConvPlan();
  TcConvParams& p = plan->p;
  p.in = view(in); p.out = view(out);
  p.cin8 = cin8 ? 1 : 0;
  p.D = in.d; p.H = in.h; p.W = in.w; p.CBin = cin8 ? 1 : cin / 8; p.CBout = cout / 8; p.dil = dil; p.kz = kz; p.nk16 = cin8 ? 1 : cin / 16;
  p.in_pad = in.pad;
  p.ccs = cout / TC_NT;
  p.tiles_x = cdiv(p.W, 128);
  // tile height: minimise waves x input rows per tile (R + 2 rows are staged and multiplied for R output rows)
  auto tiles_for = [&](int R) { return (long)in.n * p.D * cdiv(p.H, R * dil) * dil * p.tiles_x * p.ccs; };
  int R = 6;
  long best = -1;
  for (int r : {6, 4, 2}) {
    const long cost = ((tiles_for(r) + num_sms - 1) / num_sms) * (r + 2);
    if (best < 0 || cost < best) { best = cost; R = r; }
  }
  const int ns = tc_layout(p, R);
  if (ns < 1) return cudaErrorInvalidValue;
  p.nstages = ns;
  p.tiles_y = cdiv(p.H, R * dil) * dil;
  plan->smem = 128 + (size_t)p.nstages * p.stage_bytes + (size_t)(2 * p.nstages + 4) * 8 + 16;
  return cudaSuccess;
}

template <int R>
static void launch_r(const TcConvParams& p, int grid, size_t smem, cudaStream_t st) {
  static bool attr_done[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 31]) {
    cudaFuncSetAttribute(k_conv_tc<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done[dev & 31] = true;
  }
  launch_k(k_conv_tc<R>, grid, TC_THREADS, smem, st, p);
}

cudaError_t launch_conv_tc(const TcConvPlan& plan, int N, const void* w, int wlog2, const float* bias, const Tens* res, int relu,
                           int num_sms, cudaStream_t st) {
  TcConvParams p = plan.p;
  p.N = N; p.w = static_cast<const __half*>(w); p.bias = bias; p.relu = relu;
  if (res) p.res = view(*res);
  p.total_tiles = N * p.D * p.tiles_y * p.tiles_x * p.ccs;
  { static const int dbg = getenv("SNB_TC_DEBUG") ? atoi(getenv("SNB_TC_DEBUG")) : 0; p.dbg = dbg; }
  p.rzk = rz_unit();
  p.wsc = ldexpf(1.f, -wlog2);
  static const int prof = getenv("SNB_TC_PROF") ? atoi(getenv("SNB_TC_PROF")) : 0;
  static long long* d_prof = nullptr;
  if (prof && !d_prof) cudaMalloc(&d_prof, 256 * 16 * sizeof(long long));
  p.prof = prof ? d_prof : nullptr;
  if (prof) cudaMemsetAsync(d_prof, 0, 256 * 16 * sizeof(long long), st);
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  if (p.R == 6) launch_r<6>(p, grid, plan.smem, st);
  else if (p.R == 4) launch_r<4>(p, grid, plan.smem, st);
  else launch_r<2>(p, grid, plan.smem, st);
  if (prof) {     // diagnostics only: synchronous read-back, max over CTAs of each role's counters
    cudaStreamSynchronize(st);
    std::vector<long long> h(grid * 16);
    cudaMemcpy(h.data(), d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx[16] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 16; ++k) mx[k] = std::max(mx[k], h[b * 16 + k]);
    fprintf(stderr, "[tcprof] H%d W%d D%d Cin%d Cout%d dil%d res%d R%d tiles %d grid %d | producer total %lld wait_empty %lld stages %lld | "
            "issuer total %lld wait_full %lld wait_slot_empty %lld rows %lld | epilogue total %lld wait_slot_full %lld finalize %lld\n",
            p.H, p.W, p.D, p.CBin * 8, p.CBout * 8, p.dil, res ? 1 : 0, p.R, p.total_tiles, grid, mx[0], mx[1], mx[2], mx[4], mx[5], mx[6], mx[7],
            mx[8], mx[9], mx[10]);
  }
  return cudaGetLastError();
}

// Weight packing for k_conv_tc: [cc][k16][dz][kx][chunk 2][192 rows][8] fp16 from canonical
// [Cout][Cin][kz][3][3] fp32; row = part*96 + half*48 + ky*16 + (co % 16), part 0 = W_hi, 1 = W_lo,
// half = which 16 of the tile's 32 channels (one epilogue warp set each).
void tc_pack_weights(const float* W, int cout, int cin, int kz, int NT, int wlog2, std::vector<__half>& out) {
  if (NT == 8) {
    // cin8 packing: [cc][tap pair g][K half = kx - 2g][192 rows][8 ch]; the half of pair 1 that would be kx = 3 stays zero
    const int ccs = cout / TC_NT;
    out.assign((size_t)ccs * 2 * 2 * TC_SLOT_COLS * 8, __float2half(0.f));
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const float v = ldexpf(W[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx], wlog2);
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const int cc = co / TC_NT, cl = co % TC_NT;
            const int row = (cl / 16) * 48 + ky * 16 + cl % 16;
            const size_t blk = ((size_t)cc * 2 + kx / 2) * 2 + kx % 2;
            out[(blk * TC_SLOT_COLS + row) * 8 + ci] = hi;
            out[(blk * TC_SLOT_COLS + 96 + row) * 8 + ci] = lo;
          }
    return;
  }
  const int ccs = cout / TC_NT, nk16 = cin / 16;
  out.assign((size_t)ccs * nk16 * kz * 3 * 2 * TC_SLOT_COLS * 8, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int dz = 0; dz < kz; ++dz)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const float v = ldexpf(W[((((size_t)co * cin + ci) * kz + dz) * 3 + ky) * 3 + kx], wlog2);
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const int cc = co / TC_NT, cl = co % TC_NT, k16 = ci / 16, chunk = (ci % 16) / 8, e = ci % 8;
            const int row = (cl / 16) * 48 + ky * 16 + cl % 16;
            const size_t blk = ((((size_t)cc * nk16 + k16) * kz + dz) * 3 + kx) * 2 + chunk;
            out[(blk * TC_SLOT_COLS + row) * 8 + e] = hi;
            out[(blk * TC_SLOT_COLS + 96 + row) * 8 + e] = lo;
          }
}

}  // namespace snb
