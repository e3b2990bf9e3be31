// tcgen05 implicit-GEMM convolution (SNB_PREC_TC_F16X2): stride-1 3x3 (any dilation) and 3x3x3
// convolutions as shifted-window GEMMs on the 5th-gen tensor cores, with fp32-class accuracy.
//
//   GEMM view      M = 128 consecutive pixels of one image row, N = output channels, K = 16 input channels
//                  per tcgen05.mma (kind::f16, fp32 accumulate in TMEM).
//   operands       activations and weights are split fp16 (x = hi + lo).  Per (filter tap, 16 channels, row):
//                    MMA1  A_hi x [W_hi | W_lo]   N = 2*NT  -> TMEM columns [main | corr]
//                    MMA2  A_lo x  W_hi           N =   NT  -> accumulates into corr
//                  (lo*lo ~ 2^-22 is dropped).  Stacking W_hi|W_lo along N halves the shared-memory operand
//                  reads of the activation tile, which bound small-N MMAs (4 KB of A per 128x16 tile).
//   accumulation   the tensor core adds into TMEM with truncation, which biases long chains (measured
//                  -6e-6 relative over 27*Cin/16 steps: 1.4e-3 px end-point error).  So a TMEM chain
//                  covers only the 9 taps of one 16-channel chunk; the epilogue warps drain every chain
//                  with tcgen05.ld and accumulate across chunks in fp32 registers (round-to-nearest), the
//                  large hi*hi sum and the small correction sum kept apart until then.
//   A staging      activations live in HBM with a zero border (common.cuh), so the haloed pixel tile of a
//                  stage is whole rows: one cp.async.bulk per (plane, 8-channel chunk, row) lands it in the
//                  no-swizzle K-major core-matrix layout [plane][chunk][row][pixel][8ch]; every filter tap
//                  is a different start address of the same tile (pixel shift = 16 B), so the 9 taps
//                  re-read shared memory, not L2.
//   pipeline       warp 0: bulk-copy producer (all lanes issue), warp 1: MMA issuer (one lane), warp 2:
//                  TMEM allocator, warps 4-11: epilogue (two warps per TMEM lane quadrant, half the
//                  columns each).  smem ring of `nstages` stages (full/empty mbarriers); TMEM is a ring of
//                  512/(2*NT) row slots (row_full/row_empty) so draining row r overlaps the MMAs of the
//                  following rows.  Persistent: grid = min(tiles, #SM).
// Covers SURVEY.md §8a rows M1 (backbone), M3 (3-D aggregation), M5 (refinement blocks).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int TC_THREADS = 384;
constexpr int TC_EPI_WARPS = 8;

template <int NT, int R>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const TcConvParams p) {
  constexpr int S = 512 / (2 * NT);          // TMEM row slots
  constexpr int CW = NT / 2;                 // output channels per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstages;
  uint64_t* row_full = bars + 2 * p.nstages;
  uint64_t* row_empty = row_full + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(row_empty + S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < S; ++i) { mbar_init(&row_full[i], 1); mbar_init(&row_empty[i], TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kgroups = 3 / p.nky;            // stages per (k16, dz): 1 (haloed tile) or 3 (one kernel row each)
  const int zpad = p.kz >> 1;
  const uint32_t a_plane = 2 * p.a_chunk_bytes;          // hi plane -> lo plane inside a stage

  if (warp == 0) {
    // ================================ bulk-copy producer ================================
    const __half* in = static_cast<const __half*>(p.in.p);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int q = t;
      const int cc = q % p.ccs; q /= p.ccs;
      const int tx = q % p.tiles_x; q /= p.tiles_x;
      const int ty = q % p.tiles_y; q /= p.tiles_y;
      const int d = q % p.D, n = q / p.D;
      const int x0 = tx * 128, y0 = ty * R;
      for (int k16 = 0; k16 < p.nk16; ++k16) {
        for (int dz = 0; dz < p.kz; ++dz) {
          const int zin = d + dz - zpad;
          if (zin < 0 || zin >= p.D) continue;
          for (int g = 0; g < kgroups; ++g, ++it) {
            const int slot = it % p.nstages;
            uint8_t* sa = smem + (size_t)slot * p.stage_bytes;
            if (lane == 0) {
              mbar_wait(&empty[slot], ((it / p.nstages) & 1) ^ 1);
              mbar_expect_tx(&full[slot], p.tx_bytes);
            }
            __syncwarp();
            const uint32_t row_bytes = (uint32_t)p.BW * 16;
            for (int i = lane; i < 4 * p.BH; i += 32) {
              const int pc = i / p.BH, j = i - pc * p.BH;          // pc = plane*2 + chunk
              int y;
              if (p.nky == 1) y = y0 + (g - 1) * p.dil + j;
              else if (p.contig) y = y0 - p.dil + j;
              else y = y0 + (j / R - 1) * p.dil + (j % R);
              const __half* src = in + (size_t)n * p.in.ss + (size_t)(pc >> 1) * p.in.lo +
                                  ((size_t)(k16 * 2 + (pc & 1)) * p.D + zin) * p.in.slice +
                                  ((ptrdiff_t)y * p.in.ws + (x0 - p.dil)) * 8;
              bulk_load(sa + (size_t)pc * p.a_chunk_bytes + (size_t)j * row_bytes, src, row_bytes, &full[slot]);
            }
            if (lane == 0) {
              const __half* wsrc = p.w + ((((size_t)cc * p.nk16 + k16) * p.kz + dz) * 3 + g * p.nky) * (size_t)(3 * 2 * 2 * NT * 8);
              bulk_load(sa + 4 * (size_t)p.a_chunk_bytes, wsrc, p.w_bytes, &full[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, 2 * NT), idesc2 = make_idesc_f16(128, NT);
      const uint32_t a_lbo = p.a_chunk_bytes, b_lbo = (uint32_t)(2 * NT * 16);
      uint32_t it = 0, rs = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int d = (t / (p.ccs * p.tiles_x * p.tiles_y)) % p.D;
        for (int k16 = 0; k16 < p.nk16; ++k16) {
          for (int dz = 0; dz < p.kz; ++dz) {
            const int zin = d + dz - zpad;
            if (zin < 0 || zin >= p.D) continue;
            for (int g = 0; g < kgroups; ++g, ++it) {
              const int slot = it % p.nstages;
              mbar_wait(&full[slot], (it / p.nstages) & 1);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + (size_t)slot * p.stage_bytes);
              const uint64_t da_hi = make_smem_desc(sa, a_lbo, 128);
              const uint64_t da_lo = make_smem_desc(sa + a_plane, a_lbo, 128);
              const uint32_t sw = sa + 4 * p.a_chunk_bytes;
#pragma unroll 1
              for (int r = 0; r < R; ++r, ++rs) {
                const uint32_t ts = rs % S;
                mbar_wait(&row_empty[ts], ((rs / S) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + ts * (2 * NT);
                uint32_t acc = 0;
                for (int ky = 0; ky < p.nky; ++ky) {
                  const int srow = p.nky == 1 ? r : (p.contig ? r + ky * p.dil : ky * R + r);
                  for (int kx = 0; kx < 3; ++kx) {
                    const uint64_t db = make_smem_desc(sw + (uint32_t)((ky * 3 + kx) * (2 * 2 * NT * 16)), b_lbo, 128);
                    const uint32_t poff = (uint32_t)(srow * p.BW + kx * p.dil);      // 16-byte units
                    umma_f16(dcol, da_hi + poff, db, idesc1, acc);
                    umma_f16(dcol + NT, da_lo + poff, db, idesc2, 1u);
                    acc = 1u;
                  }
                }
                umma_commit(&row_full[ts]);      // this row's chain is complete once these MMAs retire
              }
              umma_commit(&empty[slot]);         // frees the smem slot once every MMA above has read it
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int wq = warp & 3;                  // TMEM lane quadrant this warp may read
    const int hf = (warp - 4) >> 2;           // which half of the NT columns
    const int m = wq * 32 + lane;             // pixel within the 128-wide row segment
    const __half* res = static_cast<const __half*>(p.res.p);
    __half* out = static_cast<__half*>(p.out.p);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t rs = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int q = t;
      const int cc = q % p.ccs; q /= p.ccs;
      const int tx = q % p.tiles_x; q /= p.tiles_x;
      const int ty = q % p.tiles_y; q /= p.tiles_y;
      const int d = q % p.D, n = q / p.D;
      const int x = tx * 128 + m;
      float acc[R][CW];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[r][c] = 0.f;

      for (int k16 = 0; k16 < p.nk16; ++k16) {
        for (int dz = 0; dz < p.kz; ++dz) {
          const int zin = d + dz - zpad;
          if (zin < 0 || zin >= p.D) continue;
          for (int g = 0; g < kgroups; ++g) {
#pragma unroll
            for (int r = 0; r < R; ++r, ++rs) {
              const uint32_t ts = rs % S;
              mbar_wait(&row_full[ts], (rs / S) & 1);
              tc_fence_after();
              const uint32_t col = lane_addr + ts * (2 * NT) + hf * CW;
#pragma unroll
              for (int c16 = 0; c16 < CW / 16; ++c16) {
                float vm[16], vc[16];
                tmem_ld_2x16(col + c16 * 16, col + NT + c16 * 16, vm, vc);
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[r][c16 * 16 + c] += vm[c] + vc[c];
              }
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&row_empty[ts]);
            }
          }
        }
      }

      // bias (+ residual) (+ ReLU), split into hi/lo, 16-byte stores (a warp writes 512 contiguous bytes)
      const int co0 = cc * NT + hf * CW;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int y = ty * R + r;
        if (y < p.H && x < p.W) {
#pragma unroll
          for (int jb = 0; jb < CW / 8; ++jb) {
            const int cbo = (co0 >> 3) + jb;
            const size_t pix = ((size_t)y * p.out.ws + x) * 8;
            const size_t o = (size_t)n * p.out.ss + ((size_t)cbo * p.D + d) * p.out.slice + pix;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = acc[r][jb * 8 + j] + __ldg(p.bias + co0 + jb * 8 + j);
            if (res) {
              const size_t ro = (size_t)n * p.res.ss + ((size_t)cbo * p.D + d) * p.res.slice + ((size_t)y * p.res.ws + x) * 8;
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
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
static const int SMEM_BUDGET = 227 * 1024 - 1024;

// Fills BW/BH/stage sizes for (NT, R, nky); returns the number of pipeline stages that fit.
static int tc_layout(TcConvParams& p, int NT, int R, int nky) {
  p.NT = NT; p.R = R; p.nky = nky;
  p.contig = (nky == 3 && p.dil < R) ? 1 : 0;
  p.BW = 128 + 2 * p.dil;
  p.BH = nky == 1 ? R : (p.contig ? R + 2 * p.dil : 3 * R);
  p.a_chunk_bytes = (uint32_t)(p.BH * p.BW * 16);
  p.w_bytes = (uint32_t)(nky * 3 * 2 * 2 * NT * 16);
  p.tx_bytes = 4 * p.a_chunk_bytes + p.w_bytes;
  p.stage_bytes = (p.tx_bytes + 127) / 128 * 128;
  const int S = 512 / (2 * NT);
  const int fixed = 128 + (2 * 8 + 2 * S) * 8 + 16;
  int ns = (SMEM_BUDGET - fixed) / (int)p.stage_bytes;
  return ns > 8 ? 8 : ns;
}

// Chooses tile shape / pipeline depth for one convolution.  in: split-fp16 tensor with pad >= dil.
cudaError_t tc_conv_plan(TcConvPlan* plan, const Tens& in, const Tens& out, int cin, int cout, int dil, int kz, int num_sms) {
  if (cin % 16 || cout % 32 || in.planes != 2 || out.planes != 2 || in.pad < dil) return cudaErrorInvalidValue;
  *plan = TcConvPlan();
  TcConvParams& p = plan->p;
  p.in = view(in); p.out = view(out);
  p.D = in.d; p.H = in.h; p.W = in.w; p.CBin = cin / 8; p.CBout = cout / 8; p.dil = dil; p.kz = kz; p.nk16 = cin / 16;
  // one (NT, R) instantiation for now: 32 output channels x 4 rows per tile
  const int NT = 32, R = 4;
  int ns = tc_layout(p, NT, R, 3);
  if (ns < 2) ns = tc_layout(p, NT, R, 1);
  if (ns < 1) return cudaErrorInvalidValue;
  p.nstages = ns;
  if (R - 1 + 2 > TAIL_ROWS) return cudaErrorInvalidValue;
  p.ccs = cout / NT;
  p.tiles_x = cdiv(p.W, 128);
  p.tiles_y = cdiv(p.H, R);
  const int S = 512 / (2 * NT);
  plan->smem = 128 + (size_t)p.nstages * p.stage_bytes + (size_t)(2 * p.nstages + 2 * S) * 8 + 16;
  (void)num_sms;
  return cudaSuccess;
}

cudaError_t launch_conv_tc(const TcConvPlan& plan, int N, const void* w, const float* bias, const Tens* res, int relu,
                           int num_sms, cudaStream_t st) {
  TcConvParams p = plan.p;
  p.N = N; p.w = static_cast<const __half*>(w); p.bias = bias; p.relu = relu;
  if (res) p.res = view(*res);
  p.total_tiles = N * p.D * p.tiles_y * p.tiles_x * p.ccs;
  if (need_attr(8)) cudaFuncSetAttribute(k_conv_tc<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  k_conv_tc<32, 4><<<grid, TC_THREADS, plan.smem, st>>>(p);
  return cudaGetLastError();
}

// Weight packing for k_conv_tc: [cc][k16][dz][ky][kx][chunk 2][W_hi NT rows | W_lo NT rows][8] fp16,
// from canonical [Cout][Cin][kz][3][3] fp32.
void tc_pack_weights(const float* W, int cout, int cin, int kz, int NT, std::vector<__half>& out) {
  const int ccs = cout / NT, nk16 = cin / 16;
  out.assign((size_t)ccs * nk16 * kz * 3 * 3 * 2 * 2 * NT * 8, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int dz = 0; dz < kz; ++dz)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const float v = W[((((size_t)co * cin + ci) * kz + dz) * 3 + ky) * 3 + kx];
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const int cc = co / NT, nn = co % NT, k16 = ci / 16, chunk = (ci % 16) / 8, e = ci % 8;
            const size_t tap = (((((size_t)cc * nk16 + k16) * kz + dz) * 3 + ky) * 3 + kx);
            const size_t base = (tap * 2 + chunk) * (size_t)(2 * NT);
            out[(base + nn) * 8 + e] = hi;
            out[(base + NT + nn) * 8 + e] = lo;
          }
}

}  // namespace snb
