// tcgen05 implicit-GEMM convolution (SNB_PREC_TC_F16X2): stride-1 3x3 (any dilation) and 3x3x3
// convolutions as shifted-window GEMMs on the 5th-gen tensor cores.
//
//   GEMM view      M = 128 consecutive pixels of one image row, N = NT output channels, K = 16 input
//                  channels per tcgen05.mma (kind::f16, fp32 accumulate in TMEM).
//   operands       activations are stored as split fp16 (x = hi + lo, two planes, layout
//                  [n][2][cb][d][h][w][8]); weights likewise.  Each product is issued as three MMAs
//                  hi*hi + hi*lo + lo*hi into ONE accumulator (lo*lo ~ 2^-22 is dropped): fp32-class
//                  accuracy (measured 1e-4 px EPE) where plain fp16 operands give 3e-2 px.
//   A staging      one TMA box per (16-channel chunk, depth tap[, kernel row]) brings the haloed
//                  pixel tile into shared memory ONCE in the no-swizzle K-major core-matrix layout
//                  [chunk][row][pixel][8ch]; every filter tap is then just a different start address
//                  of the same tile (pixel shift = 16 B), so the 9 taps re-read shared memory, not L2.
//                  TMA zero-fills out-of-bounds coordinates = the convolution's zero padding.
//   pipeline       warp 0: TMA producer, warp 1: MMA issuer (one elected thread), warp 2: TMEM
//                  allocator, warps 4-7: epilogue (tcgen05.ld -> bias/residual/ReLU -> hi/lo split ->
//                  16-byte coalesced stores).  smem ring of `nstages` stages (full/empty mbarriers),
//                  TMEM accumulators double-buffered (tmem_full/tmem_empty) so the epilogue of tile i
//                  overlaps the MMAs of tile i+1.  Persistent: grid = min(tiles, #SM).
// Covers SURVEY.md §8a rows M1 (backbone), M3 (3-D aggregation), M5 (refinement blocks).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

__global__ void __launch_bounds__(256, 1) k_conv_tc(const __grid_constant__ CUtensorMap tm_in, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = 2 * p.a_bytes + p.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstages;
  uint64_t* tmem_full = bars + 2 * p.nstages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT, R = p.R;

  if (warp == 0 && lane == 0) prefetch_tmap(&tm_in);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kgroups = 3 / p.nky;            // stages per (k16, dz): 1 (full halo) or 3 (one kernel row each)
  const int zpad = p.kz >> 1;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
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
              mbar_wait(&empty[slot], ((it / p.nstages) & 1) ^ 1);
              uint8_t* sa = smem + (size_t)slot * stage_bytes;
              mbar_expect_tx(&full[slot], p.tx_bytes);
              const int cy = y0 - p.dil + (p.nky == 1 ? g * p.dil : 0);
              tma_load_5d(sa, &tm_in, &full[slot], 0, x0 - p.dil, cy, zin, (n * 2 + 0) * p.CBin + k16 * 2);
              tma_load_5d(sa + p.a_bytes, &tm_in, &full[slot], 0, x0 - p.dil, cy, zin, (n * 2 + 1) * p.CBin + k16 * 2);
              const __half* wsrc = p.w + ((((size_t)cc * p.nk16 + k16) * p.kz + dz) * 3 + g * p.nky) * (size_t)(96 * NT);
              bulk_load(sa + 2 * p.a_bytes, wsrc, p.w_bytes, &full[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, NT);
      const uint32_t a_lbo = (uint32_t)p.BH * p.BW * 16, b_lbo = (uint32_t)NT * 16;
      uint32_t it = 0, tc = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
        const int d = (t / (p.ccs * p.tiles_x * p.tiles_y)) % p.D;
        const int as = tc & 1;
        mbar_wait(&tmem_empty[as], ((tc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(as * R * NT);
        int nst = 0;
        for (int k16 = 0; k16 < p.nk16; ++k16) {
          for (int dz = 0; dz < p.kz; ++dz) {
            const int zin = d + dz - zpad;
            if (zin < 0 || zin >= p.D) continue;
            for (int g = 0; g < kgroups; ++g, ++it, ++nst) {
              const int slot = it % p.nstages;
              mbar_wait(&full[slot], (it / p.nstages) & 1);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + (size_t)slot * stage_bytes);
              const uint64_t da_hi = make_smem_desc(sa, a_lbo, 128);
              const uint64_t da_lo = make_smem_desc(sa + p.a_bytes, a_lbo, 128);
              const uint32_t sw = sa + 2 * p.a_bytes;
              for (int ky = 0; ky < p.nky; ++ky) {
                for (int kx = 0; kx < 3; ++kx) {
                  const uint64_t db_hi = make_smem_desc(sw + (uint32_t)(((ky * 2 + 0) * 3 + kx) * 2 * NT * 16), b_lbo, 128);
                  const uint64_t db_lo = make_smem_desc(sw + (uint32_t)(((ky * 2 + 1) * 3 + kx) * 2 * NT * 16), b_lbo, 128);
                  const uint32_t first = (nst == 0 && ky == 0 && kx == 0) ? 0u : 1u;
                  for (int r = 0; r < R; ++r) {
                    const uint32_t poff = (uint32_t)((r + (p.nky == 3 ? ky * p.dil : 0)) * p.BW + kx * p.dil);   // 16-byte units
                    const uint32_t dcol = acc0 + (uint32_t)(r * NT);
                    umma_f16(dcol, da_hi + poff, db_hi, idesc, first);
                    umma_f16(dcol, da_hi + poff, db_lo, idesc, 1u);
                    umma_f16(dcol, da_lo + poff, db_hi, idesc, 1u);
                  }
                }
              }
              umma_commit(&empty[slot]);       // frees the smem slot once these MMAs have read it
            }
          }
        }
        umma_commit(&tmem_full[as]);           // accumulators of this tile complete
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int wq = warp & 3;                  // TMEM lane quadrant this warp may read
    const int m = wq * 32 + lane;             // pixel within the 128-wide row segment
    const size_t plane_out = (size_t)p.D * p.H * p.W * 8;                // one (n,hl,cb) plane, halfs
    uint32_t tc = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
      int q = t;
      const int cc = q % p.ccs; q /= p.ccs;
      const int tx = q % p.tiles_x; q /= p.tiles_x;
      const int ty = q % p.tiles_y; q /= p.tiles_y;
      const int d = q % p.D, n = q / p.D;
      const int x = tx * 128 + m;
      const int as = tc & 1;
      mbar_wait(&tmem_full[as], (tc >> 1) & 1);
      tc_fence_after();
      for (int r = 0; r < R; ++r) {
        const int y = ty * R + r;
        for (int c32 = 0; c32 < NT / 32; ++c32) {
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * R * NT + r * NT + c32 * 32), v);
          if (y < p.H && x < p.W) {
            const int co0 = cc * NT + c32 * 32;
#pragma unroll
            for (int jb = 0; jb < 4; ++jb) {
              const int cbo = (co0 >> 3) + jb;
              const size_t o = (((size_t)(n * 2) * p.CBout + cbo) * p.D + d) * (size_t)p.H * p.W * 8 + ((size_t)y * p.W + x) * 8;
              const size_t o_lo = o + (size_t)p.CBout * plane_out;
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = v[jb * 8 + j] + __ldg(p.bias + co0 + jb * 8 + j);
              if (p.res) {
                const uint4 rh = __ldg(reinterpret_cast<const uint4*>(p.res + o));
                const uint4 rl = __ldg(reinterpret_cast<const uint4*>(p.res + o_lo));
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
                const float2 hf = __half22float2(hh);
                ph[j] = hh;
                pl[j] = __floats2half2_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
              }
              *reinterpret_cast<uint4*>(p.out + o) = oh;
              *reinterpret_cast<uint4*>(p.out + o_lo) = ol;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Chooses tile shape / pipeline depth for one convolution and encodes the activation tensor map.
// in: split-fp16 tensor [n][2][cb][d][h][w][8]
cudaError_t tc_conv_plan(TcConvPlan* plan, const void* in, int nmax, int cin, int cout, int D, int H, int W, int dil, int kz, int num_sms) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  if (cin % 16 || cout % 32) return cudaErrorInvalidValue;
  TcConvParams& p = plan->p;
  memset(plan, 0, sizeof(*plan));
  p.D = D; p.H = H; p.W = W; p.CBin = cin / 8; p.CBout = cout / 8; p.dil = dil; p.kz = kz; p.nk16 = cin / 16;
  p.NT = (cout % 64 == 0 && (long)nmax * D * H * W >= 60000) ? 64 : 32;   // wide N only when there are tiles to spare
  p.nky = dil <= 2 ? 3 : 1;
  p.tiles_x = cdiv(W, 128);
  // rows per tile: as tall as TMEM (2 x R x NT <= 512) and smem allow while keeping >= ~1 wave of tiles
  const int ccs = cout / p.NT;
  int R = 512 / (2 * p.NT);
  if (R > 8) R = 8;
  const int budget = 225 * 1024 - 2048;
  for (;; R >>= 1) {
    p.R = R;
    p.BW = 128 + 2 * dil;
    p.BH = p.nky == 3 ? R + 2 * dil : R;
    p.a_bytes = (uint32_t)((2 * p.BH * p.BW * 16 + 127) / 128 * 128);   // two 8-channel chunks of one hi or lo plane; TMA wants 128 B
    p.w_bytes = (uint32_t)(p.nky * 192 * p.NT);
    p.tx_bytes = (uint32_t)(2 * (2 * p.BH * p.BW * 16)) + p.w_bytes;   // bytes the three copies of a stage really deliver
    const uint32_t stage = 2 * p.a_bytes + p.w_bytes;
    p.nstages = budget / (int)stage;
    if (p.nstages > 6) p.nstages = 6;
    const long tiles = (long)nmax * D * cdiv(H, R) * p.tiles_x * ccs;
    if (R == 1 || (p.nstages >= 2 && tiles >= num_sms)) break;
  }
  if (p.nstages < 1) return cudaErrorInvalidValue;
  p.ccs = ccs;
  p.tiles_y = cdiv(H, p.R);
  plan->smem = (size_t)p.nstages * (2 * p.a_bytes + p.w_bytes) + (2 * p.nstages + 4) * 8 + 16 + 1024;
  const cuuint64_t gdim[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)nmax * 2 * p.CBin};
  const cuuint64_t gstr[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)W * H * 16, (cuuint64_t)W * H * D * 16};
  const cuuint32_t box[5] = {8, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1, 2};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&plan->tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_conv_tc(const TcConvPlan& plan, int N, const void* w, const float* bias, const void* res, void* out, int relu,
                           int num_sms, cudaStream_t st) {
  TcConvParams p = plan.p;
  p.N = N; p.w = static_cast<const __half*>(w); p.bias = bias; p.res = static_cast<const __half*>(res);
  p.out = static_cast<__half*>(out); p.relu = relu;
  p.total_tiles = N * p.D * p.tiles_y * p.tiles_x * p.ccs;
  if (need_attr(8)) cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  k_conv_tc<<<grid, 256, plan.smem, st>>>(plan.tm_in, p);
  return cudaGetLastError();
}

// Weight packing for k_conv_tc: [cc][k16][dz][ky][hl][kx][chunk 2][NT][8] fp16, from canonical [Cout][Cin][kz][3][3] fp32.
void tc_pack_weights(const float* W, int cout, int cin, int kz, int NT, std::vector<__half>& out) {
  const int ccs = cout / NT, nk16 = cin / 16;
  out.assign((size_t)ccs * nk16 * kz * 3 * 2 * 3 * 2 * NT * 8, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int dz = 0; dz < kz; ++dz)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const float v = W[((((size_t)co * cin + ci) * kz + dz) * 3 + ky) * 3 + kx];
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const int cc = co / NT, nn = co % NT, k16 = ci / 16, chunk = (ci % 16) / 8, e = ci % 8;
            const size_t base = ((((size_t)cc * nk16 + k16) * kz + dz) * 3 + ky);
            const size_t i_hi = ((((base * 2 + 0) * 3 + kx) * 2 + chunk) * NT + nn) * 8 + e;
            const size_t i_lo = ((((base * 2 + 1) * 3 + kx) * 2 + chunk) * NT + nn) * 8 + e;
            out[i_hi] = hi; out[i_lo] = lo;
          }
}

}  // namespace snb
