// head.conv3d_alone (the last layer of the 3-D aggregation: Conv3d 32 -> 1, 3x3x3): every input row is
// read ONCE.
//
// The streaming kernel (k_conv_stream, NCO = 16) walks one output depth at a time: an input row (d', y) is fetched by the three
// output depths it feeds, the layer moves 3 x its input through L2 and is bound by that traffic (D = 192, 4 pairs: 0.84 ms for
// 1.4 GFLOP).  With ONE output channel the whole 3 x 3 (depth tap, kernel row) neighbourhood fits the N dimension of a single
// MMA: accumulator column 3*(dz*3 + ky) + t holds, for the input row just streamed, its contribution to output (d' - dz + 1,
// y - ky + 1) - t = 0: hi*hi, 1: hi*lo, 2: lo*hi - 27 columns, N = 32.  A unit = (sample, 128-pixel strip, band of rows,
// segment of depths) walks depth-major through its (rows + 2) x (depths + 2) input rows; the epilogue scatters the nine sums of
// each job into per-pixel partial outputs kept in shared memory (three live depths x band rows), and an output leaves as soon as
// its last contributor - input (d + 1, y + 1) - has been added.  Halo: (rows + 2) / rows x (depths + 2) / depths instead of 3 x.
//   pipeline   warp 0 bulk-copy producer (one ring entry per (row, 16-channel chunk)), warp 1 TMEM owner + MMA issuer
//              (12 MMAs of N = 32 per input row), warps 2-5 epilogue; 4 TMEM slots of 32 columns.
//   numerics   chains of 6 MMAs per accumulator column (the streaming kernel: 18), the nine partial sums meet in fp32; the
//              expected round-toward-zero loss is added back once per finished value (common.cuh rz_comp).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int C3_THREADS = 192;
constexpr int C3_EPI_WARPS = 4;
constexpr int C3_SLOTS = 4;
constexpr int C3_N = 32;                       // accumulator columns: 27 used
constexpr uint32_t C3_WBLK = 2 * 2 * C3_N * 16;    // weight bytes of one (k16, kx): [block][K half][32 rows][8 halfs]

struct C3Unit { int n, x0, i0, nr, d0, nd; };

__device__ __forceinline__ C3Unit c3_decode(const Cost3dParams& p, int u) {
  C3Unit r;
  const int seg = u % p.nseg; u /= p.nseg;
  const int band = u % p.nband; u /= p.nband;
  const int strip = u % p.strips;
  r.n = u / p.strips;
  r.x0 = strip * 128;
  r.i0 = band * p.rpb; r.nr = min(p.H, r.i0 + p.rpb) - r.i0;
  r.d0 = seg * p.dps; r.nd = min(p.D, r.d0 + p.dps) - r.d0;
  return r;
}

__global__ void __launch_bounds__(C3_THREADS, 1) k_cost3d(const Cost3dParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[48];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* s_w = smem;                                      // [k16][kx][block][K half][32 rows][8 halfs]
  float* s_p = reinterpret_cast<float*>(smem + p.w_bytes);  // partial outputs [3 depths][rpb][128 px]
  uint8_t* s_x = smem + p.w_bytes + p.p_bytes;              // ring of [plane][chunk][XW px][8 halfs]
  uint64_t* w_full = bars;
  uint64_t* x_full = bars + 1;
  uint64_t* x_empty = x_full + p.nxs;                       // nxs <= 16
  uint64_t* s_full = bars + 36;
  uint64_t* s_empty = s_full + C3_SLOTS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < p.nxs; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < C3_SLOTS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], C3_EPI_WARPS); }
    fence_barrier_init();
    mbar_expect_tx(w_full, p.w_bytes);                      // the weights are constants of the pass: before the wait
    bulk_load(s_w, p.w, p.w_bytes, w_full);
  }
  if (warp == 1) { tmem_alloc(&tmem_slot, 128); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (int)(p.p_bytes / 4); i += C3_THREADS) s_p[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================================ bulk-copy producer ================================
    const __half* in = static_cast<const __half*>(p.in.p);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const C3Unit un = c3_decode(p, u);
      const int dlo = max(un.d0 - 1, 0), dhi = min(un.d0 + un.nd, p.D - 1);      // input depths of this unit
      const int E = (dhi - dlo + 1) * (un.nr + 2) * 2;                              // ring entries: (depth, row, chunk)
      const int g = lane >> 2, q4 = lane & 3;                                       // 8 entries at a time, 4 copies each
      for (int e0 = 0; e0 < E; e0 += 8) {
        const int e = e0 + g;
        if (e < E) {
          const uint32_t ge = it + (uint32_t)e, slot = ge % (uint32_t)p.nxs, par = ((ge / (uint32_t)p.nxs) & 1) ^ 1;
          const int k16 = e & 1, jr = e >> 1;
          const int dd = dlo + jr / (un.nr + 2), row = un.i0 - 1 + jr % (un.nr + 2);       // row -1 .. H: inside the zero border
          if (q4 == 0) {
            mbar_wait(&x_empty[slot], par);
            mbar_expect_tx(&x_full[slot], 4 * p.sub_bytes);
          }
          __syncwarp(0xfu << (g * 4));
          const __half* src = in + (size_t)un.n * p.in.ss + (size_t)(q4 >> 1) * p.in.lo +     // q4 = plane * 2 + chunk
                              ((size_t)(k16 * 2 + (q4 & 1)) * p.D + dd) * p.in.slice + ((ptrdiff_t)row * p.in.ws + (un.x0 - 1)) * 8;
          bulk_load(s_x + (size_t)slot * p.slot_bytes + (size_t)q4 * p.sub_bytes, src, p.sub_bytes, &x_full[slot]);
        }
        __syncwarp();
      }
      it += (uint32_t)E;
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, C3_N);
    const uint32_t b_lbo = C3_N * 16;                        // bytes between the two K halves of a weight block
    const uint64_t a_desc0 = make_smem_desc(smem_u32(s_x), p.sub_bytes, 128);
    const uint64_t a_lo_off = (uint64_t)(2 * p.sub_bytes >> 4), a_slot16 = (uint64_t)(p.slot_bytes >> 4);
    const uint64_t w_desc0 = make_smem_desc(smem_u32(s_w), b_lbo, 128);
    const uint64_t blk16 = (uint64_t)(2 * b_lbo >> 4), wkx16 = (uint64_t)(C3_WBLK >> 4);
    uint32_t slot = 0, xpar = 0, ts = 0, spar = 1;
    mbar_wait(w_full, 0);
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const C3Unit un = c3_decode(p, u);
      const int dlo = max(un.d0 - 1, 0), dhi = min(un.d0 + un.nd, p.D - 1);
      const int njobs = (dhi - dlo + 1) * (un.nr + 2);
      for (int j = 0; j < njobs; ++j) {
        mbar_wait(&s_empty[ts], spar);
        const uint32_t dcol = tmem_base + ts * C3_N;
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16) {
          mbar_wait(&x_full[slot], xpar);
          tc_fence_after();
          const uint64_t a_hi = a_desc0 + (uint64_t)slot * a_slot16, a_lo = a_hi + a_lo_off;
          const uint64_t w0 = w_desc0 + (uint64_t)(k16 * 3) * wkx16;
          if (leader) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t wb = w0 + (uint64_t)kx * wkx16;
              if (k16 == 0 && kx == 0) umma_f16_zero(dcol, a_hi, wb, idesc);
              else umma_f16_acc(dcol, a_hi + (uint64_t)kx, wb, idesc);
              umma_f16_acc(dcol, a_lo + (uint64_t)kx, wb + blk16, idesc);
            }
            umma_commit(&x_empty[slot]);
          }
          __syncwarp();
          if (++slot == (uint32_t)p.nxs) { slot = 0; xpar ^= 1; }
        }
        if (leader) umma_commit(&s_full[ts]);
        __syncwarp();
        if (++ts == C3_SLOTS) { ts = 0; spar ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ================================
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* pp = s_p + m;                                    // this thread's column of the partial-output array
    const int prow = 128, pdep = p.rpb * 128;               // strides: row, depth slot
    const float kn = p.rzk * 6.f * (RZ_KAPPA_1CH_PER_MMA / RZ_KAPPA_PER_MMA);       // 6 MMAs per accumulator column
    uint32_t ts = 0, fpar = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const C3Unit un = c3_decode(p, u);
      const int dlo = max(un.d0 - 1, 0), dhi = min(un.d0 + un.nd, p.D - 1);
      const int opx = un.x0 + m;
      const bool col_ok = opx < p.W;
      for (int dd = dlo; dd <= dhi; ++dd) {
        for (int jl = -1; jl <= un.nr; ++jl) {
          mbar_wait(&s_full[ts], fpar);
          tc_fence_after();
          uint32_t r0[16], r1[16];
          tmem_ld_16(lane_addr + ts * C3_N, r0);
          tmem_ld_16(lane_addr + ts * C3_N + 16, r1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[ts]);
          if (++ts == C3_SLOTS) { ts = 0; fpar ^= 1; }
          float v[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r0[i]); v[16 + i] = __uint_as_float(r1[i]); }
          // scatter: group (dz, ky) -> output (dd - dz + 1, jl - ky + 1)
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            const int od = dd - dz + 1;
            if (od < un.d0 || od >= un.d0 + un.nd) continue;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int r = jl - ky + 1;
              if (r < 0 || r >= un.nr) continue;
              const int gq = (dz * 3 + ky) * 3;
              pp[(od % 3) * pdep + r * prow] += v[gq] + (v[gq + 1] + v[gq + 2]);
            }
          }
          // outputs whose last contributor was this input row: (dd - 1, jl - 1), and (dd, jl - 1) on the last depth of the volume
          const int r = jl - 1;
          if (r >= 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int od = k == 0 ? dd - 1 : dd;
              if (k == 1 && dd != p.D - 1) continue;
              if (od < un.d0 || od >= un.d0 + un.nd) continue;
              float* q = pp + (od % 3) * pdep + r * prow;
              const float acc = *q;
              *q = 0.f;
              if (col_ok) p.out[(((size_t)un.n * p.D + od) * p.H + un.i0 + r) * p.W + opx] = fmaf(rz_comp(acc, kn), p.wsc, p.bias);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ---- host side -----------------------------------------------------------------------------------
// Weight packing: [k16][kx][block][K half][32 rows][8]; row 3*(dz*3 + ky) + t: block 0 (x A_hi): t = 0 W_hi, t = 1 W_lo; block 1
// (x A_lo): t = 2 W_hi.  W: [1][32][3][3][3] (co, ci, dz, ky, kx), scaled by 2^wlog2.
void cost3d_pack_weights(const float* W, int cin, int wlog2, std::vector<__half>& out) {
  const int nk16 = cin / 16;
  out.assign((size_t)nk16 * 3 * 2 * 2 * C3_N * 8, __float2half(0.f));
  for (int ci = 0; ci < cin; ++ci)
    for (int dz = 0; dz < 3; ++dz)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const float v = ldexpf(W[(((size_t)ci * 3 + dz) * 3 + ky) * 3 + kx], wlog2);
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          const int k16 = ci / 16, half = (ci % 16) / 8, e = ci % 8, g = (dz * 3 + ky) * 3;
          auto at = [&](int block, int row) { return (((((size_t)k16 * 3 + kx) * 2 + block) * 2 + half) * C3_N + row) * 8 + e; };
          out[at(0, g + 0)] = hi;
          out[at(0, g + 1)] = lo;
          out[at(1, g + 2)] = hi;
        }
}

// in: split-fp16 C8 tensor [n][4 blocks][D][h][w], pad >= 1.  Chooses the (band, depth segment) cut for `n_max` samples.
cudaError_t cost3d_plan(Cost3dPlan* plan, const Tens& in, int num_sms) {
  if (in.planes != 2 || in.cb != 4 || in.pad < 1 || in.d < 3) return cudaErrorInvalidValue;
  *plan = Cost3dPlan();
  Cost3dParams& p = plan->p;
  p.in = view(in);
  p.D = in.d; p.H = in.h; p.W = in.w;
  p.strips = cdiv(p.W, 128);
  p.sub_bytes = 130 * 16;
  p.slot_bytes = 4 * p.sub_bytes;
  p.w_bytes = 2 * 3 * C3_WBLK;
  plan->num_sms = num_sms;
  return cudaSuccess;
}

cudaError_t launch_cost3d(const Cost3dPlan& plan, int N, const void* w, int wlog2, float bias, float* out, cudaStream_t st) {
  Cost3dParams p = plan.p;
  p.N = N; p.w = static_cast<const __half*>(w); p.bias = bias; p.out = out;
  p.rzk = rz_unit(); p.wsc = ldexpf(1.f, -wlog2);
  // bands of ~8 rows and segments of ~24 depths keep the halo near (10/8) x (26/24); finer cuts only when the GPU would idle
  const long cols = (long)N * p.strips;
  int nband = cdiv(p.H, 8), nseg = cdiv(p.D, 24);
  while (cols * nband * nseg < plan.num_sms && (nband < p.H || nseg < p.D)) {
    if (cdiv(p.H, nband) > 2 && (cdiv(p.H, nband) >= cdiv(p.D, nseg) || nseg >= p.D)) ++nband; else if (nseg < p.D) ++nseg; else ++nband;
  }
  p.rpb = cdiv(p.H, nband); p.nband = cdiv(p.H, p.rpb);
  p.dps = cdiv(p.D, nseg); p.nseg = cdiv(p.D, p.dps);
  p.total_units = (int)(cols * p.nband * p.nseg);
  p.p_bytes = (uint32_t)(3 * p.rpb * 128 * 4);
  const long avail = 227L * 1024 - 2048 - 128 - (long)p.w_bytes - (long)p.p_bytes;
  p.nxs = (int)std::min<long>(16, avail / p.slot_bytes);
  if (p.nxs < 4) return cudaErrorInvalidValue;
  const size_t smem = 128 + (size_t)p.w_bytes + p.p_bytes + (size_t)p.nxs * p.slot_bytes;
  if (need_attr(9)) cudaFuncSetAttribute(k_cost3d, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  const int grid = p.total_units < plan.num_sms ? p.total_units : plan.num_sms;
  cudaError_t e = launch_k(k_cost3d, grid, C3_THREADS, smem, st, p);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace snb
