// fp32 CUDA-core direct convolution over the C8 layout (SNB_PREC_FP32 path).
//
// This is the exact-arithmetic path: it anchors parity against the CPU oracle and validates the
// tcgen05 path.  One CTA = 16x16 output pixels x (4*COT) output channels; the input halo tile for
// one 8-channel block and the matching weight slab are staged in shared memory; each thread keeps
// a 4-pixel x COT-channel register tile.  Warps map to channel groups (weights broadcast inside a
// warp), lanes map to consecutive pixels (conflict-free 128-bit shared loads).
// Covers: SURVEY.md §8a rows M1 (backbone), M3 (Conv3d as kz=3 stacked depth taps), M5 (refinement).
#include "common.cuh"
#include "kernels.cuh"
#include "store.cuh"

namespace snb {

template <int COT, typename T>
__global__ void __launch_bounds__(256) k_conv_direct(ConvParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int CO = 4 * COT;               // output channels per CTA
  extern __shared__ float smem[];
  const int pad = p.dil * (p.ks / 2);
  const int IW = 15 * p.stride + 2 * pad + 1;
  const int IH = IW;
  float* s_in = smem;                       // [IH][IW][8]
  float* s_w = smem + IH * IW * 8;          // [ks*ks][8][CO]

  const int tile = blockIdx.x;
  const int tx0 = (tile % p.tiles_x) * 16, ty0 = (tile / p.tiles_x) * 16;
  const int cc = blockIdx.y;                // output-channel chunk
  const int n = blockIdx.z / p.Dout, dz = blockIdx.z % p.Dout;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cg = warp & 3, half = warp >> 2;

  float acc[4][COT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < COT; ++c) acc[i][c] = 0.f;

  const int ix0 = tx0 * p.stride - pad, iy0 = ty0 * p.stride - pad;
  const int ntap = p.ks * p.ks;
  const size_t in_slice = p.in.slice;

  for (int cb = 0; cb < p.CBin; ++cb) {
    for (int kz = 0; kz < p.kz; ++kz) {
      const int zin = dz + kz - (p.kz >> 1);
      if (zin < 0 || zin >= p.Din) continue;          // uniform across the CTA
      __syncthreads();
      const size_t src = (size_t)n * p.in.ss + ((size_t)cb * p.Din + zin) * in_slice;
      for (int pix = tid; pix < IH * IW; pix += 256) {
        const int y = iy0 + pix / IW, x = ix0 + pix % IW;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (y >= 0 && y < p.Hin && x >= 0 && x < p.Win) St<T>::ld8(p.in.p, src + ((size_t)y * p.in.ws + x) * 8, p.in.lo, v);
        reinterpret_cast<float4*>(s_in)[2 * pix] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(s_in)[2 * pix + 1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      const float* wsrc = p.w + ((((size_t)cc * p.CBin + cb) * p.kz + kz) * ntap) * 8 * CO;
      for (int i = tid; i < ntap * 8 * CO / 4; i += 256)
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wsrc) + i);
      __syncthreads();

      for (int ky = 0; ky < p.ks; ++ky) {
        for (int kx = 0; kx < p.ks; ++kx) {
          float a[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int idx = i * 64 + half * 32 + lane;
            const int py = idx >> 4, px = idx & 15;
            const float4* ip = reinterpret_cast<const float4*>(
                s_in + ((py * p.stride + ky * p.dil) * IW + px * p.stride + kx * p.dil) * 8);
            const float4 v0 = ip[0], v1 = ip[1];
            a[i][0] = v0.x; a[i][1] = v0.y; a[i][2] = v0.z; a[i][3] = v0.w;
            a[i][4] = v1.x; a[i][5] = v1.y; a[i][6] = v1.z; a[i][7] = v1.w;
          }
          const float* wp = s_w + (ky * p.ks + kx) * 8 * CO + cg * COT;
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
            float wv[COT];
#pragma unroll
            for (int c4 = 0; c4 < COT / 4; ++c4) {
              const float4 t = reinterpret_cast<const float4*>(wp + ci * CO)[c4];
              wv[c4 * 4 + 0] = t.x; wv[c4 * 4 + 1] = t.y; wv[c4 * 4 + 2] = t.z; wv[c4 * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int c = 0; c < COT; ++c) acc[i][c] = fmaf(a[i][ci], wv[c], acc[i][c]);
          }
        }
      }
    }
  }

  // epilogue: bias (+ residual) (+ ReLU).  A thread's COT channels sit inside one C8 block; for COT = 4 two
  // neighbouring channel groups (cg, cg^1) share a block, so the pair is exchanged through shared memory.
  const int co0 = cc * CO + cg * COT;       // first output channel of this thread
  float bv[COT];
#pragma unroll
  for (int c = 0; c < COT; ++c) bv[c] = __ldg(p.bias + co0 + c);
  if constexpr (COT == 8) {
    const int cbo = co0 >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 64 + half * 32 + lane;
      const int y = ty0 + (idx >> 4), x = tx0 + (idx & 15);
      if (y >= p.Hout || x >= p.Wout) continue;
      const size_t o = (size_t)n * p.out.ss + ((size_t)cbo * p.Dout + dz) * p.out.slice + ((size_t)y * p.out.ws + x) * 8;
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = acc[i][c] + bv[c];
      if (p.res.p) {
        float r[8];
        St<T>::ld8(p.res.p, (size_t)n * p.res.ss + ((size_t)cbo * p.Dout + dz) * p.res.slice + ((size_t)y * p.res.ws + x) * 8, p.res.lo, r);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] += r[c];
      }
      if (p.relu) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaxf(v[c], 0.f);
      }
      St<T>::st8(p.out.p, o, p.out.lo, v);
    }
  } else {
    // COT == 4: stage the tile through shared memory as [pixel 256][16 ch], then write whole blocks
    __syncthreads();
    float* s_o = smem;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 64 + half * 32 + lane;
#pragma unroll
      for (int c = 0; c < COT; ++c) s_o[idx * CO + cg * COT + c] = acc[i][c] + bv[c];
    }
    __syncthreads();
    for (int e = tid; e < 256 * (CO / 8); e += 256) {
      const int idx = e / (CO / 8), blk = e % (CO / 8);
      const int y = ty0 + (idx >> 4), x = tx0 + (idx & 15);
      if (y >= p.Hout || x >= p.Wout) continue;
      const int cbo = (cc * CO >> 3) + blk;
      const size_t o = (size_t)n * p.out.ss + ((size_t)cbo * p.Dout + dz) * p.out.slice + ((size_t)y * p.out.ws + x) * 8;
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = s_o[idx * CO + blk * 8 + c];
      if (p.res.p) {
        float r[8];
        St<T>::ld8(p.res.p, (size_t)n * p.res.ss + ((size_t)cbo * p.Dout + dz) * p.res.slice + ((size_t)y * p.res.ws + x) * 8, p.res.lo, r);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] += r[c];
      }
      if (p.relu) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaxf(v[c], 0.f);
      }
      St<T>::st8(p.out.p, o, p.out.lo, v);
    }
  }
}

size_t conv_direct_smem(const ConvParams& p, int co) {
  const int pad = p.dil * (p.ks / 2);
  const int IW = 15 * p.stride + 2 * pad + 1;
  return ((size_t)IW * IW * 8 + (size_t)p.ks * p.ks * 8 * co) * sizeof(float);
}

cudaError_t launch_conv_direct(ConvParams p, int cout, cudaStream_t st) {
  p.tiles_x = cdiv(p.Wout, 16);
  const int tiles = p.tiles_x * cdiv(p.Hout, 16);
  if (cout % 32 == 0) {
    const size_t sm = conv_direct_smem(p, 32);
    const dim3 g(tiles, cout / 32, p.N * p.Dout);
    if (p.half) {
      if (need_attr(3)) cudaFuncSetAttribute(k_conv_direct<8, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      launch_k(k_conv_direct<8, __half>, g, 256, sm, st, p);
    } else {
      if (need_attr(0)) cudaFuncSetAttribute(k_conv_direct<8, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      launch_k(k_conv_direct<8, float>, g, 256, sm, st, p);
    }
  } else if (cout % 16 == 0) {
    size_t sm = conv_direct_smem(p, 16);
    if (sm < 256 * 16 * sizeof(float)) sm = 256 * 16 * sizeof(float);      // epilogue staging
    const dim3 g(tiles, cout / 16, p.N * p.Dout);
    if (p.half) {
      if (need_attr(4)) cudaFuncSetAttribute(k_conv_direct<4, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      launch_k(k_conv_direct<4, __half>, g, 256, sm, st, p);
    } else {
      if (need_attr(1)) cudaFuncSetAttribute(k_conv_direct<4, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      launch_k(k_conv_direct<4, float>, g, 256, sm, st, p);
    }
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- Cout = 1: one thread per output element, weights in shared memory --------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_conv_to1(ConvTo1Params p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_w[];            // [CBin][kz][9][8]
  const int nw = p.CBin * p.kz * 9 * 8;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = __ldg(p.w + i);
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z / p.D, dz = blockIdx.z % p.D;
  if (x >= p.W || y >= p.H) return;
  float acc = p.bias;
  const size_t slice = p.in.slice;
  for (int cb = 0; cb < p.CBin; ++cb) {
    for (int kz = 0; kz < p.kz; ++kz) {
      const int zin = dz + kz - (p.kz >> 1);
      if (zin < 0 || zin >= p.D) continue;
      const size_t src = (size_t)n * p.in.ss + ((size_t)cb * p.D + zin) * slice;
      const float* wp = s_w + ((cb * p.kz + kz) * 9) * 8;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + (ky - 1) * p.dil;
        if (yy < 0 || yy >= p.H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + (kx - 1) * p.dil;
          if (xx < 0 || xx >= p.W) continue;
          float v[8];
          St<T>::ld8(p.in.p, src + ((size_t)yy * p.in.ws + xx) * 8, p.in.lo, v);
          const float* w8 = wp + (ky * 3 + kx) * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc = fmaf(v[c], w8[c], acc);
        }
      }
    }
  }
  const size_t o = (((size_t)n * p.D + dz) * p.H + y) * p.W + x;
  if (p.res.p) {
    if (p.res_c8) acc += St<T>::ld1(p.res.p, (size_t)n * p.res.ss + ((size_t)y * p.res.ws + x) * 8, p.res.lo);
    else acc += __ldg(static_cast<const float*>(p.res.p) + o);
  }
  if (p.relu) acc = fmaxf(acc, 0.f);
  p.out[o] = acc;
}

cudaError_t launch_conv_to1(const ConvTo1Params& p, cudaStream_t st) {
  const size_t sm = (size_t)p.CBin * p.kz * 9 * 8 * sizeof(float);
  const dim3 g(cdiv(p.W, 32), cdiv(p.H, 8), p.N * p.D);
  if (p.half) launch_k(k_conv_to1<__half>, g, 256, sm, st, p);
  else launch_k(k_conv_to1<float>, g, 256, sm, st, p);
  return cudaGetLastError();
}

}  // namespace snb
