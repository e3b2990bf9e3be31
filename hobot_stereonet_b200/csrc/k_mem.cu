// HBM-bound kernels of the StereoNet path: input conversion (P1-P3), cost-volume build (M2),
// soft-argmin (M4), refinement glue (M5) and output quantisation (O1).  All coalesced along x,
// 128-bit accesses where the layout allows, shared-memory staging where data is reused across D.
#include "common.cuh"
#include "kernels.cuh"
#include "store.cuh"

namespace snb {

// ------------------------------------------------------------------------------------------------
// s8 NCHW [B,6,H,W] -> C8 image [2B][1][Hp][Wp][8].  Restates the tensor semantics the BPU model
// gives its input: value = s8 * (1/128) (preprocess.cpp:1037), channels 0-2 left, 3-5 right.
// Four pixels per thread (one 4-byte load per colour plane when the rows allow it).  Split-fp16 storage: s8 / 128 is exact in
// fp16, so only the hi plane is written - the lo plane of the image tensor is zero from its allocation and nothing else
// ever writes it (k_pre_nv12 writes the same exact values).
template <typename T>
__global__ void __launch_bounds__(128) k_pre_s8(const int8_t* __restrict__ s8, TV img, int B, int H, int W, int Hp, int Wp) {
  pdl_trigger();
  pdl_wait();
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int n = blockIdx.z;                  // 0..2B-1
  if (x0 >= Wp) return;
  const int b = n % B, view = n / B;
  float v[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i][0] = v[i][1] = v[i][2] = 0.f;
  if (y < H && x0 < W) {
    const size_t plane = (size_t)H * W;
    const int8_t* src = s8 + (((size_t)b * 6 + view * 3) * H + y) * W + x0;
    if ((W & 3) == 0) {                      // rows and planes are 4-byte aligned (x0 + 3 < W follows)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const char4 q = __ldg(reinterpret_cast<const char4*>(src + c * plane));
        v[0][c] = (float)q.x * 0.0078125f; v[1][c] = (float)q.y * 0.0078125f;
        v[2][c] = (float)q.z * 0.0078125f; v[3][c] = (float)q.w * 0.0078125f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x0 + i < W) {
#pragma unroll
          for (int c = 0; c < 3; ++c) v[i][c] = (float)src[c * plane + i] * 0.0078125f;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (x0 + i >= Wp) break;
    const size_t idx = (size_t)n * img.ss + ((size_t)y * img.ws + x0 + i) * 8;
    if constexpr (sizeof(T) == 2) {
      const __half2 a = __floats2half2_rn(v[i][0], v[i][1]), c2 = __floats2half2_rn(v[i][2], 0.f);
      uint4 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&c2); o.z = 0u; o.w = 0u;
      *reinterpret_cast<uint4*>(static_cast<__half*>(img.p) + idx) = o;
    } else {
      const float o8[8] = {v[i][0], v[i][1], v[i][2], 0.f, 0.f, 0.f, 0.f, 0.f};
      St<T>::st8(img.p, idx, img.lo, o8);
    }
  }
}

cudaError_t launch_pre_s8(const int8_t* s8, Tens img, int B, int H, int W, cudaStream_t st) {
  const dim3 g(cdiv(cdiv(img.w, 4), 128), img.h, 2 * B);
  if (img.planes == 2) launch_k(k_pre_s8<__half>, g, 128, 0, st, s8, view(img), B, H, W, img.h, img.w);
  else launch_k(k_pre_s8<float>, g, 128, 0, st, s8, view(img), B, H, W, img.h, img.w);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Side-by-side NV12 frame -> the same C8 image (and optionally the s8 tensor the reference builds).
// Restates stereonet_node.cpp:702-738 (L/R split), preprocess.h:128-155 (YUV420TOYUV444 incl. the
// I420-indexing quirk on NV12 data) and preprocess.cpp:1032-1040 (x-128) in one pass.
template <typename T>
__global__ void k_pre_nv12(const PreNv12Params p) {
  pdl_trigger();
  pdl_wait();
  const uint8_t* __restrict__ frames = p.io.frames;
  int8_t* __restrict__ s8 = const_cast<int8_t*>(p.io.s8);       // this kernel PRODUCES the s8 tensor the rest of the pass reads
  const TV img = p.img;
  const int B = p.B, H = p.H, W = p.W;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (x >= p.Wp) return;
  const int b = n % B, view = n / B;
  float v0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (x < W && y < H) {
    const uint8_t* f = frames + (size_t)b * (H * 3 / 2) * (2 * W) + view * W;   // this view's column window
    const int pitch = 2 * W;
    const uint8_t yy = f[(size_t)y * pitch + x];
    uint8_t u, v;
    if (p.correct) {
      const uint8_t* row = f + (size_t)(H + (y >> 1)) * pitch;
      u = row[(x >> 1) * 2];
      v = row[(x >> 1) * 2 + 1];
    } else {
      // the view's chroma block is [H/2][W] bytes; the reference reads it as two [H/2][W/2] planes
      const int k = (y >> 1) * (W >> 1) + (x >> 1);
      const int kv = k + (W * H) / 4;
      u = f[(size_t)(H + k / W) * pitch + k % W];
      v = f[(size_t)(H + kv / W) * pitch + kv % W];
    }
    const int8_t sy = (int8_t)(yy ^ 0x80), su = (int8_t)(u ^ 0x80), sv = (int8_t)(v ^ 0x80);
    v0[0] = (float)sy * 0.0078125f;
    v0[1] = (float)su * 0.0078125f;
    v0[2] = (float)sv * 0.0078125f;
    if (s8) {
      int8_t* d = s8 + (((size_t)b * 6 + view * 3) * H + y) * W + x;
      const size_t plane = (size_t)H * W;
      d[0] = sy; d[plane] = su; d[2 * plane] = sv;
    }
  }
  if (img.p) St<T>::st8(img.p, (size_t)n * img.ss + ((size_t)y * img.ws + x) * 8, img.lo, v0);
}

// img.p == nullptr: only the s8 tensor is produced (the tensor-core path reads that tensor directly, there is no image tensor)
cudaError_t launch_pre_nv12(const uint8_t* frames, Tens img, int8_t* s8, int B, int H, int W, int correct,
                            cudaStream_t st) {
  PreNv12Params p{};
  p.io.frames = frames; p.io.s8 = s8;
  p.B = B; p.H = H; p.W = W; p.correct = correct;
  if (!img.p) {
    if (!s8) return cudaErrorInvalidValue;
    p.Hp = H; p.Wp = W;
    launch_k(k_pre_nv12<__half>, dim3(cdiv(W, 128), H, 2 * B), 128, 0, st, p);
    return cudaGetLastError();
  }
  p.img = view(img); p.Hp = img.h; p.Wp = img.w;
  const dim3 g(cdiv(img.w, 128), img.h, 2 * B);
  if (img.planes == 2) launch_k(k_pre_nv12<__half>, g, 128, 0, st, p);
  else launch_k(k_pre_nv12<float>, g, 128, 0, st, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// M2 cost-volume build.  One CTA per (row y, pair b, output block q of the gwc half): stages the
// 64 left and 64 right gwc channels of that row in shared memory once, then emits all D hypotheses.
//   vol block 0,1 : left concat feature (16 ch), zero where x < d
//   vol block 2,3 : right concat feature shifted right by d, zero where x < d
//   vol block 4..7: group-wise correlation, group = one C8 block of the 256-ch feature:
//                   mean_8( L[g][x][:] * R[g][x-d][:] ), zero where x < d
// Thread (x, j): consecutive threads write consecutive floats of [x][j] -> fully coalesced stores.
struct CostvolParams {
  TV gwc, cat, vol;
  int B, D, h, w;
  int dsplit;            // CTAs sharing one (row, pair, block): each emits D / dsplit hypotheses (more CTAs in flight)
};

// One CTA = one feature row x one (correlation block, concat block) pair q x one slice of the D hypotheses.  A thread owns
// pixel x: the eight 8-channel groups of ITS left feature live in registers for all d, the right feature row sits in shared
// memory as [group][channel half][x][4] (conflict-free 128-bit loads at x - d), and every (x, d) leaves as whole 16-byte
// vectors: consecutive threads write consecutive pixels of a [cb][d][y] row.
template <typename T>
__global__ void __launch_bounds__(128) k_costvol(CostvolParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sR[];                                   // [8 j][2 halves][w][4]
  const int y = blockIdx.x, b = blockIdx.y, q = blockIdx.z & 3, dpart = blockIdx.z >> 2;    // q in 0..3
  const int dper = (p.D + p.dsplit - 1) / p.dsplit, d_lo = dpart * dper, d_hi = min(p.D, d_lo + dper);
  const int w = p.w, D = p.D;
  const size_t goff = (size_t)(q * 8) * p.gwc.slice + (size_t)y * p.gwc.ws * 8;      // block q*8 + j adds j * slice
  for (int i = threadIdx.x; i < 8 * w; i += blockDim.x) {
    const int j = i / w, x = i - j * w;
    float v[8];
    St<T>::ld8(p.gwc.p, (size_t)(p.B + b) * p.gwc.ss + goff + (size_t)j * p.gwc.slice + (size_t)x * 8, p.gwc.lo, v);
    *reinterpret_cast<float4*>(sR + ((j * 2 + 0) * w + x) * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(sR + ((j * 2 + 1) * w + x) * 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();

  const size_t vslice = p.vol.slice;                            // one (cb, d) slice
  const size_t vg = (size_t)b * p.vol.ss + (size_t)(4 + q) * D * vslice + (size_t)y * p.vol.ws * 8;   // gwc block 4+q
  const size_t vc = (size_t)b * p.vol.ss + (size_t)q * D * vslice + (size_t)y * p.vol.ws * 8;         // concat block q
  const size_t csrc = (size_t)((q >> 1) ? p.B + b : b) * p.cat.ss + (size_t)(q & 1) * p.cat.slice + (size_t)y * p.cat.ws * 8;
  const bool shifted = (q >> 1) != 0;
  const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    float L[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) St<T>::ld8(p.gwc.p, (size_t)b * p.gwc.ss + goff + (size_t)j * p.gwc.slice + (size_t)x * 8, p.gwc.lo, L[j]);
    float cl[8];
    St<T>::ld8(p.cat.p, csrc + (size_t)x * 8, p.cat.lo, cl);       // unshifted concat block (left feature)
    for (int d = d_lo; d < d_hi; ++d) {
      if (x >= d) {
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r0 = *reinterpret_cast<const float4*>(sR + ((j * 2 + 0) * w + (x - d)) * 4);
          const float4 r1 = *reinterpret_cast<const float4*>(sR + ((j * 2 + 1) * w + (x - d)) * 4);
          float a = L[j][0] * r0.x;
          a = fmaf(L[j][1], r0.y, a); a = fmaf(L[j][2], r0.z, a); a = fmaf(L[j][3], r0.w, a);
          a = fmaf(L[j][4], r1.x, a); a = fmaf(L[j][5], r1.y, a); a = fmaf(L[j][6], r1.z, a); a = fmaf(L[j][7], r1.w, a);
          g[j] = a * 0.125f;
        }
        St<T>::st8(p.vol.p, vg + (size_t)d * vslice + (size_t)x * 8, p.vol.lo, g);
        if (shifted) {
          float cr[8];
          St<T>::ld8(p.cat.p, csrc + (size_t)(x - d) * 8, p.cat.lo, cr);
          St<T>::st8(p.vol.p, vc + (size_t)d * vslice + (size_t)x * 8, p.vol.lo, cr);
        } else {
          St<T>::st8(p.vol.p, vc + (size_t)d * vslice + (size_t)x * 8, p.vol.lo, cl);
        }
      } else {
        St<T>::st8(p.vol.p, vg + (size_t)d * vslice + (size_t)x * 8, p.vol.lo, zero8);
        St<T>::st8(p.vol.p, vc + (size_t)d * vslice + (size_t)x * 8, p.vol.lo, zero8);
      }
    }
  }
}

cudaError_t launch_costvol(Tens gwc, Tens cat, Tens vol, int B, int D, cudaStream_t st) {
  const int h = gwc.h, w = gwc.w;
  const size_t smem = (size_t)8 * 2 * w * 4 * sizeof(float);
  static const int env_split = getenv("SNB_COSTVOL_DSPLIT") ? atoi(getenv("SNB_COSTVOL_DSPLIT")) : 0;
  // measured at config 2 (D = 24, 272 CTAs at dsplit 1): 20.5 us at dsplit 1, 24.5 at 2, 23.3 at 3, 30.9 at 6 - every extra
  // slice re-stages the right feature row; long D gets one slice per 48 hypotheses
  const int dsplit = env_split > 0 ? std::min(env_split, D) : std::max(1, std::min(8, D / 48));
  CostvolParams p{view(gwc), view(cat), view(vol), B, D, h, w, dsplit};
  if (vol.planes == 2) {
    if (need_attr(5)) cudaFuncSetAttribute(k_costvol<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    launch_k(k_costvol<__half>, dim3(h, B, 4 * dsplit), 128, smem, st, p);
  } else {
    if (need_attr(2)) cudaFuncSetAttribute(k_costvol<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    launch_k(k_costvol<float>, dim3(h, B, 4 * dsplit), 128, smem, st, p);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// M4 softmax over D + soft-argmin, online (single pass over the cost column), fp32 throughout.
__global__ void k_softargmin(const float* __restrict__ cost, float* __restrict__ disp, int D, int hw, float invD) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= hw) return;
  const float* c = cost + (size_t)b * D * hw + i;
  float m = -INFINITY, s = 0.f, t = 0.f;
  for (int d0 = 0; d0 < D; d0 += 8) {          // 8 independent loads in flight, then the (serial) online-softmax update
    float v8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v8[k] = d0 + k < D ? __ldg(c + (size_t)(d0 + k) * hw) : -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = v8[k];
      if (d0 + k >= D) break;
      if (v > m) {
        const float r = expf(m - v);
        s *= r; t *= r; m = v;
      }
      const float e = expf(v - m);
      s += e;
      t = fmaf(e, (float)(d0 + k), t);
    }
  }
  disp[(size_t)b * hw + i] = t / s * invD;
}

cudaError_t launch_softargmin(Plane cost, Plane disp, cudaStream_t st) {
  const int hw = cost.h * cost.w;
  launch_k(k_softargmin, dim3(cdiv(hw, 128), cost.n), 128, 0, st, cost.p, disp.p, cost.d, hw, 1.0f / cost.d);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// M5 glue: channel 0 = x2 bilinear upsample of the disparity (align_corners=False), channels 1-3 =
// left image bilinearly resized to the stage resolution (integer factor f: mean of the central 2x2).
// first three channels of an image pixel: the image tensor (s8 / 128) is exact in fp16, so its lo plane is zero and is
// not read (half the bytes of St<__half>::ld8); fp32 storage reads one float4
template <typename T> __device__ __forceinline__ void img_ld3(const void* base, size_t idx, float (&v)[3]);
template <> __device__ __forceinline__ void img_ld3<float>(const void* base, size_t idx, float (&v)[3]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx));
  v[0] = a.x; v[1] = a.y; v[2] = a.z;
}
template <> __device__ __forceinline__ void img_ld3<__half>(const void* base, size_t idx, float (&v)[3]) {
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + idx));
  const __half2* h2 = reinterpret_cast<const __half2*>(&h);
  const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x;
}

template <typename T>
__global__ void k_refine_in(const float* __restrict__ disp, TV img, TV out, int h, int w, int f) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  const int W2 = 2 * w;
  if (x >= W2) return;
  // PyTorch upsample_bilinear2d, scale 0.5: src = max((dst+0.5)*0.5-0.5, 0)
  const float sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * 0.5f - 0.5f, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly1 = sy - y0, lx1 = sx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* dp = disp + (size_t)b * h * w;
  const float up = ly0 * (lx0 * __ldg(dp + y0 * w + x0) + lx1 * __ldg(dp + y0 * w + x1)) +
                   ly1 * (lx0 * __ldg(dp + y1 * w + x0) + lx1 * __ldg(dp + y1 * w + x1));
  float o[8] = {up, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t ib = (size_t)b * img.ss;
  const int Wf = img.ws;
  if (f == 1) {
    float v[3];
    img_ld3<T>(img.p, ib + ((size_t)y * Wf + x) * 8, v);
    o[1] = v[0]; o[2] = v[1]; o[3] = v[2];
  } else {
    const int yy = y * f + f / 2 - 1, xx = x * f + f / 2 - 1;
    float a[3], bq[3], c[3], d[3];
    img_ld3<T>(img.p, ib + ((size_t)yy * Wf + xx) * 8, a);
    img_ld3<T>(img.p, ib + ((size_t)yy * Wf + xx + 1) * 8, bq);
    img_ld3<T>(img.p, ib + ((size_t)(yy + 1) * Wf + xx) * 8, c);
    img_ld3<T>(img.p, ib + ((size_t)(yy + 1) * Wf + xx + 1) * 8, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[1 + k] = 0.5f * (0.5f * a[k] + 0.5f * bq[k]) + 0.5f * (0.5f * c[k] + 0.5f * d[k]);
  }
  St<T>::st8(out.p, (size_t)b * out.ss + ((size_t)y * out.ws + x) * 8, out.lo, o);
}

cudaError_t launch_refine_in(Plane disp, Tens img_full, Tens out, int B, cudaStream_t st) {
  const int f = img_full.h / out.h;
  const dim3 g(cdiv(out.w, 128), out.h, B);
  if (out.planes == 2) launch_k(k_refine_in<__half>, g, 128, 0, st, disp.p, view(img_full), view(out), disp.h, disp.w, f);
  else launch_k(k_refine_in<float>, g, 128, 0, st, disp.p, view(img_full), view(out), disp.h, disp.w, f);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Output tensor as the reference reads it (stereonet_node.cpp:1033): s32 NCHW [B,1,H,W], cropped
// from the padded map; q = rint(dn * qmul) so that q * 2.60443857769133e-06 * 192 = pixels.
__global__ void k_post_quant(const float* __restrict__ disp, int32_t* __restrict__ out, int H, int W, int Hp, int Wp,
                             float qmul) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  out[((size_t)b * H + y) * W + x] = __float2int_rn(disp[((size_t)b * Hp + y) * Wp + x] * qmul);
}

// four pixels per thread (W % 4 == 0; padded rows are multiples of 4 floats by construction)
__global__ void k_post_quant4(const float* __restrict__ disp, int32_t* __restrict__ out, int H, int W, int Hp, int Wp,
                              float qmul) {
  pdl_trigger();
  pdl_wait();
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(disp + ((size_t)b * Hp + y) * Wp + x));
  int4 q;
  q.x = __float2int_rn(v.x * qmul); q.y = __float2int_rn(v.y * qmul); q.z = __float2int_rn(v.z * qmul); q.w = __float2int_rn(v.w * qmul);
  *reinterpret_cast<int4*>(out + ((size_t)b * H + y) * W + x) = q;
}

cudaError_t launch_post_quant(Plane disp, int32_t* out, int H, int W, float qmul, cudaStream_t st) {
  if (W % 4 == 0 && disp.w % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    launch_k(k_post_quant4, dim3(cdiv(W, 512), H, disp.n), 128, 0, st, (const float*)disp.p, out, H, W, disp.h, disp.w, qmul);
    return cudaGetLastError();
  }
  launch_k(k_post_quant, dim3(cdiv(W, 128), H, disp.n), 128, 0, st, disp.p, out, H, W, disp.h, disp.w, qmul);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// what ParseTensor does on the CPU after the model (parser.cpp:79-118), on the GPU:
//   dis = (float)q * scale;  depth = (float)( (double)(f*B) / ((double)dis * 16.0 * 12.0) / 1000.0 )      [metres]
//   u8  = cv::convertScaleAbs(depth, alpha)   (float multiply, |.|, round-half-even, saturate; values that do not fit an
//         int32 - depth = inf at q = 0, NaN - become 0, as cv2 on x86 does via cvtps2dq)
//   bgr = cv::applyColorMap(u8, COLORMAP_JET)
// alpha = 11 in parser.cpp:115, 9 in the render tool (publisher_member_function.py:82).
__constant__ uint8_t c_jet[256][3] = {
#include "jet_lut.inc"
};

__global__ void k_post_depth_color(const int32_t* __restrict__ q, float* __restrict__ depth, uint8_t* __restrict__ bgr,
                                   size_t n, float scale, float alpha) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float fb = 527.1931762695312f * 119.89382172f;
  const float dis = (float)__ldg(q + i) * scale;
  const float z = (float)((double)fb / ((double)dis * 16.0 * 12.0) / 1000.0);
  if (depth) depth[i] = z;
  if (bgr) {
    const float v = fabsf(z * alpha);
    int iv = v < 2147483648.f ? __float2int_rn(v) : 0;       // false for NaN too
    iv = iv > 255 ? 255 : iv;
    bgr[3 * i] = c_jet[iv][0]; bgr[3 * i + 1] = c_jet[iv][1]; bgr[3 * i + 2] = c_jet[iv][2];
  }
}

cudaError_t launch_post_depth_color(const int32_t* q, float* depth, uint8_t* bgr, size_t n, float scale, float alpha,
                                    cudaStream_t st) {
  launch_k(k_post_depth_color, dim3((unsigned)((n + 255) / 256)), 256, 0, st, q, depth, bgr, n, scale, alpha);
  return cudaGetLastError();
}

}  // namespace snb
