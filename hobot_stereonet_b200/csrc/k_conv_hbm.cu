// Few-input-channel convolutions of the tensor path: NOT dense contractions.
//
//   k_conv_first<3, 2>   firstconv.0: Cin = 3 (the image), 3x3, stride 2, Cout = 32.  13 FLOP per byte moved; the tensor-core
//                        version computed it at stride 1 with 13 of 16 input channels zero (9 TFLOP/s, epilogue bound, 49 us
//                        at config 2): 19 us here.  The image, s8 / 128, is exact in fp16: its lo plane is zero and not read.
//   k_conv_first<4, 1>   refinement conv_in (Cin = 4, stride 1): measured 38.6 us at full resolution against 40.0 us on
//                        k_conv_tc (600 M FMAs: 42 % of the fp32 peak) - not dispatched, kept as the instantiation to retry.
//   One thread = one output pixel x 32 channels; the 9 x Cin x 32 weights ride in the kernel parameters, i.e. the constant
//   bank, so every FMA takes its weight as a uniform operand.  fp32 accumulation from exact (hi + lo) operands.
// (Measured and dropped: CUDA-core versions of the Cout = 1 convolutions conv_out / conv3d_alone - shared-memory tile,
// L1-cached rows with batched loads, per-warp cp.async pipeline - all landed at 44-49 us for the full-resolution conv_out
// against 42 us on k_conv_stream: converting split-fp16 inputs to fp32 costs 3 instructions per FMA.  Those layers stay
// on the tensor pipe, with hi*hi | hi*lo sharing one 16-column group: 2 MMAs per tap instead of 3.)
#include "common.cuh"
#include "kernels.cuh"
#include "store.cuh"

namespace snb {

// ---- few-input-channel convolutions: firstconv.0 (3 -> 32, stride 2) and refinement conv_in (4 -> 32) --------------
// first CIN (<= 4) channels of pixel `idx` of a split-fp16 C8 tensor as fp32; HI_ONLY: the lo plane is known to be zero
template <int CIN, bool HI_ONLY>
__device__ __forceinline__ void ld_few(const void* base, size_t idx, size_t lo, float (&v)[CIN]) {
  const __half* p = static_cast<const __half*>(base) + idx;
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(p));
  const __half2* h2 = reinterpret_cast<const __half2*>(&h);
  float f[4];
  { const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]); f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; }
  if (!HI_ONLY) {
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(p + lo));
    const __half2* l2 = reinterpret_cast<const __half2*>(&l);
    const float2 a = __half22float2(l2[0]), b = __half22float2(l2[1]);
    f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y;
  }
#pragma unroll
  for (int c = 0; c < CIN; ++c) v[c] = f[c];
}

template <int CIN, int STRIDE, bool HI_ONLY>
__global__ void __launch_bounds__(128) k_conv_first(const ConvFirstParams p) {
  pdl_trigger();
  pdl_wait();
  const int ox = blockIdx.x * 128 + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
  if (ox >= p.Wo) return;
  float acc[32];
#pragma unroll
  for (int co = 0; co < 32; ++co) acc[co] = p.b[co];
  const size_t base = (size_t)n * p.in.ss;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int y = STRIDE * oy + ky - 1;                     // -1 .. H: inside the zero border (pad >= 1)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int x = STRIDE * ox + kx - 1;
      float v[CIN];
      ld_few<CIN, HI_ONLY>(p.in.p, base + ((ptrdiff_t)y * p.in.ws + x) * 8, p.in.lo, v);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int co = 0; co < 32; ++co) acc[co] = fmaf(v[ci], p.w[(ci * 9 + ky * 3 + kx) * 32 + co], acc[co]);
    }
  }
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = p.relu ? fmaxf(acc[cb * 8 + q], 0.f) : acc[cb * 8 + q];
    St<__half>::st8(p.out.p, (size_t)n * p.out.ss + (size_t)cb * p.out.slice + ((size_t)oy * p.out.ws + ox) * 8, p.out.lo, g);
  }
}

// W: [32][cin][3][3] (co, ci, ky, kx) fp32 host weights, bias [32]; cin = 3 or 4
void conv_first_pack(const float* W, const float* bias, int cin, ConvFirstParams* p) {
  memset(p->w, 0, sizeof(p->w));
  for (int co = 0; co < 32; ++co) {
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < 9; ++t) p->w[(ci * 9 + t) * 32 + co] = W[((size_t)co * cin + ci) * 9 + t];
    p->b[co] = bias[co];
  }
  p->cin = cin;
}

// split-fp16 C8 tensors only; stride 2 with cin = 3 (the image: lo plane zero), stride 1 with cin = 4
cudaError_t launch_conv_first(const ConvFirstParams& p, int N, cudaStream_t st) {
  const dim3 g(cdiv(p.Wo, 128), p.Ho, N);
  if (p.cin == 3 && p.stride == 2) return launch_k(k_conv_first<3, 2, true>, g, 128, 0, st, p);
  if (p.cin == 4 && p.stride == 1) return launch_k(k_conv_first<4, 1, false>, g, 128, 0, st, p);
  return cudaErrorInvalidValue;
}

}  // namespace snb

namespace snb {

// ---- firstconv.0 straight from the model input tensor ---------------------------------------------------------------
// The s8 NCHW [B,6,H,W] tensor the reference hands to DnnNode::Run (preprocess.cpp:1032-1056) is read as it is: value =
// s8 / 128 (exact), channels 0-2 = left view, 3-5 = right view, zero outside the valid H x W (the bottom/right padding to a
// multiple of 2^K and the convolution's own zero padding are the same zeros).  No intermediate image tensor exists on the
// tensor-core path: the 33 MB C8 image k_pre_s8 used to write is gone, and this kernel moves 3.1 MB instead of 16.7.
__global__ void __launch_bounds__(128) k_conv_first_s8(const ConvFirstS8Params p) {
  pdl_trigger();
  pdl_wait();
  const int ox = blockIdx.x * 128 + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
  if (ox >= p.Wo) return;
  const int b = n % p.B, view = n / p.B;
  const int8_t* src = p.io.s8 + ((size_t)b * 6 + view * 3) * p.H * p.W;
  const size_t plane = (size_t)p.H * p.W;
  // (measured: staging the CTA's 3 x 3 x 257-byte window in shared memory first is slower, 35 vs 26 us at config 2)
  float acc[32];
#pragma unroll
  for (int co = 0; co < 32; ++co) acc[co] = p.b[co];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int y = 2 * oy + ky - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int x = 2 * ox + kx - 1;
      float v[3] = {0.f, 0.f, 0.f};
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        const int8_t* q = src + (size_t)y * p.W + x;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) v[ci] = (float)__ldg(q + ci * plane) * 0.0078125f;
      }
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int co = 0; co < 32; ++co) acc[co] = fmaf(v[ci], p.w[(ci * 9 + ky * 3 + kx) * 32 + co], acc[co]);
    }
  }
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = p.relu ? fmaxf(acc[cb * 8 + q], 0.f) : acc[cb * 8 + q];
    St<__half>::st8(p.out.p, (size_t)n * p.out.ss + (size_t)cb * p.out.slice + ((size_t)oy * p.out.ws + ox) * 8, p.out.lo, g);
  }
}

void conv_first_s8_pack(const float* W, const float* bias, ConvFirstS8Params* p) {
  for (int co = 0; co < 32; ++co) {
    for (int ci = 0; ci < 3; ++ci)
      for (int t = 0; t < 9; ++t) p->w[(ci * 9 + t) * 32 + co] = W[((size_t)co * 3 + ci) * 9 + t];
    p->b[co] = bias[co];
  }
}

cudaError_t launch_conv_first_s8(ConvFirstS8Params p, const IoPtrs& io, int B, cudaStream_t st) {
  p.io = io; p.B = B;
  const dim3 g(cdiv(p.Wo, 128), p.Ho, 2 * B);
  return launch_k(k_conv_first_s8, g, 128, 0, st, p);
}

// ---- M4 + the head of M5 in one kernel ------------------------------------------------------------------------------
// BASELINE.json north_star: "soft-argmin disparity regression fused with bilinear upsample + edge-aware refinement".  One
// launch per refinement stage does what used to be three (k_softargmin, k_refine_in, conv_in on k_conv_tc) and two HBM
// tensors (the 4-channel refinement input and, at stage 0, nothing but the cost tensor is read):
//   stage 0     cost [B][D][h][w] --softmax over D, sum d*p (online, fp32)--> disp0 (written for the residual of conv_out)
//   all stages  up = x2 bilinear(disp) (align_corners = False)  ||  left image at the stage's resolution, read from the
//               s8 input tensor (factor f: mean of the central 2 x 2 of every f x f block = exact bilinear)
//               -> conv_in 3x3 (4 -> 32) + bias + ReLU -> split-fp16 C8 feature tensor
// A CTA owns 4 x 32 output pixels: the coarse disparities it needs (4 x 18, clamped to the map) and the 6 x 34 x 4-channel
// input tile live in shared memory only.  Soft-argmin: DL lanes share one pixel's D hypotheses (DL = 1 for D <= 32, else
// 4) and merge their partial (max, sum, weighted sum) with warp shuffles.  The convolution is 1152 FMAs per pixel with the
// weights as uniform operands from the constant bank (the kernel parameters), fp32 throughout.
// (Measured and dropped: four output pixels per thread, so that each constant-bank weight feeds four FMAs - 70 vs 38 us at full
// resolution, and 46 vs 26 us for the same change in k_conv_first_s8: the register tile halves the resident warps.)
constexpr int RH_TY = 4, RH_TX = 32, RH_CR = 4, RH_CC = 18;

template <bool STAGE0>
__global__ void __launch_bounds__(128) k_refine_head(const RefineHeadParams p) {
  __shared__ float s_d[RH_CR][RH_CC + 2];
  __shared__ float4 s_in[RH_TY + 2][RH_TX + 2];
  pdl_trigger();
  pdl_wait();
  const int t = threadIdx.x, b = blockIdx.z;
  const int fy0 = blockIdx.y * RH_TY, fx0 = blockIdx.x * RH_TX;
  const int cy0 = fy0 / 2 - 1, cx0 = fx0 / 2 - 1;                    // coarse tile origin (may be -1: clamped below)
  const int hw = p.h * p.w;
  // ---- 1. coarse disparities of the tile
  if (STAGE0) {
    const int DL = p.D > 32 ? 4 : 1;
    const float* cost = p.src + (size_t)b * p.D * hw;
    for (int i0 = 0; i0 < RH_CR * RH_CC; i0 += 128 / DL) {
      const int i = i0 + t / DL, sub = t % DL;
      const bool act = i < RH_CR * RH_CC;
      const int r = act ? i / RH_CC : 0, cc = act ? i % RH_CC : 0;
      const int y = min(max(cy0 + r, 0), p.h - 1), x = min(max(cx0 + cc, 0), p.w - 1);
      const float* c = cost + (size_t)y * p.w + x;
      float m = -INFINITY, s = 0.f, ts = 0.f;
      for (int d0 = sub; d0 < p.D; d0 += 8 * DL) {                   // 8 independent loads in flight, then the online-softmax update
        float v8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v8[k] = d0 + k * DL < p.D ? __ldg(c + (size_t)(d0 + k * DL) * hw) : -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float v = v8[k];
          if (d0 + k * DL >= p.D) break;
          if (v > m) { const float rr = expf(m - v); s *= rr; ts *= rr; m = v; }
          const float e = expf(v - m);
          s += e;
          ts = fmaf(e, (float)(d0 + k * DL), ts);
        }
      }
      for (int o = 1; o < DL; o <<= 1) {                             // merge the DL partial soft-argmins of this pixel
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o), t2 = __shfl_xor_sync(0xffffffffu, ts, o);
        const float mm = fmaxf(m, m2);
        const float e1 = m == -INFINITY ? 0.f : expf(m - mm), e2 = m2 == -INFINITY ? 0.f : expf(m2 - mm);
        s = s * e1 + s2 * e2; ts = ts * e1 + t2 * e2; m = mm;
      }
      if (act && sub == 0) {
        const float dsp = ts / s * p.invD;
        s_d[r][cc] = dsp;
        // the CTA that owns a coarse pixel (rows 1..2, columns 1..16 of the tile) publishes it for conv_out's residual
        if (r >= 1 && r <= 2 && cc >= 1 && cc <= 16 && cy0 + r < p.h && cx0 + cc < p.w) p.disp0[(size_t)b * hw + (size_t)(cy0 + r) * p.w + cx0 + cc] = dsp;
      }
    }
  } else {
    const float* dp = p.src + (size_t)b * hw;
    if (t < RH_CR * RH_CC) {
      const int r = t / RH_CC, cc = t % RH_CC;
      s_d[r][cc] = __ldg(dp + (size_t)min(max(cy0 + r, 0), p.h - 1) * p.w + min(max(cx0 + cc, 0), p.w - 1));
    }
  }
  __syncthreads();
  // ---- 2. the 4-channel input tile with its one-pixel halo: zero outside the stage's map (the convolution's zero padding)
  const int hs = 2 * p.h, ws = 2 * p.w;
  const int8_t* img = p.io.s8 + (size_t)b * 6 * p.H * p.W;
  const size_t plane = (size_t)p.H * p.W;
  for (int i = t; i < (RH_TY + 2) * (RH_TX + 2); i += 128) {
    const int r = i / (RH_TX + 2), cc = i % (RH_TX + 2);
    const int y = fy0 + r - 1, x = fx0 + cc - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < hs && x >= 0 && x < ws) {
      // PyTorch upsample_bilinear2d, scale 0.5: src = max((dst + 0.5) * 0.5 - 0.5, 0)
      const float sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * 0.5f - 0.5f, 0.f);
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = min(y0 + 1, p.h - 1), x1 = min(x0 + 1, p.w - 1);
      const float ly1 = sy - y0, lx1 = sx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      v.x = ly0 * (lx0 * s_d[y0 - cy0][x0 - cx0] + lx1 * s_d[y0 - cy0][x1 - cx0]) + ly1 * (lx0 * s_d[y1 - cy0][x0 - cx0] + lx1 * s_d[y1 - cy0][x1 - cx0]);
      float c3[3];
      if (p.f == 1) {
        const bool in = y < p.H && x < p.W;
#pragma unroll
        for (int k = 0; k < 3; ++k) c3[k] = in ? (float)__ldg(img + k * plane + (size_t)y * p.W + x) * 0.0078125f : 0.f;
      } else {
        const int yy = y * p.f + p.f / 2 - 1, xx = x * p.f + p.f / 2 - 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int8_t* q = img + k * plane;
          const float a = yy < p.H && xx < p.W ? (float)__ldg(q + (size_t)yy * p.W + xx) * 0.0078125f : 0.f;
          const float bq = yy < p.H && xx + 1 < p.W ? (float)__ldg(q + (size_t)yy * p.W + xx + 1) * 0.0078125f : 0.f;
          const float c = yy + 1 < p.H && xx < p.W ? (float)__ldg(q + (size_t)(yy + 1) * p.W + xx) * 0.0078125f : 0.f;
          const float d = yy + 1 < p.H && xx + 1 < p.W ? (float)__ldg(q + (size_t)(yy + 1) * p.W + xx + 1) * 0.0078125f : 0.f;
          c3[k] = 0.5f * (0.5f * a + 0.5f * bq) + 0.5f * (0.5f * c + 0.5f * d);
        }
      }
      v.y = c3[0]; v.z = c3[1]; v.w = c3[2];
    }
    s_in[r][cc] = v;
  }
  __syncthreads();
  // ---- 3. conv_in 3x3 (4 -> 32) + ReLU, one output pixel per thread
  const int ty = t / RH_TX, tx = t % RH_TX;
  const int oy = fy0 + ty, ox = fx0 + tx;
  if (oy >= hs || ox >= ws) return;
  float acc[32];
#pragma unroll
  for (int co = 0; co < 32; ++co) acc[co] = p.b[co];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4 v = s_in[ty + ky][tx + kx];
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int co = 0; co < 32; ++co) acc[co] = fmaf(vv[ci], p.wgt[(ci * 9 + ky * 3 + kx) * 32 + co], acc[co]);
    }
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = fmaxf(acc[cb * 8 + q], 0.f);
    St<__half>::st8(p.out.p, (size_t)b * p.out.ss + (size_t)cb * p.out.slice + ((size_t)oy * p.out.ws + ox) * 8, p.out.lo, g);
  }
}

void refine_head_pack(const float* W, const float* bias, RefineHeadParams* p) {
  for (int co = 0; co < 32; ++co) {
    for (int ci = 0; ci < 4; ++ci)
      for (int t = 0; t < 9; ++t) p->wgt[(ci * 9 + t) * 32 + co] = W[((size_t)co * 4 + ci) * 9 + t];
    p->b[co] = bias[co];
  }
}

cudaError_t launch_refine_head(RefineHeadParams p, const IoPtrs& io, int B, cudaStream_t st) {
  p.io = io; p.B = B;
  const dim3 g(cdiv(2 * p.w, RH_TX), cdiv(2 * p.h, RH_TY), B);
  return p.stage0 ? launch_k(k_refine_head<true>, g, 128, 0, st, p) : launch_k(k_refine_head<false>, g, 128, 0, st, p);
}

}  // namespace snb

namespace snb {

// ---- 1x1 convolutions that stand alone (layer1.0's shortcut, lastconv.1) ---------------------------------------------------
// 32 -> 32 and 128 -> 16 channels at 1/4 and 1/8 resolution: 0.07 / 0.03 GFLOP.  As the centre tap of a 3x3 on the streaming
// tensor kernel they cost 9x the MMAs and a full pipeline start (15-18 us each); here one thread owns one pixel, walks its input
// channel blocks (exact hi + lo -> fp32) and keeps COUT fp32 accumulators, with the weights broadcast from shared memory.
template <int COUT>
__global__ void __launch_bounds__(64) k_conv1x1(const Conv1x1Params p) {
  extern __shared__ __align__(16) float s_w[];   // [cin][COUT]
  pdl_trigger();
  for (int i = threadIdx.x; i < p.cbin * 8 * COUT; i += 64) s_w[i] = p.wgt[i];     // constants of the pass: before the wait
  pdl_wait();
  __syncthreads();
  const int idx = blockIdx.x * 64 + threadIdx.x, n = blockIdx.y;
  if (idx >= p.h * p.w) return;
  const int y = idx / p.w, x = idx - y * p.w;
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = __ldg(p.bias + co);
  const size_t ibase = (size_t)n * p.in.ss + ((size_t)y * p.in.ws + x) * 8;
  // 16 K pixels at most: the kernel is latency-bound, so small CTAs (every SM gets a few) and four channel blocks of loads in flight
#pragma unroll 4
  for (int cb = 0; cb < p.cbin; ++cb) {
    float v[8];
    St<__half>::ld8(p.in.p, ibase + (size_t)cb * p.in.slice, p.in.lo, v);
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int co = 0; co < COUT; co += 4) {        // one 128-bit broadcast load feeds four FMAs (the scalar version was LDS-issue bound)
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[(cb * 8 + e) * COUT + co]);
        acc[co] = fmaf(v[e], w4.x, acc[co]); acc[co + 1] = fmaf(v[e], w4.y, acc[co + 1]);
        acc[co + 2] = fmaf(v[e], w4.z, acc[co + 2]); acc[co + 3] = fmaf(v[e], w4.w, acc[co + 3]);
      }
  }
  const size_t obase = (size_t)n * p.out.ss + ((size_t)y * p.out.ws + x) * 8;
#pragma unroll
  for (int cb = 0; cb < COUT / 8; ++cb) {
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = p.relu ? fmaxf(acc[cb * 8 + q], 0.f) : acc[cb * 8 + q];
    St<__half>::st8(p.out.p, obase + (size_t)cb * p.out.slice, p.out.lo, g);
  }
}

// W: [cout][cin] fp32 host weights -> [cin][cout] (cin padded to whole 8-channel blocks of the input tensor)
void conv1x1_pack(const float* W, int cout, int cin, int cbin, std::vector<float>& out) {
  out.assign((size_t)cbin * 8 * cout, 0.f);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci) out[(size_t)ci * cout + co] = W[(size_t)co * cin + ci];
}

cudaError_t launch_conv1x1(Conv1x1Params p, int cout, int N, cudaStream_t st) {
  const dim3 g(cdiv(p.h * p.w, 64), N);
  const size_t smem = (size_t)p.cbin * 8 * cout * sizeof(float);
  if (cout == 16) return launch_k(k_conv1x1<16>, g, 64, smem, st, p);
  if (cout == 32) return launch_k(k_conv1x1<32>, g, 64, smem, st, p);
  return cudaErrorInvalidValue;
}

}  // namespace snb
