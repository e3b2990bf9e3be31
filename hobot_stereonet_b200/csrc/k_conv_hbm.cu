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
