// firstconv.0 of the tensor path: NOT a dense contraction.
//
//   k_conv_first      Cin = 3 (the image), 3x3, stride 2, Cout = 32.  13 FLOP per byte moved; the tensor-core version
//                     computed it at stride 1 with 13 of 16 input channels zero (9 TFLOP/s, epilogue bound, 49 us at
//                     config 2).  One thread = one output pixel x 32 channels; the 27 x 32 weights ride in the kernel
//                     parameters, i.e. the constant bank, so every FMA takes its weight as a uniform operand: 19 us.
//                     fp32 accumulation from exact operands (the image, s8 / 128, is exact in fp16: its lo plane is zero).
// (Measured and dropped: CUDA-core versions of the Cout = 1 convolutions conv_out / conv3d_alone - shared-memory tile,
// L1-cached rows with batched loads, per-warp cp.async pipeline - all landed at 44-49 us for the full-resolution conv_out
// against 42 us on k_conv_stream: converting split-fp16 inputs to fp32 costs 3 instructions per FMA.  Those layers stay
// on the tensor pipe, with hi*hi | hi*lo sharing one 16-column group: 2 MMAs per tap instead of 3.)
#include "common.cuh"
#include "kernels.cuh"
#include "store.cuh"

namespace snb {

// ---- firstconv.0 -----------------------------------------------------------------------------------
template <typename T> struct ImgLd;
template <> struct ImgLd<float> {
  __device__ __forceinline__ static void ld3(const void* base, size_t idx, float (&v)[3]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx));
    v[0] = a.x; v[1] = a.y; v[2] = a.z;
  }
};
template <> struct ImgLd<__half> {     // the image (s8 / 128) is exact in fp16: its lo plane is zero and is not read
  __device__ __forceinline__ static void ld3(const void* base, size_t idx, float (&v)[3]) {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + idx));
    const __half2* h2 = reinterpret_cast<const __half2*>(&h);
    const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x;
  }
};

template <typename T>
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
    const int y = 2 * oy + ky - 1;                          // -1 .. H: inside the zero border (pad >= 1)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int x = 2 * ox + kx - 1;
      float v[3];
      ImgLd<T>::ld3(p.in.p, base + ((ptrdiff_t)y * p.in.ws + x) * 8, v);
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
    St<T>::st8(p.out.p, (size_t)n * p.out.ss + (size_t)cb * p.out.slice + ((size_t)oy * p.out.ws + ox) * 8, p.out.lo, g);
  }
}

// W: [32][3][3][3] (co, ci, ky, kx) fp32 host weights, bias [32]
void conv_first_pack(const float* W, const float* bias, ConvFirstParams* p) {
  for (int co = 0; co < 32; ++co) {
    for (int ci = 0; ci < 3; ++ci)
      for (int t = 0; t < 9; ++t) p->w[(ci * 9 + t) * 32 + co] = W[((size_t)co * 3 + ci) * 9 + t];
    p->b[co] = bias[co];
  }
}

cudaError_t launch_conv_first(const ConvFirstParams& p, int N, bool half, cudaStream_t st) {
  const dim3 g(cdiv(p.Wo, 128), p.Ho, N);
  if (half) return launch_k(k_conv_first<__half>, g, 128, 0, st, p);
  return launch_k(k_conv_first<float>, g, 128, 0, st, p);
}

}  // namespace snb
