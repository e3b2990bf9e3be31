// Helpers of the streaming tcgen05 convolution kernel (k_conv_stream.cu): unit decoding, TMEM drains, the split-fp16 emit.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace snb {

using namespace ptx;

struct CsUnit { int cc, n, d, x0, c, i0, nr; };

__device__ __forceinline__ CsUnit cs_decode(const CsParams& p, int u) {
  CsUnit r;
  const int chunk = u % p.nchunk; u /= p.nchunk;
  r.c = u % p.dil; u /= p.dil;
  const int strip = u % p.strips; u /= p.strips;
  r.d = u % p.D; u /= p.D;
  r.n = u % p.N;
  r.cc = u / p.N;                               // slowest index: a CTA rarely changes its weight slice
  r.x0 = strip * 128;
  const int rc = r.c < p.H ? (p.H - r.c + p.dil - 1) / p.dil : 0;
  r.i0 = chunk * p.rpc;
  r.nr = min(rc, r.i0 + p.rpc) - r.i0;
  return r;
}

__device__ __forceinline__ void cs_ld3x16(uint32_t c0, uint32_t c1, uint32_t c2, float (&v0)[16], float (&v1)[16], float (&v2)[16]) {
  uint32_t r[48];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%48];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%49];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47}, [%50];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47])
      : "r"(c0), "r"(c1), "r"(c2) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(r[i]); v1[i] = __uint_as_float(r[16 + i]); v2[i] = __uint_as_float(r[32 + i]); }
}

// One 16-channel chunk of a SPLIT job (NCOL = 96 accumulator columns per half): [main | corr] (+)= A_hi x [W_hi | W_lo] at
// N = 192 and corr += A_lo x W_hi at N = 96, for the three kernel columns (A shifted by `dil16`, weights by `wkx16`).
__device__ __forceinline__ void cs_issue_split(uint32_t dcol, uint64_t a_hi, uint64_t a_lo, uint64_t w_hi, uint64_t dil16,
                                               uint64_t wkx16, uint32_t idesc, uint32_t idesc2, uint32_t acc) {
  constexpr uint32_t NCOL = 96;
  umma_f16(dcol, a_hi, w_hi, idesc2, acc);
  umma_f16_acc(dcol + NCOL, a_lo, w_hi, idesc);
  umma_f16_acc(dcol, a_hi + dil16, w_hi + wkx16, idesc2);
  umma_f16_acc(dcol + NCOL, a_lo + dil16, w_hi + wkx16, idesc);
  umma_f16_acc(dcol, a_hi + 2 * dil16, w_hi + 2 * wkx16, idesc2);
  umma_f16_acc(dcol + NCOL, a_lo + 2 * dil16, w_hi + 2 * wkx16, idesc);
}

// Drain one SPLIT job (main at `col0`, corr 96 columns further, each [ky][32 channels]): the finished output row goes to f,
// the two partial rows roll over (a new one is born with the bias).
__device__ __forceinline__ void cs_drain_split(uint32_t col0, float (&a0)[32], float (&a1)[32], float (&f)[32], const float* s_bias) {
  constexpr uint32_t NCOL = 96;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float v0[16], v1[16], v2[16], c0[16], c1[16], c2[16];
    const uint32_t col = col0 + hf * 16;
    cs_ld3x16(col, col + 32, col + 64, v0, v1, v2);
    cs_ld3x16(col + NCOL, col + NCOL + 32, col + NCOL + 64, c0, c1, c2);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      f[hf * 16 + c] = a0[hf * 16 + c] + (v2[c] + c2[c]);
      a0[hf * 16 + c] = a1[hf * 16 + c] + (v1[c] + c1[c]);
      a1[hf * 16 + c] = (v0[c] + c0[c]) + s_bias[hf * 16 + c];
    }
  }
}

__device__ __forceinline__ void cs_split8(const float* f, uint4& oh, uint4& ol) {
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


}  // namespace snb
