// Storage policies for C8 activation tensors: fp32, or split fp16 (x = hi + lo in two planes).
// idx is the element index of an 8-channel block inside the tensor (multiple of 8);
// lo_off is the element distance from the hi plane to the lo plane (0 for fp32).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace snb {

template <typename T> struct St;

template <> struct St<float> {
  __device__ __forceinline__ static void ld8(const void* base, size_t idx, size_t, float (&v)[8]) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  struct Raw { float4 a, b; };                               // an 8-channel block as loaded, before conversion
  __device__ __forceinline__ static Raw ldraw(const void* base, size_t idx, size_t) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
    Raw r; r.a = __ldg(p); r.b = __ldg(p + 1);
    return r;
  }
  __device__ __forceinline__ static void cvt(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  __device__ __forceinline__ static void st8(void* base, size_t idx, size_t, const float (&v)[8]) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + idx);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __device__ __forceinline__ static float ld1(const void* base, size_t idx, size_t) {
    return __ldg(static_cast<const float*>(base) + idx);
  }
  __device__ __forceinline__ static void st1(void* base, size_t idx, size_t, float v) { static_cast<float*>(base)[idx] = v; }
};

template <> struct St<__half> {
  __device__ __forceinline__ static void ld8(const void* base, size_t idx, size_t lo_off, float (&v)[8]) {
    const __half* p = static_cast<const __half*>(base) + idx;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + lo_off));
    const __half2* h2 = reinterpret_cast<const __half2*>(&h);
    const __half2* l2 = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(h2[j]), b = __half22float2(l2[j]);
      v[2 * j] = a.x + b.x; v[2 * j + 1] = a.y + b.y;
    }
  }
  struct Raw { uint4 h, l; };
  __device__ __forceinline__ static Raw ldraw(const void* base, size_t idx, size_t lo_off) {
    const __half* p = static_cast<const __half*>(base) + idx;
    Raw r; r.h = __ldg(reinterpret_cast<const uint4*>(p)); r.l = __ldg(reinterpret_cast<const uint4*>(p + lo_off));
    return r;
  }
  __device__ __forceinline__ static void cvt(const Raw& r, float (&v)[8]) {
    const __half2* h2 = reinterpret_cast<const __half2*>(&r.h);
    const __half2* l2 = reinterpret_cast<const __half2*>(&r.l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(h2[j]), b = __half22float2(l2[j]);
      v[2 * j] = a.x + b.x; v[2 * j + 1] = a.y + b.y;
    }
  }
  __device__ __forceinline__ static void st8(void* base, size_t idx, size_t lo_off, const float (&v)[8]) {
    uint4 oh, ol;
    __half2* ph = reinterpret_cast<__half2*>(&oh);
    __half2* pl = reinterpret_cast<__half2*>(&ol);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      const float2 hf = __half22float2(hh);
      ph[j] = hh;
      pl[j] = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    }
    __half* p = static_cast<__half*>(base) + idx;
    *reinterpret_cast<uint4*>(p) = oh;
    *reinterpret_cast<uint4*>(p + lo_off) = ol;
  }
  __device__ __forceinline__ static float ld1(const void* base, size_t idx, size_t lo_off) {
    const __half* p = static_cast<const __half*>(base) + idx;
    return __half2float(__ldg(p)) + __half2float(__ldg(p + lo_off));
  }
  __device__ __forceinline__ static void st1(void* base, size_t idx, size_t lo_off, float v) {
    __half* p = static_cast<__half*>(base) + idx;
    const __half h = __float2half_rn(v);
    p[0] = h; p[lo_off] = __float2half_rn(v - __half2float(h));
  }
};

}  // namespace snb
