// Host-side byte formats either side of the model call, behind the C ABI (include/snb200.h).
// Same results as the reference's code, written for throughput: the reference makes three
// full-buffer copies and 5.5 M scalar float quantisations per frame (preprocess.cpp:999-1053);
// here each output byte is produced once.  Bit-exact against oracle/prepost_ref.py.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "../../include/snb200.h"

extern "C" {

// stereonet_node.cpp:702-738 — left = first w bytes of every row, right = last w bytes; rows are
// h luma rows followed by h/2 interleaved-chroma rows of the side-by-side frame.
int snb_pre_split_nv12(const uint8_t* frame, int32_t h, int32_t w2, uint8_t* left, uint8_t* right) {
  if (!frame || !left || !right || h <= 0 || w2 <= 0 || (h & 1) || (w2 & 1)) return SNB_ERR_INVALID;
  const int w = w2 / 2, rows = h + h / 2;
  for (int r = 0; r < rows; ++r) {
    memcpy(left + (size_t)r * w, frame + (size_t)r * w2, w);
    memcpy(right + (size_t)r * w, frame + (size_t)r * w2 + w, w);
  }
  return SNB_OK;
}

// preprocess.h:128-155 Tools::YUV420TOYUV444.  The reference indexes the chroma block as planar
// I420 (U plane at w*h, V plane at w*h*5/4, pitch w/2) whatever the caller passes; with NV12 input
// both planes therefore carry alternating U,V samples.  correct_chroma != 0 de-interleaves NV12.
int snb_pre_yuv420_to_yuv444(const uint8_t* in, uint8_t* out, int32_t w, int32_t h, int32_t correct_chroma) {
  if (!in || !out || w <= 0 || h <= 0 || (w & 1) || (h & 1)) return SNB_ERR_INVALID;
  const size_t wh = (size_t)w * h;
  memcpy(out, in, wh);
  const uint8_t* c = in + wh;
  uint8_t* du = out + wh;
  uint8_t* dv = du + wh;
  const int hw = w / 2;
  for (int i = 0; i < h; i += 2) {
    const uint8_t* su = correct_chroma ? c + (size_t)(i / 2) * w : c + (size_t)(i / 2) * hw;
    const uint8_t* sv = correct_chroma ? su + 1 : su + wh / 4;
    const int step = correct_chroma ? 2 : 1;
    uint8_t* u0 = du + (size_t)i * w; uint8_t* u1 = u0 + w;
    uint8_t* v0 = dv + (size_t)i * w; uint8_t* v1 = v0 + w;
    for (int j = 0; j < hw; ++j) {
      const uint8_t u = su[j * step], v = sv[j * step];
      u0[2 * j] = u0[2 * j + 1] = u1[2 * j] = u1[2 * j + 1] = u;
      v0[2 * j] = v0[2 * j + 1] = v1[2 * j] = v1[2 * j + 1] = v;
    }
  }
  return SNB_OK;
}

// preprocess.cpp:1131-1136, float arithmetic and floor rounding as written there.
int8_t snb_pre_quantize(float value, float scale, float zero_point, float lo, float hi) {
  value = floorf(value / scale + zero_point);
  value = std::min(std::max(value, lo), hi);
  return static_cast<int8_t>(value);
}

// preprocess.cpp:913-1059 — L planes then R planes (:999-1003), each byte Quantize((x-128)/128)
// (:1032-1040) which is x-128 for every byte value (tests/test_oracle_prepost.py proves the
// identity over all 256 inputs), i.e. the stored byte is x ^ 0x80.
int snb_pre_cvt_nv12_to_tensor(const uint8_t* left, const uint8_t* right, int32_t w, int32_t h,
                               int32_t correct_chroma, int8_t* out) {
  if (!left || !right || !out) return SNB_ERR_INVALID;   // "Invalid input data" (preprocess.cpp:919-922)
  if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) return SNB_ERR_INVALID;
  const size_t wh = (size_t)w * h;
  uint8_t* o = reinterpret_cast<uint8_t*>(out);
  int r = snb_pre_yuv420_to_yuv444(left, o, w, h, correct_chroma);
  if (r != SNB_OK) return r;
  r = snb_pre_yuv420_to_yuv444(right, o + 3 * wh, w, h, correct_chroma);
  if (r != SNB_OK) return r;
  const size_t n = 6 * wh;
  for (size_t i = 0; i < n; ++i) o[i] ^= 0x80;   // auto-vectorised at -O3
  return SNB_OK;
}

// stereonet_node.cpp:1033-1049
int64_t snb_post_pack(const int32_t* infer, uint64_t infer_bytes, const uint8_t* jpeg, uint64_t jpeg_bytes,
                      uint8_t* dst, uint64_t cap) {
  if (!infer || !dst || (jpeg_bytes && !jpeg)) return SNB_ERR_INVALID;
  if (cap < infer_bytes + jpeg_bytes) return SNB_ERR_NOMEM;
  memcpy(dst, infer, infer_bytes);
  if (jpeg_bytes) memcpy(dst + infer_bytes, jpeg, jpeg_bytes);
  return (int64_t)(infer_bytes + jpeg_bytes);
}

// parser.cpp:70-71,79-87: float f, B; float dis = (float)q * scale; push_back(f*B/(dis*16.0*12.0)/1000.0)
int snb_post_parse_depth(const int32_t* q, int64_t n, float scale, float* depth_m) {
  if (!q || !depth_m || n < 0) return SNB_ERR_INVALID;
  const float f = 527.1931762695312f, B = 119.89382172f;
  const float fb = f * B;
  for (int64_t i = 0; i < n; ++i) {
    const float dis = static_cast<float>(q[i]) * scale;
    depth_m[i] = static_cast<float>(fb / (dis * 16.0 * 12.0) / 1000.0);
  }
  return SNB_OK;
}

}  // extern "C"
