// Baseline JPEG of the left camera view, the second half of the payload the node publishes:
//   [s32 x H*W] || [JPEG(left)]          stereonet_node.cpp:1033-1049, consumer publisher_member_function.py:57-98
// The reference builds it on the host with OpenCV (stereonet_node.cpp:749-782): cv::cvtColor(NV12 -> BGR,
// CV_YUV2BGR_NV12) then cv::imencode(".jpg", bgr, default parameters = quality 95, 4:2:0).  OpenCV's C++ API is not a
// dependency of this build, so the same two steps are restated here:
//   1. NV12 -> BGR with OpenCV's ITU-R BT.601 fixed-point formula (bit-exact, tests/test_jpeg.py checks it against cv2);
//   2. BGR -> JFIF YCbCr, 2x2 chroma averaging, 8x8 forward DCT, IJG quality-95 quantisation, the Annex-K Huffman
//      tables, JFIF container.
// Any baseline decoder reads the result (the render tool uses cv2.imdecode); the bytes differ from libjpeg's (different
// DCT rounding), the decoded image does not beyond JPEG's own loss.  Host-only, single-threaded like the reference.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/snb200.h"

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// ITU-T T.81 Annex K.1 quantisation tables (natural order)
const uint8_t kQLuma[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                            14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                            18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                            49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kQChroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                              99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                              99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 Huffman tables: code-length counts and symbols
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct Huff { uint16_t code[256]; uint8_t len[256]; };

void build_huff(const uint8_t* bits, const uint8_t* vals, Huff* h) {
  memset(h, 0, sizeof(*h));
  uint32_t code = 0;
  int k = 0;
  for (int l = 1; l <= 16; ++l) {
    for (int i = 0; i < bits[l - 1]; ++i, ++k) { h->code[vals[k]] = (uint16_t)code++; h->len[vals[k]] = (uint8_t)l; }
    code <<= 1;
  }
}

struct BitWriter {
  uint8_t* dst; uint64_t cap, n = 0;
  uint64_t acc = 0; int nbits = 0;          // the low `nbits` bits of acc are pending, most significant first
  bool overflow = false;
  void byte(uint8_t b) { if (n < cap) dst[n] = b; else overflow = true; ++n; }
  void put(uint32_t code, int len) {
    acc = (acc << len) | (code & ((1u << len) - 1)); nbits += len;
    if (nbits >= 32) {                        // flush four bytes at once; 0xFF bytes (rare) take the byte-stuffing path
      const uint32_t w = (uint32_t)(acc >> (nbits - 32));
      nbits -= 32;
      if (n + 4 <= cap && !((w & ~(w + 0x01010101u)) & 0x80808080u)) {      // no byte of w is 0xFF
        dst[n] = (uint8_t)(w >> 24); dst[n + 1] = (uint8_t)(w >> 16); dst[n + 2] = (uint8_t)(w >> 8); dst[n + 3] = (uint8_t)w;
        n += 4;
      } else {
        for (int s = 24; s >= 0; s -= 8) { const uint8_t b = (uint8_t)(w >> s); byte(b); if (b == 0xff) byte(0); }
      }
    }
  }
  void flush() {
    while (nbits >= 8) { const uint8_t b = (uint8_t)(acc >> (nbits - 8)); byte(b); if (b == 0xff) byte(0); nbits -= 8; }
    if (nbits) { const uint8_t b = (uint8_t)(((acc << (8 - nbits)) | ((1u << (8 - nbits)) - 1)) & 0xff); byte(b); if (b == 0xff) byte(0); nbits = 0; }
  }
  void u16(uint32_t v) { byte((uint8_t)(v >> 8)); byte((uint8_t)v); }
};

// separable float forward DCT (AAN), outputs scaled by 8 * the AAN factors folded into the quantiser
void fdct8x8(float* d) {
  for (int pass = 0; pass < 2; ++pass) {
    const int s = pass == 0 ? 1 : 8, t = pass == 0 ? 8 : 1;      // pass 0: rows, pass 1: columns
    for (int i = 0; i < 8; ++i) {
      float* p = d + i * t;
      const float t0 = p[0] + p[7 * s], t7 = p[0] - p[7 * s], t1 = p[s] + p[6 * s], t6 = p[s] - p[6 * s];
      const float t2 = p[2 * s] + p[5 * s], t5 = p[2 * s] - p[5 * s], t3 = p[3 * s] + p[4 * s], t4 = p[3 * s] - p[4 * s];
      float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
      p[0] = t10 + t11; p[4 * s] = t10 - t11;
      const float z1 = (t12 + t13) * 0.707106781f;
      p[2 * s] = t13 + z1; p[6 * s] = t13 - z1;
      t10 = t4 + t5; t11 = t5 + t6; t12 = t6 + t7;
      const float z5 = (t10 - t12) * 0.382683433f, z2 = 0.541196100f * t10 + z5, z4 = 1.306562965f * t12 + z5, z3 = t11 * 0.707106781f;
      const float z11 = t7 + z3, z13 = t7 - z3;
      p[5 * s] = z13 + z2; p[3 * s] = z13 - z2; p[s] = z11 + z4; p[7 * s] = z11 - z4;
    }
  }
}

inline uint8_t sat8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

struct Encoder {
  Huff dc[2], ac[2];
  float rq[2][64];             // reciprocal quantiser incl. AAN scale factors, natural order
  uint8_t qt[2][64];           // quantisation tables as written to DQT (natural order)
  BitWriter bw;
  int last_dc[3] = {0, 0, 0};

  void init(int quality) {
    build_huff(kDcLumaBits, kDcVals, &dc[0]); build_huff(kDcChromaBits, kDcVals, &dc[1]);
    build_huff(kAcLumaBits, kAcLumaVals, &ac[0]); build_huff(kAcChromaBits, kAcChromaVals, &ac[1]);
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;          // IJG jpeg_quality_scaling
    static const double aan[8] = {1.0, 1.387039845, 1.306562965, 1.175875602, 1.0, 0.785694958, 0.541196100, 0.275899379};
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i < 64; ++i) {
        int q = ((t ? kQChroma[i] : kQLuma[i]) * scale + 50) / 100;
        q = q < 1 ? 1 : (q > 255 ? 255 : q);
        qt[t][i] = (uint8_t)q;
        rq[t][i] = (float)(1.0 / (q * aan[i >> 3] * aan[i & 7] * 8.0));
      }
  }

  void block(float* d, int comp) {
    const int t = comp ? 1 : 0;
    fdct8x8(d);
    int q[64];
    float qf[64];
    // round to nearest (ties to even, as lrintf) with the 1.5 * 2^23 trick: branch-free, so the loop vectorises
    for (int i = 0; i < 64; ++i) qf[i] = d[i] * rq[t][i] + 12582912.f;
    for (int i = 0; i < 64; ++i) { int32_t b; memcpy(&b, &qf[kZigzag[i]], 4); q[i] = b - 0x4b400000; }
    int diff = q[0] - last_dc[comp];
    last_dc[comp] = q[0];
    auto emit = [&](const Huff& h, int run, int v) {
      const int a = v < 0 ? -v : v;
      const int nb = a ? 32 - __builtin_clz((unsigned)a) : 0;
      const int sym = (run << 4) | nb;
      bw.put(h.code[sym], h.len[sym]);
      if (nb) bw.put((uint32_t)(v < 0 ? v - 1 : v), nb);
    };
    emit(dc[t], 0, diff);
    int last = 63;
    while (last > 0 && !q[last]) --last;
    int run = 0;
    for (int i = 1; i <= last; ++i) {
      if (!q[i]) { ++run; continue; }
      while (run > 15) { bw.put(ac[t].code[0xf0], ac[t].len[0xf0]); run -= 16; }
      emit(ac[t], run, q[i]);
      run = 0;
    }
    if (last < 63) bw.put(ac[t].code[0], ac[t].len[0]);
  }

  void headers(int w, int h) {
    bw.u16(0xffd8);
    bw.u16(0xffe0); bw.u16(16); for (char c : {'J', 'F', 'I', 'F', '\0'}) bw.byte((uint8_t)c);
    bw.u16(0x0101); bw.byte(0); bw.u16(1); bw.u16(1); bw.byte(0); bw.byte(0);
    for (int t = 0; t < 2; ++t) {
      bw.u16(0xffdb); bw.u16(67); bw.byte((uint8_t)t);
      for (int i = 0; i < 64; ++i) bw.byte(qt[t][kZigzag[i]]);
    }
    bw.u16(0xffc0); bw.u16(17); bw.byte(8); bw.u16((uint32_t)h); bw.u16((uint32_t)w); bw.byte(3);
    bw.byte(1); bw.byte(0x22); bw.byte(0);          // Y: 2x2 sampling, table 0
    bw.byte(2); bw.byte(0x11); bw.byte(1);          // Cb
    bw.byte(3); bw.byte(0x11); bw.byte(1);          // Cr
    auto dht = [&](int cls_id, const uint8_t* bits, const uint8_t* vals, int nvals) {
      bw.u16(0xffc4); bw.u16((uint32_t)(19 + nvals)); bw.byte((uint8_t)cls_id);
      for (int i = 0; i < 16; ++i) bw.byte(bits[i]);
      for (int i = 0; i < nvals; ++i) bw.byte(vals[i]);
    };
    dht(0x00, kDcLumaBits, kDcVals, 12); dht(0x10, kAcLumaBits, kAcLumaVals, 162);
    dht(0x01, kDcChromaBits, kDcVals, 12); dht(0x11, kAcChromaBits, kAcChromaVals, 162);
    bw.u16(0xffda); bw.u16(12); bw.byte(3);
    bw.byte(1); bw.byte(0x00); bw.byte(2); bw.byte(0x11); bw.byte(3); bw.byte(0x11);
    bw.byte(0); bw.byte(63); bw.byte(0);
  }
};

}  // namespace

extern "C" {

// cv::cvtColor(nv12, bgr, CV_YUV2BGR_NV12) (stereonet_node.cpp:775-777): OpenCV's ITU-R BT.601 limited-range
// fixed-point conversion (imgproc color_yuv: ITUR_BT_601_* constants, 20-bit shift), one chroma sample per 2x2 block.
int snb_pre_nv12_to_bgr(const uint8_t* nv12, int32_t w, int32_t h, uint8_t* bgr) {
  if (!nv12 || !bgr || w <= 0 || h <= 0 || (w & 1) || (h & 1)) return SNB_ERR_INVALID;
  const int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SHIFT = 20, HALF = 1 << (SHIFT - 1);
  const uint8_t* uvp = nv12 + (size_t)w * h;
  for (int y = 0; y < h; ++y) {
    const uint8_t* yr = nv12 + (size_t)y * w;
    const uint8_t* uvr = uvp + (size_t)(y >> 1) * w;
    uint8_t* o = bgr + (size_t)y * w * 3;
    for (int x = 0; x < w; ++x) {
      const int u = uvr[x & ~1] - 128, v = uvr[(x & ~1) + 1] - 128;
      const int ruv = HALF + CVR * v, guv = HALF + CVG * v + CUG * u, buv = HALF + CUB * u;
      const int yy = (yr[x] > 16 ? yr[x] - 16 : 0) * CY;
      o[3 * x + 0] = sat8((yy + buv) >> SHIFT);
      o[3 * x + 1] = sat8((yy + guv) >> SHIFT);
      o[3 * x + 2] = sat8((yy + ruv) >> SHIFT);
    }
  }
  return SNB_OK;
}

// stereonet_node.cpp:775-782: cvtColor(NV12 -> BGR) + imencode(".jpg") of one view.  Returns the JPEG size in bytes
// (also when dst == NULL or cap is too small: nothing is written past cap; call again with a large enough buffer),
// or a negative snb_status.  quality <= 0 selects OpenCV's default, 95.
int64_t snb_jpeg_encode_nv12(const uint8_t* nv12, int32_t w, int32_t h, int32_t quality, uint8_t* dst, uint64_t cap) {
  if (!nv12 || w <= 0 || h <= 0 || (w & 1) || (h & 1) || w > 65535 || h > 65535) return SNB_ERR_INVALID;
  Encoder e;
  e.init(quality <= 0 ? 95 : quality);
  e.bw.dst = dst; e.bw.cap = dst ? cap : 0;
  e.headers(w, h);
  const int mw = (w + 15) / 16, mh = (h + 15) / 16, wp = mw * 16;
  // one MCU row (16 image rows, edges replicated) at a time: NV12 -> BGR (OpenCV's formula, as snb_pre_nv12_to_bgr) ->
  // libjpeg's rgb_ycc_convert (16-bit fixed point) -> Y at full resolution, Cb / Cr as 2x2 sums
  std::vector<uint8_t> Y((size_t)16 * wp);
  std::vector<int> Cb((size_t)8 * mw * 8), Cr((size_t)8 * mw * 8);
  const int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SHIFT = 20, HALF = 1 << (SHIFT - 1);
  const uint8_t* uvp = nv12 + (size_t)w * h;
  for (int my = 0; my < mh; ++my) {
    std::fill(Cb.begin(), Cb.end(), 0); std::fill(Cr.begin(), Cr.end(), 0);
    for (int yy = 0; yy < 16; ++yy) {
      const int sy = my * 16 + yy < h ? my * 16 + yy : h - 1;
      const uint8_t* yr = nv12 + (size_t)sy * w;
      const uint8_t* uvr = uvp + (size_t)(sy >> 1) * w;
      uint8_t* yo = &Y[(size_t)yy * wp];
      int* cbo = &Cb[(size_t)(yy >> 1) * mw * 8];
      int* cro = &Cr[(size_t)(yy >> 1) * mw * 8];
      for (int xx = 0; xx < wp; ++xx) {
        const int sx = xx < w ? xx : w - 1;
        const int u = uvr[sx & ~1] - 128, v = uvr[(sx & ~1) + 1] - 128;
        const int yv = (yr[sx] > 16 ? yr[sx] - 16 : 0) * CY;
        const int b = sat8((yv + HALF + CUB * u) >> SHIFT), g = sat8((yv + HALF + CVG * v + CUG * u) >> SHIFT), r = sat8((yv + HALF + CVR * v) >> SHIFT);
        yo[xx] = (uint8_t)((19595 * r + 38470 * g + 7471 * b + 32768) >> 16);
        cbo[xx >> 1] += (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
        cro[xx >> 1] += (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
      }
    }
    for (int mx = 0; mx < mw; ++mx) {
      float blk[64];
      for (int by = 0; by < 2; ++by)
        for (int bx = 0; bx < 2; ++bx) {
          for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) blk[i * 8 + j] = (float)Y[(size_t)(by * 8 + i) * wp + mx * 16 + bx * 8 + j] - 128.f;
          e.block(blk, 0);
        }
      for (int c = 0; c < 2; ++c) {
        const std::vector<int>& src = c ? Cr : Cb;
        for (int i = 0; i < 8; ++i)
          for (int j = 0; j < 8; ++j)         // h2v2_downsample: 2x2 mean with the alternating 1,2 rounding bias
            blk[i * 8 + j] = (float)((src[(size_t)i * mw * 8 + mx * 8 + j] + 1 + (j & 1)) >> 2) - 128.f;
        e.block(blk, 1 + c);
      }
    }
  }
  e.bw.flush();
  e.bw.u16(0xffd9);
  return (int64_t)e.bw.n;
}

}  // extern "C"
