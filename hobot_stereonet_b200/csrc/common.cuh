// Shared definitions for the sm_100a StereoNet kernels.
//
// Activation layout in HBM ("C8"): [n][cb][d][h][w][8] — channels blocked by 8 innermost, so one
// pixel's 8-channel block is 32 B (fp32) and a row of pixels is one contiguous run.  d = 1 for 2-D
// stages; the cost volume and the 3-D aggregation use d = D.  Disparity maps and the cost tensor
// are plain fp32 planes [n][(d)][h][w].
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

namespace snb {

// Two storage types share the C8 indexing:
//   planes == 1 : fp32                                  [n][cb][d][h][w][8]      (SNB_PREC_FP32)
//   planes == 2 : split fp16, x = hi + lo, two planes   [n][2][cb][d][h][w][8]   (SNB_PREC_TC_F16X2)
// Both cost 4 bytes per element.
struct Tens {            // C8 activation tensor
  void* p = nullptr;
  int n = 0, cb = 0, d = 1, h = 0, w = 0;
  int c = 0;             // logical channels (<= cb*8)
  int planes = 1;
  float* f() const { return static_cast<float*>(p); }
  __half* hp() const { return static_cast<__half*>(p); }
  size_t plane_elems() const { return (size_t)cb * d * h * w * 8; }       // one (n[,hl]) slab
  size_t sample_stride() const { return plane_elems() * planes; }          // elements between samples
  size_t lo_off() const { return planes == 2 ? plane_elems() : 0; }
  size_t elems() const { return (size_t)n * cb * d * h * w * 8; }
  size_t bytes() const { return elems() * 4; }
};

struct Plane {           // dense fp32 [n][d][h][w]
  float* p = nullptr;
  int n = 0, d = 1, h = 0, w = 0;
  size_t elems() const { return (size_t)n * d * h * w; }
  size_t bytes() const { return elems() * sizeof(float); }
};

struct ConvParams {
  const void* in; void* out; const float* w; const float* bias;
  const void* res;       // residual added before the activation (same layout as out) or nullptr
  int N, CBin, Din, Hin, Win;
  int CBout, Dout, Hout, Wout;
  int ks;                // spatial kernel size 1 or 3
  int kz;                // depth taps 1 (2-D) or 3 (3-D, pad 1)
  int stride, dil, relu;
  int tiles_x;
  size_t in_ss, in_lo, out_ss, out_lo;   // sample strides / hi->lo plane offsets (elements), see Tens
  int half;              // 0: fp32 storage, 1: split-fp16 storage
};

struct ConvTo1Params {   // Cout = 1 convolutions (conv3d_alone, refinement conv_out)
  const void* in; float* out; const float* w; float bias;
  const void* res; int res_c8;        // residual: fp32 plane (res_c8 = 0) or channel 0 of a C8 tensor of in's storage type
  int N, CBin, D, H, W;
  int kz, dil, relu;
  size_t in_ss, in_lo;
  size_t res_ss, res_lo;              // residual C8 tensor: sample stride / hi->lo offset (elements)
  int half;
};

struct TcConvParams {    // k_conv_tc.cu
  const __half* w; const float* bias; const __half* res; __half* out;
  int N, D, H, W, CBin, CBout;
  int dil, kz, relu;
  int nky;               // kernel rows per pipeline stage: 3 (full halo tile) or 1 (row group, large dilation)
  int nk16;              // Cin / 16
  int NT, R;             // output channels / image rows per tile
  int BW, BH;            // TMA box (pixels)
  int tiles_x, tiles_y, ccs, total_tiles;
  int nstages;
  uint32_t a_bytes, w_bytes, tx_bytes;
};
struct TcConvPlan { CUtensorMap tm_in; TcConvParams p; size_t smem; };

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// true the first time kernel slot `k` is launched on the current device (opt-in smem attribute)
inline bool need_attr(int k) {
  static bool done[32][32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (done[dev & 31][k]) return false;
  done[dev & 31][k] = true;
  return true;
}

}  // namespace snb

#define SNB_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      snprintf(g_err, sizeof(g_err), "%s:%d %s: %s", __FILE__, __LINE__, #expr,       \
               cudaGetErrorString(_e));                                               \
      return SNB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)
