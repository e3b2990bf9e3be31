// Shared definitions for the sm_100a StereoNet kernels.
//
// Activation layout in HBM ("C8"): [n][cb][d][h][w][8] — channels blocked by 8 innermost, so one
// pixel's 8-channel block is 32 B (fp32) and a row of pixels is one contiguous run.  d = 1 for 2-D
// stages; the cost volume and the 3-D aggregation use d = D.  Disparity maps and the cost tensor
// are plain fp32 planes [n][(d)][h][w].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace snb {

struct Tens {            // C8 activation tensor
  float* p = nullptr;
  int n = 0, cb = 0, d = 1, h = 0, w = 0;
  int c = 0;             // logical channels (<= cb*8)
  size_t elems() const { return (size_t)n * cb * d * h * w * 8; }
  size_t bytes() const { return elems() * sizeof(float); }
};

struct Plane {           // dense fp32 [n][d][h][w]
  float* p = nullptr;
  int n = 0, d = 1, h = 0, w = 0;
  size_t elems() const { return (size_t)n * d * h * w; }
  size_t bytes() const { return elems() * sizeof(float); }
};

struct ConvParams {
  const float* in; float* out; const float* w; const float* bias;
  const float* res;      // residual added before the activation (same layout as out) or nullptr
  int N, CBin, Din, Hin, Win;
  int CBout, Dout, Hout, Wout;
  int ks;                // spatial kernel size 1 or 3
  int kz;                // depth taps 1 (2-D) or 3 (3-D, pad 1)
  int stride, dil, relu;
  int tiles_x;
};

struct ConvTo1Params {   // Cout = 1 convolutions (conv3d_alone, refinement conv_out)
  const float* in; float* out; const float* w; float bias;
  const float* res; int res_stride;   // residual element stride (8 when it is channel 0 of a C8 tensor)
  int N, CBin, D, H, W;
  int kz, dil, relu;
};

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
