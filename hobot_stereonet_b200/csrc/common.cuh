// Shared definitions for the sm_100a StereoNet kernels.
//
// Activation layout in HBM ("C8"): [n][cb][d][h][w][8] — channels blocked by 8 innermost, so one
// pixel's 8-channel block is 32 B (fp32) and a row of pixels is one contiguous run.  d = 1 for 2-D
// stages; the cost volume and the 3-D aggregation use d = D.  Disparity maps and the cost tensor
// are plain fp32 planes [n][(d)][h][w].
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace snb {

// Two storage types share the C8 indexing:
//   planes == 1 : fp32                                  [n][cb][d][hs][ws][8]      (SNB_PREC_FP32)
//   planes == 2 : split fp16, x = hi + lo, two planes   [n][2][cb][d][hs][ws][8]   (SNB_PREC_TC_F16X2)
// Both cost 4 bytes per element.  Every (cb, d) slice is stored with a zero border of `pad` pixels on all
// four sides (hs = h + 2*pad, ws = w + 2*pad): kernels write interior pixels only, so a convolution's halo
// tile is always a run of whole, in-bounds rows (k_conv_tc.cu loads it with plain bulk copies, the zero
// padding of the convolution is simply there).  Allocations carry TAIL_ROWS rows of slack for the rows a
// bottom/right tile reads past its slice.
constexpr int TAIL_ROWS = 18;
struct Tens {            // C8 activation tensor
  void* p = nullptr;
  int n = 0, cb = 0, d = 1, h = 0, w = 0;
  int c = 0;             // logical channels (<= cb*8)
  int planes = 1;
  int pad = 0;
  int pcb = 0;           // channel blocks of the parent allocation when this is a channel sub-view (0: owns its memory)
  int hs() const { return h + 2 * pad; }
  int ws() const { return w + 2 * pad; }
  float* f() const { return static_cast<float*>(p); }
  __half* hp() const { return static_cast<__half*>(p); }
  size_t esize() const { return planes == 2 ? 2 : 4; }
  size_t slice() const { return (size_t)hs() * ws() * 8; }                 // one (cb, d) slice, elements
  size_t org() const { return ((size_t)pad * ws() + pad) * 8; }            // interior pixel (0,0) inside a slice
  size_t plane_elems() const { return (size_t)(pcb ? pcb : cb) * d * slice(); }   // one (n[,hl]) slab
  size_t sample_stride() const { return plane_elems() * planes; }          // elements between samples
  size_t lo_off() const { return planes == 2 ? plane_elems() : 0; }
  size_t elems() const { return (size_t)n * cb * d * slice(); }
  size_t bytes() const { return elems() * 4 + (size_t)TAIL_ROWS * ws() * 8 * 4; }
};

// Device view of a C8 tensor: `p` points at interior pixel (0,0) of sample 0, block 0, slice 0, so
// element(n, cb, d, y, x, e) = p[n*ss + (cb*D + d)*slice + (y*ws + x)*8 + e]  (+ lo for the lo plane).
struct TV {
  void* p = nullptr;
  size_t ss = 0, lo = 0, slice = 0;
  int ws = 0;
};
// channel blocks [cb0, cb0 + ncb) of t as a tensor of its own (same strides, no copy)
inline Tens sub_view(const Tens& t, int cb0, int ncb) {
  Tens s = t;
  s.p = static_cast<char*>(t.p) + (size_t)cb0 * t.d * t.slice() * t.esize();
  s.pcb = t.pcb ? t.pcb : t.cb;
  s.cb = ncb; s.c = ncb * 8;
  return s;
}
inline TV view(const Tens& t) {
  TV v;
  v.p = static_cast<char*>(t.p) + t.org() * t.esize();
  v.ss = t.sample_stride(); v.lo = t.lo_off(); v.slice = t.slice(); v.ws = t.ws();
  return v;
}

struct Plane {           // dense fp32 [n][d][h][w]
  float* p = nullptr;
  int n = 0, d = 1, h = 0, w = 0;
  size_t elems() const { return (size_t)n * d * h * w; }
  size_t bytes() const { return elems() * sizeof(float); }
};

// The device pointers of ONE call that kernels inside the captured pass read or write: the model input tensor (s8 NCHW
// [B,6,H,W]), the model output tensor (s32 NCHW [B,1,H,W]) and, for the camera-frame entry points, the raw NV12 frames.
// Every kernel that uses them carries an IoPtrs as the FIRST member of its single by-value parameter, so a cached CUDA
// graph can be re-pointed at another call's buffers with cudaGraphExecKernelNodeSetParams (net.cu run_plan).
struct IoPtrs {
  const int8_t* s8 = nullptr;
  int32_t* q = nullptr;
  const uint8_t* frames = nullptr;
};

struct PreNv12Params {   // k_pre_nv12: camera frames (io.frames) -> the s8 model input (io.s8) and, on the older pipeline, the C8 image
  IoPtrs io;
  TV img;                // p == nullptr: no image tensor
  int B, H, W, Hp, Wp, correct;
};

struct ConvParams {
  TV in, out, res;       // res: residual added before the activation (same geometry as out) or p == nullptr
  const float* w; const float* bias;
  int N, CBin, Din, Hin, Win;
  int CBout, Dout, Hout, Wout;
  int ks;                // spatial kernel size 1 or 3
  int kz;                // depth taps 1 (2-D) or 3 (3-D, pad 1)
  int stride, dil, relu;
  int tiles_x;
  int half;              // 0: fp32 storage, 1: split-fp16 storage
};

struct ConvTo1Params {   // Cout = 1 convolutions (conv3d_alone, refinement conv_out)
  TV in;
  float* out; const float* w; float bias;
  TV res; int res_c8;    // residual: fp32 plane (res_c8 = 0, res.p = plane) or channel 0 of a C8 tensor of in's storage type
  int N, CBin, D, H, W;
  int kz, dil, relu;
  int half;
};

struct ConvFirstParams { // k_conv_first: Cin = 3 stride 2 (firstconv.0) or Cin = 4 stride 1 (refinement conv_in), 3x3, Cout = 32
  TV in, out;
  int Ho, Wo, relu, cin, stride;
  float w[36 * 32];      // [(ci*9 + ky*3 + kx)][co]: rides in the kernel parameters = the constant bank
  float b[32];
};

struct ConvFirstS8Params {   // k_conv_first_s8: firstconv.0 read straight from the s8 model input
  IoPtrs io;
  TV out;
  int B, H, W, Ho, Wo, relu;
  float w[27 * 32];      // [(ci*9 + ky*3 + kx)][co]
  float b[32];
};

struct Conv1x1Params {   // k_conv1x1: stand-alone 1x1 convolution, split-fp16 C8 in and out, stride 1
  TV in, out;
  const float* wgt;      // [cin (padded to 8-blocks)][cout] fp32
  const float* bias;
  int h, w, cbin, relu;
};

struct RefineHeadParams {    // k_refine_head: (soft-argmin +) x2 bilinear + left image + conv_in, one launch per refinement stage
  IoPtrs io;
  const float* src;      // stage 0: cost [B][D][h][w]; later stages: the previous stage's disparity [B][h][w]
  float* disp0;          // stage 0: soft-argmin result [B][h][w]
  TV out;                // [B][4 blocks][2h][2w] split-fp16 feature
  int B, D, h, w;        // coarse map
  int H, W;              // valid size of the s8 input views
  int f;                 // image decimation at this stage: padded image height / (2h)
  int stage0;
  float invD;
  float wgt[36 * 32];    // conv_in [(ci*9 + ky*3 + kx)][co], ci 0 = disparity, 1..3 = image
  float b[32];
};

struct TcConvParams {    // k_conv_tc.cu
  TV in, out, res;       // split-fp16 tensors
  const __half* w; const float* bias;
  int N, D, H, W, CBin, CBout;
  int dil, kz, relu;
  int nk16;              // Cin / 16
  int cin8;              // 1: Cin <= 8 (one channel block): K = 16 is two adjacent taps x 8 channels (A LBO = one pixel)
  int R;                 // output rows per tile (template instance)
  int BW, BH;            // shared-memory tile: pixels per row, input rows per stage (R + 2)
  int in_pad;            // zero border of the input tensor (rows below H + in_pad are never read)
  int tiles_x, tiles_y, ccs, total_tiles;
  int nstages;
  uint32_t a_chunk_bytes, w_bytes, stage_bytes;
  long long* prof;       // SNB_TC_PROF=1: per-CTA cycle counters of each pipeline role (16 per CTA)
  int dbg;               // timing experiments only (env SNB_TC_DEBUG): 1 no loads, 2 no MMAs, 4 no global stores
  float rzk;             // rz_unit(): round-toward-zero compensation per MMA of an accumulator
  float wsc;             // 2^-k: the packed weights are w * 2^k (weight_scale_log2)
};
struct TcConvPlan { TcConvParams p; size_t smem; };

struct RbParams {        // k_resblock_tc.cu: fused 32-channel residual block
  TV in, out, res;       // split-fp16 tensors, 4 channel blocks, D = 1
  const __half* wa; const __half* wb;     // tc_pack_weights layout of conv_a / conv_b
  const float* ba; const float* bb;
  int N, H, W, dil, in_pad;
  int OW, XW;            // output pixels per strip (128 - 2 dil), staged pixels per row (128 + 2 dil)
  int strips, nchunk, rpc, total_units;   // units = (sample, strip, comb, row chunk of rpc comb rows)
  int nxs;               // depth of the x-row ring
  uint32_t sub_bytes, slot_bytes;         // one (plane, chunk) row, one ring slot (8 of them)
  long long* prof;       // SNB_TC_PROF=1: per-CTA cycle counters of the issuer and epilogue roles
  int dbg;               // measurement only (env SNB_RB_DEBUG, full-resolution launches): 16 = drop the A_lo x W_hi products
  float rzk;             // rz_unit()
  float wsa, wsb;        // 2^-k of conv_a / conv_b (weight_scale_log2)
};
struct RbPlan { RbParams p; size_t smem; int num_sms; };

struct CsParams {        // k_conv_stream.cu: streaming convolution, one 32-channel (or single-channel) output slice per unit
  IoPtrs io;             // io.q != nullptr (the last conv_out): the s32 model output is written by this kernel's epilogue
  int qH, qW; float qmul;    // valid (un-padded) size of the s32 output, q = rint(disp * qmul)
  TV in, out, res;       // split-fp16 C8 tensors (out/res unused for plane output)
  const __half* w; const float* bias;
  float* out_plane; const float* res_plane;          // cout == 1: fp32 [n][D][H][W]
  int N, D, H, W, dil, kz, nk16, in_pad;
  int ncb_out, nbias;    // valid output channel blocks / biases of a slice (4 / 32 unless Cout < 32)
  int nco, ccs;          // output channels per unit (32, or 16 with one real channel), number of slices
  int XW, strips, nchunk, rpc, total_units, nxs;
  int split;             // 1: main | corr accumulators per job (k_conv_stream SPLIT)
  int nslots, tmem_cols; // TMEM accumulator slots / allocated columns (5 / 512, or 2 / 256 in the two-CTAs-per-SM configuration)
  int ostride;           // 1, or 2: computed at stride 1, rows/columns with an odd index are not stored (C8 output only)
  int relu, res_mode;    // res_mode 0: C8 tensor like out (or none), 1: channel 0 of a C8 tensor, 2: fp32 plane, 3: x2 bilinear of the fp32 plane [n][H/2][W/2]
  uint32_t sub_bytes, slot_bytes, w_bytes;
  long long* prof;       // SNB_TC_PROF=1: per-CTA cycle counters of the three roles
  float rzk;             // rz_unit()
  float wsc;             // 2^-k: the packed weights are w * 2^k (weight_scale_log2)
  int taps;              // kernel columns / rows that carry weights: 3, or 1 for a 1x1 convolution riding as the centre tap (the
                         // other MMAs add exact zeros to the accumulator and cost no rounding: common.cuh rz_comp)
  // second head (ccs2 > 0): another convolution of the SAME input with the same geometry - a BasicBlock's 1x1 shortcut next
  // to its conv_a - served by the same launch as the output slices cc >= ccs - ccs2 (same rows staged once, one launch less)
  int ccs2; const __half* w2; const float* bias2; TV out2; int relu2, taps2, ncb_out2, nbias2; float wsc2;
};
struct CsPlan { CsParams p; size_t smem; int num_sms; };

struct Cost3dParams {    // k_cost3d.cu: Conv3d 32 -> 1 (conv3d_alone) with every input row read once
  TV in;                 // split-fp16 C8 [n][4][D][h][w]
  const __half* w; float* out;      // out: fp32 plane [n][D][H][W]
  float bias, wsc, rzk;
  int N, D, H, W, strips;
  int rpb, nband, dps, nseg, total_units, nxs;    // rows per band, depths per segment
  uint32_t sub_bytes, slot_bytes, w_bytes, p_bytes;
};
struct Cost3dPlan { Cost3dParams p; int num_sms; };

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- compensation of the tensor core's round-toward-zero accumulation ------------------------------------------------
// Measured on B200 (tools/ubench/mma_round_probe.py; the model below reproduces 3440 of 3440 single-MMA results and 768 of
// 768 chains bit for bit): one tcgen05.mma kind::f16 computes acc' = RZ_fp32(sum_i RZ_u(x_i)) over x = {acc, 16 exact
// products}, u = ulp(largest addend) / 4, RZ = round TOWARD ZERO.  IEEE fp32 (the reference float model) rounds to nearest,
// with zero-mean error; the tensor core loses on average a fixed fraction of an ulp of the accumulator PER MMA, always
// toward zero - a systematic bias that adds up coherently through ~70 layers of mostly non-negative activations
// (profiles/r02a_stage_*_d192.txt: 2.9e-3 px of the 2.9e-3 px end-point error at max_disp = 1536 are bias).
// tools/tc_accum_model.py replays the exact hardware model on real layer data: the expected loss of an accumulator that
// took n MMAs is kappa * n * ulp(acc), kappa = 0.22 +- 0.03 for every accumulator structure the kernels use (3 to 24
// MMAs, main-only or merged hi/lo chains) - and the same expression with the ulp of the FINISHED output value (the sum of
// its three kernel-row accumulators, or of its 36 three-MMA chains in k_conv_tc) fits the total loss just as well.  So the
// epilogues add the expectation back ONCE per finished value, before bias, residual and activation:
//     f + copysign(kappa * n * ulp(f), f)          n = MMAs per TMEM accumulator chain (two instructions: AND, FMA)
// which removes the mean of the truncation loss (model: the per-layer bias drops 10-50x; measured on B200 at max_disp
// 1536: end-point error 2.9e-3 -> 5e-4 px, fp32 CUDA-core path 5e-4) and leaves its zero-mean part.  kappa = 0.205 is the
// value that zeroes the measured backbone bias (profiles/r02_stage_*: layer2 -8.6e-7 -> -4e-8, layer4 -1.4e-6 -> -2e-8).
constexpr float RZ_KAPPA_PER_MMA = 0.205f;
// The single-output-channel convolutions (conv3d_alone -> the cost tensor, conv_out) sum 288-864 products of mixed sign into ONE
// value: their partial sums run far above the final result, every MMA truncates at the partial sum's ulp, and the loss relative
// to the ulp of the finished value is larger.  tools/tc_accum_model.py on the same layers: 0.36 (conv3d_alone) / 0.26 (conv_out)
// per MMA; measured on B200, 0.30 takes the cost tensor's bias from -2.2e-6 to the 1e-7 class.
constexpr float RZ_KAPPA_1CH_PER_MMA = 0.30f;

// ---- power-of-two weight scaling ---------------------------------------------------------------------------------------
// Split fp16 keeps w = hi + lo with lo = fp16(w - hi) ~ 2^-12 w.  Convolution weights are small (He-normal: ~0.05), so lo
// falls into fp16's SUBNORMAL range (< 6.1e-5, spacing 6e-8) and keeps only a few bits: the weight is then represented to
// ~6e-7 relative instead of 2^-22, a FIXED perturbation that shows up as a per-channel offset (about one fp32 ulp of the
// output; several for the single-channel layers whose weights are ~1e-3).  Every tcgen05 packing therefore stores
// w * 2^k, k chosen per convolution so that max|w| * 2^k lies in [2^13, 2^14) (fp16 max is 65504): hi and lo are both
// normal numbers, the products and accumulators simply live 2^k higher (exact), and the epilogues multiply by 2^-k
// when a finished value leaves the accumulator domain (biases are pre-multiplied by 2^k: exact).
inline int weight_scale_log2(const float* w, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) { const float a = w[i] < 0 ? -w[i] : w[i]; if (a > mx) mx = a; }
  if (!(mx > 0.f) || mx > 1e30f) return 0;
  int e = 0;
  frexpf(mx, &e);                       // mx = m * 2^e, m in [0.5, 1)  ->  floor(log2 mx) = e - 1
  int k = 13 - (e - 1);
  return k < -8 ? -8 : (k > 30 ? 30 : k);
}
// kappa * 2^-23 (times SNB_RZ_COMP, default 1; 0 switches the compensation off): kernels multiply by their MMA count
inline float rz_unit() {
  static const float k = RZ_KAPPA_PER_MMA * 1.1920928955078125e-07f * (getenv("SNB_RZ_COMP") ? (float)atof(getenv("SNB_RZ_COMP")) : 1.f);
  return k;
}
#ifdef __CUDACC__
// v: a drained accumulator, kn = rz_unit() * (MMAs that accumulated into it)
__device__ __forceinline__ float rz_comp(float v, float kn) {
  return fmaf(__int_as_float(__float_as_int(v) & (int)0xff800000), kn, v);
}
#endif

// Programmatic dependent launch: every kernel of the pass is launched with the stream-serialization attribute and
// calls pdl_trigger() + pdl_wait() before it touches global memory written by its predecessors, so the next
// kernel's launch latency and prologue (barrier init, TMEM allocation, weight staging) overlap the current kernel's
// tail.  Captured into the CUDA graph as programmatic edges.  On by default (SNB_PDL=0 turns it off); it pays where
// two CTAs of consecutive kernels fit on one SM (k_conv_stream's small configuration).
inline bool pdl_enabled() {
  static const bool on = !getenv("SNB_PDL") || atoi(getenv("SNB_PDL"));
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

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
