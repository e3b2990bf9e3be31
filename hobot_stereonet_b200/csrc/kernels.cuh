// Launchers of the sm_100a kernels (one per SURVEY.md §8a row / Appendix A entry).
#pragma once
#include "common.cuh"

namespace snb {

// k_conv_direct.cu — fp32 CUDA-core convolutions (M1, M3, M5)
cudaError_t launch_conv_direct(ConvParams p, int cout, cudaStream_t st);
cudaError_t launch_conv_to1(const ConvTo1Params& p, cudaStream_t st);

// k_conv_tc.cu — tcgen05 implicit-GEMM convolution on split-fp16 tensors (M1, M3, M5)
cudaError_t tc_conv_plan(TcConvPlan* plan, const Tens& in, const Tens& out, int cin, int cout, int dil, int kz, int num_sms);
cudaError_t launch_conv_tc(const TcConvPlan& plan, int N, const void* w, int wlog2, const float* bias, const Tens* res, int relu,
                           int num_sms, cudaStream_t st);
// packs w * 2^wlog2 (common.cuh weight_scale_log2); the launchers take the same wlog2 and undo it in the epilogue
void tc_pack_weights(const float* W, int cout, int cin, int kz, int NT, int wlog2, std::vector<__half>& out);

// k_resblock_tc.cu — fused residual block (conv_a + ReLU + conv_b + residual + ReLU), 32 channels (M1 layer1, M5)
cudaError_t resblock_tc_plan(RbPlan* plan, const Tens& in, const Tens& out, const Tens& res, int dil, int num_sms);
cudaError_t launch_resblock_tc(const RbPlan& plan, int N, const void* wa, const void* wb, int wlog2a, int wlog2b, const float* ba,
                               const float* bb, cudaStream_t st);

// k_conv_stream.cu — streaming tcgen05 convolution, weights resident in shared memory (M1, M3, M5)
cudaError_t conv_stream_plan(CsPlan* plan, const Tens& in, int cin, int cout, int dil, int kz, int num_sms);
// A second convolution of the same input served by the same launch (a BasicBlock's 1x1 shortcut next to its conv_a): same Cin, same
// stride, C8 output, no residual.  Its weights are packed with cs_pack_weights like any other (1x1: the centre tap of a 3x3).
struct CsHead2 { const void* w; int wlog2; const float* bias; const Tens* out; int cout, ks, relu; };
// res_c8_ch0: 0 = C8 residual tensor `res`, 1 = channel 0 of `res`; res_plane_mode: 2 = fp32 plane like the output, 3 = x2 bilinear
// upsample of the half-resolution fp32 plane.  q (optional, single-channel output only): the s32 model output tensor.
cudaError_t launch_conv_stream(const CsPlan& plan, int N, const void* w, int wlog2, const float* bias, const Tens* out, const Tens* res,
                               float* out_plane, const float* res_plane, int res_c8_ch0, int relu, int ostride, cudaStream_t st,
                               int res_plane_mode = 2, const IoPtrs* io = nullptr, int qH = 0, int qW = 0, float qmul = 0.f,
                               const CsHead2* head2 = nullptr);
// kernel size of the convolution a plan was made for (1: a 1x1 riding as the centre tap) - set by the caller after conv_stream_plan
inline void conv_stream_set_taps(CsPlan* plan, int ks) { plan->p.taps = ks == 1 ? 1 : 3; }
void cs_pack_weights(const float* W, int cout, int cin, int kz, int ks, int NCO, int wlog2, std::vector<__half>& out);

// k_conv_hbm.cu — few-input-channel convolutions of the tensor path (firstconv.0, refinement conv_in): CUDA cores, weights in the constant bank
void conv_first_pack(const float* W, const float* bias, int cin, ConvFirstParams* p);
cudaError_t launch_conv_first(const ConvFirstParams& p, int N, cudaStream_t st);

void conv_first_s8_pack(const float* W, const float* bias, ConvFirstS8Params* p);
cudaError_t launch_conv_first_s8(ConvFirstS8Params p, const IoPtrs& io, int B, cudaStream_t st);
// M4 + head of M5: (soft-argmin over D +) x2 bilinear upsample + left image from the s8 input + conv_in, one launch per stage
void refine_head_pack(const float* W, const float* bias, RefineHeadParams* p);
cudaError_t launch_refine_head(RefineHeadParams p, const IoPtrs& io, int B, cudaStream_t st);

// k_cost3d.cu — head.conv3d_alone (Conv3d 32 -> 1) for large volumes: all nine (depth tap, kernel row) products of an input row in one
// N = 32 MMA, every input row read once (the streaming kernel reads it for each of the three output depths it feeds)
void cost3d_pack_weights(const float* W, int cin, int wlog2, std::vector<__half>& out);
cudaError_t cost3d_plan(Cost3dPlan* plan, const Tens& in, int num_sms);
cudaError_t launch_cost3d(const Cost3dPlan& plan, int N, const void* w, int wlog2, float bias, float* out, cudaStream_t st);

// stand-alone 1x1 convolutions (layer1.0's shortcut, lastconv.1): CUDA cores, one pixel per thread
void conv1x1_pack(const float* W, int cout, int cin, int cbin, std::vector<float>& out);
cudaError_t launch_conv1x1(Conv1x1Params p, int cout, int N, cudaStream_t st);

// k_mem.cu — HBM-bound kernels
// P3 tail on device: s8 NCHW [B,6,H,W] -> C8 [2B][1][Hp][Wp][8] (x/128; left n<B, right n>=B; ch 3..7 = 0)
cudaError_t launch_pre_s8(const int8_t* s8, Tens img, int B, int H, int W, cudaStream_t st);
// P1-P3 on device: side-by-side NV12 frames [B][H*3/2][2W] -> same C8 image tensor, and optionally the s8 tensor
cudaError_t launch_pre_nv12(const uint8_t* frames, Tens img, int8_t* s8_or_null, int B, int H, int W,
                            int correct_chroma, cudaStream_t st);
// M2: gwc [2B][32][h][w][8], cat [2B][2][h][w][8] -> vol [B][8][D][h][w][8]
cudaError_t launch_costvol(Tens gwc, Tens cat, Tens vol, int B, int D, cudaStream_t st);
// M4: cost [B][D][h][w] -> disp [B][h][w] = sum_d softmax(cost)_d * d / D
cudaError_t launch_softargmin(Plane cost, Plane disp, cudaStream_t st);
// M5 glue: x2 bilinear(disp) ++ left image resized to (2h,2w) -> C8 [B][1][2h][2w][8] (ch 4..7 = 0)
cudaError_t launch_refine_in(Plane disp, Tens img_full, Tens out, int B, cudaStream_t st);
// O1 head: normalised disparity [B][Hp][Wp] -> s32 NCHW [B,1,H,W] (crop), q = rint(dn * qmul)
cudaError_t launch_post_quant(Plane disp, int32_t* out, int H, int W, float qmul, cudaStream_t st);

// O2 on the GPU (next-row f3): s32 output -> depth in metres (+ JET colour map as cv::convertScaleAbs / applyColorMap)
cudaError_t launch_post_depth_color(const int32_t* q, float* depth, uint8_t* bgr, size_t n, float scale, float alpha,
                                    cudaStream_t st);

}  // namespace snb
