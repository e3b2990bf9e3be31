/*
 * snb200.h — C ABI of the B200-native StereoNet inference path.
 *
 * Drop-in boundary for the ONE span the reference runs as an opaque BPU call:
 *   PreProcess::CvtNV12Data2Tensors output (s8 NCHW [1,6,H,W])
 *     -> hobot::dnn_node::DnnNode::Run(...)            stereonet_infer/src/stereonet_node.cpp:812
 *     -> s32 NCHW [1,1,H,W] read in PostProcess         stereonet_infer/src/stereonet_node.cpp:1033-1034
 * plus the host byte formats either side of it.  Plain pointers and sizes only; no
 * torch / ROS / OpenCV types.  Every function returns 0 on success and a negative
 * snb_status on failure (the reference's convention: Init()!=0, Run()<0, preprocess -1;
 * stereonet_node.cpp:44-45,812; preprocess.cpp:919-922).  There is no CPU fallback: every
 * compute entry point fails with SNB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SNB200_H_
#define SNB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_API __attribute__((visibility("default")))

typedef enum snb_status {
  SNB_OK = 0,
  SNB_ERR_INVALID = -1,   /* bad argument / shape mismatch (reference: frame dropped) */
  SNB_ERR_MODEL = -2,     /* model_file missing or not a weight blob (SetNodePara, stereonet_node.cpp:131-134) */
  SNB_ERR_CUDA = -3,      /* CUDA runtime / no device */
  SNB_ERR_NOMEM = -4,
  SNB_ERR_BUSY = -5       /* async queue full and timeout expired */
} snb_status;

/* hbDNNTensorProperties.tensorLayout / tensorType values the node logs (stereonet_node.cpp:64-103) */
enum { SNB_LAYOUT_NHWC = 0, SNB_LAYOUT_NCHW = 2 };
enum { SNB_TENSOR_S8 = 1, SNB_TENSOR_S32 = 3 };

/* Arithmetic of the dense layers.  Both keep disparity in fp32 end to end. */
enum {
  SNB_PREC_FP32 = 0,      /* CUDA-core fp32 direct convolution (exact path, parity anchor) */
  SNB_PREC_TC_F16X2 = 1   /* tcgen05 tensor cores, split-fp16 (hi+lo) operands, fp32 accumulate in TMEM */
};

enum {
  SNB_FLAG_KEEP_STAGES = 1,   /* no scratch reuse: every stage tensor stays readable via snb_debug_read */
  SNB_FLAG_NO_GRAPH = 2,      /* launch kernels one by one instead of replaying the captured CUDA graph */
  SNB_FLAG_CORRECT_CHROMA = 4, /* snb_infer_nv12: de-interleave NV12 chroma properly (NOT the reference behaviour) */
  SNB_FLAG_NO_TENSOR = 8,     /* diagnostics: SNB_PREC_TC_F16X2 storage, but every convolution on the CUDA-core kernel */
  SNB_FLAG_NO_FUSE = 16,      /* diagnostics: residual blocks as two separate convolution launches */
  SNB_FLAG_NO_STREAM = 32,    /* diagnostics: tiled k_conv_tc / CUDA-core kernels instead of the streaming convolution */
  SNB_FLAG_NO_HEADFUSE = 64,  /* diagnostics: the older glue - pre-process kernel -> C8 image tensor, soft-argmin / refine_in / conv_in / post-quantise as separate launches - instead of the fused refinement heads reading the s8 input directly */
  SNB_FLAG_NO_HBMCONV = 128,  /* diagnostics: firstconv.0 on the tcgen05 streaming kernel instead of k_conv_first (k_conv_hbm.cu) */
  SNB_FLAG_NO_COALESCE = 256, /* snb_infer_async: one pass per call even when max_batch > 1 (default: queued calls are merged into passes of up to max_batch pairs) */
  SNB_FLAG_DEFER_WEIGHTS = 1024  /* snb_create without a model: the weights arrive through snb_set_weights (multi-GPU init, snb_pool_create); until then every infer call returns SNB_ERR_MODEL */
};

/* Replaces dnn_node_para_ptr_->{model_file, model_task_type, task_num} (stereonet_node.cpp:136-144)
 * and the compiled-in model geometry (preprocess.h:220-221 model_in_w_/h_ = 1280/720). */
typedef struct snb_config {
  int32_t struct_size;      /* = sizeof(snb_config) */
  int32_t height, width;    /* valid model input H x W (one view) */
  int32_t K;                /* x2 refinement stages == log2(cost-volume stride), 2..4; deployed model: 4 */
  int32_t D;                /* disparity hypotheses at cost-volume resolution; deployed model: 12 */
  int32_t max_batch;        /* stereo pairs processed per pass; larger calls are chunked */
  int32_t device;           /* CUDA ordinal */
  int32_t task_num;         /* async tasks in flight (reference: 4, stereonet_node.cpp:144) */
  int32_t precision;        /* SNB_PREC_* */
  int32_t flags;            /* SNB_FLAG_* */
  const char* model_file;   /* weight blob path ("SNB2WGT1"); may be NULL when `weights` is given */
  const void* weights;      /* in-memory weight blob (host memory), or NULL */
  uint64_t weights_bytes;
} snb_config;

/* The fields of hbDNNTensorProperties / DNNTensor the reference reads
 * (preprocess.cpp:924-945; stereonet_node.cpp:64-103; parser.cpp:58,169-188). */
typedef struct snb_tensor_props {
  int32_t valid_shape[4];
  int32_t aligned_shape[4];
  int32_t tensor_layout;    /* SNB_LAYOUT_NCHW */
  int32_t tensor_type;      /* SNB_TENSOR_S8 (input) / SNB_TENSOR_S32 (output) */
  int32_t scale_len;        /* 1 */
  float scale;              /* input 1/128; output 2.60443857769133e-06 (hbm@0x27be8) */
  uint64_t mem_size;        /* bytes of one tensor (sysMem[0].memSize) */
} snb_tensor_props;

/* rt_stat of the dnn_node runtime (stereonet_node.cpp:1071-1086) plus device timing. */
typedef struct snb_rt_stat {
  float input_fps, output_fps;
  int32_t infer_time_ms;
  int32_t fps_updated;
  float gpu_ms;             /* device time of the pass (CUDA events) */
  float h2d_ms, d2h_ms;
  int32_t kernel_launches;  /* kernels of this library launched for the call */
} snb_rt_stat;

typedef struct snb_ctx snb_ctx;
typedef void (*snb_done_fn)(void* user, int status, const snb_rt_stat* stat);

/* ---- model lifetime: DnnNode::Init() / GetModel() / ~DnnNode ------------------------------------ */
SNB_API int snb_create(snb_ctx** out, const snb_config* cfg);
SNB_API void snb_destroy(snb_ctx* ctx);
/* Replace the weights of a live context (multi-GPU init: rank 0 loads, NCCL-broadcasts the blob,
 * every rank installs it).  `blob` may be a device pointer when is_device != 0. */
SNB_API int snb_set_weights(snb_ctx* ctx, const void* blob, uint64_t bytes, int is_device);

/* Seeded synthetic weight blob for refinement-stage count K (the reference's float weights are not
 * recoverable from its BPU binary).  dst == NULL: returns the size needed.  Host-only, no GPU. */
SNB_API int64_t snb_weights_synthesize(int32_t K, uint64_t seed, void* dst, uint64_t cap);

/* Host-only validation of a weight blob (no GPU): magic, table bounds, byte counts, and every convolution of the K-stage
 * topology present with its shape - exactly what snb_create / snb_set_weights accept.  K <= 0: use the K in the blob.
 * Returns SNB_OK or SNB_ERR_MODEL (reason: snb_last_error(NULL)). */
SNB_API int snb_weights_validate(const void* blob, uint64_t bytes, int32_t K);

/* ---- hbSysAllocCachedMem / hbSysFreeMem (preprocess.cpp:956-960,972): host buffers for tensors ----- */
/* Page-locked host memory when a CUDA device is present (so snb_infer's copies are true async DMA),
 * plain aligned host memory otherwise.  hbSysFlushMem has no equivalent: nothing to flush. */
SNB_API int snb_sys_alloc(void** ptr, uint64_t bytes);
SNB_API void snb_sys_free(void* ptr);

/* ---- hbDNNGet{Input,Output}TensorProperties / GetModelInputSize (stereonet_node.cpp:45,78,94) --- */
SNB_API int snb_get_io(const snb_ctx* ctx, snb_tensor_props* in, snb_tensor_props* out);
SNB_API int snb_get_model_input_size(const snb_ctx* ctx, int32_t input_index, int32_t* w, int32_t* h);

/* ---- DnnNode::Run (stereonet_node.cpp:812; sync variants :177,584,968) -------------------------- */
/* in: s8 NCHW [batch,6,H,W] host memory; out: s32 NCHW [batch,1,H,W] host memory.
 * value * 2.60443857769133e-06 * 192 = left-view disparity in pixels. */
SNB_API int snb_infer(snb_ctx* ctx, const int8_t* in, int32_t* out, int32_t batch);
/* is_sync_mode=false: returns after enqueueing; `done` fires on a library-owned thread (the
 * reference's PostProcess thread).  Blocks up to timeout_ms (-1: forever) for a free task slot.
 * On a context with max_batch > 1 the library may serve several queued calls with ONE pass over the network (up to
 * max_batch pairs; SNB_FLAG_NO_COALESCE turns that off).  Nothing changes per call: own buffers, own callback,
 * callbacks in submission order, results bit-identical to snb_infer. */
SNB_API int snb_infer_async(snb_ctx* ctx, const int8_t* in, int32_t* out, int32_t batch,
                            snb_done_fn done, void* user, int32_t timeout_ms);
SNB_API int snb_wait_all(snb_ctx* ctx);
/* Device-resident variant: both pointers are device memory; runs on `cuda_stream` (a cudaStream_t,
 * NULL = the context's own stream) and does not synchronise when a stream is supplied. */
SNB_API int snb_infer_device(snb_ctx* ctx, const int8_t* d_in, int32_t* d_out, int32_t batch,
                             void* cuda_stream);
/* Raw camera frames in (stereonet_node.cpp:657-738 + preprocess.cpp:913-1059 done on the GPU):
 * frames: batch x side-by-side NV12 [H*3/2, 2W] host memory. */
SNB_API int snb_infer_nv12(snb_ctx* ctx, const uint8_t* frames, int32_t* out, int32_t batch);

/* The asynchronous form of snb_infer_nv12: the node uploads the raw camera frame (3 bytes per pixel pair position instead of
 * the 6-plane s8 tensor: half the host->device bytes) and the L/R split, the chroma "444" step and the x-128 of
 * stereonet_node.cpp:702-738 + preprocess.cpp:913-1059 run on the GPU inside the pass.  Same queue, task slots, callback
 * thread and merging of queued calls as snb_infer_async (calls of the two kinds are never merged into one pass). */
SNB_API int snb_infer_nv12_async(snb_ctx* ctx, const uint8_t* frames, int32_t* out, int32_t batch,
                                 snb_done_fn done, void* user, int32_t timeout_ms);
/* PreProcess::CvtNV12Data2Tensors (preprocess.cpp:913-1059) on the GPU, as a call of its own: batch side-by-side NV12
 * frames (host) -> the s8 NCHW [batch,6,H,W] tensor (host), byte-identical to snb_pre_cvt_nv12_to_tensor on the split
 * views (SNB_FLAG_CORRECT_CHROMA selects the de-interleaving variant). */
SNB_API int snb_pre_nv12_gpu(snb_ctx* ctx, const uint8_t* frames, int32_t batch, int8_t* s8_out);

/* Whole-network passes launched so far (each = kernel_launches kernels).  With coalescing a pass can serve several
 * snb_infer_async calls, so passes <= calls. */
SNB_API int64_t snb_get_pass_count(const snb_ctx* ctx);
SNB_API int snb_get_rt_stat(const snb_ctx* ctx, snb_rt_stat* stat);
SNB_API const char* snb_last_error(const snb_ctx* ctx);   /* ctx may be NULL: last create error */
SNB_API const char* snb_version(void);

/* ---- one node process, N GPUs: replicas + batch sharding ------------------------------------------
 * The reference drives ONE BPU (stereonet_node.cpp:44 Init, :812 Run); a B200 box has eight GPUs.  snb_pool_create makes one
 * replica of the model per listed device (devices == NULL: every visible GPU; cfg->device is ignored): replica 0 loads
 * model_file, ONE ncclBroadcast over NVLink puts the blob into every other GPU's memory and each replica installs it from
 * there (snb_set_weights, is_device = 1).  Nothing is exchanged on the per-frame path.
 *   snb_pool_infer_async / _nv12_async: one call = one Run(); goes to the replica with the fewest calls in flight.
 *   snb_pool_infer: a batch in one call; contiguous shards (snb_shard_range), all replicas at once; returns when done. */
typedef struct snb_pool snb_pool;
typedef struct snb_pool_stat {
  int32_t n_devices;
  int32_t device[16];       /* CUDA ordinals */
  int64_t calls[16];        /* calls (or shards) served by each replica */
  uint64_t weight_bytes;    /* size of the broadcast blob */
  float broadcast_ms;       /* device time of the NCCL broadcast (0 for a pool of one) */
} snb_pool_stat;
SNB_API int snb_pool_create(snb_pool** out, const snb_config* cfg, const int32_t* devices, int32_t n_devices);
SNB_API void snb_pool_destroy(snb_pool* pool);
SNB_API int32_t snb_pool_size(const snb_pool* pool);
SNB_API snb_ctx* snb_pool_ctx(const snb_pool* pool, int32_t replica);      /* for snb_get_io, snb_get_rt_stat ... */
SNB_API int snb_pool_infer_async(snb_pool* pool, const int8_t* in, int32_t* out, int32_t batch,
                                 snb_done_fn done, void* user, int32_t timeout_ms);
SNB_API int snb_pool_infer_nv12_async(snb_pool* pool, const uint8_t* frames, int32_t* out, int32_t batch,
                                      snb_done_fn done, void* user, int32_t timeout_ms);
SNB_API int snb_pool_infer(snb_pool* pool, const int8_t* in, int32_t* out, int32_t batch);
SNB_API int snb_pool_wait_all(snb_pool* pool);
SNB_API int snb_pool_get_stat(const snb_pool* pool, snb_pool_stat* stat);
SNB_API const char* snb_pool_last_error(const snb_pool* pool);              /* pool may be NULL: last create error */
/* Contiguous [start, stop) of n_pairs owned by `rank` of `world`, remainder to the lowest ranks (host-only arithmetic). */
SNB_API int snb_shard_range(int64_t n_pairs, int32_t world, int32_t rank, int64_t* start, int64_t* stop);

/* ---- stage taps for parity tests (needs SNB_FLAG_KEEP_STAGES) ----------------------------------- */
/* Copies stage `name` of the last pass to host as dense fp32 in [N,C,(D,)H,W] order.
 * shape[5] receives N,C,D,H,W (D=1 for 2-D stages).  Returns elements written or <0. */
SNB_API int64_t snb_debug_read(snb_ctx* ctx, const char* name, float* dst, uint64_t cap, int32_t shape[5]);
/* Per-kernel device times of the last pass run with SNB_FLAG_NO_GRAPH: fills up to cap entries. */
typedef struct snb_kernel_time { char name[48]; float ms; double flops; double bytes; } snb_kernel_time;
SNB_API int snb_profile_pass(snb_ctx* ctx, int32_t batch, snb_kernel_time* out, int32_t cap);

/* ---- host-side byte formats either side of the model call (single-threaded, as the reference) --- */
/* stereonet_node.cpp:702-738: split one side-by-side NV12 frame [h*3/2, 2w] into left/right [h*3/2, w]. */
SNB_API int snb_pre_split_nv12(const uint8_t* frame, int32_t h, int32_t w2, uint8_t* left, uint8_t* right);
/* preprocess.h:128-155 Tools::YUV420TOYUV444, quirk included (see oracle/prepost_ref.py). */
SNB_API int snb_pre_yuv420_to_yuv444(const uint8_t* in, uint8_t* out, int32_t w, int32_t h, int32_t correct_chroma);
/* preprocess.cpp:913-1059 CvtNV12Data2Tensors: two NV12 views -> s8 NCHW [1,6,h,w]. */
SNB_API int snb_pre_cvt_nv12_to_tensor(const uint8_t* left, const uint8_t* right, int32_t w, int32_t h,
                                       int32_t correct_chroma, int8_t* out);
/* preprocess.cpp:1131-1136 PreProcess::Quantize with preprocess.h:236-240 defaults. */
SNB_API int8_t snb_pre_quantize(float value, float scale, float zero_point, float lo, float hi);
/* stereonet_node.cpp:1033-1049: payload = s32 output || jpeg; returns bytes written or <0. */
SNB_API int64_t snb_post_pack(const int32_t* infer, uint64_t infer_bytes, const uint8_t* jpeg, uint64_t jpeg_bytes,
                              uint8_t* dst, uint64_t cap);
/* stereonet_node.cpp:775-777 cv::cvtColor(nv12, bgr, CV_YUV2BGR_NV12): one NV12 view [h*3/2, w] -> BGR u8 [h, w, 3]
 * (OpenCV's BT.601 fixed-point formula, bit-exact).  w and h even. */
SNB_API int snb_pre_nv12_to_bgr(const uint8_t* nv12, int32_t w, int32_t h, uint8_t* bgr);
/* stereonet_node.cpp:775-782 cvtColor + cv::imencode(".jpg"): baseline JFIF (4:2:0, quality <= 0: OpenCV's default 95)
 * of one NV12 view, the part of the payload the render tool hands to cv2.imdecode (publisher_member_function.py:93-95).
 * Returns the JPEG size; nothing is written past cap (dst may be NULL to size the buffer).  Host-only. */
SNB_API int64_t snb_jpeg_encode_nv12(const uint8_t* nv12, int32_t w, int32_t h, int32_t quality, uint8_t* dst, uint64_t cap);
/* parser.cpp:79-87 ParseTensor: s32 -> depth in metres (float), f = 527.19..., B = 119.89... mm. */
SNB_API int snb_post_parse_depth(const int32_t* q, int64_t n, float scale, float* depth_m);

/* parser.cpp:79-118 on the GPU: s32 model output -> depth in metres and the JET colour map
 * (cv::convertScaleAbs(depth, alpha) + cv::applyColorMap; alpha = 11 in parser.cpp:115, 9 in the render tool).
 * q: [batch,1,H,W] s32; depth_m: [batch,H,W] f32 or NULL; bgr: [batch,H,W,3] u8 or NULL.  is_device != 0: all three
 * pointers are device memory and the call only enqueues on the context's stream + synchronises. */
SNB_API int snb_post_depth_color(snb_ctx* ctx, const int32_t* q, int32_t batch, float alpha, float* depth_m, uint8_t* bgr,
                                 int32_t is_device);

#ifdef __cplusplus
}
#endif
#endif  /* SNB200_H_ */
