// extern "C" entry points declared in include/snb200.h.
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <fstream>

#include "kernels.cuh"
#include "net.h"

using namespace snb;

static thread_local char g_err[512] = {0};

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int fail(snb_ctx* c, int code, const char* msg) {
  if (c) snprintf(c->err, sizeof(c->err), "%s", msg);
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

#define CK(c, expr)                                                                                 \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      snprintf((c)->err, sizeof((c)->err), "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      snprintf(g_err, sizeof(g_err), "%s", (c)->err);                                               \
      return SNB_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

static void worker_main(snb_ctx* c);

// one eager pass over zeros: sets kernel attributes outside graph capture and surfaces launch errors at init
static int warm_up(snb_ctx* c) {
  cudaMemsetAsync(c->d_in, 0, c->in_bytes * c->maxB, c->stream);
  if (!c->direct_io) launch_pre_s8(c->d_in, c->img, c->maxB, c->H, c->W, c->stream);
  IoPtrs io;
  io.s8 = c->d_in; io.q = c->d_out;
  int r = run_plan(c, c->maxB, io, c->stream, false);
  if (r == SNB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "warm-up pass failed: %s", cudaGetErrorString(cudaGetLastError()));
    r = SNB_ERR_CUDA;
  }
  if (r == SNB_OK) c->warmed = true;
  return r;
}

extern "C" {

const char* snb_version(void) { return "snb200 0.1 (sm_100a)"; }

const char* snb_last_error(const snb_ctx* ctx) { return ctx ? ctx->err : g_err; }

// Host-only check of a weight blob against the topology for K refinement stages (K <= 0: the K stored in the blob):
// what snb_create / snb_set_weights would accept.  The message of a rejection is in snb_last_error(NULL).
int snb_weights_validate(const void* blob, uint64_t bytes, int32_t K) {
  std::map<std::string, HostTensor> wts;
  int blob_K = -1;
  int r = parse_blob(blob, bytes, &wts, &blob_K, g_err, sizeof(g_err));
  if (r == SNB_OK && K > 0 && blob_K != K) {
    snprintf(g_err, sizeof(g_err), "weight blob was generated for K=%d, asked K=%d", blob_K, K);
    r = SNB_ERR_MODEL;
  }
  return r;
}

int snb_create(snb_ctx** out, const snb_config* cfg) {
  if (!out || !cfg || cfg->struct_size != (int32_t)sizeof(snb_config)) return fail(nullptr, SNB_ERR_INVALID, "snb_create: bad config struct");
  *out = nullptr;
  if (cfg->height <= 0 || cfg->width <= 0)
    return fail(nullptr, SNB_ERR_INVALID, "snb_create: height and width must be positive");
  // the backbone reduces by 4 (K = 2), 8 (K = 3) or 16 (K = 4) - layer_strides() in layers.h - and the cost volume sits at 1 / 2^K
  if (cfg->K < 2 || cfg->K > 4 || cfg->D < 1 || cfg->D > 512 || cfg->max_batch < 1)
    return fail(nullptr, SNB_ERR_INVALID, "snb_create: need 2<=K<=4, 1<=D<=512, max_batch>=1");
  if (cfg->precision != SNB_PREC_FP32 && cfg->precision != SNB_PREC_TC_F16X2)
    return fail(nullptr, SNB_ERR_INVALID, "snb_create: unknown precision");

  // model file check first, as SetNodePara does (stereonet_node.cpp:131-134)
  const bool defer = (cfg->flags & SNB_FLAG_DEFER_WEIGHTS) != 0;
  std::vector<char> blob;
  if (defer) {
    // multi-GPU init (pool.cu): the weights arrive later from another GPU, through snb_set_weights
  } else if (cfg->weights && cfg->weights_bytes) {
    blob.assign((const char*)cfg->weights, (const char*)cfg->weights + cfg->weights_bytes);
  } else {
    if (!cfg->model_file || access(cfg->model_file, F_OK) != 0) {
      snprintf(g_err, sizeof(g_err), "File is not exist! model_file: %s", cfg->model_file ? cfg->model_file : "(null)");
      return SNB_ERR_MODEL;
    }
    std::ifstream f(cfg->model_file, std::ios::binary);
    blob.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, SNB_ERR_CUDA, "snb_create: no CUDA device (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SNB_ERR_INVALID, "snb_create: bad device ordinal");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  const int num_sms = prop.multiProcessorCount;
  if (prop.major != 10) {
    snprintf(g_err, sizeof(g_err), "snb_create: device %d is sm_%d%d; this build carries sm_100a code only", cfg->device, prop.major, prop.minor);
    return SNB_ERR_CUDA;
  }

  snb_ctx* c = new snb_ctx();
  c->cfg = *cfg;
  if (cfg->model_file) c->model_file = cfg->model_file;
  c->cfg.model_file = nullptr; c->cfg.weights = nullptr;
  c->num_sms = num_sms;
  c->H = cfg->height; c->W = cfg->width; c->K = cfg->K; c->D = cfg->D; c->maxB = cfg->max_batch;
  const int s = 1 << c->K;
  c->Hp = (c->H + s - 1) / s * s; c->Wp = (c->W + s - 1) / s * s;
  c->h = c->Hp / s; c->w = c->Wp / s;
  c->qmul = (float)((double)(s * c->D) / (192.0 * 2.60443857769133e-06));
  c->in_bytes = (size_t)6 * c->H * c->W;
  c->out_bytes = (size_t)4 * c->H * c->W;
  c->frame_bytes = (size_t)c->H * 3 / 2 * 2 * c->W;

  auto bail = [&](int code) { snprintf(g_err, sizeof(g_err), "%s", c->err); free_ctx(c); delete c; return code; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { snprintf(c->err, sizeof(c->err), "cudaSetDevice failed"); return bail(SNB_ERR_CUDA); }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { snprintf(c->err, sizeof(c->err), "stream create failed"); return bail(SNB_ERR_CUDA); }
  for (auto& e : c->ev) cudaEventCreate(&e);
  cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming);
  int r = SNB_OK;
  if (!defer) {
    r = parse_blob(blob.data(), blob.size(), &c->wts, &c->blob_K, c->err, sizeof(c->err));
    if (r != SNB_OK) return bail(r);
    r = upload_weights(c);
    if (r != SNB_OK) return bail(r);
  }
  if (cudaMalloc(&c->d_in, c->in_bytes * c->maxB) != cudaSuccess || cudaMalloc(&c->d_out, c->out_bytes * c->maxB) != cudaSuccess ||
      cudaMalloc(&c->d_frames, c->frame_bytes * c->maxB) != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "cudaMalloc io staging failed");
    return bail(SNB_ERR_NOMEM);
  }
  // asynchronous calls: task_num slots with their own staging buffers + two copy streams
  cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking);
  c->slots.resize(std::max(1, std::min(cfg->task_num, 8)));
  for (auto& sl : c->slots) {
    if (cudaMalloc(&sl.d_in, c->in_bytes * c->maxB) != cudaSuccess || cudaMalloc(&sl.d_out, c->out_bytes * c->maxB) != cudaSuccess ||
        cudaMalloc(&sl.d_frames, c->frame_bytes * c->maxB) != cudaSuccess) {
      snprintf(c->err, sizeof(c->err), "cudaMalloc async staging failed");
      return bail(SNB_ERR_NOMEM);
    }
    cudaEventCreate(&sl.e_in); cudaEventCreate(&sl.e_done);
    cudaEventCreateWithFlags(&sl.e_out, cudaEventBlockingSync);
  }
  if (defer) {
    c->broken = true;                  // no model yet: every infer call returns SNB_ERR_MODEL until snb_set_weights succeeds
  } else {
    r = build_plan(c);
    if (r == SNB_OK) r = warm_up(c);
    if (r != SNB_OK) return bail(r);
  }
  c->fps_t0 = now_s();
  c->worker = std::thread(worker_main, c);
  *out = c;
  return SNB_OK;
}

void snb_destroy(snb_ctx* c) {
  if (!c) return;
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->stop = true;
  }
  c->cv_push.notify_all(); c->cv_pop.notify_all();
  if (c->worker.joinable()) c->worker.join();
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  free_ctx(c);
  delete c;
}

// Wait until the most recent pass - on whatever stream it was enqueued - has finished.
static void sync_last_pass(snb_ctx* c) {
  cudaStreamSynchronize(c->stream);
  if (c->last_stream) cudaEventSynchronize(c->ev_last);
}

int snb_set_weights(snb_ctx* c, const void* blob, uint64_t bytes, int is_device) {
  if (!c || !blob || !bytes) return fail(c, SNB_ERR_INVALID, "snb_set_weights: bad arguments");
  std::lock_guard<std::mutex> run(c->run_mu);
  cudaSetDevice(c->cfg.device);
  std::vector<char> host;
  if (is_device) {
    host.resize(bytes);
    CK(c, cudaMemcpy(host.data(), blob, bytes, cudaMemcpyDeviceToHost));
    blob = host.data();
  }
  // everything that can be wrong with the blob is found here, before the live model is touched
  std::map<std::string, HostTensor> wts;
  int blob_K = -1;
  int r = parse_blob(blob, bytes, &wts, &blob_K, c->err, sizeof(c->err));
  if (r == SNB_OK && blob_K != c->K) {
    snprintf(c->err, sizeof(c->err), "weight blob was generated for K=%d, this context runs K=%d", blob_K, c->K);
    r = SNB_ERR_MODEL;
  }
  if (r != SNB_OK) { snprintf(g_err, sizeof(g_err), "%s", c->err); return r; }     // the old model stays installed
  sync_last_pass(c);                                   // no pass may still read the buffers replaced below
  // device pointers baked into the plan/graphs change: rebuild both
  for (auto& g : c->graphs) { cudaGraphExecDestroy(g.second.exec); if (g.second.graph) cudaGraphDestroy(g.second.graph); }
  c->graphs.clear();
  c->wts = std::move(wts); c->blob_K = blob_K;
  r = upload_weights(c);
  if (r == SNB_OK) r = build_plan(c);
  if (r == SNB_OK && !c->warmed) r = warm_up(c);       // a context created with SNB_FLAG_DEFER_WEIGHTS runs its first pass here
  // only a device allocation can fail at this point: the old plan is gone, so the context refuses to run until a later
  // snb_set_weights succeeds
  c->broken = r != SNB_OK;
  if (r != SNB_OK) snprintf(g_err, sizeof(g_err), "%s", c->err);
  return r;
}

// Host tensor memory.  A small header in front of the block records which allocator owns it.
int snb_sys_alloc(void** ptr, uint64_t bytes) {
  if (!ptr || !bytes) return SNB_ERR_INVALID;
  *ptr = nullptr;
  void* base = nullptr;
  uint64_t tag = 1;                            // 1: cudaHostAlloc, 2: posix_memalign
  if (cudaHostAlloc(&base, bytes + 64, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();                        // no device / driver: ordinary host memory
    if (posix_memalign(&base, 64, bytes + 64) != 0) return SNB_ERR_NOMEM;
    tag = 2;
  }
  memcpy(base, &tag, sizeof(tag));
  *ptr = static_cast<char*>(base) + 64;
  return SNB_OK;
}

void snb_sys_free(void* ptr) {
  if (!ptr) return;
  void* base = static_cast<char*>(ptr) - 64;
  uint64_t tag = 0;
  memcpy(&tag, base, sizeof(tag));
  if (tag == 1) cudaFreeHost(base);
  else if (tag == 2) free(base);
}

int snb_get_io(const snb_ctx* c, snb_tensor_props* in, snb_tensor_props* out) {
  if (!c) return SNB_ERR_INVALID;
  if (in) {
    memset(in, 0, sizeof(*in));
    const int32_t s[4] = {1, 6, c->H, c->W};
    memcpy(in->valid_shape, s, sizeof(s)); memcpy(in->aligned_shape, s, sizeof(s));
    in->tensor_layout = SNB_LAYOUT_NCHW; in->tensor_type = SNB_TENSOR_S8;
    in->scale_len = 1; in->scale = 1.0f / 128.0f; in->mem_size = c->in_bytes;
  }
  if (out) {
    memset(out, 0, sizeof(*out));
    const int32_t s[4] = {1, 1, c->H, c->W};
    memcpy(out->valid_shape, s, sizeof(s)); memcpy(out->aligned_shape, s, sizeof(s));
    out->tensor_layout = SNB_LAYOUT_NCHW; out->tensor_type = SNB_TENSOR_S32;
    out->scale_len = 1; out->scale = 2.60443857769133e-06f; out->mem_size = c->out_bytes;
  }
  return SNB_OK;
}

int snb_get_model_input_size(const snb_ctx* c, int32_t idx, int32_t* w, int32_t* h) {
  if (!c || idx != 0 || !w || !h) return SNB_ERR_INVALID;
  *w = c->W; *h = c->H;
  return SNB_OK;
}

// one chunk (<= maxB pairs) on stream st.  d_in: the s8 tensor (input of the pass; with d_frames != nullptr it is the buffer
// the NV12 kernel writes the tensor into first).
static int run_chunk(snb_ctx* c, int B, int8_t* d_in, const uint8_t* d_frames, int32_t* d_out, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (c->broken) { snprintf(c->err, sizeof(c->err), "no model installed: the last snb_set_weights failed"); return SNB_ERR_MODEL; }
  // the scratch set is shared by every pass: order this pass behind the previous one when that ran on another stream
  if (c->last_stream && c->last_stream != st && cudaStreamWaitEvent(st, c->ev_last, 0) != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "cudaStreamWaitEvent: %s", cudaGetErrorString(cudaGetLastError()));
    return SNB_ERR_CUDA;
  }
  const int correct = (c->cfg.flags & SNB_FLAG_CORRECT_CHROMA) ? 1 : 0;
  IoPtrs io;
  io.s8 = d_in; io.q = d_out;
  if (c->direct_io) {
    // the pass reads the s8 tensor and writes the s32 tensor itself; camera frames become that tensor by the first kernel of
    // the pass (P1-P3 on the GPU, inside the captured graph: run_plan)
    io.frames = d_frames;
  } else if (d_frames) {
    e = launch_pre_nv12(d_frames, c->img, nullptr, B, c->H, c->W, correct, st);
  } else {
    e = launch_pre_s8(d_in, c->img, B, c->H, c->W, st);
  }
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "pre: %s", cudaGetErrorString(e)); return SNB_ERR_CUDA; }
  int r = run_plan(c, B, io, st, !(c->cfg.flags & (SNB_FLAG_NO_GRAPH | SNB_FLAG_KEEP_STAGES)));
  if (r != SNB_OK) return r;
  if (!c->direct_io) {
    Plane d = c->disp_final; d.n = B;
    e = launch_post_quant(d, d_out, c->H, c->W, c->qmul, st);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "post: %s", cudaGetErrorString(e)); return SNB_ERR_CUDA; }
  }
  cudaEventRecord(c->ev_last, st);
  c->last_stream = st;
  c->last_B = B;
  ++c->n_passes;
  return SNB_OK;
}

// kernels of this library per pass: the plan, plus the pre / post kernels where the pass does not contain them
static int launches_per_pass(const snb_ctx* c, bool nv12) { return (int)c->ops.size() + (c->direct_io ? (nv12 ? 1 : 0) : 2); }

static int infer_host(snb_ctx* c, const int8_t* in, const uint8_t* frames, int32_t* out, int batch) {
  if (!c || (!in && !frames) || !out || batch < 1) return fail(c, SNB_ERR_INVALID, "snb_infer: bad arguments");
  std::lock_guard<std::mutex> run(c->run_mu);
  CK(c, cudaSetDevice(c->cfg.device));
  const double t0 = now_s();
  cudaStream_t st = c->stream;
  float gpu_ms = 0.f;
  for (int b0 = 0; b0 < batch; b0 += c->maxB) {
    const int B = std::min(c->maxB, batch - b0);
    if (frames) CK(c, cudaMemcpyAsync(c->d_frames, frames + (size_t)b0 * c->frame_bytes, c->frame_bytes * B, cudaMemcpyHostToDevice, st));
    else CK(c, cudaMemcpyAsync(c->d_in, in + (size_t)b0 * c->in_bytes, c->in_bytes * B, cudaMemcpyHostToDevice, st));
    cudaEventRecord(c->ev[0], st);
    int r = run_chunk(c, B, c->d_in, frames ? c->d_frames : nullptr, c->d_out, st);
    if (r != SNB_OK) { snprintf(g_err, sizeof(g_err), "%s", c->err); return r; }
    cudaEventRecord(c->ev[1], st);
    CK(c, cudaMemcpyAsync(out + (size_t)b0 * c->H * c->W, c->d_out, c->out_bytes * B, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); gpu_ms += ms;
  }
  const double t1 = now_s();
  c->stat.gpu_ms = gpu_ms;
  c->stat.infer_time_ms = (int)((t1 - t0) * 1e3 + 0.5);
  c->stat.kernel_launches = launches_per_pass(c, frames != nullptr);
  c->fps_in += batch; c->fps_out += batch;
  c->stat.fps_updated = 0;
  if (t1 - c->fps_t0 >= 1.0) {     // dnn_node refreshes its fps statistics about once a second
    c->stat.input_fps = (float)(c->fps_in / (t1 - c->fps_t0));
    c->stat.output_fps = (float)(c->fps_out / (t1 - c->fps_t0));
    c->fps_in = c->fps_out = 0; c->fps_t0 = t1; c->stat.fps_updated = 1;
  }
  return SNB_OK;
}

int snb_infer(snb_ctx* c, const int8_t* in, int32_t* out, int32_t batch) { return infer_host(c, in, nullptr, out, batch); }

int snb_infer_nv12(snb_ctx* c, const uint8_t* frames, int32_t* out, int32_t batch) {
  // NV12 has 2x2 chroma blocks: odd model sizes (e.g. KITTI's 375 rows) exist only at the s8 tensor entry points
  if (c && ((c->H & 1) || (c->W & 1))) return fail(c, SNB_ERR_INVALID, "snb_infer_nv12: NV12 frames need even height and width");
  return infer_host(c, nullptr, frames, out, batch);
}

int snb_infer_device(snb_ctx* c, const int8_t* d_in, int32_t* d_out, int32_t batch, void* cuda_stream) {
  if (!c || !d_in || !d_out || batch < 1) return fail(c, SNB_ERR_INVALID, "snb_infer_device: bad arguments");
  std::lock_guard<std::mutex> run(c->run_mu);
  CK(c, cudaSetDevice(c->cfg.device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  for (int b0 = 0; b0 < batch; b0 += c->maxB) {
    const int B = std::min(c->maxB, batch - b0);
    int r = run_chunk(c, B, const_cast<int8_t*>(d_in) + (size_t)b0 * c->in_bytes, nullptr, d_out + (size_t)b0 * c->H * c->W, st);
    if (r != SNB_OK) { snprintf(g_err, sizeof(g_err), "%s", c->err); return r; }
  }
  c->stat.kernel_launches = launches_per_pass(c, false);
  if (!cuda_stream) CK(c, cudaStreamSynchronize(st));
  return SNB_OK;
}

static int submit_async(snb_ctx* c, const int8_t* in, const uint8_t* frames, int32_t* out, int32_t batch, snb_done_fn done, void* user,
                        int32_t timeout_ms) {
  std::unique_lock<std::mutex> lk(c->mu);
  const int cap = std::max(1, c->cfg.task_num);
  auto room = [&] { return c->stop || c->inflight < cap; };
  if (timeout_ms < 0) c->cv_push.wait(lk, room);
  else if (!c->cv_push.wait_for(lk, std::chrono::milliseconds(timeout_ms), room)) return fail(c, SNB_ERR_BUSY, "snb_infer_async: no free task slot");
  if (c->stop) return SNB_ERR_INVALID;
  c->queue.push_back(Task{in, frames, out, batch, done, user});
  ++c->inflight;
  lk.unlock();
  c->cv_pop.notify_one();
  return SNB_OK;
}

int snb_infer_async(snb_ctx* c, const int8_t* in, int32_t* out, int32_t batch, snb_done_fn done, void* user, int32_t timeout_ms) {
  if (!c || !in || !out || batch < 1) return fail(c, SNB_ERR_INVALID, "snb_infer_async: bad arguments");
  return submit_async(c, in, nullptr, out, batch, done, user, timeout_ms);
}

int snb_infer_nv12_async(snb_ctx* c, const uint8_t* frames, int32_t* out, int32_t batch, snb_done_fn done, void* user, int32_t timeout_ms) {
  if (!c || !frames || !out || batch < 1) return fail(c, SNB_ERR_INVALID, "snb_infer_nv12_async: bad arguments");
  if ((c->H & 1) || (c->W & 1)) return fail(c, SNB_ERR_INVALID, "snb_infer_nv12_async: NV12 frames need even height and width");
  return submit_async(c, nullptr, frames, out, batch, done, user, timeout_ms);
}

// preprocess.cpp:913-1059 on the GPU: the s8 tensor CvtNV12Data2Tensors builds, from raw camera frames
int snb_pre_nv12_gpu(snb_ctx* c, const uint8_t* frames, int32_t batch, int8_t* s8_out) {
  if (!c || !frames || !s8_out || batch < 1) return fail(c, SNB_ERR_INVALID, "snb_pre_nv12_gpu: bad arguments");
  if ((c->H & 1) || (c->W & 1)) return fail(c, SNB_ERR_INVALID, "snb_pre_nv12_gpu: NV12 frames need even height and width");
  std::lock_guard<std::mutex> run(c->run_mu);
  if (c->broken) return fail(c, SNB_ERR_MODEL, "no model installed");
  CK(c, cudaSetDevice(c->cfg.device));
  sync_last_pass(c);
  cudaStream_t st = c->stream;
  for (int b0 = 0; b0 < batch; b0 += c->maxB) {
    const int B = std::min(c->maxB, batch - b0);
    CK(c, cudaMemcpyAsync(c->d_frames, frames + (size_t)b0 * c->frame_bytes, c->frame_bytes * B, cudaMemcpyHostToDevice, st));
    cudaError_t e = launch_pre_nv12(c->d_frames, c->direct_io ? Tens() : c->img, c->d_in, B, c->H, c->W, (c->cfg.flags & SNB_FLAG_CORRECT_CHROMA) ? 1 : 0, st);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "pre_nv12: %s", cudaGetErrorString(e)); return SNB_ERR_CUDA; }
    CK(c, cudaMemcpyAsync(s8_out + (size_t)b0 * c->in_bytes, c->d_in, c->in_bytes * B, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
  }
  return SNB_OK;
}

int snb_wait_all(snb_ctx* c) {
  if (!c) return SNB_ERR_INVALID;
  std::unique_lock<std::mutex> lk(c->mu);
  c->cv_push.wait(lk, [&] { return (c->inflight == 0 && c->in_callbacks == 0) || c->stop; });
  return SNB_OK;
}

int64_t snb_get_pass_count(const snb_ctx* c) { return c ? (int64_t)c->n_passes : SNB_ERR_INVALID; }

int snb_get_rt_stat(const snb_ctx* c, snb_rt_stat* s) {
  if (!c || !s) return SNB_ERR_INVALID;
  *s = c->stat;
  return SNB_OK;
}

int64_t snb_debug_read(snb_ctx* c, const char* name, float* dst, uint64_t cap, int32_t shape[5]) {
  if (!c || !name) return SNB_ERR_INVALID;
  if (!(c->cfg.flags & SNB_FLAG_KEEP_STAGES)) return fail(c, SNB_ERR_INVALID, "snb_debug_read needs SNB_FLAG_KEEP_STAGES");
  auto it = c->stages.find(name);
  if (it == c->stages.end()) return fail(c, SNB_ERR_INVALID, "snb_debug_read: unknown stage");
  std::lock_guard<std::mutex> run(c->run_mu);
  cudaSetDevice(c->cfg.device);
  sync_last_pass(c);
  const Stage& s = it->second;
  const int B = std::max(1, c->last_B);
  if (s.is_plane) {
    const Plane& p = s.p;
    const size_t n = (size_t)B * p.d * p.h * p.w;
    if (shape) { shape[0] = B; shape[1] = 1; shape[2] = p.d; shape[3] = p.h; shape[4] = p.w; }
    if (!dst) return (int64_t)n;
    if (cap < n) return SNB_ERR_INVALID;
    CK(c, cudaMemcpy(dst, p.p, n * 4, cudaMemcpyDeviceToHost));
    return (int64_t)n;
  }
  const Tens& t = s.t;
  const int N = s.nmul * B;
  const size_t n = (size_t)N * t.c * t.d * t.h * t.w;
  if (shape) { shape[0] = N; shape[1] = t.c; shape[2] = t.d; shape[3] = t.h; shape[4] = t.w; }
  if (!dst) return (int64_t)n;
  if (cap < n) return SNB_ERR_INVALID;
  const size_t sp = (size_t)t.d * t.h * t.w;
  const size_t slice = t.slice();
  const int ws = t.ws(), pad = t.pad;
  // element (nn, hl, ch, dd, y, x) of the padded C8 layout
  const size_t ss = t.sample_stride(), lo = t.lo_off();
  const size_t span = (size_t)(N - 1) * ss + (t.planes - 1) * lo + (size_t)t.cb * t.d * slice;   // sub-views included
  auto at = [&](int nn, int hl, int ch, int dd, int y, int x) {
    return (size_t)nn * ss + (size_t)hl * lo + ((size_t)(ch / 8) * t.d + dd) * slice + ((size_t)(y + pad) * ws + x + pad) * 8 + ch % 8;
  };
  if (t.planes == 2) {
    std::vector<__half> tmp(span);
    CK(c, cudaMemcpy(tmp.data(), t.p, tmp.size() * 2, cudaMemcpyDeviceToHost));
    for (int nn = 0; nn < N; ++nn)
      for (int ch = 0; ch < t.c; ++ch) {
        float* d = dst + ((size_t)nn * t.c + ch) * sp;
        for (int dd = 0; dd < t.d; ++dd)
          for (int y = 0; y < t.h; ++y)
            for (int x = 0; x < t.w; ++x)
              d[((size_t)dd * t.h + y) * t.w + x] = __half2float(tmp[at(nn, 0, ch, dd, y, x)]) + __half2float(tmp[at(nn, 1, ch, dd, y, x)]);
      }
    return (int64_t)n;
  }
  std::vector<float> tmp(span);
  CK(c, cudaMemcpy(tmp.data(), t.p, tmp.size() * 4, cudaMemcpyDeviceToHost));
  for (int nn = 0; nn < N; ++nn)
    for (int ch = 0; ch < t.c; ++ch) {
      float* d = dst + ((size_t)nn * t.c + ch) * sp;
      for (int dd = 0; dd < t.d; ++dd)
        for (int y = 0; y < t.h; ++y)
          for (int x = 0; x < t.w; ++x) d[((size_t)dd * t.h + y) * t.w + x] = tmp[at(nn, 0, ch, dd, y, x)];
    }
  return (int64_t)n;
}

int snb_post_depth_color(snb_ctx* c, const int32_t* q, int32_t batch, float alpha, float* depth_m, uint8_t* bgr, int32_t is_device) {
  if (!c || !q || batch < 1 || (!depth_m && !bgr)) return fail(c, SNB_ERR_INVALID, "snb_post_depth_color: bad arguments");
  std::lock_guard<std::mutex> run(c->run_mu);
  CK(c, cudaSetDevice(c->cfg.device));
  cudaStream_t st = c->stream;
  const size_t n = (size_t)batch * c->H * c->W;
  const float scale = 2.60443857769133e-06f;
  if (is_device) {
    cudaError_t e = launch_post_depth_color(q, depth_m, bgr, n, scale, alpha, st);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "post_depth_color: %s", cudaGetErrorString(e)); return SNB_ERR_CUDA; }
    CK(c, cudaStreamSynchronize(st));
    return SNB_OK;
  }
  int32_t* dq = nullptr; float* dd = nullptr; uint8_t* dc = nullptr;
  auto cleanup = [&] { if (dq) cudaFree(dq); if (dd) cudaFree(dd); if (dc) cudaFree(dc); };
  if (cudaMalloc(&dq, n * 4) != cudaSuccess || (depth_m && cudaMalloc(&dd, n * 4) != cudaSuccess) || (bgr && cudaMalloc(&dc, n * 3) != cudaSuccess)) {
    cleanup();
    return fail(c, SNB_ERR_NOMEM, "snb_post_depth_color: device allocation failed");
  }
  cudaError_t e = cudaMemcpyAsync(dq, q, n * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = launch_post_depth_color(dq, dd, dc, n, scale, alpha, st);
  if (e == cudaSuccess && depth_m) e = cudaMemcpyAsync(depth_m, dd, n * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && bgr) e = cudaMemcpyAsync(bgr, dc, n * 3, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "snb_post_depth_color: %s", cudaGetErrorString(e)); return SNB_ERR_CUDA; }
  return SNB_OK;
}

int snb_profile_pass(snb_ctx* c, int32_t batch, snb_kernel_time* out, int32_t cap) {
  if (!c || batch < 1 || batch > c->maxB) return fail(c, SNB_ERR_INVALID, "snb_profile_pass: bad batch");
  std::lock_guard<std::mutex> run(c->run_mu);
  if (c->broken) return fail(c, SNB_ERR_MODEL, "no model installed");
  CK(c, cudaSetDevice(c->cfg.device));
  cudaStream_t st = c->stream;
  sync_last_pass(c);
  const bool wrap = !c->direct_io;                      // the older pipeline: pre_s8 before and post_quant after the plan
  const int nops = (int)c->ops.size(), n = nops + (wrap ? 2 : 0), o0 = wrap ? 1 : 0;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  IoPtrs io;
  io.s8 = c->d_in; io.q = c->d_out;
  cudaEventRecord(ev[0], st);
  if (wrap) {
    launch_pre_s8(c->d_in, c->img, batch, c->H, c->W, st);
    cudaEventRecord(ev[1], st);
  }
  for (int i = 0; i < nops; ++i) {
    cudaError_t e = c->ops[i].fn(batch, io, st);
    if (e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "%s: %s", c->ops[i].name.c_str(), cudaGetErrorString(e)); return SNB_ERR_CUDA; }
    cudaEventRecord(ev[o0 + i + 1], st);
  }
  if (wrap) {
    Plane d = c->disp_final; d.n = batch;
    launch_post_quant(d, c->d_out, c->H, c->W, c->qmul, st);
    cudaEventRecord(ev[n], st);
  }
  CK(c, cudaStreamSynchronize(st));
  for (int i = 0; i < n && i < cap; ++i) {
    memset(&out[i], 0, sizeof(out[i]));
    const bool is_pre = wrap && i == 0, is_post = wrap && i == n - 1;
    const char* nm = is_pre ? "pre_s8" : (is_post ? "post_quant" : c->ops[i - o0].name.c_str());
    snprintf(out[i].name, sizeof(out[i].name), "%s", nm);
    cudaEventElapsedTime(&out[i].ms, ev[i], ev[i + 1]);
    if (is_pre) out[i].bytes = (double)batch * (c->in_bytes + 2.0 * c->Hp * c->Wp * 32);
    else if (is_post) out[i].bytes = (double)batch * (4.0 * c->Hp * c->Wp + c->out_bytes);
    else { out[i].flops = c->ops[i - o0].flops * batch; out[i].bytes = c->ops[i - o0].bytes * batch; }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return n;
}

}  // extern "C"

// Retire the oldest in-flight pass: wait for its device->host copies, fire the callbacks, free its task slots.
static void retire_oldest(snb_ctx* c) {
  AsyncSlot& sl = c->slots[c->n_ret % c->slots.size()];
  int status = SNB_OK;
  if (cudaEventSynchronize(sl.e_out) != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "async call failed: %s", cudaGetErrorString(cudaGetLastError()));
    status = SNB_ERR_CUDA;
  }
  snb_rt_stat st;
  {
    std::lock_guard<std::mutex> run(c->run_mu);
    float ms = 0.f;
    if (status == SNB_OK) cudaEventElapsedTime(&ms, sl.e_in, sl.e_done);
    const double t1 = now_s();
    c->stat.gpu_ms = ms;
    c->stat.infer_time_ms = (int)((t1 - sl.t0) * 1e3 + 0.5);
    c->stat.kernel_launches = launches_per_pass(c, !sl.tasks.empty() && sl.tasks[0].frames != nullptr);
    c->fps_in += sl.batch; c->fps_out += sl.batch;
    c->stat.fps_updated = 0;
    if (t1 - c->fps_t0 >= 1.0) {
      c->stat.input_fps = (float)(c->fps_in / (t1 - c->fps_t0));
      c->stat.output_fps = (float)(c->fps_out / (t1 - c->fps_t0));
      c->fps_in = c->fps_out = 0; c->fps_t0 = t1; c->stat.fps_updated = 1;
    }
    st = c->stat;
  }
  // The pass is over: its staging slot and its task slots are free again BEFORE the callbacks run, so a callback may
  // submit the next frame (snb_infer_async with timeout -1) without deadlocking on a full task table.  snb_wait_all
  // counts callbacks still running through `in_callbacks`.
  std::vector<Task> tasks;
  tasks.swap(sl.tasks);
  sl.busy = false;
  ++c->n_ret;
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->inflight -= (int)tasks.size();
    c->in_callbacks += (int)tasks.size();
  }
  c->cv_push.notify_all();
  for (const Task& t : tasks)
    if (t.done) t.done(t.user, status, &st);
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->in_callbacks -= (int)tasks.size();
  }
  c->cv_push.notify_all();
}

// Enqueue one pass (sum of the calls' batches <= max_batch) without waiting for it: H2D on st_in, kernels on the compute
// stream, D2H on st_out, chained by events.
static int enqueue_async(snb_ctx* c, const std::vector<Task>& group) {
  AsyncSlot& sl = c->slots[c->n_enq % c->slots.size()];
  std::lock_guard<std::mutex> run(c->run_mu);
  CK(c, cudaSetDevice(c->cfg.device));
  sl.tasks = group; sl.t0 = now_s();
  int B = 0;
  const bool nv12 = group[0].frames != nullptr;        // a pass is fed by ONE pre-process kernel: the worker never mixes kinds
  for (const Task& t : group) {
    if (nv12) CK(c, cudaMemcpyAsync(sl.d_frames + (size_t)B * c->frame_bytes, t.frames, c->frame_bytes * t.batch, cudaMemcpyHostToDevice, c->st_in));
    else CK(c, cudaMemcpyAsync(sl.d_in + (size_t)B * c->in_bytes, t.in, c->in_bytes * t.batch, cudaMemcpyHostToDevice, c->st_in));
    B += t.batch;
  }
  sl.batch = B;
  CK(c, cudaEventRecord(sl.e_in, c->st_in));
  CK(c, cudaStreamWaitEvent(c->stream, sl.e_in, 0));
  int r = run_chunk(c, B, sl.d_in, nv12 ? sl.d_frames : nullptr, sl.d_out, c->stream);
  if (r != SNB_OK) return r;
  CK(c, cudaEventRecord(sl.e_done, c->stream));
  CK(c, cudaStreamWaitEvent(c->st_out, sl.e_done, 0));
  B = 0;
  for (const Task& t : group) {
    CK(c, cudaMemcpyAsync(t.out, sl.d_out + (size_t)B * c->H * c->W, c->out_bytes * t.batch, cudaMemcpyDeviceToHost, c->st_out));
    B += t.batch;
  }
  CK(c, cudaEventRecord(sl.e_out, c->st_out));
  sl.busy = true;
  ++c->n_enq;
  return SNB_OK;
}

// The worker thread (= the reference's PostProcess thread: callbacks fire here).
// Coalescing (max_batch > 1, not SNB_FLAG_NO_COALESCE): at configs[1] one pair per pass leaves the GPU latency-bound
// (1.58 ms), two pairs per pass cost 2.68 ms and three 3.81 ms (tools/batch_probe.py), so queued calls are merged into
// one pass of up to max_batch pairs.  To let calls queue up the worker keeps only TWO passes on the GPU (one running,
// one behind it - enough to hide copies and launch latency) and, while a pass is still running, gives the caller a
// moment (200 us) to submit the rest of a batch before it launches the next pass.  Per-call semantics do not change:
// own input / output buffers, own callback, callbacks in submission order.
static void worker_main(snb_ctx* c) {
  const bool coalesce = c->maxB > 1 && !(c->cfg.flags & SNB_FLAG_NO_COALESCE);
  const uint64_t max_passes = coalesce ? std::min<uint64_t>(2, c->slots.size()) : c->slots.size();
  const int cap = std::max(1, c->cfg.task_num);
  for (;;) {
    std::vector<Task> group;
    {
      std::unique_lock<std::mutex> lk(c->mu);
      if (c->queue.empty()) {
        if (c->n_enq == c->n_ret) {
          c->cv_pop.wait(lk, [&] { return c->stop || !c->queue.empty(); });
          if (c->queue.empty()) return;          // stop requested, nothing queued, nothing in flight
        } else if (coalesce) {
          // passes in flight and nothing queued: sleep until a call arrives, but keep an eye on the oldest pass
          c->cv_pop.wait_for(lk, std::chrono::microseconds(50), [&] { return c->stop || !c->queue.empty(); });
        }
      }
      if (!c->queue.empty()) { group.push_back(c->queue.front()); c->queue.pop_front(); }
    }
    if (group.empty()) {
      if (c->n_enq == c->n_ret) continue;
      if (!coalesce || c->stop || cudaEventQuery(c->slots[c->n_ret % c->slots.size()].e_out) != cudaErrorNotReady) retire_oldest(c);
      continue;
    }
    while (c->n_enq - c->n_ret >= max_passes) retire_oldest(c);
    int total = group[0].batch;
    int r;
    if (total <= c->maxB) {
      if (coalesce) {
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(200);
        std::unique_lock<std::mutex> lk(c->mu);
        for (;;) {
          while (!c->queue.empty() && total + c->queue.front().batch <= c->maxB &&
                 (c->queue.front().frames != nullptr) == (group[0].frames != nullptr)) {
            group.push_back(c->queue.front()); total += c->queue.front().batch; c->queue.pop_front();
          }
          // stop gathering: the pass is full, the next call does not fit, the caller cannot submit more (every task
          // slot is taken), or the GPU has nothing left to do
          if (total >= c->maxB || !c->queue.empty() || c->inflight >= cap || c->n_enq == c->n_ret || c->stop) break;
          if (!c->cv_pop.wait_until(lk, deadline, [&] { return c->stop || !c->queue.empty(); })) break;
        }
      }
      r = enqueue_async(c, group);
      if (r == SNB_OK) continue;
    } else {
      while (c->n_enq != c->n_ret) retire_oldest(c);   // larger than one pass: chunked synchronous path
      r = infer_host(c, group[0].in, group[0].frames, group[0].out, group[0].batch);
    }
    snb_rt_stat st = c->stat;
    for (const Task& t : group)
      if (t.done) t.done(t.user, r, &st);
    {
      std::unique_lock<std::mutex> lk(c->mu);
      c->inflight -= (int)group.size();
    }
    c->cv_push.notify_all();
  }
}
