// Context of one B200-resident StereoNet model instance: weights, scratch arena, static op plan.
// Replaces the state the closed dnn_node runtime keeps behind DnnNode::{Init,GetModel,Run}
// (stereonet_node.cpp:44,51,812).
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/snb200.h"
#include "common.cuh"

namespace snb {

struct HostTensor {
  std::vector<int> shape;
  std::vector<float> data;
};

struct ConvW {             // device-resident, kernel-specific packing of one convolution
  float* w = nullptr;      // direct: [cc][cb][kz][tap][8][CO]; to1: [cb][kz][9][8]
  std::map<int, __half*> w_tc;   // k_conv_tc packing by NT: [cc][k16][dz][ky][hl][kx][2][NT][8]
  float* b = nullptr;      // [cout]
  float b0 = 0.f;          // bias of Cout=1 convs
  int cout = 0, cin = 0, ks = 1, kz = 1;
  int wlog2 = 0;           // the tcgen05 packings hold w * 2^wlog2 (common.cuh weight_scale_log2)
};

struct Op {
  std::string name;
  std::function<cudaError_t(int B, const IoPtrs& io, cudaStream_t)> fn;
  double flops = 0;        // algorithmic FLOPs per stereo pair
  double bytes = 0;        // algorithmic HBM bytes per stereo pair
  int io_bytes = 0;        // > 0: the kernel reads / writes the call's own buffers; = sizeof its parameter struct, IoPtrs first
};

// A captured pass (one per batch size) and what is needed to point it at another call's buffers without re-capturing:
// the kernel nodes whose parameter struct starts with an IoPtrs, each with a private copy of that struct.
struct GraphEntry {
  cudaGraph_t graph = nullptr;                 // kept: node handles are only valid while their graph lives
  cudaGraphExec_t exec = nullptr;
  IoPtrs io;                                   // what the nodes currently point at
  struct IoNode { cudaGraphNode_t node; cudaKernelNodeParams kp; std::vector<char> args; void* argv[1]; };
  std::vector<std::unique_ptr<IoNode>> nodes;
};

struct Stage {             // named tap for snb_debug_read
  bool is_plane = false;
  Tens t;
  Plane p;
  int nmul = 1;            // tensor batch = nmul * B
};

struct Arena {
  struct Blk { void* p; size_t bytes; bool free; uint64_t key; };
  std::vector<Blk> blks;
  bool reuse = true;
  size_t total = 0;
  void* get(size_t bytes, uint64_t key);
  void put(void* p);
  void release_all();
};

struct Task {
  const int8_t* in;        // s8 tensor [batch,6,H,W] (host) ...
  const uint8_t* frames;   // ... or raw side-by-side NV12 frames [batch][H*3/2][2W] (host); exactly one of the two is set
  int32_t* out; int batch; snb_done_fn done; void* user;
};

// One in-flight asynchronous pass: its own device staging buffers, so the host->device copy of pass k+1 and the
// device->host copy of pass k-1 overlap the kernels of pass k (three streams, ordered by events).
struct AsyncSlot {
  int8_t* d_in = nullptr;
  uint8_t* d_frames = nullptr;         // NV12 staging of snb_infer_nv12_async passes
  int32_t* d_out = nullptr;
  cudaEvent_t e_in = nullptr, e_done = nullptr, e_out = nullptr;
  std::vector<Task> tasks;             // the calls this pass serves (> 1: coalesced, see worker_main)
  int batch = 0;                       // pairs of the pass = sum of the calls' batches
  bool busy = false;
  double t0 = 0;
};

}  // namespace snb

struct snb_ctx {
  snb_config cfg{};
  std::string model_file;
  int H = 0, W = 0, K = 0, D = 0, Hp = 0, Wp = 0, h = 0, w = 0, maxB = 1;
  float qmul = 0.f;
  int planes = 1;                      // activation storage: 1 fp32, 2 split fp16 (SNB_PREC_TC_F16X2)
  int num_sms = 148;
  int n_tc_convs = 0, n_direct_convs = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // Every pass uses the same scratch (image tensor, arena, graphs): a pass enqueued on one stream must not start before
  // the previous pass - possibly on another stream (snb_infer_device with a caller's stream) - has finished.
  cudaEvent_t ev_last = nullptr;       // recorded after the last kernel of the most recent pass
  cudaStream_t last_stream = nullptr;  // the stream it was recorded on (nullptr: no pass yet)
  bool warmed = false;                 // the eager warm-up pass has run (kernel attributes set outside graph capture)
  bool broken = false;                 // a failed snb_set_weights left no usable plan: every infer call returns SNB_ERR_MODEL

  int blob_K = -1;
  std::map<std::string, snb::HostTensor> wts;
  std::map<std::string, snb::ConvW> convs;
  std::vector<void*> wallocs;

  snb::Arena arena;
  std::vector<snb::Op> ops;
  std::map<std::string, snb::Stage> stages;
  snb::Tens img;                       // C8 input image [2B][1][Hp][Wp][8]
  snb::Plane disp_final;               // [B][Hp][Wp]
  int8_t* d_in = nullptr;              // s8 [maxB][6][H][W]
  int32_t* d_out = nullptr;            // s32 [maxB][H][W]
  uint8_t* d_frames = nullptr;         // NV12 frames [maxB][H*3/2][2W]
  size_t in_bytes = 0, out_bytes = 0, frame_bytes = 0;   // per pair

  std::map<int, snb::GraphEntry> graphs;   // by batch (x2 + 1 for the camera-frame entry: its pass starts with the NV12 kernel)
  bool direct_io = false;              // tensor-core path: kernels of the pass read the s8 input / write the s32 output themselves (no image tensor, no pre / post launch)
  int last_B = 0;
  uint64_t n_passes = 0;               // whole-network passes launched so far (snb_get_pass_count)

  // async tasks (task_num in flight; callbacks on the worker thread = the reference's PostProcess thread)
  std::thread worker;
  std::vector<snb::AsyncSlot> slots;   // task_num entries
  cudaStream_t st_in = nullptr, st_out = nullptr;
  uint64_t n_enq = 0, n_ret = 0;       // calls enqueued / retired by the worker (slot = n % slots.size())
  std::mutex mu, run_mu;
  std::condition_variable cv_push, cv_pop;
  std::deque<snb::Task> queue;
  int inflight = 0;
  int in_callbacks = 0;                // done-callbacks currently running on the worker thread (snb_wait_all waits for them too)
  bool stop = false;

  snb_rt_stat stat{};
  double fps_t0 = 0; int fps_in = 0, fps_out = 0;
  mutable char err[512] = {0};
};

namespace snb {
int parse_blob(const void* blob, size_t bytes, std::map<std::string, HostTensor>* out, int* blob_K, char* err, size_t errlen);
int upload_weights(snb_ctx* c);
int build_plan(snb_ctx* c);
int run_plan(snb_ctx* c, int B, const IoPtrs& io, cudaStream_t st, bool use_graph);
void free_ctx(snb_ctx* c);
}  // namespace snb
