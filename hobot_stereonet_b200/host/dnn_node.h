// Stand-in for the closed tros `hobot::dnn_node` runtime, implemented on the snb200 C ABI
// (include/snb200.h).  Only the surface the reference's StereonetNode touches is mirrored, with the
// same names and argument meaning, so the node code above it reads like the reference's:
//   DnnNode::{Init, GetModel, GetModelInputSize, Run}       stereonet_node.cpp:44,45,51,812
//   virtual SetNodePara / PostProcess                        stereonet_node.h:68-71
//   DnnNodePara{model_file, model_task_type, task_num}      stereonet_node.cpp:136-144
//   DNNTensor{properties, sysMem[0].{virAddr,memSize}}      preprocess.cpp:952-973, stereonet_node.cpp:1033
//   DnnNodeOutput{msg_header, output_tensors, rt_stat}      stereonet_node.cpp:693-696,1033,1071-1086
//   Model::{GetInputCount, GetOutputCount, Get*TensorProperties}   stereonet_node.cpp:56-103
// There is no ROS dependency here; a ROS 2 shim only has to forward parameters and messages.
#pragma once
#include <stdint.h>

#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "../../include/snb200.h"

namespace hobot {
namespace dnn_node {

enum class ModelTaskType { InvalidType = 0, ModelInferType = 1, ModelRoiInferType = 2 };

// hbSysMem: where a tensor's bytes live.  virAddr is pinned host memory from snb_sys_alloc.
struct SysMem {
  uint64_t phyAddr = 0;
  void* virAddr = nullptr;
  uint32_t memSize = 0;
};

// hbDNNTensorShape / hbDNNTensorProperties: the fields the node reads.
struct TensorShape {
  int32_t dimensionSize[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int32_t numDimensions = 4;
};
struct QuantiScale {
  int32_t scaleLen = 0;
  float* scaleData = nullptr;
};
enum { HB_DNN_LAYOUT_NHWC = SNB_LAYOUT_NHWC, HB_DNN_LAYOUT_NCHW = SNB_LAYOUT_NCHW };
struct hbDNNTensorProperties {
  TensorShape validShape, alignedShape;
  int32_t tensorLayout = HB_DNN_LAYOUT_NCHW;
  int32_t tensorType = 0;
  QuantiScale scale;
  float scale_storage = 0.f;      // scale.scaleData points here (scaleLen == 1 for this model)
};

struct DNNTensor {
  SysMem sysMem[4];
  hbDNNTensorProperties properties;
};

struct MsgHeader {                 // std_msgs/Header
  std::string frame_id;
  int32_t stamp_sec = 0;
  uint32_t stamp_nanosec = 0;
};

struct DnnNodeRtStat {             // stereonet_node.cpp:1071-1086
  float input_fps = 0.f, output_fps = 0.f;
  int infer_time_ms = 0;
  bool fps_updated = false;
};

struct DnnNodeOutput {
  virtual ~DnnNodeOutput() = default;
  std::shared_ptr<MsgHeader> msg_header;
  std::vector<std::shared_ptr<DNNTensor>> output_tensors;
  std::shared_ptr<DnnNodeRtStat> rt_stat;
};

struct DnnNodePara {
  std::string model_file;
  ModelTaskType model_task_type = ModelTaskType::InvalidType;
  int task_num = 2;
  // The reference's geometry is compiled into its .hbm (hbm@0x1b60 records 0-1); a weight blob
  // carries K only, so the instance shape is a parameter here.  Defaults = the deployed model.
  int model_in_h = 720, model_in_w = 1280, D = 12, K = 4;
  int device = 0;
  std::vector<int> devices;        // more than one entry: one replica per GPU behind snb_pool_* (weights by one NCCL broadcast)
  int precision = SNB_PREC_TC_F16X2;
};

class DnnNode;

class Model {
 public:
  int GetInputCount() const { return 1; }
  int GetOutputCount() const { return 1; }
  int GetInputTensorProperties(hbDNNTensorProperties& p, int idx) const;
  int GetOutputTensorProperties(hbDNNTensorProperties& p, int idx) const;
  snb_ctx* GetDNNHandle() const { return ctx_; }

 private:
  friend class DnnNode;
  snb_ctx* ctx_ = nullptr;
};

class DnnNode {
 public:
  explicit DnnNode(const std::string& node_name);
  virtual ~DnnNode();
  DnnNode(const DnnNode&) = delete;
  DnnNode& operator=(const DnnNode&) = delete;

  // Calls SetNodePara(), then loads the model onto the GPU.  0 on success, -1 on failure.
  int Init();
  Model* GetModel();
  int GetModelInputSize(int32_t input_index, int& w, int& h);
  // inputs[0]: s8 NCHW [1,6,H,W] tensor.  is_sync_mode=false enqueues and returns; PostProcess then
  // runs on the runtime's own thread (at most task_num calls in flight; alloctask_timeout_ms bounds
  // the wait for a free task, -1 = forever).  <0 on failure.
  int Run(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output,
          bool is_sync_mode = false, int alloctask_timeout_ms = -1, int infer_timeout_ms = -1);
  // The B200 form of the same call: inputs[0] holds ONE raw side-by-side NV12 camera frame ([H*3/2][2W] bytes) instead of
  // the s8 tensor; the L/R split, chroma step and x-128 (stereonet_node.cpp:702-738, preprocess.cpp:913-1059) run on the GPU
  // inside the pass and the upload is half the bytes.  Same output, same callback, same error convention.
  int RunNv12(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output,
              bool is_sync_mode = false, int alloctask_timeout_ms = -1);
  // Blocks until every enqueued Run has been post-processed (the reference relies on rclcpp::spin).
  int WaitAll();
  // Drains the in-flight calls (their PostProcess still runs) and releases the GPU side.  A derived class MUST call this
  // from its own destructor: once ~DnnNode runs, the derived part - and with it PostProcess - is gone.
  void Shutdown();
  int device_count() const;
  const std::string& node_name() const { return node_name_; }
  std::string LastError() const;

 protected:
  virtual int SetNodePara() = 0;
  virtual int PostProcess(const std::shared_ptr<DnnNodeOutput>& node_output) = 0;
  std::shared_ptr<DnnNodePara> dnn_node_para_ptr_;

 private:
  struct Task;
  static void OnDone(void* user, int status, const snb_rt_stat* stat);
  std::shared_ptr<DNNTensor> AllocOutput();
  int Submit(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output, bool nv12,
             bool is_sync_mode, int alloctask_timeout_ms);
  std::string node_name_;
  Model model_;
  std::atomic<bool> derived_gone_{false};   // set by ~DnnNode: results of calls still in flight are dropped, PostProcess is not called
  snb_pool* pool_ = nullptr;       // set when the node drives more than one GPU; model_.ctx_ is then replica 0
};

// hbSysAllocCachedMem / hbSysFreeMem (preprocess.cpp:956-960,972).  The reference allocates and frees BPU memory per
// frame; page-locked host memory is expensive to map and its release synchronises the device, so blocks are recycled
// through a process-wide pool keyed by size (at most kPoolBlocksPerSize idle blocks per size).
constexpr int kPoolBlocksPerSize = 16;
std::shared_ptr<DNNTensor> AllocTensor(const hbDNNTensorProperties& props, uint32_t bytes);
void ReleaseTensorPool();          // frees the idle blocks (tensors still alive free their block when they die)

}  // namespace dnn_node
}  // namespace hobot
