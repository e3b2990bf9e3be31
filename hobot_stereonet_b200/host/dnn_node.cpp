#include "dnn_node.h"

#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>

namespace hobot {
namespace dnn_node {

static void fill_props(const snb_tensor_props& s, hbDNNTensorProperties& p) {
  for (int i = 0; i < 4; ++i) {
    p.validShape.dimensionSize[i] = s.valid_shape[i];
    p.alignedShape.dimensionSize[i] = s.aligned_shape[i];
  }
  p.validShape.numDimensions = p.alignedShape.numDimensions = 4;
  p.tensorLayout = s.tensor_layout;
  p.tensorType = s.tensor_type;
  p.scale_storage = s.scale;
  p.scale.scaleLen = s.scale_len;
  p.scale.scaleData = &p.scale_storage;
}

int Model::GetInputTensorProperties(hbDNNTensorProperties& p, int idx) const {
  snb_tensor_props in;
  if (idx != 0 || snb_get_io(ctx_, &in, nullptr) != SNB_OK) return -1;
  fill_props(in, p);
  return 0;
}

int Model::GetOutputTensorProperties(hbDNNTensorProperties& p, int idx) const {
  snb_tensor_props out;
  if (idx != 0 || snb_get_io(ctx_, nullptr, &out) != SNB_OK) return -1;
  fill_props(out, p);
  return 0;
}

// ---- pinned tensor memory, recycled -------------------------------------------------------------
namespace {
struct TensorPool {
  std::mutex mu;
  std::map<uint32_t, std::vector<void*>> idle;
  void* get(uint32_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto it = idle.find(bytes);
      if (it != idle.end() && !it->second.empty()) { void* p = it->second.back(); it->second.pop_back(); return p; }
    }
    void* p = nullptr;
    return snb_sys_alloc(&p, bytes) == SNB_OK ? p : nullptr;
  }
  void put(void* p, uint32_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      auto& v = idle[bytes];
      if ((int)v.size() < kPoolBlocksPerSize) { v.push_back(p); return; }
    }
    snb_sys_free(p);
  }
  void release() {
    std::map<uint32_t, std::vector<void*>> take;
    { std::lock_guard<std::mutex> lk(mu); take.swap(idle); }
    for (auto& kv : take) for (void* p : kv.second) snb_sys_free(p);
  }
};
TensorPool& tensor_pool() { static TensorPool* p = new TensorPool(); return *p; }   // never destroyed: outlives the CUDA runtime's teardown order
}  // namespace

std::shared_ptr<DNNTensor> AllocTensor(const hbDNNTensorProperties& props, uint32_t bytes) {
  void* p = tensor_pool().get(bytes);
  if (!p) return nullptr;
  auto t = std::shared_ptr<DNNTensor>(new DNNTensor(), [](DNNTensor* t) {
    if (t->sysMem[0].virAddr) tensor_pool().put(t->sysMem[0].virAddr, t->sysMem[0].memSize);
    delete t;
  });
  t->properties = props;
  t->properties.scale.scaleData = &t->properties.scale_storage;
  t->sysMem[0].virAddr = p;
  t->sysMem[0].memSize = bytes;
  return t;
}

void ReleaseTensorPool() { tensor_pool().release(); }

// ---- the node -----------------------------------------------------------------------------------
DnnNode::DnnNode(const std::string& node_name)
    : dnn_node_para_ptr_(std::make_shared<DnnNodePara>()), node_name_(node_name) {}

DnnNode::~DnnNode() {
  // A derived class calls Shutdown() from its own destructor.  If it did not, calls may still be in flight whose
  // completion would invoke the (now gone) derived PostProcess: release without running callbacks into a dead object
  // is impossible, so at least say so.
  if (model_.ctx_ || pool_) {
    fprintf(stderr, "[%s] DnnNode destroyed without Shutdown(): results of calls still in flight are dropped\n", node_name_.c_str());
    derived_gone_ = true;
    Shutdown();
  }
}

void DnnNode::Shutdown() {
  if (pool_) {
    snb_pool_wait_all(pool_);
    snb_pool_destroy(pool_);
    pool_ = nullptr;
    model_.ctx_ = nullptr;
  } else if (model_.ctx_) {
    snb_wait_all(model_.ctx_);
    snb_destroy(model_.ctx_);
    model_.ctx_ = nullptr;
  }
  ReleaseTensorPool();
}

int DnnNode::device_count() const { return pool_ ? snb_pool_size(pool_) : (model_.ctx_ ? 1 : 0); }

int DnnNode::Init() {
  if (model_.ctx_) return 0;
  if (SetNodePara() != 0) return -1;
  const DnnNodePara& p = *dnn_node_para_ptr_;
  if (p.model_task_type != ModelTaskType::ModelInferType) return -1;
  snb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = sizeof(cfg);
  cfg.height = p.model_in_h; cfg.width = p.model_in_w; cfg.K = p.K; cfg.D = p.D;
  // one frame per Run(), as the reference - but up to task_num frames may be in flight (stereonet_node.cpp:144), and the
  // library merges the queued ones into one pass (capi.cu worker_main): room for that many pairs per pass
  cfg.max_batch = p.task_num > 1 ? (p.task_num < 4 ? p.task_num : 4) : 1;
  cfg.device = p.device;
  cfg.task_num = p.task_num;
  cfg.precision = p.precision;
  cfg.model_file = p.model_file.c_str();
  if (p.devices.size() > 1) {
    std::vector<int32_t> devs(p.devices.begin(), p.devices.end());
    if (snb_pool_create(&pool_, &cfg, devs.data(), (int32_t)devs.size()) != SNB_OK) { pool_ = nullptr; return -1; }
    model_.ctx_ = snb_pool_ctx(pool_, 0);
    return 0;
  }
  if (p.devices.size() == 1) cfg.device = p.devices[0];
  return snb_create(&model_.ctx_, &cfg) == SNB_OK ? 0 : -1;
}

Model* DnnNode::GetModel() { return model_.ctx_ ? &model_ : nullptr; }

int DnnNode::GetModelInputSize(int32_t input_index, int& w, int& h) {
  int32_t ww = 0, hh = 0;
  if (!model_.ctx_ || snb_get_model_input_size(model_.ctx_, input_index, &ww, &hh) != SNB_OK) return -1;
  w = ww; h = hh;
  return 0;
}

std::string DnnNode::LastError() const {
  if (!model_.ctx_) return pool_ ? snb_pool_last_error(pool_) : (std::string(snb_last_error(nullptr)) + " " + snb_pool_last_error(nullptr));
  return pool_ ? snb_pool_last_error(pool_) : snb_last_error(model_.ctx_);
}

struct DnnNode::Task {
  DnnNode* node;
  std::shared_ptr<DnnNodeOutput> output;
  std::vector<std::shared_ptr<DNNTensor>> inputs;   // keeps the caller's tensors alive until done
};

std::shared_ptr<DNNTensor> DnnNode::AllocOutput() {
  hbDNNTensorProperties props;
  if (model_.GetOutputTensorProperties(props, 0) != 0) return nullptr;
  snb_tensor_props o;
  snb_get_io(model_.ctx_, nullptr, &o);
  return AllocTensor(props, (uint32_t)o.mem_size);
}

static std::shared_ptr<DnnNodeRtStat> make_stat(const snb_rt_stat& st) {
  auto rs = std::make_shared<DnnNodeRtStat>();
  rs->input_fps = st.input_fps; rs->output_fps = st.output_fps;
  rs->infer_time_ms = st.infer_time_ms; rs->fps_updated = st.fps_updated != 0;
  return rs;
}

void DnnNode::OnDone(void* user, int status, const snb_rt_stat* stat) {
  std::unique_ptr<Task> t(static_cast<Task*>(user));
  if (stat) t->output->rt_stat = make_stat(*stat);
  if (t->node->derived_gone_) return;
  if (status == SNB_OK) t->node->PostProcess(t->output);
  else fprintf(stderr, "[%s] inference failed: %s\n", t->node->node_name_.c_str(), t->node->LastError().c_str());
}

int DnnNode::Submit(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output, bool nv12,
                    bool is_sync_mode, int alloctask_timeout_ms) {
  if (!model_.ctx_ || !output || inputs.size() != 1 || !inputs[0] || !inputs[0]->sysMem[0].virAddr) return -1;
  snb_tensor_props in;
  snb_get_io(model_.ctx_, &in, nullptr);
  const uint64_t need = nv12 ? in.mem_size / 2 : in.mem_size;          // NV12: 2 views x 1.5 bytes per pixel; s8: 6 planes
  if (inputs[0]->sysMem[0].memSize < need) return -1;
  auto out = AllocOutput();
  if (!out) return -1;
  output->output_tensors.assign(1, out);
  const void* src = inputs[0]->sysMem[0].virAddr;
  int32_t* dst = static_cast<int32_t*>(out->sysMem[0].virAddr);
  if (is_sync_mode) {
    const int r = nv12 ? snb_infer_nv12(model_.ctx_, static_cast<const uint8_t*>(src), dst, 1)
                       : snb_infer(model_.ctx_, static_cast<const int8_t*>(src), dst, 1);
    if (r != SNB_OK) return -1;
    snb_rt_stat st;
    snb_get_rt_stat(model_.ctx_, &st);
    output->rt_stat = make_stat(st);
    return PostProcess(output);
  }
  Task* t = new Task{this, output, inputs};
  int r;
  if (pool_)
    r = nv12 ? snb_pool_infer_nv12_async(pool_, static_cast<const uint8_t*>(src), dst, 1, &DnnNode::OnDone, t, alloctask_timeout_ms)
             : snb_pool_infer_async(pool_, static_cast<const int8_t*>(src), dst, 1, &DnnNode::OnDone, t, alloctask_timeout_ms);
  else
    r = nv12 ? snb_infer_nv12_async(model_.ctx_, static_cast<const uint8_t*>(src), dst, 1, &DnnNode::OnDone, t, alloctask_timeout_ms)
             : snb_infer_async(model_.ctx_, static_cast<const int8_t*>(src), dst, 1, &DnnNode::OnDone, t, alloctask_timeout_ms);
  if (r != SNB_OK) { delete t; return -1; }
  return 0;
}

int DnnNode::Run(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output,
                 bool is_sync_mode, int alloctask_timeout_ms, int /*infer_timeout_ms*/) {
  return Submit(inputs, output, false, is_sync_mode, alloctask_timeout_ms);
}

int DnnNode::RunNv12(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output,
                     bool is_sync_mode, int alloctask_timeout_ms) {
  return Submit(inputs, output, true, is_sync_mode, alloctask_timeout_ms);
}

int DnnNode::WaitAll() {
  if (pool_) return snb_pool_wait_all(pool_) == SNB_OK ? 0 : -1;
  return model_.ctx_ && snb_wait_all(model_.ctx_) == SNB_OK ? 0 : -1;
}

}  // namespace dnn_node
}  // namespace hobot
