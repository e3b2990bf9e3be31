#include "dnn_node.h"

#include <stdio.h>
#include <string.h>

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

std::shared_ptr<DNNTensor> AllocTensor(const hbDNNTensorProperties& props, uint32_t bytes) {
  void* p = nullptr;
  if (snb_sys_alloc(&p, bytes) != SNB_OK) return nullptr;
  auto t = std::shared_ptr<DNNTensor>(new DNNTensor(), [](DNNTensor* t) {
    if (t->sysMem[0].virAddr) snb_sys_free(t->sysMem[0].virAddr);
    delete t;
  });
  t->properties = props;
  t->properties.scale.scaleData = &t->properties.scale_storage;
  t->sysMem[0].virAddr = p;
  t->sysMem[0].memSize = bytes;
  return t;
}

DnnNode::DnnNode(const std::string& node_name)
    : dnn_node_para_ptr_(std::make_shared<DnnNodePara>()), node_name_(node_name) {}

DnnNode::~DnnNode() {
  if (model_.ctx_) snb_destroy(model_.ctx_);   // drains in-flight tasks first
}

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
  return snb_create(&model_.ctx_, &cfg) == SNB_OK ? 0 : -1;
}

Model* DnnNode::GetModel() { return model_.ctx_ ? &model_ : nullptr; }

int DnnNode::GetModelInputSize(int32_t input_index, int& w, int& h) {
  int32_t ww = 0, hh = 0;
  if (!model_.ctx_ || snb_get_model_input_size(model_.ctx_, input_index, &ww, &hh) != SNB_OK) return -1;
  w = ww; h = hh;
  return 0;
}

std::string DnnNode::LastError() const { return snb_last_error(model_.ctx_); }

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

void DnnNode::OnDone(void* user, int status, const snb_rt_stat* stat) {
  std::unique_ptr<Task> t(static_cast<Task*>(user));
  if (stat) {
    auto rs = std::make_shared<DnnNodeRtStat>();
    rs->input_fps = stat->input_fps; rs->output_fps = stat->output_fps;
    rs->infer_time_ms = stat->infer_time_ms; rs->fps_updated = stat->fps_updated != 0;
    t->output->rt_stat = rs;
  }
  if (status == SNB_OK) t->node->PostProcess(t->output);
  else fprintf(stderr, "[%s] inference failed: %s\n", t->node->node_name_.c_str(), snb_last_error(t->node->model_.ctx_));
}

int DnnNode::Run(std::vector<std::shared_ptr<DNNTensor>>& inputs, const std::shared_ptr<DnnNodeOutput>& output,
                 bool is_sync_mode, int alloctask_timeout_ms, int /*infer_timeout_ms*/) {
  if (!model_.ctx_ || !output || inputs.size() != 1 || !inputs[0] || !inputs[0]->sysMem[0].virAddr) return -1;
  snb_tensor_props in;
  snb_get_io(model_.ctx_, &in, nullptr);
  if (inputs[0]->sysMem[0].memSize < in.mem_size) return -1;
  auto out = AllocOutput();
  if (!out) return -1;
  output->output_tensors.assign(1, out);
  const int8_t* src = static_cast<const int8_t*>(inputs[0]->sysMem[0].virAddr);
  int32_t* dst = static_cast<int32_t*>(out->sysMem[0].virAddr);
  if (is_sync_mode) {
    if (snb_infer(model_.ctx_, src, dst, 1) != SNB_OK) return -1;
    snb_rt_stat st;
    snb_get_rt_stat(model_.ctx_, &st);
    auto rs = std::make_shared<DnnNodeRtStat>();
    rs->input_fps = st.input_fps; rs->output_fps = st.output_fps;
    rs->infer_time_ms = st.infer_time_ms; rs->fps_updated = st.fps_updated != 0;
    output->rt_stat = rs;
    return PostProcess(output);
  }
  Task* t = new Task{this, output, inputs};
  const int r = snb_infer_async(model_.ctx_, src, dst, 1, &DnnNode::OnDone, t, alloctask_timeout_ms);
  if (r != SNB_OK) { delete t; return -1; }
  return 0;
}

int DnnNode::WaitAll() { return model_.ctx_ && snb_wait_all(model_.ctx_) == SNB_OK ? 0 : -1; }

}  // namespace dnn_node
}  // namespace hobot
