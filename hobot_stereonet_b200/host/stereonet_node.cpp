#include "stereonet_node.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <chrono>
#include <thread>

namespace hobot {
namespace stereonet {

static int now_ms_since(const std::chrono::steady_clock::time_point& t0) {
  return (int)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
}

static std::string param(const Params& p, const char* key, const std::string& def) {
  auto it = p.find(key);
  return it == p.end() ? def : it->second;
}

StereonetNode::StereonetNode(const std::string& node_name, const Params& params)
    : hobot::dnn_node::DnnNode(node_name), params_(params) {
  config_file_ = param(params, "config_file", config_file_);          // declared, never read (as the reference)
  model_file_ = param(params, "model_file", model_file_);
  sub_hbmem_topic_name_ = param(params, "sub_hbmem_topic_name", sub_hbmem_topic_name_);
  ros_img_topic_name_ = param(params, "ros_img_topic_name", ros_img_topic_name_);
  fprintf(stderr, "[stereonet_node]\n config_file: %s\n model_file: %s\n sub_hbmem_topic_name: %s\n ros_img_topic_name: %s\n",
          config_file_.c_str(), model_file_.c_str(), sub_hbmem_topic_name_.c_str(), ros_img_topic_name_.c_str());
  if (Init() != 0 || GetModelInputSize(0, model_input_width_, model_input_height_) < 0) {
    fprintf(stderr, "[stereonet_node] Node init fail! %s\n", LastError().c_str());
    return;
  }
  model_ = GetModel();
  if (!model_) {
    fprintf(stderr, "[stereonet_node] Invalid model\n");
    return;
  }
  sp_preprocess_ = std::make_shared<PreProcess>("");
  gpu_preprocess_ = param(params, "preprocess", "gpu") != "cpu" && model_input_width_ % 2 == 0 && model_input_height_ % 2 == 0;
  jpeg_on_ = param(params, "jpeg", "on") != "off";
  {
    const int hw = (int)std::thread::hardware_concurrency();
    const int def = hw > 8 ? hw / 2 : 4;
    jpeg_threads_ = atoi(param(params, "jpeg_threads", std::to_string(def)).c_str());
    if (jpeg_threads_ < 1) jpeg_threads_ = 1;
  }
  // the reference's two steps (cv::cvtColor NV12 -> BGR, cv::imencode ".jpg"; stereonet_node.cpp:775-782) restated in the library
  jpeg_encoder_ = [](const uint8_t* nv12, int w, int h, std::vector<uint8_t>& jpeg) {
    jpeg.resize((size_t)w * h * 3 + 4096);                 // one pass: more than any baseline stream of this picture can take
    int64_t n = snb_jpeg_encode_nv12(nv12, w, h, 0, jpeg.data(), jpeg.size());
    if (n > (int64_t)jpeg.size()) {                        // (it reports the size it needs, nothing was overrun)
      jpeg.resize((size_t)n);
      n = snb_jpeg_encode_nv12(nv12, w, h, 0, jpeg.data(), jpeg.size());
    }
    if (n <= 0) return false;
    jpeg.resize((size_t)n);
    return true;
  };
  ok_ = true;
}

// In-flight calls complete into PostProcess, which uses this object's publisher and members: drain them while the
// object is still whole (the base destructor would be too late).
StereonetNode::~StereonetNode() { Shutdown(); }

int StereonetNode::SetNodePara() {
  if (!dnn_node_para_ptr_) return -1;
  if (access(model_file_.c_str(), F_OK) != 0) {
    fprintf(stderr, "[hobot_stereonet] File is not exist! model_file: %s\n", model_file_.c_str());
    return -1;
  }
  dnn_node_para_ptr_->model_file = model_file_;
  dnn_node_para_ptr_->model_task_type = hobot::dnn_node::ModelTaskType::ModelInferType;
  dnn_node_para_ptr_->task_num = 4;
  dnn_node_para_ptr_->model_in_h = atoi(param(params_, "model_in_h", "720").c_str());
  dnn_node_para_ptr_->model_in_w = atoi(param(params_, "model_in_w", "1280").c_str());
  dnn_node_para_ptr_->K = atoi(param(params_, "K", "4").c_str());
  dnn_node_para_ptr_->D = atoi(param(params_, "D", "12").c_str());
  dnn_node_para_ptr_->device = atoi(param(params_, "device", "0").c_str());
  {
    const std::string devs = param(params_, "devices", "");
    size_t pos = 0;
    while (pos < devs.size()) {
      size_t end = devs.find(',', pos);
      if (end == std::string::npos) end = devs.size();
      if (end > pos) dnn_node_para_ptr_->devices.push_back(atoi(devs.substr(pos, end - pos).c_str()));
      pos = end + 1;
    }
  }
  dnn_node_para_ptr_->precision = param(params_, "precision", "tc") == "fp32" ? SNB_PREC_FP32 : SNB_PREC_TC_F16X2;
  return 0;
}

void StereonetNode::FeedImg(const HbmMsg1080P& img_msg) {
  if (!ok_ || !img_msg.data) return;
  // 1. frame gating, unchanged (stereonet_node.cpp:672-690): bad frames are dropped with an error log
  if (img_msg.encoding != "nv12") {
    fprintf(stderr, "[stereonet_node] Only support nv12 img encoding! got %s\n", img_msg.encoding.c_str());
    ++dropped_;
    return;
  }
  if (img_msg.height != (uint32_t)model_input_height_ || img_msg.width != (uint32_t)model_input_width_ * 2) {
    fprintf(stderr, "[stereonet_node] recved img msg h: %u, w: %u is unmatch with model_input_width: %d, model_input_height: %d\n",
            img_msg.height, img_msg.width, model_input_width_, model_input_height_);
    ++dropped_;
    return;
  }
  const int w = img_msg.width / 2, h = img_msg.height;
  if (img_msg.data_size && img_msg.data_size < (uint32_t)(h * 3 / 2) * img_msg.width) { ++dropped_; return; }

  // 2. output object carrying the header through the asynchronous call (:693-696)
  auto dnn_output = std::make_shared<StereonetNodeOutput>();
  dnn_output->msg_header = std::make_shared<hobot::dnn_node::MsgHeader>();
  dnn_output->msg_header->frame_id = std::to_string(img_msg.index);
  dnn_output->msg_header->stamp_sec = img_msg.time_stamp_sec;
  dnn_output->msg_header->stamp_nanosec = img_msg.time_stamp_nanosec;

  // 3. pre-process.  Reference: L/R split (:702-738) then CvtNV12Data2Tensors, on the host.  Default here: the raw frame
  //    goes into a (recycled) page-locked tensor and the same bytes are produced on the GPU inside the pass.
  const auto tp_start = std::chrono::steady_clock::now();
  std::vector<std::shared_ptr<DNNTensor>> input_tensors;
  const size_t frame_bytes = (size_t)(h * 3 / 2) * img_msg.width;
  if (gpu_preprocess_) {
    hobot::dnn_node::hbDNNTensorProperties props;
    model_->GetInputTensorProperties(props, 0);
    auto frame = hobot::dnn_node::AllocTensor(props, (uint32_t)frame_bytes);
    if (!frame) { ++dropped_; return; }
    memcpy(frame->sysMem[0].virAddr, img_msg.data, frame_bytes);
    input_tensors.emplace_back(frame);
  } else {
    std::vector<uint8_t> left_buf((size_t)w * h * 3 / 2), right_buf((size_t)w * h * 3 / 2);
    if (snb_pre_split_nv12(img_msg.data, h, img_msg.width, left_buf.data(), right_buf.data()) != SNB_OK) { ++dropped_; return; }
    if (sp_preprocess_->CvtNV12Data2Tensors(input_tensors, model_, left_buf.data(), right_buf.data()) < 0) {
      fprintf(stderr, "[stereonet_node] Preprocess fail\n");
      ok_ = false;                       // the reference shuts the node down here (:741-744)
      return;
    }
  }
  if (enable_pub_output_) {
    auto bin = std::make_shared<BinDataType>();
    bin->w = w; bin->h = h;            // the reference leaves the 1280x720 defaults (stereonet_node.h:43-44)
    if (jpeg_on_ && jpeg_encoder_) {
      // left view of the side-by-side frame (:752-766), then NV12 -> BGR -> JPEG (:775-782).  The encode does not feed
      // the model, so it runs beside the GPU pass (jpeg_threads_ at once; beyond that the feeding thread does it itself).
      auto left = std::make_shared<std::vector<uint8_t>>((size_t)w * h * 3 / 2);
      for (int r = 0; r < h * 3 / 2; ++r) memcpy(left->data() + (size_t)r * w, img_msg.data + (size_t)r * img_msg.width, w);
      JpegEncoder enc = jpeg_encoder_;
      auto job = [this, enc, left, w, h]() {
        std::vector<uint8_t> jpeg;
        if (!enc(left->data(), w, h, jpeg)) jpeg.clear();
        --jpeg_inflight_;
        return jpeg;
      };
      const bool spawn = ++jpeg_inflight_ <= jpeg_threads_;
      bin->jpeg_future = std::async(spawn ? std::launch::async : std::launch::deferred, job).share();
      if (!spawn) bin->jpeg_future.wait();
    }
    dnn_output->sp_left_nv12 = bin;
  }
  dnn_output->preprocess_time_ms = now_ms_since(tp_start);

  // 4. the hot-path entry (:812): asynchronous, PostProcess fires on the runtime's thread
  if ((gpu_preprocess_ ? RunNv12(input_tensors, dnn_output, false, -1) : Run(input_tensors, dnn_output, false, -1, -1)) < 0) {
    fprintf(stderr, "[stereonet_node] Run infer fail! %s\n", LastError().c_str());
    return;
  }
}

int StereonetNode::PostProcess(const std::shared_ptr<hobot::dnn_node::DnnNodeOutput>& node_output) {
  const auto tp_start = std::chrono::steady_clock::now();
  auto out = std::dynamic_pointer_cast<StereonetNodeOutput>(node_output);
  if (!out) {
    fprintf(stderr, "[stereonet_node] Cast dnn node output fail!\n");
    return -1;
  }
  if (enable_pub_output_ && out->sp_left_nv12 && !out->output_tensors.empty()) {
    ImageMsg msg;
    msg.height = out->sp_left_nv12->h;
    msg.width = out->sp_left_nv12->w;
    msg.encoding = "jpeg";
    msg.header = *out->msg_header;
    const auto& mem = out->output_tensors[0]->sysMem[0];
    if (out->sp_left_nv12->jpeg_future.valid()) out->sp_left_nv12->jpeg = out->sp_left_nv12->jpeg_future.get();
    const auto& jpeg = out->sp_left_nv12->jpeg;
    msg.data.resize((size_t)mem.memSize + jpeg.size());
    const int64_t n = snb_post_pack(static_cast<const int32_t*>(mem.virAddr), mem.memSize, jpeg.data(), jpeg.size(),
                                    msg.data.data(), msg.data.size());
    if (n < 0) return -1;
    msg.step = (uint32_t)n;
    if (ros_img_publisher_) ros_img_publisher_(std::move(msg));
  }
  if (node_output->rt_stat && node_output->rt_stat->fps_updated) {
    fprintf(stderr, "[stereonet_node] input fps: %.2f, out fps: %.2f, preprocess time ms: %d, infer time ms: %d, "
            "msg preparation for pub time cost ms: %d\n", node_output->rt_stat->input_fps, node_output->rt_stat->output_fps,
            out->preprocess_time_ms, node_output->rt_stat->infer_time_ms, now_ms_since(tp_start));
  }
  return 0;
}

}  // namespace stereonet
}  // namespace hobot
