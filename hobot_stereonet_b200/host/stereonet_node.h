// The stereonet_infer node with its ROS 2 shell removed: same class, parameters, callbacks and wire
// format as the reference (stereonet_infer/include/stereonet_node.h:40-127, src/stereonet_node.cpp),
// with the BPU model call replaced by the B200 path behind hobot::dnn_node::DnnNode (dnn_node.h).
// A ROS 2 shim forwards the four parameters, feeds HbmMsg1080P messages into FeedImg and publishes
// what the publish callback receives; nothing else of the reference's surface changes.
#pragma once
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "dnn_node.h"
#include "parser.h"
#include "preprocess.h"

namespace hobot {
namespace stereonet {

using hobot::dnn_node::DNNTensor;
using hobot::dnn_node::DnnNodeOutput;

// hbm_img_msgs/msg/HbmMsg1080P: the fields FeedImg reads (stereonet_node.cpp:663-668,707)
struct HbmMsg1080P {
  uint32_t index = 0;
  int32_t time_stamp_sec = 0;
  uint32_t time_stamp_nanosec = 0;
  uint32_t height = 0, width = 0;      // side-by-side frame: width = 2 * model width
  std::string encoding;                // "nv12"
  const uint8_t* data = nullptr;       // height*3/2 rows of `width` bytes
  uint32_t data_size = 0;
};

// sensor_msgs/msg/Image as PostProcess fills it (stereonet_node.cpp:1026-1049)
struct ImageMsg {
  hobot::dnn_node::MsgHeader header;
  uint32_t height = 0, width = 0;
  std::string encoding;                // "jpeg"
  uint32_t step = 0;                   // = data.size()
  std::vector<uint8_t> data;           // [s32 LE x H*W] || [JPEG of the left view]
};

struct BinDataType {                   // stereonet_node.h:40-47
  char* data = nullptr;
  int len = 0;
  int w = 1280;
  int h = 720;
  std::vector<uint8_t> jpeg;
  std::shared_future<std::vector<uint8_t>> jpeg_future;   // the encode runs beside the GPU pass; PostProcess collects it
};

struct StereonetNodeOutput : public hobot::dnn_node::DnnNodeOutput {   // stereonet_node.h:49-59
  float ratio = 1.0;
  std::shared_ptr<BinDataType> sp_left_nv12 = nullptr;
  int preprocess_time_ms = 0;
};

using Params = std::map<std::string, std::string>;
// left view as NV12 (w*h*3/2 bytes) -> JPEG bytes.  The reference uses cv::cvtColor + cv::imencode
// (stereonet_node.cpp:775-782); the default here is the library's own baseline encoder (snb_jpeg_encode_nv12: the same
// colour conversion bit for bit, quality 95, 4:2:0), so the payload's JPEG half is always present and the untouched render
// tool can cv2.imdecode it (publisher_member_function.py:93-95).  set_jpeg_encoder replaces it (e.g. by OpenCV's).
using JpegEncoder = std::function<bool(const uint8_t* nv12, int w, int h, std::vector<uint8_t>& jpeg)>;
using Publisher = std::function<void(ImageMsg&&)>;

class StereonetNode : public hobot::dnn_node::DnnNode {
 public:
  // Parameters (same names and defaults as stereonet_node.cpp:27-35): config_file, model_file,
  // sub_hbmem_topic_name, ros_img_topic_name.  Extra, B200-only: model_in_h, model_in_w, K, D,
  // device, precision ("tc"|"fp32") — the geometry the reference compiles into its .hbm.
  // More B200-only parameters: devices ("0,1,2,3": one replica per GPU, frames go to the least busy one),
  // preprocess ("gpu" default: the raw NV12 frame is uploaded and split/converted on the GPU | "cpu": the reference's host
  // path through PreProcess::CvtNV12Data2Tensors), jpeg ("on" default | "off": empty JPEG part).
  explicit StereonetNode(const std::string& node_name = "stereonet_node", const Params& params = Params());
  ~StereonetNode() override;

  bool ok() const { return ok_; }      // false: "Node init fail!" (the reference calls rclcpp::shutdown())
  void FeedImg(const HbmMsg1080P& img_msg);
  void set_publisher(Publisher p) { ros_img_publisher_ = std::move(p); }
  void set_jpeg_encoder(JpegEncoder e) { jpeg_encoder_ = std::move(e); }
  const std::string& sub_hbmem_topic_name() const { return sub_hbmem_topic_name_; }
  const std::string& ros_img_topic_name() const { return ros_img_topic_name_; }
  int model_input_width() const { return model_input_width_; }
  int model_input_height() const { return model_input_height_; }
  int dropped_frames() const { return dropped_; }

 protected:
  int SetNodePara() override;
  int PostProcess(const std::shared_ptr<hobot::dnn_node::DnnNodeOutput>& node_output) override;

 private:
  hobot::dnn_node::Model* model_ = nullptr;
  int model_input_width_ = -1;
  int model_input_height_ = -1;
  std::string sub_hbmem_topic_name_ = "hbmem_stereo_img";
  std::string ros_img_topic_name_ = "/stereonet_node_output";
  bool enable_pub_output_ = true;
  std::string config_file_ = "config/hobot_stereonet_config.json";
  std::string model_file_ = "config/hobot_stereonet.hbm";
  std::shared_ptr<PreProcess> sp_preprocess_ = nullptr;
  Params params_;
  Publisher ros_img_publisher_;
  JpegEncoder jpeg_encoder_;
  bool ok_ = false;
  bool gpu_preprocess_ = true;
  bool jpeg_on_ = true;
  std::atomic<int> jpeg_inflight_{0};
  int jpeg_threads_ = 4;               // JPEG encodes running beside the GPU passes (parameter jpeg_threads; default: half the host cores)
  int dropped_ = 0;
};

}  // namespace stereonet
}  // namespace hobot
