// ROS 2 shell of the B200 stereonet node: the same parameters, topics and message types as the
// reference's stereonet_infer node (stereonet_infer/src/stereonet_node.cpp:24-127, src/main.cpp:17-22), around
// hobot::stereonet::StereonetNode (hobot_stereonet_b200/host/stereonet_node.h), which carries everything else.
//   parameters   config_file, model_file, sub_hbmem_topic_name, ros_img_topic_name          (stereonet_node.cpp:27-35)
//                + B200-only, optional: model_in_h, model_in_w, K, D, device, precision       (the geometry an .hbm compiles in)
//   subscribes   hbm_img_msgs/msg/HbmMsg1080P on sub_hbmem_topic_name, side-by-side NV12      (stereonet_node.cpp:108-112)
//   publishes    sensor_msgs/msg/Image on ros_img_topic_name: [s32 LE x H*W] || [JPEG], "jpeg" (stereonet_node.cpp:116-118,1026-1064)
// Built only inside a ROS 2 workspace (ros2/CMakeLists.txt); this image has no rclcpp, so tests/test_host_node.py
// compile-checks this file against minimal stand-in headers (tests/ros_stubs) instead.
#include <map>
#include <memory>
#include <string>
#include <utility>

#include "hbm_img_msgs/msg/hbm_msg1080_p.hpp"
#include "rclcpp/rclcpp.hpp"
#include "sensor_msgs/msg/image.hpp"
#include "stereonet_node.h"

namespace {

class StereonetRos : public rclcpp::Node {
 public:
  StereonetRos() : rclcpp::Node("stereonet_node") {
    hobot::stereonet::Params params;
    for (const auto& kv : defaults_) params[kv.first] = this->declare_parameter<std::string>(kv.first, kv.second);
    for (const char* k : {"model_in_h", "model_in_w", "K", "D", "device", "devices", "precision", "preprocess", "jpeg"}) {
      const std::string v = this->declare_parameter<std::string>(k, "");
      if (!v.empty()) params[k] = v;
    }
    node_ = std::make_unique<hobot::stereonet::StereonetNode>("stereonet_node", params);
    if (!node_->ok()) {                                                   // stereonet_node.cpp:44-48
      RCLCPP_ERROR(this->get_logger(), "Node init fail!");
      rclcpp::shutdown();
      return;
    }
    pub_ = this->create_publisher<sensor_msgs::msg::Image>(node_->ros_img_topic_name(), 10);
    node_->set_publisher([this](hobot::stereonet::ImageMsg&& m) {         // called on the runtime thread (PostProcess)
      auto msg = std::make_unique<sensor_msgs::msg::Image>();
      msg->header.frame_id = m.header.frame_id;
      msg->header.stamp.sec = m.header.stamp_sec;
      msg->header.stamp.nanosec = m.header.stamp_nanosec;
      msg->height = m.height;
      msg->width = m.width;
      msg->encoding = m.encoding;
      msg->step = m.step;
      msg->data = std::move(m.data);
      pub_->publish(std::move(msg));                                      // stereonet_node.cpp:1064
    });
    sub_ = this->create_subscription<hbm_img_msgs::msg::HbmMsg1080P>(
        node_->sub_hbmem_topic_name(), 10, [this](hbm_img_msgs::msg::HbmMsg1080P::ConstSharedPtr in) {
          hobot::stereonet::HbmMsg1080P m;
          m.index = in->index;
          m.height = in->height;
          m.width = in->width;
          m.time_stamp_sec = in->time_stamp.sec;
          m.time_stamp_nanosec = in->time_stamp.nanosec;
          m.encoding = std::string(reinterpret_cast<const char*>(in->encoding.data()));
          m.data = in->data.data();
          m.data_size = in->data_size;
          node_->FeedImg(m);                                              // stereonet_node.cpp:657
        });
  }

 private:
  const std::map<std::string, std::string> defaults_{{"config_file", "config/hobot_stereonet_config.json"},
                                                     {"model_file", "config/hobot_stereonet.snb"},
                                                     {"sub_hbmem_topic_name", "hbmem_stereo_img"},
                                                     {"ros_img_topic_name", "/stereonet_node_output"}};
  rclcpp::Publisher<sensor_msgs::msg::Image>::SharedPtr pub_;
  rclcpp::Subscription<hbm_img_msgs::msg::HbmMsg1080P>::SharedPtr sub_;
  // declared last = destroyed first: ~StereonetNode drains the in-flight calls, whose PostProcess publishes through pub_
  std::unique_ptr<hobot::stereonet::StereonetNode> node_;
};

}  // namespace

int main(int argc, char** argv) {                                          // main.cpp:17-22
  rclcpp::init(argc, argv);
  rclcpp::spin(std::make_shared<StereonetRos>());
  rclcpp::shutdown();
  return 0;
}
