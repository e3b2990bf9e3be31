// Stand-in for hbm_img_msgs/msg/HbmMsg1080P (fields the reference reads: stereonet_node.cpp:663-668,707).  Compile check only.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
namespace hbm_img_msgs { namespace msg {
struct Time { int32_t sec = 0; uint32_t nanosec = 0; };
struct HbmMsg1080P {
  using ConstSharedPtr = std::shared_ptr<const HbmMsg1080P>;
  uint32_t index = 0;
  Time time_stamp;
  uint32_t height = 0, width = 0, data_size = 0;
  std::array<uint8_t, 12> encoding{};
  std::array<uint8_t, 16> data{};      // 1920*1080*3 in the real message
};
}}  // namespace hbm_img_msgs::msg
