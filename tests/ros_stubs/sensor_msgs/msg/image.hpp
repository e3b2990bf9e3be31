// Stand-in for sensor_msgs/msg/Image.  Compile check only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace sensor_msgs { namespace msg {
struct Stamp { int32_t sec = 0; uint32_t nanosec = 0; };
struct Header { Stamp stamp; std::string frame_id; };
struct Image {
  Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  uint32_t step = 0;
  std::vector<uint8_t> data;
};
}}  // namespace sensor_msgs::msg
