// Minimal stand-in for rclcpp, ONLY to compile-check hobot_stereonet_b200/ros2/src/stereonet_ros_node.cpp in an image
// without ROS 2 (tests/test_host_node.py).  Signatures follow the rclcpp API the shim uses; nothing here runs.
#pragma once
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
namespace rclcpp {
struct Logger {};
template <typename M> struct Publisher {
  using SharedPtr = std::shared_ptr<Publisher<M>>;
  void publish(std::unique_ptr<M> msg) { (void)msg; }
};
template <typename M> struct Subscription { using SharedPtr = std::shared_ptr<Subscription<M>>; };
class Node {
 public:
  explicit Node(const std::string& name) { (void)name; }
  virtual ~Node() = default;
  template <typename T> T declare_parameter(const std::string& name, const T& def) { (void)name; return def; }
  Logger get_logger() const { return Logger(); }
  template <typename M> typename Publisher<M>::SharedPtr create_publisher(const std::string& topic, int qos) {
    (void)topic; (void)qos; return std::make_shared<Publisher<M>>();
  }
  template <typename M, typename F> typename Subscription<M>::SharedPtr create_subscription(const std::string& topic, int qos, F&& cb) {
    (void)topic; (void)qos;
    std::function<void(typename M::ConstSharedPtr)> f = cb; (void)f;
    return std::make_shared<Subscription<M>>();
  }
};
inline void init(int, char**) {}
inline void spin(std::shared_ptr<Node>) {}
inline void shutdown() {}
}  // namespace rclcpp
#define RCLCPP_ERROR(logger, ...) do { (void)(logger); std::fprintf(stderr, __VA_ARGS__); } while (0)
