// Output parsing of the reference (stereonet_infer/include/parser.h:19-40, src/parser.cpp:33-199):
// s32 disparity tensor -> depth in metres.  The reference node keeps this call disabled
// (stereonet_node.cpp:999-1003) and publishes raw s32; it is kept because it defines what the
// integers mean.  The colour-map rendering half of ParseTensor is the render tool's job.
#pragma once
#include <memory>
#include <vector>

#include "dnn_node.h"

namespace hobot {
namespace stereonet {

struct StereonetResult {
  std::vector<float> results;   // depth, metres, H*W (parser.h:22-25)
};

// 0 on success, -1 on failure (parser.h:37-39).
int32_t Parse(const std::shared_ptr<hobot::dnn_node::DnnNodeOutput>& node_output,
              std::vector<std::shared_ptr<StereonetResult>>& results);

}  // namespace stereonet
}  // namespace hobot
