#include "parser.h"

namespace hobot {
namespace stereonet {

int32_t Parse(const std::shared_ptr<hobot::dnn_node::DnnNodeOutput>& node_output,
              std::vector<std::shared_ptr<StereonetResult>>& results) {
  if (!node_output) return -1;
  results.clear();
  for (auto& t : node_output->output_tensors) {
    if (!t || !t->sysMem[0].virAddr || t->properties.scale.scaleLen < 1) return -1;
    const auto& s = t->properties.validShape;      // NCHW: [1,1,H,W] (parser.cpp:169-188)
    const int64_t n = (int64_t)s.dimensionSize[2] * s.dimensionSize[3];
    auto r = std::make_shared<StereonetResult>();
    r->results.resize(n);
    if (snb_post_parse_depth(static_cast<const int32_t*>(t->sysMem[0].virAddr), n, t->properties.scale.scaleData[0],
                             r->results.data()) != SNB_OK)
      return -1;
    results.push_back(r);
  }
  return 0;
}

}  // namespace stereonet
}  // namespace hobot
