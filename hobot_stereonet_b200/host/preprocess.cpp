#include "preprocess.h"

#include <stdio.h>

namespace hobot {
namespace stereonet {

PreProcess::PreProcess(const std::string&) {}

int PreProcess::CvtNV12Data2Tensors(std::vector<std::shared_ptr<DNNTensor>>& input_tensors, Model* pmodel,
                                    const unsigned char* img_l, const unsigned char* img_r) {
  if (!pmodel || !img_l || !img_r) {
    fprintf(stderr, "[hobot_stereonet] Invalid input data\n");
    return -1;
  }
  hobot::dnn_node::hbDNNTensorProperties properties;
  if (pmodel->GetInputTensorProperties(properties, 0) != 0) return -1;
  // layout-dependent index lookup, as the reference does (:926-939)
  int h_index = 1, w_index = 2, c_index = 3;
  if (properties.tensorLayout == hobot::dnn_node::HB_DNN_LAYOUT_NCHW) { c_index = 1; h_index = 2; w_index = 3; }
  const int in_h = properties.validShape.dimensionSize[h_index];
  const int in_w = properties.validShape.dimensionSize[w_index];
  const int c_stride = properties.validShape.dimensionSize[c_index];
  if (c_stride != 6) return -1;
  auto dnn_tensor = hobot::dnn_node::AllocTensor(properties, (uint32_t)(in_h * in_w * c_stride));
  if (!dnn_tensor) return -1;
  // one pass: chroma upsample + L|R plane merge + (x-128) quantise (snb200.h cites the lines)
  if (snb_pre_cvt_nv12_to_tensor(img_l, img_r, in_w, in_h, correct_chroma_ ? 1 : 0,
                                 static_cast<int8_t*>(dnn_tensor->sysMem[0].virAddr)) != SNB_OK)
    return -1;
  input_tensors.emplace_back(dnn_tensor);
  return 0;
}

}  // namespace stereonet
}  // namespace hobot
