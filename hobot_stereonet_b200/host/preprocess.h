// Host pre-process of the live path, mirroring the reference's class and method names
// (stereonet_infer/include/preprocess.h:28,214-219; src/preprocess.cpp:913-1059).  The developer-only
// file feeders (CvtImgData2Tensors, CvtBinData2Tensors, CvtNV12File2Tensors) are out of scope
// (SURVEY.md §2.1 item 2): they are not called from FeedImg.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "dnn_node.h"

namespace hobot {
namespace stereonet {

using hobot::dnn_node::DNNTensor;
using hobot::dnn_node::Model;

class PreProcess {
 public:
  explicit PreProcess(const std::string& config_file);   // the string is ignored, as in the reference (:35-36)
  // img_l / img_r: one NV12 view each (w*h*3/2 bytes).  Appends one s8 NCHW [1,6,h,w] tensor.
  // Returns 0, or -1 on null arguments ("Invalid input data", :919-922).
  int CvtNV12Data2Tensors(std::vector<std::shared_ptr<DNNTensor>>& input_tensors, Model* pmodel,
                          const unsigned char* img_l, const unsigned char* img_r);
  // false (default): the reference's I420-style chroma indexing of NV12 data (preprocess.h:128-155)
  void set_correct_chroma(bool v) { correct_chroma_ = v; }

 private:
  bool correct_chroma_ = false;
};

}  // namespace stereonet
}  // namespace hobot
