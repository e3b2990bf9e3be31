// The layer table of the network (names, shapes): one definition for the weight synthesiser (weights_host.cpp) and
// the weight-blob validation (net.cu); tests/test_capi_host.py checks it against oracle/arch.py name by name.
#pragma once
#include <string>
#include <vector>

namespace snb {

struct Spec { std::string name; int cout, cin, kd, ks; float gain; };   // kd = 0: 2-D convolution
std::vector<Spec> conv_specs(int K);

}  // namespace snb
