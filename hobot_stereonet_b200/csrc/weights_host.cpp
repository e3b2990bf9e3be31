// Layer table of the network and a seeded synthetic weight generator ("SNB2WGT1" blobs).
//
// The reference's float weights cannot be recovered from its BPU binary, so a
// deployment of this build is driven by a weight blob named by `model_file`.  This generator makes
// such a blob without any Python dependency (He-normal weights, small biases, splitmix64 +
// Box-Muller); tests/test_capi_host.py checks its table against oracle/arch.py name by name.
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/snb200.h"
#include "layers.h"

namespace snb {

std::vector<Spec> conv_specs(int K) {
  std::vector<Spec> v;
  v.push_back({"backbone.firstconv.0", 32, 3, 0, 3, 1.f});
  v.push_back({"backbone.firstconv.1", 32, 32, 0, 3, 1.f});
  v.push_back({"backbone.firstconv.2", 32, 32, 0, 3, 1.f});
  const int blocks[4] = {3, 16, 3, 3}, ch[4] = {32, 64, 128, 128};
  int cin = 32;
  for (int li = 1; li <= 4; ++li) {
    for (int b = 0; b < blocks[li - 1]; ++b) {
      const std::string p = "backbone.layer" + std::to_string(li) + "." + std::to_string(b);
      v.push_back({p + ".conv_a", ch[li - 1], b == 0 ? cin : ch[li - 1], 0, 3, 1.f});
      v.push_back({p + ".conv_b", ch[li - 1], ch[li - 1], 0, 3, 0.35f});
      if (b == 0 && li <= 3) v.push_back({p + ".downsample", ch[li - 1], cin, 0, 1, 0.7f});
    }
    cin = ch[li - 1];
  }
  v.push_back({"backbone.lastconv.0", 128, 256, 0, 3, 1.f});
  v.push_back({"backbone.lastconv.1", 16, 128, 0, 1, 0.7f});
  v.push_back({"head.filter.0", 32, 64, 3, 3, 1.f});
  for (int i = 1; i < 5; ++i) v.push_back({"head.filter." + std::to_string(i), 32, 32, 3, 3, 1.f});
  v.push_back({"head.conv3d_alone", 1, 32, 3, 3, 6.f});
  for (int s = 0; s < K; ++s) {
    const std::string p = "head.refine." + std::to_string(s);
    v.push_back({p + ".conv_in", 32, 4, 0, 3, 1.f});
    for (int b = 0; b < 6; ++b) {
      v.push_back({p + ".blocks." + std::to_string(b) + ".conv_a", 32, 32, 0, 3, 1.f});
      v.push_back({p + ".blocks." + std::to_string(b) + ".conv_b", 32, 32, 0, 3, 0.35f});
    }
    v.push_back({p + ".conv_out", 1, 32, 0, 3, 0.02f});
  }
  return v;
}

}  // namespace snb

namespace {

using snb::Spec;
using snb::conv_specs;

struct Rng {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  double uniform() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double normal() { return sqrt(-2.0 * log(uniform())) * cos(6.283185307179586 * uniform()); }
};

}  // namespace

extern "C" int64_t snb_weights_synthesize(int32_t K, uint64_t seed, void* dst, uint64_t cap) {
  if (K < 1 || K > 5) return SNB_ERR_INVALID;
  const std::vector<Spec> specs = conv_specs(K);
  const size_t n = specs.size() * 2;
  const size_t head = (24 + n * 104 + 63) / 64 * 64;
  size_t data = 0;
  std::vector<size_t> offs;
  for (const Spec& s : specs) {
    const size_t wn = (size_t)s.cout * s.cin * (s.kd ? s.kd : 1) * s.ks * s.ks;
    offs.push_back(data); data += (wn * 4 + 63) / 64 * 64;
    offs.push_back(data); data += ((size_t)s.cout * 4 + 63) / 64 * 64;
  }
  const size_t total = head + data;
  if (!dst) return (int64_t)total;
  if (cap < total) return SNB_ERR_NOMEM;
  uint8_t* p = static_cast<uint8_t*>(dst);
  memset(p, 0, total);
  memcpy(p, "SNB2WGT1", 8);
  const uint32_t hdr[4] = {1u, (uint32_t)K, (uint32_t)n, 0u};
  memcpy(p + 8, hdr, 16);
  Rng rng{seed * 0x2545f4914f6cdd1dull + 0x1234567ull};
  for (size_t i = 0; i < specs.size(); ++i) {
    const Spec& s = specs[i];
    const int kd = s.kd ? s.kd : 1;
    const size_t wn = (size_t)s.cout * s.cin * kd * s.ks * s.ks;
    const float stdv = s.gain * sqrtf(2.0f / (float)(s.cin * kd * s.ks * s.ks));
    const float bstd = 0.05f * (s.gain < 1.f ? s.gain : 1.f);
    for (int t = 0; t < 2; ++t) {
      uint8_t* e = p + 24 + (2 * i + t) * 104;
      const std::string nm = s.name + (t == 0 ? ".weight" : ".bias");
      memcpy(e, nm.c_str(), nm.size());
      uint32_t ndim, dims[5] = {1, 1, 1, 1, 1};
      if (t == 1) { ndim = 1; dims[0] = s.cout; }
      else if (s.kd) { ndim = 5; dims[0] = s.cout; dims[1] = s.cin; dims[2] = s.kd; dims[3] = s.ks; dims[4] = s.ks; }
      else { ndim = 4; dims[0] = s.cout; dims[1] = s.cin; dims[2] = s.ks; dims[3] = s.ks; }
      const uint64_t off = offs[2 * i + t], nb = (t == 0 ? wn : (size_t)s.cout) * 4;
      memcpy(e + 64, &ndim, 4); memcpy(e + 68, dims, 20); memcpy(e + 88, &off, 8); memcpy(e + 96, &nb, 8);
      float* d = reinterpret_cast<float*>(p + head + off);
      const size_t cnt = nb / 4;
      for (size_t k = 0; k < cnt; ++k) d[k] = (float)rng.normal() * (t == 0 ? stdv : bstd);
    }
  }
  return (int64_t)total;
}
