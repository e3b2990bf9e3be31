// stereonet_infer without ROS: feeds side-by-side NV12 frames from a file through StereonetNode and
// writes every published payload ([s32 x H*W] || [JPEG]) to an output file, in arrival order.
// Stands where the reference's main() spins the node (stereonet_infer/src/main.cpp:17-22).
//   stereonet_infer --model_file blob --frames in.nv12 --out out.bin [--model_in_h 720 --model_in_w 1280 --K 4 --D 12]
//                   [--precision tc|fp32] [--encoding nv12] [--repeat N]
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <fstream>
#include <mutex>

#include "stereonet_node.h"

using namespace hobot::stereonet;

int main(int argc, char** argv) {
  Params params;
  std::string frames_path, out_path, encoding = "nv12";
  int repeat = 1;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k.rfind("--", 0) != 0) { fprintf(stderr, "bad argument %s\n", k.c_str()); return 2; }
    if (k == "--frames") frames_path = v;
    else if (k == "--out") out_path = v;
    else if (k == "--encoding") encoding = v;
    else if (k == "--repeat") repeat = atoi(v.c_str());
    else params[k.substr(2)] = v;
  }
  StereonetNode node("stereonet_node", params);
  if (!node.ok()) return 1;
  const int w2 = node.model_input_width() * 2, h = node.model_input_height();
  const size_t frame_bytes = (size_t)h * 3 / 2 * w2;
  std::ifstream f(frames_path, std::ios::binary);
  std::vector<uint8_t> all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (all.empty() || all.size() % frame_bytes) {
    fprintf(stderr, "frames file must hold whole %dx%d NV12 frames (%zu bytes each)\n", w2, h, frame_bytes);
    return 2;
  }
  std::ofstream out(out_path, std::ios::binary);
  std::mutex mu;
  int published = 0;
  node.set_publisher([&](ImageMsg&& m) {
    std::lock_guard<std::mutex> lk(mu);
    const uint32_t hdr[4] = {(uint32_t)atoi(m.header.frame_id.c_str()), m.height, m.width, m.step};
    out.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
    out.write(reinterpret_cast<const char*>(m.data.data()), m.data.size());
    ++published;
  });
  const int n = (int)(all.size() / frame_bytes);
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < repeat; ++r)
    for (int i = 0; i < n; ++i) {
      HbmMsg1080P msg;
      msg.index = r * n + i; msg.height = h; msg.width = w2; msg.encoding = encoding;
      msg.data = all.data() + (size_t)i * frame_bytes; msg.data_size = (uint32_t)frame_bytes;
      node.FeedImg(msg);
    }
  node.WaitAll();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stderr, "fed %d frame(s), published %d, dropped %d\n", n * repeat, published, node.dropped_frames());
  fprintf(stderr, "%d GPU(s), %.3f s, %.1f frames/s end to end (FeedImg -> Run -> PostProcess -> publish callback)\n", node.device_count(),
          secs, n * repeat / secs);
  return published + node.dropped_frames() == n * repeat ? 0 : 1;
}
