// Host-side dispatch arithmetic of the multi-GPU front: stereo pairs are independent units, so one
// node process drives N replicas of the model, one per GPU.  Two policies, both pure functions of integers (unit-tested
// on the CPU by lib/test_dispatch, built from host/test_dispatch.cpp):
//   * shard_range  - a batch of n pairs in ONE call is cut into contiguous chunks, remainder to the lowest ranks
//                    (config 4: 32 pairs over 8 GPUs -> 4 each; hobot_stereonet_b200/shard.py is the Python twin);
//   * Dispatcher   - a stream of one-pair calls (the reference's Run() per camera frame, stereonet_node.cpp:812) goes to
//                    the replica with the fewest calls in flight, round-robin among equals.
// The reference has ONE device and no dispatch at all; this replaces "the BPU" by "the least busy B200".
#pragma once
#include <stdint.h>

#include <vector>

namespace snb {

// [start, stop) of `n_pairs` owned by `rank` of `world`; false on bad arguments.
inline bool shard_range(int64_t n_pairs, int world, int rank, int64_t* start, int64_t* stop) {
  if (world < 1 || rank < 0 || rank >= world || n_pairs < 0 || !start || !stop) return false;
  const int64_t base = n_pairs / world, rem = n_pairs % world;
  *start = rank * base + (rank < rem ? rank : rem);
  *stop = *start + base + (rank < rem ? 1 : 0);
  return true;
}

class Dispatcher {
 public:
  explicit Dispatcher(int n_replicas) : inflight_(n_replicas > 0 ? n_replicas : 1, 0) {}
  int size() const { return (int)inflight_.size(); }
  // replica for the next call: fewest calls in flight; among equals the one after the last choice (round-robin)
  int pick() {
    const int n = size();
    int best = -1;
    for (int k = 1; k <= n; ++k) {
      const int i = (last_ + k) % n;
      if (best < 0 || inflight_[i] < inflight_[best]) best = i;
    }
    last_ = best;
    ++inflight_[best];
    return best;
  }
  void done(int replica) { if (replica >= 0 && replica < size() && inflight_[replica] > 0) --inflight_[replica]; }
  int inflight(int replica) const { return inflight_[replica]; }
  int total_inflight() const { int s = 0; for (int v : inflight_) s += v; return s; }

 private:
  std::vector<int> inflight_;
  int last_ = -1;
};

}  // namespace snb
