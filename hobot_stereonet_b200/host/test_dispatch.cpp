// CPU unit test of the multi-GPU dispatch arithmetic (host/dispatch.h), no CUDA, no torch:
//   lib/test_dispatch            exits 0 and prints "dispatch ok" when every property holds.
// Run by tests/test_dispatch.py (not gpu) next to the Python twin's gloo tests (tests/test_shard_gloo.py).
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "dispatch.h"

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) { fprintf(stderr, "%s:%d: CHECK(%s) failed\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main() {
  using snb::Dispatcher;
  using snb::shard_range;
  // shard_range: contiguous, ordered, covers [0, n) exactly once, sizes differ by at most one, remainder to the lowest ranks
  for (int world = 1; world <= 9; ++world)
    for (int64_t n = 0; n <= 70; ++n) {
      int64_t expect = 0, mn = 1 << 30, mx = 0;
      for (int r = 0; r < world; ++r) {
        int64_t a, b;
        CHECK(shard_range(n, world, r, &a, &b));
        CHECK(a == expect && b >= a);
        expect = b;
        mn = b - a < mn ? b - a : mn; mx = b - a > mx ? b - a : mx;
        if (r > 0) { int64_t pa, pb; shard_range(n, world, r - 1, &pa, &pb); CHECK(pb - pa >= b - a); }
      }
      CHECK(expect == n && mx - mn <= 1);
    }
  { int64_t a, b; CHECK(shard_range(32, 8, 3, &a, &b) && a == 12 && b == 16);      // BASELINE config 4: 4 pairs per GPU
    CHECK(shard_range(64, 8, 7, &a, &b) && a == 56 && b == 64);                    // config 5: 8 per GPU
    CHECK(shard_range(10, 4, 0, &a, &b) && b - a == 3); CHECK(shard_range(10, 4, 3, &a, &b) && b - a == 2);
    CHECK(!shard_range(4, 0, 0, &a, &b) && !shard_range(4, 2, 2, &a, &b) && !shard_range(-1, 2, 0, &a, &b)); }

  // Dispatcher, nothing ever completes: plain round-robin, loads differ by at most one at every moment
  for (int n = 1; n <= 8; ++n) {
    Dispatcher d(n);
    for (int i = 0; i < 5 * n + 3; ++i) {
      CHECK(d.pick() == i % n);
      int mn = 1 << 30, mx = 0;
      for (int r = 0; r < n; ++r) { mn = d.inflight(r) < mn ? d.inflight(r) : mn; mx = d.inflight(r) > mx ? d.inflight(r) : mx; }
      CHECK(mx - mn <= 1);
    }
    CHECK(d.total_inflight() == 5 * n + 3);
  }
  // a slow replica gets fewer calls: replica 0 never completes, the others complete immediately
  {
    Dispatcher d(4);
    std::vector<int> served(4, 0);
    for (int i = 0; i < 400; ++i) {
      const int r = d.pick();
      ++served[r];
      if (r != 0) d.done(r);
    }
    CHECK(served[0] == 1 && served[1] + served[2] + served[3] == 399);
    CHECK(abs(served[1] - served[2]) <= 1 && abs(served[2] - served[3]) <= 1);
  }
  // steady state with `depth` calls in flight per replica (task_num = 4): every completion is followed by a call to that replica
  {
    Dispatcher d(8);
    std::vector<int> order;
    for (int i = 0; i < 32; ++i) order.push_back(d.pick());
    for (int r = 0; r < 8; ++r) CHECK(d.inflight(r) == 4);
    for (int i = 0; i < 1000; ++i) {
      const int fin = order[i];
      d.done(fin);
      const int r = d.pick();
      CHECK(r == fin);
      order.push_back(r);
    }
  }
  // done() on an idle or bad replica is harmless
  { Dispatcher d(2); d.done(0); d.done(5); d.done(-1); CHECK(d.total_inflight() == 0); }
  printf("dispatch ok\n");
  return 0;
}
