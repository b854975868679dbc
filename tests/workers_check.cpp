// CPU check of srb::DeviceWorkers (csrc/srb_workers.h): fork/join rounds, hot (spinning) and cold (sleeping)
// helpers, partial rounds, ordering between rounds.  Built and run by tests/test_workers.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../super-resolution_b200/csrc/srb_workers.h"

int main(int argc, char** argv) {
  const int helpers = argc > 1 ? atoi(argv[1]) : 3;
  const int rounds = argc > 2 ? atoi(argv[2]) : 20000;
  srb::DeviceWorkers pool(helpers);
  const int G = pool.capacity();
  std::vector<long long> slot(G, 0), input(G, 0), echo(G, 0);
  long long expect_total = 0, bad = 0;
  for (int r = 0; r < rounds; ++r) {
    if (r % 500 == 0) pool.set_hot((r / 500) % 2 == 0);       // alternate spinning and sleeping helpers
    const int count = 1 + (r % G);                              // partial rounds leave the last helpers idle
    for (int k = 0; k < G; ++k) input[k] = (long long)r * G + k;  // written here before the fork ...
    pool.run(count, [&](int i) {
      long long sum = 0;
      for (int k = 0; k < G; ++k) sum += input[k];            // ... must be what every job of the round reads
      echo[i] = sum;
      slot[i] += i + 1;
    });
    long long want = 0;
    for (int k = 0; k < G; ++k) want += input[k];
    for (int i = 0; i < count; ++i) {                          // and what the jobs wrote is visible after the join
      expect_total += i + 1;
      if (echo[i] != want) ++bad;
    }
    long long total = 0;
    for (int k = 0; k < G; ++k) total += slot[k];
    if (total != expect_total) ++bad;
  }
  pool.set_hot(false);
  pool.run(G, [&](int i) { slot[i] = -1; });
  for (int k = 0; k < G; ++k)
    if (slot[k] != -1) ++bad;
  printf("%s helpers=%d rounds=%d total=%lld\n", bad ? "FAIL" : "OK", helpers, rounds, expect_total);
  return bad ? 1 : 0;
}
