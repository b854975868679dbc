// CPU check of srb::plan_row_bands (csrc/srb_row_bands.h): the bands tile the active range in order without gaps
// or overlap, fall on unit boundaries, their halos never leave the channels they touch, and the pulls cover exactly
// the halo outside the band, each piece inside its owner's band.  Built and run by tests/test_row_bands.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../super-resolution_b200/csrc/srb_row_bands.h"

static long long bad = 0;
#define EXPECT(cond)                                                                         \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      if (bad < 10) fprintf(stderr, "FAILED %s (line %d): G=%d Ca=%d H=%d W=%d TH=%d halo=%d r=%d\n", #cond, __LINE__, G, Ca, H, W, TH, halo, r); \
      ++bad;                                                                                 \
    }                                                                                        \
  } while (0)

int main() {
  long long cases = 0;
  const int Hs[] = {1, 31, 32, 33, 64, 100, 288, 2048};
  const int Ws[] = {1, 64, 80};
  for (int G = 1; G <= 8; ++G)
    for (int Ca : {1, 2, 3, 7})
      for (int H : Hs)
        for (int W : Ws)
          for (int TH : {32, 64})
            for (int halo : {0, 3, 7, 16, 40}) {
              ++cases;
              const std::vector<srb::RowBand> b = srb::plan_row_bands(G, Ca, H, W, TH, halo);
              const long long P = (long long)H * W, n = P * Ca;
              const int tr = (H + TH - 1) / TH;
              int r = 0;
              EXPECT((int)b.size() == G);
              long long at = 0;
              int unit = 0;
              for (r = 0; r < G; ++r) {
                EXPECT(b[r].u0 == unit && b[r].u1 >= b[r].u0);
                EXPECT(b[r].begin == at && b[r].end >= b[r].begin);
                unit = b[r].u1;
                at = b[r].end;
                // unit boundaries: a whole number of rows of one channel
                EXPECT(b[r].begin % W == 0 && b[r].end % W == 0);
                if (b[r].end > b[r].begin) {
                  const long long row0 = b[r].begin % P / W;
                  EXPECT(row0 % TH == 0);
                  // balanced: no band holds more than ceil(units / G) units
                  EXPECT(b[r].u1 - b[r].u0 <= (tr * Ca + G - 1) / G);
                  // halo: inside the vector, inside the first / last channel the band touches, at most halo rows
                  EXPECT(b[r].halo_begin <= b[r].begin && b[r].halo_end >= b[r].end);
                  EXPECT(b[r].halo_begin >= b[r].begin / P * P && b[r].halo_end <= (b[r].end + P - 1) / P * P);
                  EXPECT(b[r].begin - b[r].halo_begin <= (long long)halo * W && b[r].halo_end - b[r].end <= (long long)halo * W);
                  // ... and exactly halo rows unless the channel ends first
                  EXPECT(b[r].begin - b[r].halo_begin == std::min((long long)halo * W, b[r].begin - b[r].begin / P * P));
                  EXPECT(b[r].halo_end - b[r].end == std::min((long long)halo * W, (b[r].end + P - 1) / P * P - b[r].end));
                  // pulls: disjoint pieces inside their owners' bands that cover the halo outside the band exactly
                  std::vector<char> covered((size_t)(b[r].halo_end - b[r].halo_begin), 0);
                  for (const srb::RowBandPull& h : b[r].pulls) {
                    EXPECT(h.from >= 0 && h.from < G && h.from != r && h.end > h.begin);
                    EXPECT(h.begin >= b[h.from].begin && h.end <= b[h.from].end);
                    EXPECT(h.begin >= b[r].halo_begin && h.end <= b[r].halo_end);
                    EXPECT(h.end <= b[r].begin || h.begin >= b[r].end);
                    for (long long i = h.begin; i < h.end; ++i) {
                      EXPECT(!covered[(size_t)(i - b[r].halo_begin)]);
                      covered[(size_t)(i - b[r].halo_begin)] = 1;
                    }
                  }
                  for (long long i = b[r].halo_begin; i < b[r].halo_end; ++i) {
                    const bool inside = i >= b[r].begin && i < b[r].end;
                    EXPECT(inside ? !covered[(size_t)(i - b[r].halo_begin)] : covered[(size_t)(i - b[r].halo_begin)]);
                  }
                } else {
                  EXPECT(b[r].pulls.empty());
                }
              }
              r = G;
              EXPECT(unit == tr * Ca && at == n);
            }
  printf("%s cases=%lld\n", bad ? "FAIL" : "OK", cases);
  return bad ? 1 : 0;
}
