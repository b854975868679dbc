// srb_row_bands.h -- the row-band partition of the solver vectors over the devices of a multi-GPU context
// (srb_multi_solver.cuh).  Pure host arithmetic, no CUDA: unit-tested on the CPU (tests/test_row_bands.py).
//
// The active range [Ca][H][W] is cut, in memory order, into units of TH rows of one channel (the last unit of a
// channel may be shorter); units [0, Ca * ceil(H / TH)) are dealt to the G devices in contiguous bands, so a
// device's band is ONE contiguous element range [begin, end) of every vector.  An evaluation of the band reads the
// estimate on the band plus `halo_rows` rows either side, clipped to the channels the band touches (no stencil
// crosses a channel boundary): [halo_begin, halo_end).  What lies outside the band belongs to other devices:
// `pulls` lists it by owner.
#pragma once
#include <algorithm>
#include <vector>

namespace srb {

struct RowBandPull {  // elements [begin, end) are owned by device `from`
  int from;
  long long begin, end;
};
struct RowBand {
  int u0 = 0, u1 = 0;                   // units [u0, u1)
  long long begin = 0, end = 0;         // = elements [begin, end)
  long long halo_begin = 0, halo_end = 0;
  std::vector<RowBandPull> pulls;
};

inline std::vector<RowBand> plan_row_bands(int G, int Ca, int H, int W, int TH, int halo_rows) {
  std::vector<RowBand> bands(G > 0 ? G : 0);
  if (G <= 0 || Ca <= 0 || H <= 0 || W <= 0 || TH <= 0) return bands;
  const int tr = (H + TH - 1) / TH, nu = tr * Ca;
  const long long P = (long long)H * W, n = P * Ca;
  auto first_elem = [&](int u) -> long long {
    if (u >= nu) return n;
    const int ch = u / tr, t = u - ch * tr;
    const int row = t * TH < H ? t * TH : H;
    return (long long)ch * P + (long long)row * W;
  };
  for (int r = 0; r < G; ++r) {
    RowBand& b = bands[r];
    b.u0 = (int)((long long)nu * r / G);
    b.u1 = (int)((long long)nu * (r + 1) / G);
    b.begin = first_elem(b.u0);
    b.end = first_elem(b.u1);
    b.halo_begin = b.begin;
    b.halo_end = b.end;
    if (b.end <= b.begin) continue;
    const long long halo = (long long)(halo_rows > 0 ? halo_rows : 0) * W;
    b.halo_begin = std::max(b.begin - halo, b.begin / P * P);
    b.halo_end = std::min(b.end + halo, (b.end + P - 1) / P * P);
  }
  for (int r = 0; r < G; ++r) {
    RowBand& b = bands[r];
    if (b.end <= b.begin) continue;
    for (int q = 0; q < G; ++q) {
      if (q == r || bands[q].end <= bands[q].begin) continue;
      const long long a0 = std::max(b.halo_begin, bands[q].begin), a1 = std::min(b.begin, bands[q].end);  // before the band
      if (a1 > a0) b.pulls.push_back({q, a0, a1});
      const long long c0 = std::max(b.end, bands[q].begin), c1 = std::min(b.halo_end, bands[q].end);      // after it
      if (c1 > c0) b.pulls.push_back({q, c0, c1});
    }
  }
  return bands;
}

}  // namespace srb
