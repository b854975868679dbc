// srb_geometry.h -- pure host geometry of the image formation model: cv::resize's nearest index map
// and cv::warpAffine's fixed-point translation.  No CUDA; shared by the C ABI (srb_api.cu), the tile
// planner (srb_tile_plan.cuh) and, through srb_plan(), the CPU tests.
#pragma once
#include <cmath>

#include "srb_common.cuh"

namespace srb {

// ---- geometry (host) ---------------------------------------------------------------------------
// cv::resize INTER_NEAREST index map (reference call: image_data.cpp:341-347): bit-exact fp64.
inline int nearest_index(int q, int n_src, int n_dst) {
  const double inv_scale = (double)n_dst / (double)n_src;
  const double ifx = 1.0 / inv_scale;
  int s = (int)std::floor(q * ifx);
  if (s > n_src - 1) s = n_src - 1;
  return s;
}
// ImageData::ResizeImage(scale factor) output size (image_data.cpp:353-364).
inline void lr_size(int s, int H, int W, int* h, int* w) {
  const double f = 1.0 / (double)s;
  *w = (int)(W * f);
  *h = (int)(H * f);
}
// Fixed-point translation of cv::warpAffine for the matrix [1 0 dx; 0 1 dy] (motion_module.cpp:
// 18-24): AB_BITS = 10, round_delta = 16, INTER_BITS = 5.
inline WarpQ quantize_warp(double dx, double dy, int H, int* rowY) {
  WarpQ q;
  const double m2 = -dx, m5 = -dy;
  const long X0 = std::lrint(m2 * 1024.0) + 16;
  q.nX = (int)(X0 >> 5);
  q.uniform = true;
  q.nY = 0;
  for (int y = 0; y < H; ++y) {
    const long Y0 = std::lrint((1.0 * y + m5) * 1024.0) + 16;
    const int Y = (int)(Y0 >> 5);
    if (y == 0) q.nY = Y;
    if (Y != 32 * y + q.nY) q.uniform = false;
    if (rowY) rowY[y] = Y;
  }
  return q;
}


}  // namespace srb
