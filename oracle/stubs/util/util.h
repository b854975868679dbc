// Stand-in for src/util/util.h: the reference's util.cpp drags in gflags/glog/dirent, but the
// regularizers only need GetPixelIndex (util.cpp:81-89), restated here.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_UTIL_UTIL_H_
#define ORACLE_STUB_UTIL_UTIL_H_
#include "opencv2/core/core.hpp"
namespace super_resolution {
namespace util {
inline int GetPixelIndex(const cv::Size& image_size, const int channel, const int row,
                         const int col) {
  const int channel_index = channel * (image_size.width * image_size.height);
  return channel_index + (row * image_size.width + col);
}
}  // namespace util
}  // namespace super_resolution
#endif
