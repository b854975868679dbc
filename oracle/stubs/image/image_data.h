// Stand-in for src/image/image_data.h (whose real implementation needs OpenCV imgproc): a planar
// fp64 container exposing exactly the members map_solver.cpp / irls_map_solver.cpp use, so those
// reference files compile UNMODIFIED into oracle/_ref.  Resizing goes through the oracle's
// restatement of cv::resize INTER_NEAREST.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_IMAGE_IMAGE_DATA_H_
#define ORACLE_STUB_IMAGE_IMAGE_DATA_H_
#include <cstdlib>
#include <vector>
#include "opencv2/core/core.hpp"
#include "sr_oracle.h"
namespace super_resolution {
enum ResizeInterpolationMethod {
  INTERPOLATE_NEAREST, INTERPOLATE_LINEAR, INTERPOLATE_CUBIC, INTERPOLATE_ADDITIVE
};
class ImageData {
 public:
  ImageData() {}
  ImageData(const double* pixel_values, const cv::Size& size, const int num_channels)
      : image_size_(size) {
    const int n = size.width * size.height;
    for (int c = 0; c < num_channels; ++c)
      channels_.push_back(std::vector<double>(pixel_values + c * n, pixel_values + (c + 1) * n));
  }
  void AddChannel(const double* pixel_values, const cv::Size& size) {
    image_size_ = size;
    channels_.push_back(std::vector<double>(pixel_values, pixel_values + size.width * size.height));
  }
  void ResizeImage(const cv::Size& new_size,
                   const ResizeInterpolationMethod method = INTERPOLATE_LINEAR) {
    if (method != INTERPOLATE_NEAREST) std::abort();
    for (auto& ch : channels_) {
      std::vector<double> out((size_t)new_size.width * new_size.height);
      sro_resize_nearest(ch.data(), image_size_.height, image_size_.width, out.data(),
                         new_size.height, new_size.width);
      ch.swap(out);
    }
    image_size_ = new_size;
  }
  int GetNumChannels() const { return (int)channels_.size(); }
  cv::Size GetImageSize() const { return image_size_; }
  int GetNumPixels() const { return image_size_.width * image_size_.height; }
  const double* GetChannelData(const int index) const { return channels_[index].data(); }
 private:
  cv::Size image_size_;
  std::vector<std::vector<double>> channels_;
};
}  // namespace super_resolution
#endif
