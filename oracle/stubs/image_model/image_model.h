// Stand-in for src/image_model/image_model.h (the real one needs OpenCV imgproc): carries the
// oracle's model description; the solver files only call GetDownsamplingScale().
// TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_IMAGE_MODEL_IMAGE_MODEL_H_
#define ORACLE_STUB_IMAGE_MODEL_IMAGE_MODEL_H_
#include <vector>
#include "image/image_data.h"
#include "sr_oracle.h"
namespace super_resolution {
class ImageModel {
 public:
  explicit ImageModel(const int downsampling_scale) : downsampling_scale_(downsampling_scale) {}
  int GetDownsamplingScale() const { return downsampling_scale_; }
  // oracle-side payload
  std::vector<double> psf;     // K*K or empty
  int psf_size = 0;
  std::vector<double> shifts;  // 2*N or empty
  int num_frames = 0;
  sro_model AsOracleModel() const {
    sro_model m;
    m.scale = downsampling_scale_;
    m.psf_size = psf_size;
    m.psf = psf.empty() ? nullptr : psf.data();
    m.num_frames = num_frames;
    m.shifts = shifts.empty() ? nullptr : shifts.data();
    return m;
  }
 private:
  const int downsampling_scale_;
};
}  // namespace super_resolution
#endif
