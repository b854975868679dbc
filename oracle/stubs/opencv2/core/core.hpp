// Minimal stand-in for the one OpenCV type the reference's regularizer / solver translation units
// use (cv::Size) so they compile UNMODIFIED into oracle/_ref.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_OPENCV_CORE_HPP_
#define ORACLE_STUB_OPENCV_CORE_HPP_
namespace cv {
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  int area() const { return width * height; }
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
}  // namespace cv
#endif
