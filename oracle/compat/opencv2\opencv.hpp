// oracle/compat/opencv2\opencv.hpp — TEST INFRASTRUCTURE ONLY.  (The file name contains a literal
// backslash: the reference includes <opencv2\opencv.hpp>, Common/Utils.h:17.)  Just enough of the
// cv:: names that the static helpers of Common/Utils.h:188-243 (RenderDepthMap, SSD — debug image
// output, not on the hot path) mention, so that the header parses; OpenCV's C++ headers are not in
// this image.  Nothing on the compared path calls these.
#ifndef MVS_COMPAT_OPENCV_HPP
#define MVS_COMPAT_OPENCV_HPP
#include <string>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
namespace cv {
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
class Mat {
 public:
  int rows, cols;
  Mat() : rows(0), cols(0), type_(0) {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), d_((size_t)r * c * 4, 0) {}
  Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), type_(m.type_),
                                       d_((size_t)roi.height * roi.width * 4, 0) {}
  template <typename T>
  T& at(int i, int j) { return *reinterpret_cast<T*>(&d_[((size_t)i * cols + j) * 4]); }
  template <typename T>
  const T& at(int i, int j) const { return *reinterpret_cast<const T*>(&d_[((size_t)i * cols + j) * 4]); }
 private:
  int type_;
  std::vector<unsigned char> d_;
};
enum { COLOR_RGB2GRAY = 7 };
static inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src; }
static inline bool imwrite(const std::string&, const Mat&) { return false; }
}  // namespace cv
#endif
