// oracle/compat/opencv2\opencv.hpp — TEST INFRASTRUCTURE ONLY.  (The file name contains a literal
// backslash: the reference includes <opencv2\opencv.hpp>, Common/Utils.h:17.)  Just enough of the
// cv:: names that the static helpers of Common/Utils.h:188-243 (RenderDepthMap, SSD — debug image
// output, not on the hot path) mention, so that the header parses; OpenCV's C++ headers are not in
// this image.  Nothing on the compared path calls these.
//
// cv::flann::Index (PartRecognition/PartRecognition.cpp:50-77): the reference asks FLANN's randomised
// kd-trees (checks = 200) for the nearest template vertex — an APPROXIMATE search whose contract is
// the exact nearest neighbour.  The stand-in below IS that exact answer (float32 squared distances
// summed in x, y, z order, lowest index on ties), so that Alignment::Align can be compiled and run
// as a whole; it pins the alignment arithmetic around the search, not FLANN itself.
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
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  template <typename T>
  T& at(int i, int j) { return *reinterpret_cast<T*>(&d_[((size_t)i * cols + j) * 4]); }
  template <typename T>
  const T& at(int i, int j) const { return *reinterpret_cast<const T*>(&d_[((size_t)i * cols + j) * 4]); }
 private:
  int type_;
  std::vector<unsigned char> d_;
};
namespace flann {
struct KDTreeIndexParams {
  explicit KDTreeIndexParams(int = 4) {}
};
struct SearchParams {
  explicit SearchParams(int = 32, float = 0.f, bool = true) {}
};
class Index {
 public:
  Index(const Mat& features, const KDTreeIndexParams&) : n_(features.rows), p_((size_t)features.rows * 3) {
    for (int i = 0; i < n_; ++i)
      for (int j = 0; j < 3; ++j) p_[(size_t)3 * i + j] = features.at<float>(i, j);
  }
  void knnSearch(const std::vector<float>& q, std::vector<int>& indices, std::vector<float>& dists,
                 int knn, const SearchParams&) const {
    (void)knn;  // the reference only asks for one neighbour
    int best = -1;
    float bd = 0.f;
    for (int i = 0; i < n_; ++i) {
      volatile float dx = q[0] - p_[(size_t)3 * i], dy = q[1] - p_[(size_t)3 * i + 1],
                     dz = q[2] - p_[(size_t)3 * i + 2];
      volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
      volatile float xy = xx + yy;
      float d = xy + zz;
      if (best < 0 || d < bd) {
        best = i;
        bd = d;
      }
    }
    indices.assign(1, best);
    dists.assign(1, bd);
  }

 private:
  int n_;
  std::vector<float> p_;
};
}  // namespace flann
enum { COLOR_RGB2GRAY = 7 };
static inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src; }
static inline bool imwrite(const std::string&, const Mat&) { return false; }
}  // namespace cv
#endif
