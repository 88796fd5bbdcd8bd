// oracle/compat/opencv2\opencv.hpp — TEST INFRASTRUCTURE ONLY.  (The file name contains a literal
// backslash: the reference includes <opencv2\opencv.hpp>, Common/Utils.h:17.)  Just enough of the
// cv:: names that the static helpers of Common/Utils.h:188-243 (RenderDepthMap, SSD — debug image
// output, not on the hot path) mention, so that the header parses; OpenCV's C++ headers are not in
// this image.  Nothing on the compared path calls these.
//
// cv::flann::Index (PartRecognition/PartRecognition.cpp:50-77, Deformation/Deformation.cpp:63-153,
// 155-230, 232-297): the reference asks FLANN's randomised kd-trees (checks = 32..200) for nearest
// neighbours — an APPROXIMATE search whose contract is the exact answer.  The stand-in below IS that
// exact answer (float32 squared distances summed in x, y, z order, ascending (distance, index)
// order, radius search in squared units as FLANN's L2 functor defines it), so that
// Alignment::Align and Deformation::Deform can be compiled and run as a whole; it pins the
// arithmetic around the searches, not FLANN itself.
#ifndef MVS_COMPAT_OPENCV_HPP
#define MVS_COMPAT_OPENCV_HPP
#include <algorithm>
#include <string>
#include <utility>
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
  int type() const { return type_; }
  void copyTo(Mat& dst) const { dst = *this; }
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
// exact searches in the feature space of the matrix rows: float32 squared distances accumulated
// dimension by dimension, results in ascending (distance, index) order
class Index {
 public:
  Index(const Mat& features, const KDTreeIndexParams&)
      : n_(features.rows), dim_(features.cols), p_((size_t)features.rows * features.cols) {
    for (int i = 0; i < n_; ++i)
      for (int j = 0; j < dim_; ++j) p_[(size_t)dim_ * i + j] = features.at<float>(i, j);
  }
  void knnSearch(const std::vector<float>& q, std::vector<int>& indices, std::vector<float>& dists,
                 int knn, const SearchParams&) const {
    std::vector<std::pair<float, int> > all = scan(q);
    size_t k = std::min<size_t>((size_t)(knn > 0 ? knn : 0), all.size());
    std::partial_sort(all.begin(), all.begin() + k, all.end());
    if (indices.size() < (size_t)knn) indices.resize((size_t)knn, -1);
    if (dists.size() < (size_t)knn) dists.resize((size_t)knn, 0.f);
    for (size_t j = 0; j < k; ++j) {
      indices[j] = all[j].second;
      dists[j] = all[j].first;
    }
  }
  // radius in the units of the distances above (squared); writes the first min(count, maxResults)
  // entries and leaves the rest of the output vectors as the caller filled them
  int radiusSearch(const std::vector<float>& q, std::vector<int>& indices, std::vector<float>& dists,
                   double radius, int maxResults, const SearchParams&) const {
    std::vector<std::pair<float, int> > all = scan(q), in;
    for (size_t j = 0; j < all.size(); ++j)
      if (all[j].first <= (float)radius) in.push_back(all[j]);
    std::sort(in.begin(), in.end());
    size_t k = std::min<size_t>(in.size(), (size_t)(maxResults > 0 ? maxResults : 0));
    if (indices.size() < k) indices.resize(k, -1);
    if (dists.size() < k) dists.resize(k, 0.f);
    for (size_t j = 0; j < k; ++j) {
      indices[j] = in[j].second;
      dists[j] = in[j].first;
    }
    return (int)k;
  }

 private:
  std::vector<std::pair<float, int> > scan(const std::vector<float>& q) const {
    std::vector<std::pair<float, int> > all((size_t)n_);
    for (int i = 0; i < n_; ++i) {
      float d = 0.f;
      for (int j = 0; j < dim_; ++j) {
        float e = q[(size_t)j] - p_[(size_t)dim_ * i + j];
        float ee = e * e;
        d = j ? d + ee : ee;
      }
      all[(size_t)i] = std::make_pair(d, i);
    }
    return all;
  }
  int n_, dim_;
  std::vector<float> p_;
};
}  // namespace flann
// debug drawing / display of Processor/Processor.cpp (match and key-point pictures): no-ops
struct Point2f {
  float x, y;
  Point2f(float x_ = 0.f, float y_ = 0.f) : x(x_), y(y_) {}
};
struct Scalar {
  double v[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
};
struct RNG {
  double uniform(double a, double) { return a; }
};
static inline Mat imread(const std::string&, int = 1) { return Mat(); }
static inline void vconcat(const Mat&, const Mat&, Mat&) {}
static inline void circle(Mat&, Point2f, int, const Scalar&, int = 1) {}
static inline void line(Mat&, Point2f, Point2f, const Scalar&, int = 1) {}
static inline void imshow(const std::string&, const Mat&) {}
static inline int waitKey(int = 0) { return -1; }
enum { COLOR_RGB2GRAY = 7 };
static inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src; }
static inline bool imwrite(const std::string&, const Mat&) { return false; }
}  // namespace cv
#endif
