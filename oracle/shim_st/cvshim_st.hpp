// Minimal OpenCV stand-in so that the reference's STMatching/StereoHelper.cpp compiles UNMODIFIED from
// /root/reference (test infrastructure only).  Only what that file and STMatching/Toolkit.h touch:
// Mat (owning or wrapping, clone, size, type, data), Mat_<T> views with (y, x) access, Size, Scalar,
// InputArray / OutputArray, CV_Assert and the three type codes.  (reference: StereoHelper.cpp:29-35 includes
// <opencv2/core/core.hpp>, <opencv2/highgui/highgui.hpp>, <opencv2/imgproc/imgproc.hpp>.)
#ifndef GSM_ORACLE_CVSHIM_ST_HPP
#define GSM_ORACLE_CVSHIM_ST_HPP
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_32F 5
#define CV_8UC3 16
#define CV_Assert(expr) assert(expr)
namespace cv {
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  int area() const { return width * height; }
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Vec3b {
  uchar v[3];
  uchar& operator[](int i) { return v[i]; }
  const uchar& operator[](int i) const { return v[i]; }
};
struct Scalar {
  double v;
  Scalar(double x = 0) : v(x) {}
};
struct Mat {
  int rows, cols, type_;
  uchar* data;
  std::shared_ptr<std::vector<uchar> > own;
  static size_t esz(int t) { return (size_t)((t & 7) == CV_32F ? 4 : 1) * (size_t)((t >> 3) + 1); }
  Mat() : rows(0), cols(0), type_(0), data(0) {}
  Mat(int r, int c, int t) { create(r, c, t); }
  Mat(Size s, int t) { create(s.height, s.width, t); }
  Mat(Size s, int t, const Scalar& v) {
    create(s.height, s.width, t);
    assert(t == CV_32F || t == CV_8U);
    if (t == CV_32F) for (size_t i = 0; i < (size_t)rows * cols; ++i) ((float*)data)[i] = (float)v.v;
    else memset(data, (int)v.v, (size_t)rows * cols);
  }
  Mat(int r, int c, int t, void* d) : rows(r), cols(c), type_(t), data((uchar*)d) {}
  void create(int r, int c, int t) {
    rows = r; cols = c; type_ = t;
    own.reset(new std::vector<uchar>((size_t)r * c * esz(t)));
    data = own->data();
  }
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t step1() const { return (size_t)cols * channels(); }  // continuous rows, in elements (Toolkit.cpp:39-41)
  bool empty() const { return data == 0; }
  void copyTo(Mat dst) const {  // destination already has the size and type (all uses in STMatching)
    assert(dst.rows == rows && dst.cols == cols && dst.type_ == type_);
    if (dst.data != data) memcpy(dst.data, data, (size_t)rows * cols * esz(type_));
  }
  Mat& operator*=(double s) {  // disparity *= scale (StereoDisparity.cpp:87,158): CV_8U, saturating
    assert(type_ == CV_8U);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
      const double v = data[i] * s;
      data[i] = (uchar)(v > 255 ? 255 : (v < 0 ? 0 : (int)(v + 0.5)));
    }
    return *this;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    memcpy(m.data, data, (size_t)rows * cols * esz(type_));
    return m;
  }
};
template <class T> struct Mat_ : Mat {
  Mat_() {}
  Mat_(const Mat& m) : Mat(m) {}
  T& operator()(int y, int x) { return ((T*)data)[(size_t)y * cols + x]; }
  const T& operator()(int y, int x) const { return ((const T*)data)[(size_t)y * cols + x]; }
};
typedef Mat_<uchar> Mat1b;
typedef Mat_<Vec3b> Mat3b;
typedef Mat_<float> Mat1f;
struct _InputArray {
  Mat m;
  _InputArray(const Mat& x) : m(x) {}
  Mat getMat() const { return m; }
  // OutputArray::create is only reached with the size and type the array already has (Toolkit.cpp:43-45,
  // StereoDisparity.cpp:66,110), where OpenCV's create() is a no-op
  void create(Size s, int t) const { assert(m.rows == s.height && m.cols == s.width && m.type_ == t); }
};
// file I/O of the command-line wrappers (StereoDisparity.cpp:43-44,54): never reached from the test drivers
inline Mat imread(const char*) { abort(); return Mat(); }
inline bool imwrite(const char*, const Mat&) { abort(); return false; }
typedef const _InputArray& InputArray;
typedef const _InputArray& OutputArray;
}  // namespace cv
#endif
