// Minimal cv::Mat / cv::Point3i stand-in so that the reference's BlockMatching.cpp
// compiles UNMODIFIED from /root/reference (test infrastructure only; see oracle/README).
// Only what BlockMatching.cpp touches: Mat{rows, cols, data}, Mat(r,c,type,ptr), ptr<T>(row),
// Point3i{x,y,z}, uchar, CV_8UC1.  (reference: BlockMatching/BlockMatching.h:4-5 includes
// <opencv2\core\core.hpp> / <opencv2\highgui\highgui.hpp> with literal backslashes.)
#ifndef GSM_ORACLE_CVSHIM_HPP
#define GSM_ORACLE_CVSHIM_HPP
#include <cstdlib>
#include <cstring>
#include <cmath>
typedef unsigned char uchar;
#ifndef CV_8UC1
#define CV_8UC1 0
#endif
namespace cv {
struct Point3i {
  int x, y, z;
  Point3i() : x(0), y(0), z(0) {}
  Point3i(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
};
struct Mat {
  int rows, cols;
  uchar* data;
  Mat() : rows(0), cols(0), data(0) {}
  Mat(int r, int c, int /*type*/, void* d) : rows(r), cols(c), data((uchar*)d) {}
  template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * cols); }
  template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * cols); }
};
}  // namespace cv
#endif
