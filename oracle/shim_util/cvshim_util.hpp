// OpenCV stand-in so that the reference's BlockMatching/Utility.cpp compiles UNMODIFIED from /root/reference
// (test infrastructure only).  The two functions the oracle pins -- CPU_Remap / CPU_BilinearInterpolation
// (Utility.cpp:236-264) and cvtColor_cpu (:289-298) -- only need Mat{rows, cols, data, ptr<T>(row)} and
// saturate_cast<uchar>(float), which is OpenCV's documented cvRound (round half to even) + clamp.  Everything else
// the file mentions (capture, calibration, file storage, StereoBM, display) is declared so the file compiles and
// aborts if it is ever called.
#ifndef GSM_ORACLE_CVSHIM_UTIL_HPP
#define GSM_ORACLE_CVSHIM_UTIL_HPP
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
typedef unsigned char uchar;
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH_MASK 7
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_BGR2GRAY 6
#define CV_MINMAX 32
#define CV_CALIB_CB_ADAPTIVE_THRESH 1
#define CV_CALIB_CB_FILTER_QUADS 4
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS 2
#define CV_CALIB_ZERO_DISPARITY 1024
namespace cv {
using std::string;
using std::vector;
static inline void shim_unavailable(const char* what) {
  fprintf(stderr, "oracle shim: %s is not available (only CPU_Remap / cvtColor_cpu are exercised)\n", what);
  abort();
}
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f { float x, y, z; Point3f() : x(0), y(0), z(0) {} Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct TermCriteria { TermCriteria(int, int, double) {} };
struct _IOArray { template <class T> _IOArray(const T&) {} };
typedef const _IOArray& InputArray;
typedef const _IOArray& OutputArray;
typedef const _IOArray& InputOutputArray;
typedef const _IOArray& InputArrayOfArrays;
typedef const _IOArray& OutputArrayOfArrays;
struct Mat {
  int rows, cols, type_;
  uchar* data;
  std::shared_ptr<std::vector<uchar> > own;
  static size_t esz(int t) {
    static const size_t d[7] = {1, 1, 2, 2, 4, 4, 8};
    return d[t & CV_MAT_DEPTH_MASK] * (size_t)((t >> CV_CN_SHIFT) + 1);
  }
  Mat() : rows(0), cols(0), type_(0), data(0) {}
  Mat(int r, int c, int t) : rows(r), cols(c), type_(t) {
    own.reset(new std::vector<uchar>((size_t)r * c * esz(t)));
    data = own->data();
  }
  Mat(int r, int c, int t, void* d) : rows(r), cols(c), type_(t), data((uchar*)d) {}
  Size size() const { return Size(cols, rows); }
  template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * cols * esz(type_)); }
  template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * cols * esz(type_)))[c]; }
  void convertTo(OutputArray, int) const { shim_unavailable("Mat::convertTo"); }
};
template <class T> static inline T saturate_cast(float v);
template <> inline uchar saturate_cast<uchar>(float v) {  // OpenCV: cvRound (nearest, ties to even), then clamp to 0..255
  const int iv = (int)lrintf(v);
  return (uchar)((unsigned)iv <= 255 ? iv : iv > 0 ? 255 : 0);
}
struct FileNode { void operator>>(Mat&) const { shim_unavailable("FileNode"); } };
struct FileStorage {
  enum { READ = 0, WRITE = 1 };
  FileStorage(const string&, int) {}
  FileNode operator[](const string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  void release() {}
};
template <class T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { shim_unavailable("FileStorage"); return fs; }
struct VideoCapture {
  VideoCapture(int) {}
  VideoCapture& operator>>(Mat&) { shim_unavailable("VideoCapture"); return *this; }
  void release() {}
};
struct _BMState {
  int SADWindowSize, numberOfDisparities, preFilterSize, preFilterCap, minDisparity, textureThreshold, uniquenessRatio,
      speckleWindowSize, speckleRange, disp12MaxDiff;
};
struct StereoBM {
  _BMState s_, *state;
  StereoBM() : state(&s_) {}
  void operator()(InputArray, InputArray, OutputArray) { shim_unavailable("StereoBM"); }
};
static inline void imshow(const string&, InputArray) { shim_unavailable("imshow"); }
static inline int waitKey(int = 0) { shim_unavailable("waitKey"); return 0; }
static inline void destroyAllWindows() {}
static inline void cvtColor(InputArray, OutputArray, int) { shim_unavailable("cvtColor"); }
static inline void normalize(InputArray, OutputArray, double, double, int, int) { shim_unavailable("normalize"); }
static inline bool findChessboardCorners(InputArray, Size, OutputArray, int) { shim_unavailable("findChessboardCorners"); return false; }
static inline void cornerSubPix(InputArray, InputOutputArray, Size, Size, TermCriteria) { shim_unavailable("cornerSubPix"); }
static inline void drawChessboardCorners(InputOutputArray, Size, InputArray, bool) { shim_unavailable("drawChessboardCorners"); }
static inline double calibrateCamera(InputArrayOfArrays, InputArrayOfArrays, Size, InputOutputArray, InputOutputArray,
                                     OutputArrayOfArrays, OutputArrayOfArrays) { shim_unavailable("calibrateCamera"); return 0; }
static inline void undistort(InputArray, OutputArray, InputArray, InputArray) { shim_unavailable("undistort"); }
static inline void resize(InputArray, OutputArray, Size) { shim_unavailable("resize"); }
static inline bool imwrite(const string&, InputArray) { shim_unavailable("imwrite"); return false; }
static inline void stereoRectify(InputArray, InputArray, InputArray, InputArray, Size, InputArray, InputArray, OutputArray,
                                 OutputArray, OutputArray, OutputArray, OutputArray, int) { shim_unavailable("stereoRectify"); }
static inline void initUndistortRectifyMap(InputArray, InputArray, InputArray, InputArray, Size, int, OutputArray,
                                           OutputArray) { shim_unavailable("initUndistortRectifyMap"); }
}  // namespace cv
#endif
