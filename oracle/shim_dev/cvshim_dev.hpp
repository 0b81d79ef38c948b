// OpenCV stand-in so that the reference's BlockMatching/Device.cu compiles UNMODIFIED from /root/reference with nvcc
// (dev tool only: tools/ref_gpu_compare.py runs the reference's OWN CUDA kernels on the B200 next to libgsm.so).
// Device.cuh:10-18 includes the OpenCV headers with literal backslashes; the files next to this one carry those names.
#ifndef GSM_ORACLE_CVSHIM_DEV_HPP
#define GSM_ORACLE_CVSHIM_DEV_HPP
#include "../shim_util/cvshim_util.hpp"
namespace cv {
struct Point3i {
  int x, y, z;
  Point3i() : x(0), y(0), z(0) {}
  Point3i(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
};
}  // namespace cv
#endif
