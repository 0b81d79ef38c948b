// extern "C" driver around the reference's OWN BlockMatching/Utility.cpp (compiled unmodified from
// /root/reference): the CPU twins of remap_gpu and cvtColor_gpu.  TEST INFRASTRUCTURE ONLY: pins orc_remap /
// orc_cvtcolor in oracle/stereo_oracle.c.  Never linked into the product library.
#include "Utility.h"  // reference header, BlockMatching/Utility.h:44-54
#include <cstdint>

extern "C" {
// Utility.cpp:236-264 (note the reference's own argument order: BilinearInterpolation(src, ycoo, xcoo))
void ref_cpu_remap(const uint8_t* src, int rows, int cols, const float* mapx, const float* mapy, uint8_t* dst) {
  Mat s(rows, cols, CV_8UC1, (void*)src), mx(rows, cols, CV_32FC1, (void*)mapx), my(rows, cols, CV_32FC1, (void*)mapy);
  CPU_Remap(s, dst, mx, my);
}
// Utility.cpp:289-298 (truncating conversion, RGB weights applied to the bytes as stored)
void ref_cvtcolor_cpu(const uint8_t* src3, uint8_t* dst, int rows, int cols) {
  cvtColor_cpu((uchar3*)src3, dst, rows, cols);
}
}
