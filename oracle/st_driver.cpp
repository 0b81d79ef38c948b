// extern "C" driver around the reference's OWN STMatching/StereoHelper.cpp (compiled unmodified from
// /root/reference): float winner-take-all and the right-view cost volume.  TEST INFRASTRUCTURE ONLY: pins the
// restatements in oracle/stereo_oracle.c.  Never linked into the product library.
#include "StereoHelper.h"  // reference header, STMatching/StereoHelper.h:34-40
#include <cstdint>

extern "C" {
// StereoHelper.cpp:131-154; vol is float [h][w][D] pixel-major (Toolkit.h:76-79)
void ref_wta_float(const float* vol, int w, int h, int D, uint8_t* out) {
  CDisparityHelper hlp;
  cv::Mat d = hlp.GetDisparity_WTA((float*)vol, w, h, D);
  memcpy(out, d.data, (size_t)w * h);
}
// StereoHelper.cpp:156-180; left / right are float [h][w][D]
void ref_right_from_left(const float* left, int w, int h, int D, float* right) {
  CDisparityHelper hlp;
  cv::Mat lv(1, w * h * D, CV_32F, (void*)left);
  cv::Mat rv = hlp.GetRightMatchingCostFromLeft(lv, w, h, D);
  memcpy(right, rv.data, sizeof(float) * (size_t)w * h * D);
}
}
