// extern "C" driver around the reference's OWN CPU functions (compiled unmodified from
// /root/reference/BlockMatching/BlockMatching.cpp and /root/reference/STMatching/ctmf.c).
// TEST INFRASTRUCTURE ONLY: used to pin oracle/stereo_oracle.c and as the CPU baseline
// (bench.py cpu_baseline.kind == "reference").  Never linked into the product library.
#include "cvshim.hpp"
#include <cstdint>

// reference declarations: BlockMatching/BlockMatching.h:8-15
void testBM(const cv::Mat&, const cv::Mat&, cv::Mat&, int, int);
void PreCal(const cv::Mat&, const cv::Mat&, uchar*, int, int);
void getDisp(const cv::Mat&, const cv::Mat&, uchar*, int, int);
void getAllSAD(const cv::Mat&, const cv::Mat&, uchar*, int, int);
void compareDisp(const cv::Mat&, const cv::Mat&, uchar*, int, int, int, int);
void compareDiff(const cv::Mat&, const cv::Mat&, uchar*, int, int, int);
// reference declaration: STMatching/ctmf.h:39-45
extern "C" void ctmf(const unsigned char* src, unsigned char* dst, int width, int height,
                     int src_step, int dst_step, int r, int cn, unsigned long memsize);

extern "C" {

// BlockMatching.cpp:89-109; caller pre-zeroes `diff` ([D][H][W] u8) like :39,143
void ref_PreCal(const uint8_t* L, const uint8_t* R, int rows, int cols, int radius, int D, uint8_t* diff) {
  cv::Mat l(rows, cols, CV_8UC1, (void*)L), r(rows, cols, CV_8UC1, (void*)R);
  PreCal(l, r, diff, radius, D);
}
// BlockMatching.cpp:111-189
void ref_getDisp(const uint8_t* L, const uint8_t* R, int rows, int cols, int radius, int D, uint8_t* disp) {
  cv::Mat l(rows, cols, CV_8UC1, (void*)L), r(rows, cols, CV_8UC1, (void*)R);
  getDisp(l, r, disp, radius, D);
}
// BlockMatching.cpp:191-261; out is [H*W][D] u8 (caller pre-fills with 255 like compareSAD :299)
void ref_getAllSAD(const uint8_t* L, const uint8_t* R, int rows, int cols, int radius, int D, uint8_t* out) {
  cv::Mat l(rows, cols, CV_8UC1, (void*)L), r(rows, cols, CV_8UC1, (void*)R);
  getAllSAD(l, r, out, radius, D);
}
// BlockMatching.cpp:278-293 -- the reference's own acceptance check (prints mismatches to stdout)
void ref_compareDisp(const uint8_t* L, const uint8_t* R, uint8_t* gpu, int rows, int cols, int radius, int D) {
  cv::Mat l(rows, cols, CV_8UC1, (void*)L), r(rows, cols, CV_8UC1, (void*)R);
  compareDisp(l, r, gpu, radius, D, cols, rows);
}
// Toolkit.cpp:33-48 (MeanFilter == (2r+1)^2 median through ctmf, memsize = area*channels)
void ref_median(const uint8_t* src, uint8_t* dst, int rows, int cols, int r) {
  ctmf(src, dst, cols, rows, cols, cols, r, 1, (unsigned long)rows * cols);
}
}
