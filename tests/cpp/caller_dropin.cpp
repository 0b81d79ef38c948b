// Drop-in check in the reference's own terms: the compute part of singleFrame()
// (BlockMatching/Caller.cpp:9-25) through the kept signature blockMatching_gpu(g1, g2, disp, 5, 64),
// followed by the reference's own acceptance hook compareDisp (BlockMatching.cpp:278-293; commented out at
// Device.cu:296-297) when the test links oracle/_ref/libref.so (-DWITH_REF).  compareDisp prints one block
// per mismatching pixel and nothing when GPU == CPU.
// With an eighth argument it also calls ::cvtColor_gpu((uchar3*)..., gray, rows, cols) exactly like Caller.cpp:106 on a
// 3-channel image built from the two inputs and writes the gray result there.
//   usage: caller_dropin left.gray right.gray rows cols [radius=5] [searchRange=64] [out.gray] [cvt_out.gray]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cvshim.hpp"  // stands in for <opencv2/core/core.hpp> in this repository's tests
#define GSM_COMPAT_REFERENCE_NAMES
#define GSM_COMPAT_DEFINE_UCHAR3  // this translation unit has no CUDA headers: the wrapper brings its own uchar3
#include "gsm_compat.hpp"

using namespace cv;

#ifdef WITH_REF
void compareDisp(const cv::Mat& left, const cv::Mat& right, uchar* GPUresult, int SADWindowSize, int searchRange,
                 int cols, int rows);  // BlockMatching.h:14
#endif

static std::vector<uchar> slurp(const char* path, size_t n) {
  std::vector<uchar> v(n);
  FILE* f = std::fopen(path, "rb");
  if (!f || std::fread(v.data(), 1, n, f) != n) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
  std::fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: %s left right rows cols [radius] [range] [out]\n", argv[0]); return 2; }
  const int rows = std::atoi(argv[3]), cols = std::atoi(argv[4]);
  const int radius = argc > 5 ? std::atoi(argv[5]) : 5, range = argc > 6 ? std::atoi(argv[6]) : 64;
  std::vector<uchar> l = slurp(argv[1], (size_t)rows * cols), r = slurp(argv[2], (size_t)rows * cols);
  Mat g1(rows, cols, CV_8UC1, l.data()), g2(rows, cols, CV_8UC1, r.data()), disp;
  gsm_compat::set_device(0);                       // the wrappers run on a selectable GPU (default 0)
  blockMatching_gpu(g1, g2, disp, radius, range);  // Caller.cpp:19
  std::printf("GPU_DONE %d %d\n", disp.rows, disp.cols);
  std::fflush(stdout);
#ifdef WITH_REF
  compareDisp(g1, g2, disp.data, radius, range, cols, rows);
  std::printf("\nCOMPARE_DONE\n");
#endif
  if (argc > 7) {
    FILE* f = std::fopen(argv[7], "wb");
    std::fwrite(disp.data, 1, (size_t)rows * cols, f);
    std::fclose(f);
  }
  if (argc > 8) {
    std::vector<uchar> bgr((size_t)3 * rows * cols), gray((size_t)rows * cols);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) { bgr[3 * i] = l[i]; bgr[3 * i + 1] = r[i]; bgr[3 * i + 2] = (uchar)(l[i] ^ r[i]); }
    cvtColor_gpu((uchar3*)bgr.data(), gray.data(), rows, cols);  // Caller.cpp:106, reference name and signature
    FILE* f = std::fopen(argv[8], "wb");
    std::fwrite(gray.data(), 1, gray.size(), f);
    std::fclose(f);
  }
  return 0;
}
