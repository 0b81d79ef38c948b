// File-to-file drop-in check: the GUI-free singleFrame of include/gsm_caller.hpp (imread + cvtColor + blockMatching_gpu
// + "imshow" into a file; reference: BlockMatching/Caller.cpp:9-25) on the reference's own demo pair read FROM DISK,
// followed by the reference's own acceptance hook compareDisp (BlockMatching.cpp:278-293) when linked with
// oracle/_ref/libref.so (-DWITH_REF): compareDisp prints one block per mismatching pixel and nothing when GPU == CPU.
//   usage: caller_files left.png right.png disp.pgm
#include <cstdio>

#include "cvshim.hpp"  // stands in for <opencv2/core/core.hpp> in this repository's tests
#include "gsm_caller.hpp"

#ifdef WITH_REF
void compareDisp(const cv::Mat& left, const cv::Mat& right, uchar* GPUresult, int SADWindowSize, int searchRange,
                 int cols, int rows);  // BlockMatching.h:14
#endif

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  if (gsm_caller::singleFrame(argv[1], argv[2], argv[3])) return 1;  // SAD r=5, 64 disparities: Caller.cpp:19
  gsm_io::Image L, R, D;
  std::string err;
  if (!gsm_io::read_image(argv[1], L, err) || !gsm_io::read_image(argv[2], R, err) || !gsm_io::read_image(argv[3], D, err)) {
    std::fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  std::printf("GPU_DONE %d %d\n", D.rows, D.cols);
  std::fflush(stdout);
#ifdef WITH_REF
  std::vector<unsigned char> g1 = gsm_io::to_gray(L), g2 = gsm_io::to_gray(R);
  cv::Mat m1(L.rows, L.cols, CV_8UC1, g1.data()), m2(L.rows, L.cols, CV_8UC1, g2.data());
  compareDisp(m1, m2, D.data.data(), 5, 64, L.cols, L.rows);
  std::printf("\nCOMPARE_DONE\n");
#endif
  return 0;
}
