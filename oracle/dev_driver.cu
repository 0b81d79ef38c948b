// extern "C" driver around the reference's OWN CUDA code (BlockMatching/Device.cu compiled unmodified from
// /root/reference with nvcc for sm_100a).  DEV TOOL ONLY (tools/ref_gpu_compare.py): lets the reference's GPU path
// run on the B200 next to libgsm.so.  Never linked into the product library, not used by tests or bench.
#include "Device.cuh"  // reference header, BlockMatching/Device.cuh:49-52
#include <cstdint>

extern "C" {
// Device.cu:173-301.  The reference's launch geometry only covers 256 x 320 images (PreCal_V2 grid (8, 10, D) x
// block (32, 32), FindCorr <<<rows, cols>>>) -- the size of its demo pair (Caller.cpp:12-19).
int devref_block_matching(const uint8_t* L, const uint8_t* R, int rows, int cols, int radius, int D, uint8_t* disp) {
  Mat l(rows, cols, CV_8UC1, (void*)L), r(rows, cols, CV_8UC1, (void*)R), d;
  blockMatching_gpu(l, r, d, radius, D);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess || !d.data) return (int)e ? (int)e : -1;
  memcpy(disp, d.data, (size_t)rows * cols);
  return 0;
}
// Device.cu:303-342 (returns the LEFT remapped image only, like the reference)
int devref_remap(const uint8_t* left, const uint8_t* right, const float* mx1, const float* my1, const float* mx2,
                 const float* my2, int rows, int cols, uint8_t* result) {
  Mat l(rows, cols, CV_8UC1, (void*)left), r(rows, cols, CV_8UC1, (void*)right);
  Mat a(rows, cols, CV_32FC1, (void*)mx1), b(rows, cols, CV_32FC1, (void*)my1), c(rows, cols, CV_32FC1, (void*)mx2),
      e(rows, cols, CV_32FC1, (void*)my2);
  remap_gpu(l, r, a, b, c, e, rows, cols, rows * cols, result);
  return (int)cudaDeviceSynchronize();
}
// Device.cu:344-367 (1000 launches of kernalCvtColor, as shipped)
int devref_cvtcolor(const uint8_t* src3, uint8_t* dst, int rows, int cols) {
  cvtColor_gpu((uchar3*)src3, dst, rows, cols);
  return (int)cudaDeviceSynchronize();
}
}
