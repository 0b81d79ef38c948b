// gsm_caller.hpp -- GUI-free replacements of the reference's demo entry points (BlockMatching/Caller.h:8-10,
// Caller.cpp:9-112; called from Main.cpp:5-7): the same three steps -- singleFrame, remapTest, cvtColorTest -- with
// file paths (or buffers) in and image files out instead of cv::imread / cv::imshow / cv::waitKey
// (Caller.cpp:12-13,23-24,70-73,108-111), plus depth from disparity and a batch front-end.  Capture and display stay
// decoupled from the compute path (north_star); image decoding is host code (gsm_imageio.hpp), every pixel operation
// runs in libgsm.so.  Header-only; link with -lgsm -lz.  gpu_stereo_matching_b200/csrc/gsm_caller.cpp is the CLI.
#ifndef GSM_CALLER_HPP
#define GSM_CALLER_HPP

#include <chrono>
#include <cstdio>
#include <string>
#include <vector>

#include "gsm.h"
#include "gsm_imageio.hpp"

namespace gsm_caller {

struct Options {
  int device = 0;
  int mode = GSM_MODE_SAD;  // singleFrame() of the reference is SAD r=5, 64 disparities (Caller.cpp:19)
  int radius = 5;
  int num_disp = 64;
  int lr_check = 0;
  int median_radius = 0;
  float eps = 0.f;
};

inline int fail(const char* where, const std::string& msg) {
  std::fprintf(stderr, "%s: %s\n", where, msg.c_str());
  return 1;
}

struct Ctx {  // RAII around gsm_create / gsm_destroy
  gsm_ctx* c = nullptr;
  Ctx(int device, int rows, int cols, int disp, int batch) {
    if (gsm_create(&c, device, rows, cols, disp, batch) != GSM_OK) c = nullptr;
  }
  ~Ctx() { if (c) gsm_destroy(c); }
  Ctx(const Ctx&) = delete;
  Ctx& operator=(const Ctx&) = delete;
};

inline gsm_params params_of(const Options& o) {
  gsm_params p = {};
  p.mode = o.mode;
  p.radius = o.radius;
  p.num_disp = o.num_disp;
  p.lr_check = o.lr_check;
  p.median_radius = o.median_radius;
  p.eps = o.eps;
  return p;
}

// ---- singleFrame (Caller.cpp:9-25) on buffers: gray pair in, disparity out --------------------------------------------
inline int singleFrame(const unsigned char* left_gray, const unsigned char* right_gray, unsigned char* disp,
                       unsigned char* mask, int rows, int cols, const Options& o = Options(), double* seconds = nullptr) {
  Ctx ctx(o.device, rows, cols, o.num_disp, 1);
  if (!ctx.c) return fail("singleFrame", gsm_last_error());
  const gsm_params p = params_of(o);
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = (o.mode == GSM_MODE_SAD && !o.lr_check && !o.median_radius)
                     ? gsm_block_matching(ctx.c, left_gray, right_gray, disp, rows, cols, o.radius, o.num_disp)  // Caller.cpp:19
                     : gsm_stereo_batch(ctx.c, &p, 1, left_gray, right_gray, disp, mask, rows, cols);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc == GSM_OK ? 0 : fail("singleFrame", gsm_last_error());
}

// ---- singleFrame on files: imread + cvtColor(BGR2GRAY) + blockMatching_gpu + "imshow" to a file ------------------------
// left / right: PNG or binary PGM / PPM; disp_path: .png, .pgm or raw; mask_path (optional, with lr_check) likewise.
inline int singleFrame(const char* left_path, const char* right_path, const char* disp_path, const Options& o = Options(),
                       const char* mask_path = nullptr) {
  gsm_io::Image L, R;
  std::string err;
  if (!gsm_io::read_image(left_path, L, err) || !gsm_io::read_image(right_path, R, err)) return fail("singleFrame", err);
  if (L.rows != R.rows || L.cols != R.cols) return fail("singleFrame", "left and right images differ in size");
  const std::vector<unsigned char> g1 = gsm_io::to_gray(L), g2 = gsm_io::to_gray(R);  // Caller.cpp:15-16
  std::vector<unsigned char> disp((size_t)L.rows * L.cols), mask(o.lr_check ? disp.size() : 0);
  double sec = 0;
  if (singleFrame(g1.data(), g2.data(), disp.data(), o.lr_check ? mask.data() : nullptr, L.rows, L.cols, o, &sec)) return 1;
  std::printf("GPU : %g\n", sec);  // Caller.cpp:21
  if (!gsm_io::write_gray(disp_path, disp.data(), L.rows, L.cols, err)) return fail("singleFrame", err);
  if (mask_path && o.lr_check && !gsm_io::write_gray(mask_path, mask.data(), L.rows, L.cols, err)) return fail("singleFrame", err);
  return 0;
}

// the reference's argument-less entry point with its own relative paths (Caller.cpp:12-13); the map goes to disp.png
inline int singleFrame() { return singleFrame("./../Images/Art/view1_.png", "./../Images/Art/view5_.png", "disp.png"); }

// ---- remapTest (Caller.cpp:27-74): rectify a pair through the four CV_32FC1 maps Rectify() builds ----------------------
// (Utility.cpp:228-234; building them needs OpenCV's calibration module and stays with the host application).
// maps_path: raw float32, [mapX1][mapY1][mapX2][mapY2], each rows x cols.  Like remap_gpu (Device.cu:303-342) both
// images are remapped; the left result is always written, the right one when out_right_path is given.
inline int remapTest(const char* left_path, const char* right_path, const char* maps_path, const char* out_left_path,
                     const char* out_right_path = nullptr, int device = 0) {
  gsm_io::Image L, R;
  std::string err;
  if (!gsm_io::read_image(left_path, L, err) || !gsm_io::read_image(right_path, R, err)) return fail("remapTest", err);
  if (L.rows != R.rows || L.cols != R.cols) return fail("remapTest", "left and right images differ in size");
  const std::vector<unsigned char> g1 = gsm_io::to_gray(L), g2 = gsm_io::to_gray(R);  // Caller.cpp:41-42
  const size_t n = (size_t)L.rows * L.cols;
  std::vector<unsigned char> mb;
  if (!gsm_io::read_file(maps_path, mb, err)) return fail("remapTest", err);
  if (mb.size() != 4 * n * sizeof(float)) return fail("remapTest", "maps file must hold 4 float32 maps of the image size");
  const float* m = reinterpret_cast<const float*>(mb.data());
  Ctx ctx(device, L.rows, L.cols, 1, 1);
  if (!ctx.c) return fail("remapTest", gsm_last_error());
  std::vector<unsigned char> outL(n), outR(n);
  if (gsm_remap(ctx.c, g1.data(), m, m + n, outL.data(), L.rows, L.cols) != GSM_OK ||
      gsm_remap(ctx.c, g2.data(), m + 2 * n, m + 3 * n, outR.data(), L.rows, L.cols) != GSM_OK)
    return fail("remapTest", gsm_last_error());
  if (!gsm_io::write_gray(out_left_path, outL.data(), L.rows, L.cols, err)) return fail("remapTest", err);
  if (out_right_path && !gsm_io::write_gray(out_right_path, outR.data(), L.rows, L.cols, err)) return fail("remapTest", err);
  return 0;
}

// ---- cvtColorTest (Caller.cpp:76-112): 3-channel image -> gray through cvtColor_gpu ------------------------------------
// The weights .299/.587/.114 apply to the channels as stored (kernalCvtColor, Device.cu:136-143).  truncate selects
// cvtColor_cpu's rounding (Utility.cpp:289-298).  The reference's 1000-iteration timing loops are not reproduced.
inline int cvtColorTest(const char* src_path, const char* gray_path, bool truncate = false, int device = 0) {
  gsm_io::Image S;
  std::string err;
  if (!gsm_io::read_image(src_path, S, err)) return fail("cvtColorTest", err);
  if (S.channels < 3) return fail("cvtColorTest", "a 3-channel image is expected");
  const size_t n = (size_t)S.rows * S.cols;
  std::vector<unsigned char> rgb(3 * n), gray(n);
  for (size_t i = 0; i < n; ++i)
    for (int ch = 0; ch < 3; ++ch) rgb[3 * i + ch] = S.data[i * S.channels + ch];
  Ctx ctx(device, S.rows, S.cols, 1, 1);
  if (!ctx.c) return fail("cvtColorTest", gsm_last_error());
  if (gsm_cvtcolor(ctx.c, rgb.data(), gray.data(), S.rows, S.cols, truncate ? 1 : 0) != GSM_OK)
    return fail("cvtColorTest", gsm_last_error());
  return gsm_io::write_gray(gray_path, gray.data(), S.rows, S.cols, err) ? 0 : fail("cvtColorTest", err);
}

// ---- depth = f * B / d (the Q matrix of stereoRectify, Utility.cpp:228-234, for a rectified rig) ----------------------
// disparity image in, raw float32 depth out (0 where d == 0: occluded / no match).  fB = focal length in pixels x
// baseline (52,554 px*mm for the rig of Calib_Data_OpenCV.yml at 1280x720).
inline int depthFromDisparity(const char* disp_path, float fB, const char* depth_f32_path, int device = 0) {
  gsm_io::Image D;
  std::string err;
  if (!gsm_io::read_image(disp_path, D, err)) return fail("depthFromDisparity", err);
  const std::vector<unsigned char> d = gsm_io::to_gray(D);
  std::vector<float> depth(d.size());
  Ctx ctx(device, D.rows, D.cols, 1, 1);
  if (!ctx.c) return fail("depthFromDisparity", gsm_last_error());
  if (gsm_disparity_to_depth(ctx.c, d.data(), depth.data(), D.rows, D.cols, fB) != GSM_OK)
    return fail("depthFromDisparity", gsm_last_error());
  FILE* f = std::fopen(depth_f32_path, "wb");
  if (!f || std::fwrite(depth.data(), sizeof(float), depth.size(), f) != depth.size()) return fail("depthFromDisparity", "cannot write the depth file");
  std::fclose(f);
  return 0;
}

// ---- STMatching/main.cpp:37-70 (stereo_routine, StereoDisparity.cpp:41-56): segment-tree stereo, files in, file out --
// leftImgPath rightImgPath dispImgPath [maxLevel = 60] [scale = 4] [sigma = 0.1] [method = 0 (ST-1) | 1 (ST-2)]
inline int segmentTreeStereo(const char* left_path, const char* right_path, const char* disp_path, int max_level = 60,
                             int scale = 4, float sigma = 0.1f, int method = 0, int device = 0) {
  gsm_io::Image L, R;
  std::string err;
  if (!gsm_io::read_image(left_path, L, err) || !gsm_io::read_image(right_path, R, err)) return fail("segmentTreeStereo", err);
  if (L.rows != R.rows || L.cols != R.cols) return fail("segmentTreeStereo", "left and right images differ in size");
  if (L.channels < 3 || R.channels < 3) return fail("segmentTreeStereo", "3-channel images are expected (cv::imread default)");
  const size_t n = (size_t)L.rows * L.cols;
  std::vector<unsigned char> l3(3 * n), r3(3 * n), disp(n);
  for (size_t i = 0; i < n; ++i)
    for (int ch = 0; ch < 3; ++ch) {  // file order R, G, B -> cv::imread's B, G, R; alpha dropped
      l3[3 * i + ch] = L.data[i * L.channels + 2 - ch];
      r3[3 * i + ch] = R.data[i * R.channels + 2 - ch];
    }
  Ctx ctx(device, L.rows, L.cols, max_level, 1);
  if (!ctx.c) return fail("segmentTreeStereo", gsm_last_error());
  gsm_st_params p = {};
  p.num_disp = max_level;
  p.sigma = sigma;
  p.tau = 1200.f;  // TAU, Toolkit.h:33
  p.median_radius = 3;
  p.scale = scale;
  p.refined = method ? 1 : 0;
  const auto t0 = std::chrono::steady_clock::now();
  if (gsm_segment_tree_stereo(ctx.c, &p, l3.data(), r3.data(), disp.data(), L.rows, L.cols) != GSM_OK)
    return fail("segmentTreeStereo", gsm_last_error());
  std::printf("GPU : %g\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  return gsm_io::write_gray(disp_path, disp.data(), L.rows, L.cols, err) ? 0 : fail("segmentTreeStereo", err);
}

// ---- batch front-end: a list of stereo pairs (any mix of sizes) in, one disparity file per pair out -------------------
// list_path: one pair per line, "left right out"; the whole list runs as mixed-size batches (gsm_stereo_batch_v).
inline int batchFrames(const char* list_path, const Options& o = Options(), int max_batch = 16) {
  FILE* f = std::fopen(list_path, "r");
  if (!f) return fail("batchFrames", std::string("cannot open ") + list_path);
  std::vector<std::string> lp, rp, op;
  char a[1024], b[1024], c[1024];
  while (std::fscanf(f, "%1023s %1023s %1023s", a, b, c) == 3) { lp.push_back(a); rp.push_back(b); op.push_back(c); }
  std::fclose(f);
  const int n = (int)lp.size();
  if (!n) return fail("batchFrames", "empty list");
  std::vector<std::vector<unsigned char>> L(n), R(n), D(n), M(n);
  std::vector<int> rows(n), cols(n);
  int mr = 0, mc = 0;
  std::string err;
  for (int i = 0; i < n; ++i) {
    gsm_io::Image il, ir;
    if (!gsm_io::read_image(lp[i].c_str(), il, err) || !gsm_io::read_image(rp[i].c_str(), ir, err)) return fail("batchFrames", err);
    if (il.rows != ir.rows || il.cols != ir.cols) return fail("batchFrames", lp[i] + ": left and right differ in size");
    L[i] = gsm_io::to_gray(il);
    R[i] = gsm_io::to_gray(ir);
    rows[i] = il.rows;
    cols[i] = il.cols;
    D[i].resize(L[i].size());
    if (o.lr_check) M[i].resize(L[i].size());
    mr = mr > il.rows ? mr : il.rows;
    mc = mc > il.cols ? mc : il.cols;
  }
  Ctx ctx(o.device, mr, mc, o.num_disp, n < max_batch ? n : max_batch);
  if (!ctx.c) return fail("batchFrames", gsm_last_error());
  std::vector<const unsigned char*> lptr(n), rptr(n);
  std::vector<unsigned char*> dptr(n), mptr(n);
  for (int i = 0; i < n; ++i) { lptr[i] = L[i].data(); rptr[i] = R[i].data(); dptr[i] = D[i].data(); mptr[i] = o.lr_check ? M[i].data() : nullptr; }
  const gsm_params p = params_of(o);
  const auto t0 = std::chrono::steady_clock::now();
  if (gsm_stereo_batch_v(ctx.c, &p, n, lptr.data(), rptr.data(), dptr.data(), o.lr_check ? mptr.data() : nullptr, rows.data(),
                         cols.data()) != GSM_OK)
    return fail("batchFrames", gsm_last_error());
  std::printf("GPU : %g (%d pairs)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), n);
  for (int i = 0; i < n; ++i)
    if (!gsm_io::write_gray(op[i].c_str(), D[i].data(), rows[i], cols[i], err)) return fail("batchFrames", err);
  return 0;
}

}  // namespace gsm_caller
#endif  // GSM_CALLER_HPP
