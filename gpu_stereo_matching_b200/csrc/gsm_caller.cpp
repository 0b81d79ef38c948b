// gsm_caller -- command-line front-end that replaces the GUI demos of the reference (BlockMatching/Main.cpp:3-9 calls
// singleFrame(); remapTest(); cvtColorTest(); -- Caller.cpp:9-112): files in, files out, no imshow / waitKey.
//   gsm_caller singleFrame left right disp [--gf r] [--disp D] [--radius r] [--lr] [--median m] [--mask file] [--device i]
//   gsm_caller remapTest left right maps.f32 out_left [out_right]
//   gsm_caller cvtColorTest src gray [--truncate]
//   gsm_caller depth disp fB depth.f32
//   gsm_caller stmatching left right disp [maxLevel=60] [scale=4] [sigma=0.1] [method=0|1]   (STMatching/main.cpp:37-70)
//   gsm_caller gray src gray                          (host only: imread + cvtColor(BGR2GRAY) of Caller.cpp:12-16)
//   gsm_caller batch list.txt [same options as singleFrame]
//   gsm_caller                      (no arguments: the reference's singleFrame() with its own relative paths)
#include <cstdlib>
#include <cstring>

#include "../../include/gsm_caller.hpp"

static gsm_caller::Options parse(int argc, char** argv, int first, const char** mask) {
  gsm_caller::Options o;
  for (int i = first; i < argc; ++i) {
    auto val = [&](int dflt) { return i + 1 < argc ? std::atoi(argv[++i]) : dflt; };
    if (!std::strcmp(argv[i], "--gf")) { o.mode = GSM_MODE_GF; o.radius = val(9); }
    else if (!std::strcmp(argv[i], "--radius")) o.radius = val(5);
    else if (!std::strcmp(argv[i], "--disp")) o.num_disp = val(64);
    else if (!std::strcmp(argv[i], "--lr")) o.lr_check = 1;
    else if (!std::strcmp(argv[i], "--median")) o.median_radius = val(3);
    else if (!std::strcmp(argv[i], "--device")) o.device = val(0);
    else if (!std::strcmp(argv[i], "--mask") && i + 1 < argc && mask) *mask = argv[++i];
  }
  return o;
}

int main(int argc, char** argv) {
  if (argc < 2) return gsm_caller::singleFrame();
  const char* cmd = argv[1];
  if (!std::strcmp(cmd, "singleFrame") && argc >= 5) {
    const char* mask = nullptr;
    const gsm_caller::Options o = parse(argc, argv, 5, &mask);
    return gsm_caller::singleFrame(argv[2], argv[3], argv[4], o, mask);
  }
  if (!std::strcmp(cmd, "remapTest") && argc >= 6) return gsm_caller::remapTest(argv[2], argv[3], argv[4], argv[5], argc > 6 ? argv[6] : nullptr);
  if (!std::strcmp(cmd, "cvtColorTest") && argc >= 4)
    return gsm_caller::cvtColorTest(argv[2], argv[3], argc > 4 && !std::strcmp(argv[4], "--truncate"));
  if (!std::strcmp(cmd, "depth") && argc >= 5) return gsm_caller::depthFromDisparity(argv[2], (float)std::atof(argv[3]), argv[4]);
  if (!std::strcmp(cmd, "stmatching") && argc >= 5)
    return gsm_caller::segmentTreeStereo(argv[2], argv[3], argv[4], argc > 5 ? std::atoi(argv[5]) : 60,
                                         argc > 6 ? std::atoi(argv[6]) : 4, argc > 7 ? (float)std::atof(argv[7]) : 0.1f,
                                         argc > 8 ? std::atoi(argv[8]) : 0);
  if (!std::strcmp(cmd, "gray") && argc >= 4) {
    gsm_io::Image img;
    std::string err;
    if (!gsm_io::read_image(argv[2], img, err)) return gsm_caller::fail("gray", err);
    const std::vector<unsigned char> g = gsm_io::to_gray(img);
    return gsm_io::write_gray(argv[3], g.data(), img.rows, img.cols, err) ? 0 : gsm_caller::fail("gray", err);
  }
  if (!std::strcmp(cmd, "batch") && argc >= 3) return gsm_caller::batchFrames(argv[2], parse(argc, argv, 3, nullptr));
  std::fprintf(stderr,
               "usage: gsm_caller singleFrame left right disp [--gf r] [--disp D] [--radius r] [--lr] [--median m] [--mask f]\n"
               "       gsm_caller remapTest left right maps.f32 out_left [out_right]\n"
               "       gsm_caller cvtColorTest src gray [--truncate]\n"
               "       gsm_caller depth disp fB depth.f32\n"
               "       gsm_caller stmatching left right disp [maxLevel] [scale] [sigma] [method]\n"
               "       gsm_caller gray src gray\n"
               "       gsm_caller batch list.txt [options]\n");
  return 2;
}
