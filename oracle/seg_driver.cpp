// extern "C" driver around the reference's OWN segment-tree stereo (STMatching/SegmentTree.cpp, segment-graph.h,
// disjoint-set.h, StereoHelper.cpp, Toolkit.cpp, StereoDisparity.cpp and ctmf.c, all compiled UNMODIFIED from
// /root/reference against the stand-in oracle/shim_st/cvshim_st.hpp).  TEST INFRASTRUCTURE ONLY: pins the segment-tree
// path of the product.  Never linked into the product library.
#define private public  // the ordered tree (CSegmentTree::m_tree, SegmentTree.h:64) is exported for the tree-builder test
#include "SegmentTree.h"
#undef private
#include "StereoDisparity.h"
#include "StereoHelper.h"
#include "Toolkit.h"

#include <cstdint>
#include <cstring>

extern "C" {
// GetMatchingCost, StereoHelper.cpp:75-129: colour + gradient cost, float [h][w][D] pixel-major
void ref_st_matching_cost(const uint8_t* bgrL, const uint8_t* bgrR, int w, int h, int D, float* out) {
  CDisparityHelper hlp;
  cv::Mat l(h, w, CV_8UC3, (void*)bgrL), r(h, w, CV_8UC3, (void*)bgrR);
  cv::Mat vol = hlp.GetMatchingCost(l, r, D);
  memcpy(out, vol.data, sizeof(float) * (size_t)w * h * D);
}
// CColorWeight + BuildSegmentTree + Filter (SegmentTree.cpp:38-139,148-195): aggregates cost [h][w][D] in place.
// order / father / fdist (optional, w*h entries each): the ordered tree -- node id, father id and quantised edge
// weight in breadth-first order (m_tree).
void ref_st_filter(const uint8_t* bgr, int w, int h, int D, float sigma, float tau, float* cost, int* order, int* father,
                   uint8_t* fdist) {
  cv::Mat img(h, w, CV_8UC3, (void*)bgr);
  CColorWeight cw(img);
  CSegmentTree st;
  st.BuildSegmentTree(cv::Size(w, h), sigma, tau, cw);
  if (cost) {
    cv::Mat vol(1, w * h * D, CV_32F, (void*)cost);
    st.Filter(vol, D);
  }
  if (order)
    for (int i = 0; i < w * h; ++i) {
      order[i] = st.m_tree[i].id;
      father[i] = st.m_tree[i].father.id;
      fdist[i] = st.m_tree[i].father.dist;
    }
}
// stereo_disparity_normal, StereoDisparity.cpp:58-90: cost -> segment-tree aggregation -> WTA -> 7x7 median -> * scale
void ref_st_routine(const uint8_t* bgrL, const uint8_t* bgrR, int w, int h, int D, int scale, float sigma, uint8_t* disp) {
  cv::Mat l(h, w, CV_8UC3, (void*)bgrL), r(h, w, CV_8UC3, (void*)bgrR), d(h, w, CV_8U, (void*)disp);
  stereo_disparity_normal(l, r, d, D, scale, sigma);
}
// stereo_disparity_iteration, StereoDisparity.cpp:92-160: two-pass version with the L-R check (:128-147) in the middle
void ref_st_iteration(const uint8_t* bgrL, const uint8_t* bgrR, int w, int h, int D, int scale, float sigma, uint8_t* disp) {
  cv::Mat l(h, w, CV_8UC3, (void*)bgrL), r(h, w, CV_8UC3, (void*)bgrR), d(h, w, CV_8U, (void*)disp);
  stereo_disparity_iteration(l, r, d, D, scale, sigma);
}
}
