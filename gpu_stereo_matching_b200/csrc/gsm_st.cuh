// gsm_st.cuh -- segment-tree stereo: the pipeline of the reference's STMatching project,
// stereo_disparity_normal (STMatching/StereoDisparity.cpp:58-90), on the GPU wherever it is data parallel:
//   GetMatchingCost (StereoHelper.cpp:75-129)           st_gray_grad_kernel + st_cost_kernel
//   CColorWeight (SegmentTree.cpp:183-195)              st_median3_kernel (3x3 median per channel) + st_edge_weight_kernel
//   BuildSegmentTree (SegmentTree.cpp:38-139)           HOST (gsm_st_host.hpp): Kruskal with Felzenszwalb's adaptive threshold
//                                                       is defined by its sequential edge order; O(N) here (counting sort)
//   Filter (SegmentTree.cpp:148-181)                    st_filter_kernel: two level-synchronous passes over the ordered tree
//   GetDisparity_WTA (StereoHelper.cpp:131-154)         st_wta_kernel
//   MeanFilter(disparity, 3), disparity *= scale        median_kernel (gsm_util.cuh) + st_scale_kernel
// Every floating-point expression is evaluated with the reference's operand order and one rounding per operation
// (explicit __f*_rn / __d*_rn: no FMA contraction), so the stages are bit-exact to the reference compiled with
// -ffp-contract=off.
#pragma once
#include "gsm_common.cuh"

namespace gsm {

// rgb_2_gray (StereoHelper.cpp:36) and GetGradient (:38-73): gray = (uchar)(0.299 c2 + 0.587 c1 + 0.114 c0 + 0.5) in
// double; gradient = 0.5 (g[x+1] - g[x-1]) + 127.5, one-sided (no 0.5) at the two border columns.
__global__ void st_gray_grad_kernel(const u8* __restrict__ bgr, float* __restrict__ grad, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  auto gray = [&](int xx) {
    const u8* p = bgr + ((size_t)y * W + xx) * 3;
    const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)p[2]), __dmul_rn(0.587, (double)p[1])),
                                         __dmul_rn(0.114, (double)p[0])), 0.5);
    return (float)(u8)v;
  };
  float g;
  if (x == 0) g = gray(1) - gray(0) + 127.5f;
  else if (x == W - 1) g = gray(W - 1) - gray(W - 2) + 127.5f;
  else g = 0.5f * (gray(x + 1) - gray(x - 1)) + 127.5f;
  grad[(size_t)y * W + x] = g;
}

// GetMatchingCost (StereoHelper.cpp:75-129): right image and gradient shifted by d (first column replicated),
// cost = float(0.11 min(sum_c |L - S| / 3, 7) + (1.0 - 0.11) min(|gL - gS|, 2)) in double.  Output at
// out[d * dstride + pos[pixel]] (pos == nullptr: pixel-major [pixel][d], the reference layout).
__global__ void st_cost_kernel(const u8* __restrict__ L, const u8* __restrict__ R, const float* __restrict__ gL,
                               const float* __restrict__ gR, float* __restrict__ out, const int* __restrict__ pos,
                               size_t dstride, int H, int W, int D) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  const u8* l = L + p * 3;
  const float gl = gL[p];
  const size_t o = pos ? (size_t)pos[p] : p * (size_t)D;
  const double wc = 0.11, wg = 1.0 - 0.11;
  for (int d = 0; d < D; ++d) {
    const int xs = x >= d ? x - d : 0;
    const u8* s = R + ((size_t)y * W + xs) * 3;
    const int sad = abs((int)l[0] - (int)s[0]) + abs((int)l[1] - (int)s[1]) + abs((int)l[2] - (int)s[2]);
    const double cc = fmin(__ddiv_rn((double)sad, 3.0), 7.0);
    const double cg = fmin((double)fabsf(__fsub_rn(gl, gR[(size_t)y * W + xs])), 2.0);
    const float c = (float)__dadd_rn(__dmul_rn(wc, cc), __dmul_rn(wg, cg));
    if (pos) out[(size_t)d * dstride + o] = c; else out[o + d] = c;
  }
}

// GetRightMatchingCostFromLeft (StereoHelper.cpp:156-180) applied to GetMatchingCost: CR(y, x, d) = CL(y, x + dd, dd) with
// dd = min(d, W - 1 - x) (past the right border the last valid disparity is repeated), and CL(y, x + dd, dd) compares
// left pixel x + dd with right pixel x.  Same arithmetic as st_cost_kernel.  Needs D <= W like the reference.
__global__ void st_cost_right_kernel(const u8* __restrict__ L, const u8* __restrict__ R, const float* __restrict__ gL,
                                     const float* __restrict__ gR, float* __restrict__ out, const int* __restrict__ pos,
                                     size_t dstride, int H, int W, int D) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  const u8* r = R + p * 3;
  const float gr = gR[p];
  const size_t o = (size_t)pos[p];
  const double wc = 0.11, wg = 1.0 - 0.11;
  for (int d = 0; d < D; ++d) {
    const int xl = x + min(d, W - 1 - x);
    const u8* l = L + ((size_t)y * W + xl) * 3;
    const int sad = abs((int)l[0] - (int)r[0]) + abs((int)l[1] - (int)r[1]) + abs((int)l[2] - (int)r[2]);
    const double cc = fmin(__ddiv_rn((double)sad, 3.0), 7.0);
    const double cg = fmin((double)fabsf(__fsub_rn(gL[(size_t)y * W + xl], gr)), 2.0);
    out[(size_t)d * dstride + o] = (float)__dadd_rn(__dmul_rn(wc, cc), __dmul_rn(wg, cg));
  }
}

// MeanFilter(img, img, 1) of CColorWeight (SegmentTree.cpp:185; ctmf with r = 1 on 3 interleaved channels): 3x3 median
// per channel, replicate border.
__global__ void st_median3_kernel(const u8* __restrict__ src, u8* __restrict__ dst, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = min(H - 1, max(0, y + dy)), xx = min(W - 1, max(0, x + dx));
        v[(dy + 1) * 3 + dx + 1] = src[((size_t)yy * W + xx) * 3 + c];
      }
    // median of 9 = the value with exactly 4 smaller-or-equal-ranked elements: 5th smallest by partial selection
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int j = i + 1; j < 9; ++j) {
        const int lo = min(v[i], v[j]), hi = max(v[i], v[j]);
        v[i] = lo;
        v[j] = hi;
      }
    }
    dst[((size_t)y * W + x) * 3 + c] = (u8)v[4];
  }
}

// CColorWeight::GetWeight (SegmentTree.cpp:189-195): max over the channels of |a - b| between 4-neighbours of the
// median-filtered image.  wr[p]: edge (p, p+1); wu[p]: edge (p, p-W).  (255 where the edge does not exist.)
__global__ void st_edge_weight_kernel(const u8* __restrict__ img, u8* __restrict__ wr, u8* __restrict__ wu, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  const u8* a = img + p * 3;
  auto wgt = [&](const u8* b) { return max(max(abs((int)a[0] - (int)b[0]), abs((int)a[1] - (int)b[1])), abs((int)a[2] - (int)b[2])); };
  wr[p] = x + 1 < W ? (u8)wgt(a + 3) : (u8)255;
  wu[p] = y >= 1 ? (u8)wgt(a - (size_t)W * 3) : (u8)255;
}

// CColorDepthWeight::GetWeight (SegmentTree.cpp:204-218): where both pixels passed the L-R check,
// 0.5 |d0 - d1| / level + (1 - 0.5) colour / 255, else colour / 255 (colour = max channel difference of the
// median-filtered image); float, one rounding per operation.
__global__ void st_edge_weight_depth_kernel(const u8* __restrict__ img, const u8* __restrict__ disp,
                                            const u8* __restrict__ mask, float level, float* __restrict__ wr,
                                            float* __restrict__ wu, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  const u8* a = img + p * 3;
  auto wgt = [&](size_t q) {
    const u8* b = img + q * 3;
    const int col = max(max(abs((int)a[0] - (int)b[0]), abs((int)a[1] - (int)b[1])), abs((int)a[2] - (int)b[2]));
    const float cv = __fdiv_rn((float)col, 255.0f);
    if (mask[p] && mask[q]) {
      const float dv = __fdiv_rn((float)abs((int)disp[p] - (int)disp[q]), level);
      return __fadd_rn(__fmul_rn(0.5f, dv), __fmul_rn(1.0f - 0.5f, cv));
    }
    return cv;
  };
  wr[p] = x + 1 < W ? wgt(p + 1) : 0.f;
  wu[p] = y >= 1 ? wgt(p - W) : 0.f;
}

// The ordered tree (breadth-first from pixel 0, SegmentTree.cpp:97-131) as arrays indexed by BFS position:
//   father[i]  BFS position of the father (root: -1)      fw[i]  m_table[father.dist] = exp(-dist / (255 sigma))
//   child0[i]  BFS position of the first child             nchild[i]  number of children
// packed two per 8-byte word so that a node costs one load per pass
// (children of a node are consecutive in BFS order, and in the order the reference's Filter visits them);
// level_off[l] .. level_off[l+1] = the nodes of depth l.
struct StTree {
  const int2* up;    // [i] = {child0, nchild}: pass 1 reads one 8-byte word per node
  const int2* down;  // [i] = {father, float bits of the edge weight to the father}: pass 2 likewise
  const float* fw;   // [i] = weight of the edge to the father (children's weights are contiguous: fw[child0 + z])
  const int* level_off;
  int levels, n;
};

// CSegmentTree::Filter (SegmentTree.cpp:148-181) for one disparity channel per CTA: buf / fin are [D][n] in BFS order.
// Pass 1, leaves to root, level by level:  buf[i] += sum_z buf[child_z] * w_z   (children in list order, mul and add
// rounded separately).  Pass 2, root to leaves:  fin[i] = w (fin[father] - w buf[i]) + buf[i].
// One __syncthreads per level: a level only depends on the next / previous one, and the channels are independent.
// Each level is two dependent rounds of L2 loads (node word + own value, then the children / the father).
__global__ void __launch_bounds__(256) st_filter_kernel(float* __restrict__ buf, float* __restrict__ fin, StTree t) {
  float* b = buf + (size_t)blockIdx.x * t.n;
  float* f = fin + (size_t)blockIdx.x * t.n;
  for (int l = t.levels - 2; l >= 0; --l) {  // the deepest level has no children
    const int lo = t.level_off[l], hi = t.level_off[l + 1];
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const int2 u = t.up[i];
      float c = b[i];
      if (u.y) {
        for (int z = 0; z < u.y; ++z) c = __fadd_rn(c, __fmul_rn(b[u.x + z], t.fw[u.x + z]));
        b[i] = c;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) f[0] = b[0];
  __syncthreads();
  for (int l = 1; l < t.levels; ++l) {
    const int lo = t.level_off[l], hi = t.level_off[l + 1];
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const int2 dn = t.down[i];
      const float w = __int_as_float(dn.y), cur = b[i];
      f[i] = __fadd_rn(__fmul_rn(w, __fsub_rn(f[dn.x], __fmul_rn(w, cur))), cur);
    }
    __syncthreads();
  }
}

// [D][n] BFS order -> [pixel][D] (the reference's volume layout), for the stage export
__global__ void st_unpermute_kernel(const float* __restrict__ fin, const int* __restrict__ order, float* __restrict__ out,
                                    int n, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
  if (i >= n) return;
  out[(size_t)order[i] * D + d] = fin[(size_t)d * n + i];
}
// [pixel][D] -> [D][n] BFS order
__global__ void st_permute_kernel(const float* __restrict__ vol, const int* __restrict__ order, float* __restrict__ buf,
                                  int n, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
  if (i >= n) return;
  buf[(size_t)d * n + i] = vol[(size_t)order[i] * D + d];
}

// GetDisparity_WTA (StereoHelper.cpp:131-154): argmin over d, strict '<', first minimum wins; thread = BFS position
__global__ void st_wta_kernel(const float* __restrict__ fin, const int* __restrict__ order, u8* __restrict__ disp, int n,
                              int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float best = fin[i];
  int bd = 0;
  for (int d = 1; d < D; ++d) {
    const float v = fin[(size_t)d * n + i];
    if (v < best) { best = v; bd = d; }
  }
  disp[order[i]] = (u8)bd;
}

// disparity *= scale on CV_8U (StereoDisparity.cpp:87): saturating
__global__ void st_scale_kernel(u8* __restrict__ d, size_t n, int scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (u8)min(255, (int)d[i] * scale);
}

}  // namespace gsm
