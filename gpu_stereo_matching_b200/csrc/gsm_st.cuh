// gsm_st.cuh -- segment-tree stereo: the pipeline of the reference's STMatching project,
// stereo_disparity_normal (STMatching/StereoDisparity.cpp:58-90), on the GPU wherever it is data parallel:
//   GetMatchingCost (StereoHelper.cpp:75-129)           st_gray_grad_kernel + st_cost_kernel
//   CColorWeight (SegmentTree.cpp:183-195)              st_median3_kernel (3x3 median per channel) + st_edge_weight_kernel
//   BuildSegmentTree (SegmentTree.cpp:38-139)           edge sort: st_enumerate_edges_kernel + radix sort; Kruskal with Felzenszwalb's
//                                                       adaptive threshold (defined by its sequential edge order) and the
//                                                       breadth-first ordering: HOST (gsm_st_host.hpp); between them
//                                                       st_records_kernel, after them st_pack_kernel
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

// The first phase of the tree builder on the GPU: the grid's edges enumerated in (b, a) order -- for every pixel b in
// raster order the edge from its left neighbour, then the edge from the pixel below (segment-graph.h:33-41 sorts by
// (w, b, a); a stable sort by weight of this enumeration is that order).  code = (lower / left pixel a) << 1 | direction
// (0: (a, a+1), 1: (a, a-W)), key = the weight (u8 value, or the bits of the non-negative float).
template <typename WT>
__global__ void st_enumerate_edges_kernel(const WT* __restrict__ wr, const WT* __restrict__ wu, u32* __restrict__ key,
                                          u32* __restrict__ code, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int b = y * W + x;
  const bool down = y + 1 < H;
  // edges of the pixels before b: rows above have W - 1 left edges and W down edges each
  int e = y * (2 * W - 1) + (x > 0 ? x - 1 : 0) + (down ? x : 0);
  auto bits = [](WT v) -> u32 { return sizeof(WT) == 1 ? (u32)v : __float_as_uint((float)v + 0.0f); };  // (-0 -> +0)
  if (x > 0) { key[e] = bits(wr[b - 1]); code[e] = (u32)(b - 1) << 1; ++e; }
  if (down) { key[e] = bits(wu[b + W]); code[e] = (u32)(b + W) << 1 | 1u; }
}
// sorted keys -> the float weights the Kruskal passes compare
template <typename WT>
__global__ void st_edge_weights_from_keys_kernel(const u32* __restrict__ key, float* __restrict__ ws, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) ws[i] = sizeof(WT) == 1 ? (float)key[i] : __uint_as_float(key[i]);
}

// The middle, data-parallel phase of the tree builder (gsm_st_host.hpp: kruskal -> RECORDS -> bfs) on the GPU: per pixel
// its kept grid edges in edge-list order (SegmentTree.cpp:70-94), i.e. sorted by (w, b, a): by weight, ties in the order
// up, left, down, right.  flags[p]: bit 0 / 1 = p's right / up edge is in the tree, bit 2 / 3 = that edge carries the
// cross-segment penalty (segment-graph.h: w += 5).  wr[p] / wu[p] = weight of p's right / up edge (u8: CColorWeight,
// float: CColorDepthWeight).  rec[p] = deg | direction codes (2 bits each) << 8 | quantised distances
// min((int)(w * scale + 0.5f), 255) << 32 -- the same words detail::records() produces on the host.
template <typename WT>
__global__ void st_records_kernel(const u8* __restrict__ flags, const WT* __restrict__ wr, const WT* __restrict__ wu,
                                  unsigned long long* __restrict__ rec, int H, int W, float scale) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int p = y * W + x;
  typedef unsigned long long u64;
  const int f = flags[p], fl = x > 0 ? flags[p - 1] : 0, fd = y + 1 < H ? flags[p + W] : 0;
  auto item = [&](int used, int pen, WT wq, u64 dcode) -> u64 {
    float w = (float)wq;
    const u64 key = sizeof(WT) == 1 ? (u64)wq : (u64)__float_as_uint(w);  // weights >= 0: the bits order like the value
    if (pen) w = __fadd_rn(w, 5.0f);
    const u64 dq = (u64)min((int)__fadd_rn(__fmul_rn(w, scale), 0.5f), 255);
    return used ? (key << 12 | dcode << 10 | dcode << 8 | dq) : ~0ull;
  };
  u64 i0 = item(f & 2, f & 8, wu[p], 0);
  u64 i1 = item(fl & 1, fl & 4, x > 0 ? wr[p - 1] : wr[p], 1);
  u64 i2 = item(fd & 2, fd & 8, y + 1 < H ? wu[p + W] : wu[p], 2);
  u64 i3 = item(f & 1, f & 4, wr[p], 3);
  auto cswap = [](u64& a, u64& b) { const u64 lo = min(a, b), hi = max(a, b); a = lo; b = hi; };
  cswap(i0, i1); cswap(i2, i3); cswap(i0, i2); cswap(i1, i3); cswap(i1, i2);
  const u64 deg = (u64)((f & 2) != 0) + ((fl & 1) != 0) + ((fd & 2) != 0) + ((f & 1) != 0);
  rec[p] = deg | ((i0 >> 8) & 3) << 8 | ((i1 >> 8) & 3) << 10 | ((i2 >> 8) & 3) << 12 | ((i3 >> 8) & 3) << 14 |
           (i0 & 255) << 32 | (i1 & 255) << 40 | (i2 & 255) << 48 | (i3 & 255) << 56;
}

// The ordered tree as the host's breadth-first pass leaves it (per BFS position: pixel, father, first child, number of
// children, quantised weight of the edge to the father) -> the words the filter kernels read + the pixel -> position map.
// wtab[q] = float bits of m_table[q] = exp(-q / (255 sigma)) (CSegmentTree::UpdateTable, computed on the host).
__global__ void st_pack_kernel(const int* __restrict__ order, const int* __restrict__ father, const int* __restrict__ child0,
                               const u8* __restrict__ fdist, const u8* __restrict__ nchild, const int* __restrict__ wtab,
                               int2* __restrict__ up, int2* __restrict__ down, int* __restrict__ pos, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pos[order[i]] = i;
  up[i] = make_int2(child0[i], (int)nchild[i]);
  down[i] = make_int2(father[i], wtab[fdist[i]]);
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
  const int* level_off;
  int levels, n;
  int max_width;  // nodes of the widest level
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
        for (int z = 0; z < u.y; ++z) c = __fadd_rn(c, __fmul_rn(b[u.x + z], __int_as_float(t.down[u.x + z].y)));
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

// The same two passes with the level-to-level dependency kept ON CHIP.  A level only ever reads the level next to it
// (its children in pass 1, its fathers in pass 2), and the nodes of a level are contiguous in BFS order, so:
//   * everything a level needs that does NOT depend on the neighbouring level -- its node words, its own values, the
//     edge weights -- is a set of contiguous ranges, fetched RING - 2 levels ahead with cp.async into a ring of RING
//     level slots in shared memory (the level offsets themselves ride one step ahead in registers);
//   * the values the neighbouring level produced are read from that level's ring slot, not from L2.
// The per-level critical path drops from two dependent L2 round trips to a shared-memory read and one barrier.
// `cap` = slot capacity in nodes (>= the widest level; the host picks RING, or st_filter_kernel when nothing fits).
// Arithmetic and its order are those of st_filter_kernel (bit-identical results).
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warps 0-3 compute, warps 4-7 fetch (the two halves meet at the one barrier per level): what bounds a level is the
// instruction chain of the slowest warp, so the address arithmetic of the fetches is kept off the computing warps.
constexpr int ST_PAD = 4;  // a slot is cap + ST_PAD entries: the branch-free child loop reads up to 3 entries past a level
template <int RING>
__global__ void __launch_bounds__(256) st_filter_ring_kernel(float* __restrict__ buf, float* __restrict__ fin, StTree t,
                                                             int cap) {
  static_assert(RING >= 4, "ring = the neighbouring level + the current one + at least two in flight");
  extern __shared__ __align__(16) unsigned char st_smem[];
  const int slot = cap + ST_PAD;
  int2* s_meta = reinterpret_cast<int2*>(st_smem);                     // [RING][slot] node words
  float* s_val = reinterpret_cast<float*>(s_meta + (size_t)RING * slot);  // [RING][slot] own value -> pass-1 sum
  float* s_aux = s_val + (size_t)RING * slot;                          // [RING][slot] pass 1: edge weights; pass 2: results
  const u32 a_meta = smem_u32(s_meta), a_val = smem_u32(s_val), a_aux = smem_u32(s_aux);
  float* b = buf + (size_t)blockIdx.x * t.n;
  float* f = fin + (size_t)blockIdx.x * t.n;
  const int L = t.levels;
  const int* __restrict__ off = t.level_off;
  constexpr int AHEAD = RING - 2, HALF = 128;
  const bool producer = threadIdx.x >= HALF;
  const int tid = threadIdx.x & (HALF - 1);
  auto cp8 = [](u32 dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); };
  auto cp4 = [](u32 dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); };

  // ---- pass 1: leaves to root.  Levels run downwards, so level li is [off[li], hi) with hi = the previous level's lo.
  if (producer) {
    auto fetch = [&](int li, int lo, int hi) {
      if (li >= 0) {
        const int so = (li % RING) * slot;
        for (int i = tid; i < hi - lo; i += HALF) {
          cp8(a_meta + 8 * (so + i), t.up + lo + i);
          cp4(a_val + 4 * (so + i), b + lo + i);
          cp4(a_aux + 4 * (so + i), &t.down[lo + i].y);
        }
      }
      cp_async_commit();
    };
    int f_hi = off[L], f_lo = off[L - 1];  // the level being fetched
    for (int k = 0; k < AHEAD; ++k) {
      fetch(L - 1 - k, f_lo, f_hi);
      f_hi = f_lo;
      f_lo = L - 2 - k >= 0 ? off[L - 2 - k] : 0;
    }
    int f_nlo = L - 2 - AHEAD >= 0 ? off[L - 2 - AHEAD] : 0;
    for (int l = L - 1; l >= 0; --l) {
      cp_async_wait<AHEAD - 1>();  // this thread's share of level l has landed
      __syncthreads();             // level l+1 is complete, the slot of level l+2 is free
      const int f_nn = l - AHEAD - 2 >= 0 ? off[l - AHEAD - 2] : 0;  // used two steps on
      fetch(l - AHEAD, f_lo, f_hi);
      f_hi = f_lo; f_lo = f_nlo; f_nlo = f_nn;
    }
    cp_async_wait<0>();
  } else {
    int c_hi = off[L], c_lo = off[L - 1], c_nlo = L >= 2 ? off[L - 2] : 0;  // the level being computed
    for (int l = L - 1; l >= 0; --l) {
      __syncthreads();
      const int c_nn = l - 2 >= 0 ? off[l - 2] : 0;
      const int so = (l % RING) * slot, co = ((l + 1) % RING) * slot - c_hi;  // children (level l+1) start at position c_hi
      for (int i = tid; i < c_hi - c_lo; i += HALF) {
        const int2 u = s_meta[so + i];
        const float* cv = s_val + co + u.x;
        const float* cw = s_aux + co + u.x;
        float c = s_val[so + i];
#pragma unroll
        for (int z = 0; z < 4; ++z) {  // a grid pixel has at most 4 neighbours; entries past u.y are read and dropped
          const float cz = __fadd_rn(c, __fmul_rn(cv[z], cw[z]));
          c = z < u.y ? cz : c;
        }
        if (u.y) {
          s_val[so + i] = c;
          b[c_lo + i] = c;
        }
      }
      c_hi = c_lo; c_lo = c_nlo; c_nlo = c_nn;
    }
  }
  __syncthreads();  // pass 1's global writes are visible to pass 2's fetches; every ring slot is free

  // ---- pass 2: root to leaves
  auto offc = [&](int li) { return off[min(li, L)]; };
  if (producer) {
    auto fetch = [&](int li, int lo, int hi) {
      if (li < L) {
        const int so = (li % RING) * slot;
        for (int i = tid; i < hi - lo; i += HALF) {
          cp8(a_meta + 8 * (so + i), t.down + lo + i);
          cp4(a_val + 4 * (so + i), b + lo + i);
        }
      }
      cp_async_commit();
    };
    int f_lo = 0, f_hi = offc(1);
    for (int k = 0; k < AHEAD; ++k) {
      fetch(k, f_lo, f_hi);
      f_lo = f_hi;
      f_hi = offc(k + 2);
    }
    int f_nhi = offc(AHEAD + 2);
    for (int l = 0; l < L; ++l) {
      cp_async_wait<AHEAD - 1>();
      __syncthreads();
      const int f_nn = offc(l + AHEAD + 3);
      fetch(l + AHEAD, f_lo, f_hi);
      f_lo = f_hi; f_hi = f_nhi; f_nhi = f_nn;
    }
    cp_async_wait<0>();
  } else {
    int c_lo = 0, c_hi = offc(1), c_nhi = offc(2), p_lo = 0;
    for (int l = 0; l < L; ++l) {
      __syncthreads();
      const int c_nn = offc(l + 3);
      const int so = (l % RING) * slot, po = ((l + RING - 1) % RING) * slot - p_lo;  // fathers (level l-1) start at p_lo
      for (int i = tid; i < c_hi - c_lo; i += HALF) {
        const int2 dn = s_meta[so + i];
        const float w = __int_as_float(dn.y), cur = s_val[so + i];
        const float r = l ? __fadd_rn(__fmul_rn(w, __fsub_rn(s_aux[po + dn.x], __fmul_rn(w, cur))), cur) : cur;
        s_aux[so + i] = r;
        f[c_lo + i] = r;
      }
      p_lo = c_lo; c_lo = c_hi; c_hi = c_nhi; c_nhi = c_nn;
    }
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
