// gsm_st_host.hpp -- host side of the segment-tree stereo (gsm_st.cuh): the tree itself.
//
// CSegmentTree::BuildSegmentTree (STMatching/SegmentTree.cpp:38-139) with segment_graph (segment-graph.h:48-101) is
// Kruskal's algorithm with Felzenszwalb's adaptive merge threshold: whether an edge joins two components depends on
// the sizes of the components all lighter edges have formed, i.e. on the sequential order of the sorted edge list.
// That sequential core -- the two Kruskal passes -- and the breadth-first ordering are the host's part; the sort before
// them and the per-pixel step between them are data parallel and, in the product pipeline (st_build in gsm_api.cu), run
// on the GPU (gsm_st.cuh: st_enumerate_edges_kernel + radix sort, st_records_kernel).  This header holds the two host
// phases plus host versions of the other two (build_tree: the host-only gsm_st_build_tree_host, which the CPU tests
// compare with the reference's tree).  Everything is O(N) and built around the fact that the graph is the pixel GRID:
//   * an edge is one 32-bit word (lower/left pixel << 1 | direction); the reference's std::sort by (w, b, a)
//     (segment-graph.h:33-41) is a stable sort by weight of the edges enumerated in (b, a) order (here: a counting sort
//     over CColorWeight's integer weights);
//   * which edges the two Kruskal passes keep (and which the second one penalises) is four bits per pixel;
//   * the reference's adjacency lists "in edge order" (SegmentTree.cpp:70-94) are, per pixel, its <= 4 kept grid edges
//     sorted by (w, b, a) -- a local sort done in one raster pass, no scattered list building;
//   * the breadth-first ordering (SegmentTree.cpp:97-131) reads one 8-byte record per node; in a tree the only visited
//     neighbour is the father, so there is no visited map.
// The result is the reference's ordered tree node for node: same father, same quantised edge weight, same order of the
// children (tests/test_oracle.py compares it with the reference's own tree).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

namespace gsm_st {

struct Tree {
  std::vector<int> order;       // BFS position -> pixel id                  (m_tree[i].id)
  std::vector<int> father;      // BFS position -> BFS position of the father, -1 for the root
  std::vector<int> father_id;   // BFS position -> pixel id of the father     (m_tree[i].father.id, root: 0)
  std::vector<uint8_t> fdist;   // BFS position -> quantised weight of the edge to the father (m_tree[i].father.dist)
  std::vector<int> child0;      // BFS position -> BFS position of the first child
  std::vector<uint8_t> nchild;  // BFS position -> number of children
  std::vector<int> level_off;   // level l = BFS positions [level_off[l], level_off[l+1])
};

namespace detail {

// Union-find with path compression (disjoint-set.h:36-82).  The reference unions by rank; only connectivity and the
// component sizes are observable, and those do not depend on the variant -- union by size needs no rank array.
// The chains stay short (0.7 hops per find on natural images), so what a textbook loop costs here is its branch
// mispredictions, not its loads: find() always takes two hops (a root points at itself, extra hops are harmless) and
// falls into a loop only when that was not enough.
struct Sets {
  int* p;
  int* size;
  float* thr;  // the component's merge threshold, read at roots only
  int find(int x) const {
    const int x1 = p[x];
    int y = p[x1];
    if (__builtin_expect(p[y] != y, 0)) {
      while (y != p[y]) y = p[y];
      int z = x1;
      while (p[z] != y) { const int nz = p[z]; p[z] = y; z = nz; }
    }
    p[x] = y;
    return y;
  }
};

struct Work {  // reused across calls: fresh multi-megabyte allocations page-fault every time
  std::vector<uint32_t> code;
  std::vector<float> ws;
  std::vector<int> p, size;
  std::vector<float> thr;
  std::vector<uint8_t> flags;
  std::vector<uint64_t> rec;
};
inline Work& work() {
  static thread_local Work w;
  return w;
}
// the cache is for streams of ordinary frames (tens of bytes per pixel); past 4 Mpx it is given back after the call
inline void release_if_large(Work& k, int n) {
  if (n > (1 << 22)) k = Work();
}

enum : uint8_t { USED_R = 1, USED_U = 2, PEN_R = 4, PEN_U = 8 };

// The builder runs in three phases so that the middle one -- per-pixel, data parallel -- can run on the GPU
// (st_records_kernel in gsm_st.cuh computes the same records from the same inputs):
//   kruskal()  edges (code[i], ws[i]) sorted by (w, b, a)    ->  k.flags, four bits per pixel
//   records()  flags + weights                               ->  k.rec, per pixel its kept edges in edge-list order
//   bfs()      records                                       ->  the ordered tree
inline void kruskal(Work& k, const uint32_t* code, const float* ws, int H, int Wd, int m, float tau) {
  const int n = H * Wd;
  k.p.resize(n); k.size.resize(n); k.thr.resize(n);
  k.flags.assign(n, 0);
  uint8_t* flags = k.flags.data();
  Sets u{k.p.data(), k.size.data(), k.thr.data()};
  for (int i = 0; i < n; ++i) { u.p[i] = i; u.size[i] = 1; u.thr[i] = tau / 1.0f; }
  // ---- segment_graph (segment-graph.h:48-101), pass 1: adaptive-threshold Kruskal
  // (branch-free body: whether an edge merges is a coin flip the predictor loses; a rejected edge "joins" a root to itself)
  int sets = n;
  for (int i = 0; i < m; ++i) {
    const int a = (int)(code[i] >> 1), dir = (int)(code[i] & 1), b = dir ? a - Wd : a + 1;
    const int ra = u.find(a), rb = u.find(b);
    const float w = ws[i];
    const bool ok = (ra != rb) & (w <= u.thr[ra]) & (w <= u.thr[rb]);
    const bool sw = u.size[ra] > u.size[rb];
    const int sm = sw ? rb : ra, bg = sw ? ra : rb;  // the smaller root goes under the bigger one
    u.p[sm] = ok ? bg : sm;
    const int sz = u.size[bg] + (ok ? u.size[sm] : 0);
    u.size[bg] = sz;
    const float th = w + tau / (float)sz;
    u.thr[bg] = ok ? th : u.thr[bg];
    flags[a] |= (uint8_t)((int)ok << dir);  // USED_R / USED_U
    sets -= (int)ok;
  }
  // ---- pass 2: the remaining edges, in the same order, join the segments into ONE tree; an edge between two segments
  // of more than MIN_SIZE_SEG pixels is penalised.  Labels first: with every pixel pointing at its root, an edge inside
  // a segment (most of them) is rejected by one comparison, and the finds that remain are one hop.
  if (sets > 1) {
    for (int i = 0; i < n; ++i) u.p[i] = u.find(i);
    for (int i = 0; i < m && sets > 1; ++i) {  // once one component is left no further edge can join anything
      const int a = (int)(code[i] >> 1), dir = (int)(code[i] & 1), b = dir ? a - Wd : a + 1;
      if (u.p[a] == u.p[b]) continue;  // same segment (this also covers every edge pass 1 used)
      const int ra = u.find(a), rb = u.find(b);
      if (ra != rb) {
        const int size_min = std::min(u.size[ra], u.size[rb]);
        const bool sw = u.size[ra] > u.size[rb];
        const int sm = sw ? rb : ra, bg = sw ? ra : rb;
        u.p[sm] = bg;
        u.size[bg] += u.size[sm];
        --sets;
        flags[a] |= (uint8_t)((USED_R << dir) | (size_min > 50 ? PEN_R << dir : 0));  // MIN_SIZE_SEG, PENALTY_CROSS_SEG
      }
    }
  }
}

// wgt(p, dir) = weight of pixel p's right (0) / up (1) edge.
template <class W>
inline void records(Work& k, int H, int Wd, float scale, const W& wgt) {
  const int n = H * Wd;
  const uint8_t* flags = k.flags.data();
  // ---- per pixel: its kept edges in edge-list order (SegmentTree.cpp:70-94).  Ties in w are ordered by (b, a): the
  // up neighbour (b = p - W), the left one (b = p, a = p - 1), the one below (b = p, a = p + W), the right one (b = p + 1).
  // rec = deg | dir codes (2 bits each) << 8 | quantised distances << 32
  k.rec.resize(n);
  uint64_t* rec = k.rec.data();
  uint8_t qlut[261];  // integer weights: min((int)((w [+ 5]) * scale + 0.5f), 255), w + 5 is exact in float
  for (int v = 0; v < 261; ++v) qlut[v] = (uint8_t)std::min((int)((float)v * scale + 0.5f), 255);
  // item = sort key (weight bits, then tie rank) << 10 | direction << 8 | quantised distance; an absent edge sorts last
  // (weights are >= 0: the bit pattern of the float orders like the value)
  auto item = [&](bool used, bool pen, int q, int dir, uint64_t dcode) -> uint64_t {
    const auto wq = wgt(q, dir);
    uint64_t key, dq;
    if (std::is_integral<decltype(wq)>::value) {
      key = (uint64_t)wq;
      dq = qlut[(int)wq + (pen ? 5 : 0)];
    } else {
      float w = (float)wq;
      uint32_t wb;
      std::memcpy(&wb, &w, 4);
      key = wb;
      if (pen) w += 5.0f;
      dq = (uint64_t)std::min((int)(w * scale + 0.5f), 255);
    }
    const uint64_t it = key << 12 | dcode << 10 | dcode << 8 | dq;
    return used ? it : ~0ull;
  };
  auto cswap = [](uint64_t& x, uint64_t& y) { const uint64_t lo = std::min(x, y), hi = std::max(x, y); x = lo; y = hi; };
  for (int y = 0, p = 0; y < H; ++y)
    for (int x = 0; x < Wd; ++x, ++p) {
      const int f = flags[p], fl = x > 0 ? flags[p - 1] : 0, fd = y + 1 < H ? flags[p + Wd] : 0;
      uint64_t i0 = item(f & USED_U, f & PEN_U, p, 1, 0);
      uint64_t i1 = item(fl & USED_R, fl & PEN_R, x > 0 ? p - 1 : p, 0, 1);
      uint64_t i2 = item(fd & USED_U, fd & PEN_U, y + 1 < H ? p + Wd : p, 1, 2);
      uint64_t i3 = item(f & USED_R, f & PEN_R, p, 0, 3);
      cswap(i0, i1); cswap(i2, i3); cswap(i0, i2); cswap(i1, i3); cswap(i1, i2);
      const uint64_t deg = (uint64_t)((f & USED_U) != 0) + ((fl & USED_R) != 0) + ((fd & USED_U) != 0) + ((f & USED_R) != 0);
      rec[p] = deg | ((i0 >> 8) & 3) << 8 | ((i1 >> 8) & 3) << 10 | ((i2 >> 8) & 3) << 12 | ((i3 >> 8) & 3) << 14 |
               (i0 & 255) << 32 | (i1 & 255) << 40 | (i2 & 255) << 48 | (i3 & 255) << 56;
    }
}

// rec[p] = deg | dir codes (2 bits each, U L D R = 0 1 2 3) << 8 | quantised distances << 32 (records() / st_records_kernel)
inline void bfs(const uint64_t* rec, int H, int Wd, Tree& t) {
  const int n = H * Wd;
  // ---- ordered tree: breadth-first from pixel 0 (SegmentTree.cpp:97-131)
  // (branch-free body: all four slots of a record are written, `end` only moves past the real children; the arrays
  // carry 4 spare entries for the writes past the last node)
  t.order.resize(n + 4); t.father.resize(n + 4); t.fdist.resize(n + 4);
  t.child0.resize(n); t.nchild.resize(n); t.level_off.clear();
  int* order = t.order.data();
  int* father = t.father.data();
  uint8_t* fdist = t.fdist.data();
  const int delta[4] = {-Wd, -1, Wd, 1};
  order[0] = 0; father[0] = -1; fdist[0] = 0;
  t.level_off.push_back(0);
  int end = 1, level_end = 1;
  for (int pos = 0; pos < n; ++pos) {
    if (pos == level_end) { t.level_off.push_back(pos); level_end = end; }
    if (pos + 8 < end) __builtin_prefetch(&rec[order[pos + 8]]);
    const int id = order[pos], fid = pos ? order[father[pos]] : -1;  // (the father was ordered long ago: a cache hit)
    const uint64_t r = rec[id];
    const int deg = (int)(r & 7), e0 = end;
    for (int j = 0; j < 4; ++j) {
      const int c = id + delta[(r >> (8 + 2 * j)) & 3];
      order[end] = c;
      father[end] = pos;
      fdist[end] = (uint8_t)(r >> (32 + 8 * j));
      end += (int)((j < deg) & (c != fid));
    }
    t.child0[pos] = e0;
    t.nchild[pos] = (uint8_t)(end - e0);
  }
  t.order.resize(n); t.father.resize(n); t.fdist.resize(n);
  t.level_off.push_back(n);
}
// m_tree[i].father.id for the callers that report it (the pipeline itself works on BFS positions)
inline void father_ids(Tree& t) {
  const size_t n = t.order.size();
  t.father_id.resize(n);
  if (n) t.father_id[0] = 0;
  for (size_t i = 1; i < n; ++i) t.father_id[i] = t.order[t.father[i]];
}

template <class W>
inline void finish(Work& k, int H, int Wd, int m, float tau, float scale, const W& wgt, Tree& t) {
  kruskal(k, k.code.data(), k.ws.data(), H, Wd, m, tau);
  records(k, H, Wd, scale, wgt);
  bfs(k.rec.data(), H, Wd, t);
  father_ids(t);
}

}  // namespace detail

// wr[p]: weight of edge (p, p+1) for x < W-1; wu[p]: weight of edge (p, p-W) for y >= 1 (st_edge_weight_kernel).
// tau: the constant c of the threshold function c / size (TAU = 1200 in Toolkit.h:33); scale: CWeightProvider::GetScale().
// k: the builder's work space (default: one per calling thread; callers that build on short-lived threads pass their own).
// edges of the grid in the reference's sorted order -> k.code / k.ws; returns their number
inline int sort_edges(const uint8_t* wr, const uint8_t* wu, int H, int W, detail::Work& k) {
  // ---- edges in the reference's sorted order: by weight, then by b, then by a (segment-graph.h:33-41): a counting
  // sort whose buckets are filled in (b, a) order by construction
  int cnt[257] = {0};
  for (int y = 0, p = 0; y < H; ++y)
    for (int x = 0; x < W; ++x, ++p) {
      if (x < W - 1) cnt[wr[p] + 1]++;
      if (y >= 1) cnt[wu[p] + 1]++;
    }
  for (int i = 0; i < 256; ++i) cnt[i + 1] += cnt[i];
  const int m = cnt[256];
  k.code.resize(m);
  k.ws.resize(m);
  uint32_t* code = k.code.data();
  float* ws = k.ws.data();
  for (int y = 0, b = 0; y < H; ++y)  // edges whose second endpoint is b, by increasing first endpoint a
    for (int x = 0; x < W; ++x, ++b) {
      if (x >= 1) { const int a = b - 1, i = cnt[wr[a]]++; code[i] = (uint32_t)a << 1; ws[i] = (float)wr[a]; }            // (a, a+1)
      if (y + 1 < H) { const int a = b + W, i = cnt[wu[a]]++; code[i] = (uint32_t)a << 1 | 1u; ws[i] = (float)wu[a]; }   // (a, a-W)
    }
  return m;
}
inline void build_tree(const uint8_t* wr, const uint8_t* wu, int H, int W, float tau, float scale, Tree& t,
                       detail::Work& k = detail::work()) {
  const int m = sort_edges(wr, wu, H, W, k);
  detail::finish(k, H, W, m, tau, scale, [&](int p, int dir) { return dir ? wu[p] : wr[p]; }, t);
  detail::release_if_large(k, H * W);
}

// m_table of CSegmentTree::UpdateTable (SegmentTree.cpp:141-146): exp(-i / (255 sigma)) in float, sigma >= 0.01
inline void weight_table(float sigma, float (&table)[256]) {
  sigma = std::max(0.01f, sigma);
  for (int i = 0; i <= 255; ++i) table[i] = std::exp(-float(i) / (255 * sigma));
}

}  // namespace gsm_st
