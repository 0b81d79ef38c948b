// gsm_st_host.hpp -- host side of the segment-tree stereo (gsm_st.cuh): the tree itself.
//
// CSegmentTree::BuildSegmentTree (STMatching/SegmentTree.cpp:38-139) with segment_graph (segment-graph.h:48-101) is
// Kruskal's algorithm with Felzenszwalb's adaptive merge threshold: whether an edge joins two components depends on
// the sizes of the components all lighter edges have formed, i.e. on the sequential order of the sorted edge list.
// It is therefore built on the host -- but in O(N): the edge weights of CColorWeight are integers 0..255, so the
// reference's std::sort by (w, b, a) (segment-graph.h:33-41) is a counting sort over w whose buckets are filled in
// (b, a) order by construction.  The result is the reference's ordered tree (breadth-first from pixel 0,
// SegmentTree.cpp:97-131), node for node: same father, same quantised edge weight, same order of the children.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
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

struct DisjointSets {  // union by rank with path compression (disjoint-set.h:36-82); only connectivity and component
  std::vector<int> p;   // sizes are observable, and those do not depend on the variant.  The parent array is on its own
  struct Info { int size, rank; float thr; };  // (find touches nothing else); size, rank and the component's merge
  std::vector<Info> info;                       // threshold are only read at roots
  DisjointSets(int n, float thr0) : p(n), info(n) {
    for (int i = 0; i < n; ++i) { p[i] = i; info[i].size = 1; info[i].rank = 0; info[i].thr = thr0; }
  }
  int find(int x) {
    int y = p[x];
    if (y == x) return x;
    while (y != p[y]) y = p[y];
    while (p[x] != y) { const int nx = p[x]; p[x] = y; x = nx; }
    return y;
  }
  int join(int a, int b) {  // a, b roots; returns the new root
    if (info[a].rank > info[b].rank) std::swap(a, b);
    p[a] = b;
    info[b].size += info[a].size;
    if (info[a].rank == info[b].rank) info[b].rank++;
    return b;
  }
};

struct Edge {
  int a, b;
  float w;
};

inline void finish_tree(std::vector<Edge>& e, int H, int W, float tau, float scale, Tree& t);

// wr[p]: weight of edge (p, p+1) for x < W-1; wu[p]: weight of edge (p, p-W) for y >= 1 (st_edge_weight_kernel).
// tau: the constant c of the threshold function c / size (TAU = 1200 in Toolkit.h:33); scale: CWeightProvider::GetScale().
inline void build_tree(const uint8_t* wr, const uint8_t* wu, int H, int W, float tau, float scale, Tree& t) {
  const int n = H * W;
  // ---- edges in the reference's sorted order: by weight, then by b, then by a (segment-graph.h:33-41)
  static thread_local std::vector<Edge> e;  // work space reused across calls: fresh 4 MB allocations page-fault every time
  std::vector<int> cnt(257, 0);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int p = y * W + x;
      if (x < W - 1) cnt[wr[p] + 1]++;
      if (y >= 1) cnt[wu[p] + 1]++;
    }
  for (int i = 0; i < 256; ++i) cnt[i + 1] += cnt[i];
  const int m = cnt[256];
  e.resize(m);
  {
    std::vector<int> at(cnt.begin(), cnt.begin() + 256);
    for (int y = 0, b = 0; y < H; ++y)  // edges whose second endpoint is b, by increasing first endpoint a
      for (int x = 0; x < W; ++x, ++b) {
        if (x >= 1) { const int a = b - 1; Edge& d = e[at[wr[a]]++]; d.a = a; d.b = b; d.w = (float)wr[a]; }          // (a, a+1)
        if (y + 1 < H) { const int a = b + W; Edge& d = e[at[wu[a]]++]; d.a = a; d.b = b; d.w = (float)wu[a]; }       // (a, a-W)
      }
  }
  finish_tree(e, H, W, tau, scale, t);
}

// The same for real-valued weights (CColorDepthWeight, SegmentTree.cpp:204-218): a comparison sort by (w, b, a).
inline void build_tree_f(const float* wr, const float* wu, int H, int W, float tau, float scale, Tree& t) {
  std::vector<Edge> e;
  e.reserve(2 * (size_t)H * W);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int p = y * W + x;
      if (x < W - 1) e.push_back(Edge{p, p + 1, wr[p]});
      if (y >= 1) e.push_back(Edge{p, p - W, wu[p]});
    }
  std::sort(e.begin(), e.end(), [](const Edge& x, const Edge& y) {
    if (x.w != y.w) return x.w < y.w;
    if (x.b != y.b) return x.b < y.b;
    return x.a < y.a;
  });
  finish_tree(e, H, W, tau, scale, t);
}

inline void finish_tree(std::vector<Edge>& e, int H, int W, float tau, float scale, Tree& t) {
  const int n = H * W;
  const int m = (int)e.size();
  // ---- segment_graph (segment-graph.h:48-101): adaptive-threshold Kruskal, then the remaining edges in the same order
  // join the segments into ONE tree; an edge between two segments of more than MIN_SIZE_SEG pixels is penalised
  static thread_local std::vector<uint8_t> used, adjd, deg, seen;
  static thread_local std::vector<int> adj, depth;
  used.assign(m, 0);
  {
    DisjointSets u(n, tau / 1.0f);
    int sets = n;
    for (int i = 0; i < m; ++i) {
      const int a = u.find(e[i].a), b = u.find(e[i].b);
      if (a != b && e[i].w <= u.info[a].thr && e[i].w <= u.info[b].thr) {
        used[i] = 1;
        const int r = u.join(a, b);
        u.info[r].thr = e[i].w + tau / (float)u.info[r].size;
        --sets;
      }
    }
    for (int i = 0; i < m && sets > 1; ++i) {  // once one component is left no further edge can join anything
      if (used[i]) continue;                   // its endpoints are already connected
      const int a = u.find(e[i].a), b = u.find(e[i].b);
      if (a != b) {
        const int size_min = std::min(u.info[a].size, u.info[b].size);
        u.join(a, b);
        used[i] = 1;
        --sets;
        if (size_min > 50) e[i].w += 5.0f;  // MIN_SIZE_SEG, PENALTY_CROSS_SEG
      }
    }
  }
  // ---- adjacency in edge order (SegmentTree.cpp:70-94): at most 4 neighbours per pixel
  adj.resize(4 * (size_t)n);
  adjd.resize(4 * (size_t)n);
  deg.assign(n, 0);
  for (int i = 0; i < m; ++i) {
    if (!used[i]) continue;
    const int dis = std::min((int)(e[i].w * scale + 0.5f), 255);
    const int a = e[i].a, b = e[i].b;
    adj[4 * (size_t)a + deg[a]] = b; adjd[4 * (size_t)a + deg[a]++] = (uint8_t)dis;
    adj[4 * (size_t)b + deg[b]] = a; adjd[4 * (size_t)b + deg[b]++] = (uint8_t)dis;
  }
  // ---- ordered tree: breadth-first from pixel 0 (SegmentTree.cpp:97-131)
  t.order.assign(n, 0); t.father.assign(n, -1); t.father_id.assign(n, 0); t.fdist.assign(n, 0);
  t.child0.assign(n, 0); t.nchild.assign(n, 0); t.level_off.clear();
  depth.assign(n, 0);
  seen.assign(n, 0);
  seen[0] = 1;
  int start = 0, end = 1;
  while (start < end) {
    const int pos = start++, id = t.order[pos];
    t.child0[pos] = end;
    for (int k = 0; k < deg[id]; ++k) {
      const int c = adj[4 * (size_t)id + k];
      if (seen[c]) continue;  // the father
      seen[c] = 1;
      t.order[end] = c;
      t.father[end] = pos;
      t.father_id[end] = id;
      t.fdist[end] = adjd[4 * (size_t)id + k];
      depth[end] = depth[pos] + 1;
      ++end;
    }
    t.nchild[pos] = (uint8_t)(end - t.child0[pos]);
  }
  for (int i = 0; i < n; ++i)
    if (i == 0 || depth[i] != depth[i - 1]) t.level_off.push_back(i);
  t.level_off.push_back(n);
}

// m_table of CSegmentTree::UpdateTable (SegmentTree.cpp:141-146): exp(-i / (255 sigma)) in float, sigma >= 0.01
inline void weight_table(float sigma, float (&table)[256]) {
  sigma = std::max(0.01f, sigma);
  for (int i = 0; i <= 255; ++i) table[i] = std::exp(-float(i) / (255 * sigma));
}

}  // namespace gsm_st
