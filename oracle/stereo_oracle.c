/*
 * stereo_oracle.c -- CPU restatement of the reference's BlockMatching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library (libgsm.so)
 * never links, imports or calls anything under oracle/.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - orc_ad_volume / orc_sad_* / orc_all_sad : PINNED. Checked bit-exact against the
 *     reference's own BlockMatching.cpp compiled unmodified (oracle/_ref/libref.so) on the
 *     Art pairs and on seeded random inputs, and against the committed digests in
 *     tests/golden/ (generated from oracle/_ref).
 *   - orc_median : PINNED against the reference's ctmf.c compiled unmodified.
 *   - right-view costs (orc_ad_slice view 1) and the float WTA rule of orc_gf_wta : PINNED against the
 *     reference's STMatching/StereoHelper.cpp compiled unmodified (oracle/_ref/libstref.so, a small
 *     cv::Mat stand-in under oracle/shim_st/).
 *   - orc_lr_check : restatement of a loop inside stereo_disparity_iteration
 *     (StereoDisparity.cpp:136-147), which cannot be called in isolation; pinned only by a
 *     literal-loop test.
 *   - orc_remap / orc_cvtcolor(truncate) : PINNED against the reference's BlockMatching/Utility.cpp compiled
 *     unmodified (oracle/_ref/libutilref.so, stand-in under oracle/shim_util/).
 *   - orc_gf_* (guided filter) : PARITY UNPINNED.  The reference contains no guided filter;
 *     this file is the de-facto definition ("GF-v1").
 *
 * All file:line citations are relative to /root/reference.
 * Layouts: images u8 [H][W] row-major contiguous; volumes [D][H][W] unless stated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------
 * A.1  absolute-difference slice.
 * view 0 (left reference image):  BlockMatching/BlockMatching.cpp:101-108 (PreCal)
 *     AD_d[y][x] = |L[y][x] - R[y][x-d]| if x-d >= 0 else 0 (buffer pre-zeroed :39,:143)
 * view 1 (right reference image): STMatching/StereoHelper.cpp:156-180
 *     pR_d[y][x] = pL_d[y][x+d] if x+d < W else pR_{d-1}[y][x]
 *     (for x+d >= W this recursion bottoms out at d' = W-1-x, i.e. |L[W-1] - R[x]|)
 * ---------------------------------------------------------------------------------- */
void orc_ad_slice(const u8* L, const u8* R, int H, int W, int d, int view, u8* out) {
  for (int y = 0; y < H; ++y) {
    const u8* l = L + (size_t)y * W;
    const u8* r = R + (size_t)y * W;
    u8* o = out + (size_t)y * W;
    if (view == 0) {
      for (int x = 0; x < W; ++x) o[x] = (x - d >= 0) ? (u8)abs((int)l[x] - (int)r[x - d]) : 0;
    } else {
      for (int x = 0; x < W; ++x) {
        int dd = d;
        while (x + dd >= W && dd > 0) --dd; /* literal restatement of the d-1 fallback */
        o[x] = (u8)abs((int)l[x + dd] - (int)r[x]);
      }
    }
  }
}

/* BlockMatching.cpp:89-109 -- whole volume, [D][H][W] */
void orc_ad_volume(const u8* L, const u8* R, int H, int W, int D, u8* out) {
#pragma omp parallel for schedule(static)
  for (int d = 0; d < D; ++d) orc_ad_slice(L, R, H, W, d, 0, out + (size_t)d * H * W);
}

/* clipped (2r+1)^2 window SUM of an int32 plane, exact, separable running sums */
static void box_sum_i32(const int32_t* src, int H, int W, int r, int32_t* dst, int32_t* tmp) {
  /* horizontal */
  for (int y = 0; y < H; ++y) {
    const int32_t* s = src + (size_t)y * W;
    int32_t* t = tmp + (size_t)y * W;
    int64_t acc = 0;
    for (int x = 0; x <= imin(r, W - 1); ++x) acc += s[x];
    for (int x = 0; x < W; ++x) {
      t[x] = (int32_t)acc;
      if (x + r + 1 < W) acc += s[x + r + 1];
      if (x - r >= 0) acc -= s[x - r];
    }
  }
  /* vertical */
  for (int x = 0; x < W; ++x) {
    int64_t acc = 0;
    for (int y = 0; y <= imin(r, H - 1); ++y) acc += tmp[(size_t)y * W + x];
    for (int y = 0; y < H; ++y) {
      dst[(size_t)y * W + x] = (int32_t)acc;
      if (y + r + 1 < H) acc += tmp[(size_t)(y + r + 1) * W + x];
      if (y - r >= 0) acc -= tmp[(size_t)(y - r) * W + x];
    }
  }
}

/* un-truncated SAD slice: clipped-window, un-normalised (BlockMatching.cpp:167-177) */
void orc_sad_slice(const u8* L, const u8* R, int H, int W, int r, int d, int view, int32_t* out) {
  size_t n = (size_t)H * W;
  u8* ad = (u8*)malloc(n);
  int32_t* a32 = (int32_t*)malloc(n * 4);
  int32_t* tmp = (int32_t*)malloc(n * 4);
  orc_ad_slice(L, R, H, W, d, view, ad);
  for (size_t i = 0; i < n; ++i) a32[i] = ad[i];
  box_sum_i32(a32, H, W, r, out, tmp);
  free(ad); free(a32); free(tmp);
}

/* ------------------------------------------------------------------------------------
 * A.2  SAD + WTA, literal loop structure of getDisp, BlockMatching.cpp:156-185
 * (minus the early-out at :174-176 which only prunes and never changes the result).
 * O((2r+1)^2) per evaluation: use for small cases / cross-checks.
 * ---------------------------------------------------------------------------------- */
void orc_sad_wta_direct(const u8* L, const u8* R, int H, int W, int r, int D, u8* disp) {
  size_t total = (size_t)H * W;
  u8* dif = (u8*)calloc(total * D, 1);
  orc_ad_volume(L, R, H, W, D, dif);
  int dnum = (2 * r + 1) * (2 * r + 1);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      int best = 50 * dnum; /* :157 */
      int dm = -256;        /* :158 */
      for (int d = 0; d < D; ++d) {
        if (x + d > W) break; /* :166 (sic) */
        int temp = 0;
        for (int dy = -r; dy <= r; ++dy) {
          int yy = y + dy;
          if (yy >= H || yy < 0) continue; /* :171-172 */
          for (int dx = -r; dx <= r; ++dx) {
            int xx = x + dx;
            if (xx >= W || xx < 0) continue; /* :169-170 */
            temp += dif[(size_t)d * total + (size_t)yy * W + xx];
          }
        }
        if (temp < best) { dm = d; best = temp; } /* :178-181 strict < */
      }
      disp[(size_t)y * W + x] = (u8)dm; /* :184 (-256 -> 0) */
    }
  }
  free(dif);
}

/* Same result through exact box sums, O(1) per evaluation.  best_cost (optional, int32
 * [H][W]) receives the winning SAD, or 50*(2r+1)^2 where nothing was accepted. */
void orc_sad_wta(const u8* L, const u8* R, int H, int W, int r, int D, u8* disp, int32_t* best_cost) {
  size_t n = (size_t)H * W;
  int thr = 50 * (2 * r + 1) * (2 * r + 1);
  int32_t* best = (int32_t*)malloc(n * 4);
  int* bd = (int*)malloc(n * sizeof(int));
  for (size_t i = 0; i < n; ++i) { best[i] = thr; bd[i] = -256; }
#pragma omp parallel
  {
    int32_t* sad = (int32_t*)malloc(n * 4);
#pragma omp for schedule(dynamic, 1) ordered
    for (int d = 0; d < D; ++d) {
      orc_sad_slice(L, R, H, W, r, d, 0, sad);
#pragma omp ordered
      {
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            if (x + d > W) continue; /* the reference breaks; d ascending => same set */
            size_t i = (size_t)y * W + x;
            if (sad[i] < best[i]) { best[i] = sad[i]; bd[i] = d; }
          }
      }
    }
    free(sad);
  }
  for (size_t i = 0; i < n; ++i) disp[i] = (u8)bd[i];
  if (best_cost) memcpy(best_cost, best, n * 4);
  free(best); free(bd);
}

/* a3: getAllSAD, BlockMatching.cpp:191-261.  out[p*D+d] = (u8)SAD, 255 where x+d > W.
 * Unlike the reference (which leaves positions it never writes untouched) every cell is
 * written, so the caller need not pre-fill. */
void orc_all_sad(const u8* L, const u8* R, int H, int W, int r, int D, u8* out) {
  size_t n = (size_t)H * W;
#pragma omp parallel
  {
    int32_t* sad = (int32_t*)malloc(n * 4);
#pragma omp for schedule(dynamic, 1)
    for (int d = 0; d < D; ++d) {
      orc_sad_slice(L, R, H, W, r, d, 0, sad);
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          size_t i = (size_t)y * W + x;
          out[i * D + d] = (x + d > W) ? 255 : (u8)sad[i]; /* :244-248, :258 (truncation) */
        }
    }
    free(sad);
  }
}

/* ------------------------------------------------------------------------------------
 * A.3  guided-filter aggregation "GF-v1" (NOT in the reference; parity unpinned).
 *   guide I (u8), input p (u8 AD slice), clipped (2r+1)^2 window, N = in-image count.
 *   S_I=box(I) S_II=box(I*I) S_p=box(p) S_Ip=box(I*p)            (exact integers)
 *   a = (N*S_Ip - S_I*S_p) / (N*S_II - S_I^2 + eps*N^2)          (== cov/(var+eps))
 *   b = (S_p - a*S_I) / N                                        (== mean_p - a*mean_I)
 *   q = (box(a)*I + box(b)) / N
 * float64, stage-2 window sums are direct (non-running) separable sums.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int H, W, r;
  double eps;
  int32_t* N;    /* window pixel count */
  int32_t* S_I;  /* box(I) */
  int64_t* den;  /* N*S_II - S_I^2 (exact) */
} gf_guide;

static void box_sum_f64_direct(const double* src, int H, int W, int r, double* dst, double* tmp) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      double acc = 0.0;
      for (int xx = imax(0, x - r); xx <= imin(W - 1, x + r); ++xx) acc += src[(size_t)y * W + xx];
      tmp[(size_t)y * W + x] = acc;
    }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      double acc = 0.0;
      for (int yy = imax(0, y - r); yy <= imin(H - 1, y + r); ++yy) acc += tmp[(size_t)yy * W + x];
      dst[(size_t)y * W + x] = acc;
    }
}

static gf_guide* gf_guide_new(const u8* I, int H, int W, int r, double eps) {
  size_t n = (size_t)H * W;
  gf_guide* g = (gf_guide*)malloc(sizeof(gf_guide));
  g->H = H; g->W = W; g->r = r; g->eps = eps;
  g->N = (int32_t*)malloc(n * 4);
  g->S_I = (int32_t*)malloc(n * 4);
  g->den = (int64_t*)malloc(n * 8);
  int32_t* t0 = (int32_t*)malloc(n * 4);
  int32_t* t1 = (int32_t*)malloc(n * 4);
  int32_t* sii = (int32_t*)malloc(n * 4);
  for (size_t i = 0; i < n; ++i) t0[i] = 1;
  box_sum_i32(t0, H, W, r, g->N, t1);
  for (size_t i = 0; i < n; ++i) t0[i] = I[i];
  box_sum_i32(t0, H, W, r, g->S_I, t1);
  for (size_t i = 0; i < n; ++i) t0[i] = (int32_t)I[i] * I[i];
  box_sum_i32(t0, H, W, r, sii, t1);
  for (size_t i = 0; i < n; ++i)
    g->den[i] = (int64_t)g->N[i] * sii[i] - (int64_t)g->S_I[i] * g->S_I[i];
  free(t0); free(t1); free(sii);
  return g;
}

static void gf_guide_free(gf_guide* g) {
  free(g->N); free(g->S_I); free(g->den); free(g);
}

/* one slice; scratch = 6 planes of n doubles worth of bytes */
static void gf_slice(const gf_guide* g, const u8* I, const u8* p, double* q, void* scratch) {
  int H = g->H, W = g->W, r = g->r;
  size_t n = (size_t)H * W;
  int32_t* t0 = (int32_t*)scratch;
  int32_t* t1 = t0 + n;
  int32_t* sp = t1 + n;
  int32_t* sip = sp + n;
  double* a = (double*)(sip + n);
  double* b = a + n;
  double* sa = b + n;
  double* sb = sa + n;
  double* tmp = sb + n;
  for (size_t i = 0; i < n; ++i) t0[i] = p[i];
  box_sum_i32(t0, H, W, r, sp, t1);
  for (size_t i = 0; i < n; ++i) t0[i] = (int32_t)I[i] * p[i];
  box_sum_i32(t0, H, W, r, sip, t1);
  for (size_t i = 0; i < n; ++i) {
    double Nn = (double)g->N[i];
    int64_t num = (int64_t)g->N[i] * sip[i] - (int64_t)g->S_I[i] * sp[i];
    double den = (double)g->den[i] + g->eps * Nn * Nn;
    a[i] = (double)num / den;
    b[i] = ((double)sp[i] - a[i] * (double)g->S_I[i]) / Nn;
  }
  box_sum_f64_direct(a, H, W, r, sa, tmp);
  box_sum_f64_direct(b, H, W, r, sb, tmp);
  for (size_t i = 0; i < n; ++i) q[i] = (sa[i] * (double)I[i] + sb[i]) / (double)g->N[i];
}

static size_t gf_scratch_bytes(size_t n) { return n * (4 * 4 + 5 * 8); }

/* q for disparities d0..d0+nd-1 -> out [nd][H][W] float64.  view 0: guide L; view 1: guide R */
void orc_gf_cost_slices(const u8* L, const u8* R, int H, int W, int r, double eps, int view,
                        int d0, int nd, double* out) {
  size_t n = (size_t)H * W;
  const u8* I = view == 0 ? L : R;
  gf_guide* g = gf_guide_new(I, H, W, r, eps);
#pragma omp parallel
  {
    u8* p = (u8*)malloc(n);
    void* scratch = malloc(gf_scratch_bytes(n));
#pragma omp for schedule(dynamic, 1)
    for (int k = 0; k < nd; ++k) {
      orc_ad_slice(L, R, H, W, d0 + k, view, p);
      gf_slice(g, I, p, out + (size_t)k * n, scratch);
    }
    free(p); free(scratch);
  }
  gf_guide_free(g);
}

/* A.3 + A.4: guided-filter aggregation then WTA over ALL D candidates, strict '<', first
 * minimum wins, no threshold (STMatching/StereoHelper.cpp:137-150).  best_cost optional. */
void orc_gf_wta(const u8* L, const u8* R, int H, int W, int r, int D, double eps, int view,
                u8* disp, double* best_cost) {
  size_t n = (size_t)H * W;
  const u8* I = view == 0 ? L : R;
  gf_guide* g = gf_guide_new(I, H, W, r, eps);
  double* best = (double*)malloc(n * 8);
  int* bd = (int*)malloc(n * sizeof(int));
  for (size_t i = 0; i < n; ++i) { best[i] = INFINITY; bd[i] = 0; }
#pragma omp parallel
  {
    u8* p = (u8*)malloc(n);
    double* q = (double*)malloc(n * 8);
    void* scratch = malloc(gf_scratch_bytes(n));
#pragma omp for schedule(dynamic, 1)
    for (int d = 0; d < D; ++d) {
      orc_ad_slice(L, R, H, W, d, view, p);
      gf_slice(g, I, p, q, scratch);
#pragma omp critical
      {
        for (size_t i = 0; i < n; ++i)
          if (q[i] < best[i] || (q[i] == best[i] && d < bd[i])) { best[i] = q[i]; bd[i] = d; }
      }
    }
    free(p); free(q); free(scratch);
  }
  for (size_t i = 0; i < n; ++i) disp[i] = (u8)bd[i];
  if (best_cost) memcpy(best_cost, best, n * 8);
  free(best); free(bd);
  gf_guide_free(g);
}

/* ------------------------------------------------------------------------------------
 * A.5  left-right consistency check, STMatching/StereoDisparity.cpp:136-147
 *   d = DL(y,x); if x-d >= 0: occ = (d == 0 || |d - DR(y,x-d)| > 1) else occ = 1; mask = !occ
 * ---------------------------------------------------------------------------------- */
void orc_lr_check(const u8* DL, const u8* DR, int H, int W, u8* occ, u8* mask) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      size_t i = (size_t)y * W + x;
      int d = DL[i];
      u8 o;
      if (x - d >= 0) {
        int dc = DR[(size_t)y * W + (x - d)];
        o = (u8)(d == 0 || abs(d - dc) > 1);
      } else {
        o = 1;
      }
      if (occ) occ[i] = o;
      if (mask) mask[i] = (u8)!o;
    }
}

/* ------------------------------------------------------------------------------------
 * A.6  (2r+1)^2 median, u8, REPLICATE border: value v at which the cumulative histogram
 * first exceeds t = 2r^2+2r (STMatching/ctmf.c:281,288-294,319-325; border behaviour from
 * the first-row init :228-232 and the MIN/MAX clamps :244,252,285,315).
 * ---------------------------------------------------------------------------------- */
void orc_median(const u8* src, u8* dst, int H, int W, int r) {
  int t = 2 * r * r + 2 * r;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y) {
    int hist[256];
    for (int x = 0; x < W; ++x) {
      memset(hist, 0, sizeof(hist));
      for (int dy = -r; dy <= r; ++dy) {
        int yy = imin(H - 1, imax(0, y + dy));
        for (int dx = -r; dx <= r; ++dx) {
          int xx = imin(W - 1, imax(0, x + dx));
          hist[src[(size_t)yy * W + xx]]++;
        }
      }
      int sum = 0, v = 0;
      for (v = 0; v < 256; ++v) {
        sum += hist[v];
        if (sum > t) break;
      }
      dst[(size_t)y * W + x] = (u8)v;
    }
  }
}

/* ------------------------------------------------------------------------------------
 * SURVEY 8(f)-1  bilinear remap, BlockMatching/Utility.cpp:236-264 (CPU_Remap + CPU_BilinearInterpolation;
 * GPU twin kernalRemap + BilinearInterpolation, Device.cu:127-134,152-167).  Note the argument order quirk:
 * the interpolator is called as (src, ycoo, xcoo), so its `x` is the ROW coordinate.  Out-of-range
 * (x1 < 0 || x2 >= rows || y1 < 0 || y2 >= cols) -> 0.  Result rounded to nearest-even and saturated
 * (saturate_cast<uchar> == cvt.rni.sat.u8.f32, Device.cu:148).  Every product and sum is rounded to float
 * separately (no FMA contraction; this file is built with -ffp-contract=off).
 * ---------------------------------------------------------------------------------- */
static inline u8 sat_rne_u8(float v) {
  float r = nearbyintf(v); /* default rounding mode: to nearest, ties to even */
  if (!(r > 0.0f)) return 0;
  if (r > 255.0f) return 255;
  return (u8)r;
}

void orc_remap(const u8* src, const float* mapx, const float* mapy, int rows, int cols, u8* dst) {
  for (int row = 0; row < rows; ++row)
    for (int col = 0; col < cols; ++col) {
      size_t i = (size_t)row * cols + col;
      float x = mapy[i], y = mapx[i]; /* CPU_BilinearInterpolation(src, ycoo, xcoo) */
      int x1 = (int)floorf(x), y1 = (int)floorf(y), x2 = x1 + 1, y2 = y1 + 1;
      float result = 0.0f;
      if (!(x1 < 0 || x2 >= rows || y1 < 0 || y2 >= cols)) {
        u8 Q11 = src[(size_t)x1 * cols + y1], Q12 = src[(size_t)x1 * cols + y2];
        u8 Q21 = src[(size_t)x2 * cols + y1], Q22 = src[(size_t)x2 * cols + y2];
        float left = ((float)x2 - x) * (float)Q11 + (x - (float)x1) * (float)Q21;
        float right = ((float)x2 - x) * (float)Q12 + (x - (float)x1) * (float)Q22;
        result = ((float)y2 - y) * left + (y - (float)y1) * right;
      }
      dst[i] = sat_rne_u8(result);
    }
}

/* SURVEY 8(f)-2  3-channel -> gray, weights .299/.587/.114 applied to channels 0/1/2 as stored
 * (the reference feeds OpenCV BGR data, Caller.cpp:106).  truncate = 1: cvtColor_cpu, Utility.cpp:289-298
 * ((uchar)channelSum), host code: no contraction; truncate = 0: kernalCvtColor, Device.cu:136-143
 * (round-nearest-even, saturate) with the contraction nvcc applies to the kernel's expression:
 * FMUL(.587 c1), FFMA(.299 c0), FFMA(.114 c2) -- read off the SASS of the reference kernel compiled unmodified
 * and confirmed by running it (tools/ref_gpu_compare.py). */
void orc_cvtcolor(const u8* src3, int rows, int cols, int truncate, u8* dst) {
  size_t n = (size_t)rows * cols;
  for (size_t i = 0; i < n; ++i) {
    const float c0 = (float)src3[3 * i], c1 = (float)src3[3 * i + 1], c2 = (float)src3[3 * i + 2];
    if (truncate) {
      float sum = .299f * c0 + .587f * c1 + .114f * c2;
      dst[i] = (u8)sum;
    } else {
      dst[i] = sat_rne_u8(fmaf(.114f, c2, fmaf(.299f, c0, .587f * c1)));
    }
  }
}

/* FNV-1a 64 over a byte buffer: digest format of tests/golden/  */
uint64_t orc_fnv1a64(const u8* p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ULL;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
  return h;
}
