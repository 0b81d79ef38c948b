// gsm_gf.cuh -- fused AD -> guided-filter aggregation -> WTA kernel (GSM_MODE_GF) and its guide-statistics
// pre-pass.  The reference has no guided filter; the arithmetic follows "GF-v1"
// (oracle/stereo_oracle.c gf_slice) and the WTA follows STMatching/StereoHelper.cpp:131-154.
//
// Per disparity d, guide I, p = AD_d, clipped (2r+1)^2 window sums S_x = box(x), N = window pixel count:
//     a = (N*S_Ip - S_I*S_p) / (N*S_II - S_I^2 + eps*N^2)      b = (S_p - a*S_I) / N
//     q = (box(a)*I + box(b)) / N
// Nothing of the D x H x W volume is materialised.  A CTA owns (strip of columns) x (32 disparities) and
// marches down the rows.  Because box(a), box(b) need a and b on the 2r+1 rows around the output row, two
// instances of the integer stage-1 pipeline run 2r+1 rows apart ("lead" adds a row of (a,b) to the stage-2
// running sums, "trail" recomputes the row that leaves) -- that trades ~35% more arithmetic for not keeping a
// (2r+1)-row ring of float (a,b) rows, which would not fit on chip for more than ~8 disparities per CTA.
//
//   stage 1 (exact, int32):  V_p, V_Ip vertical running sums via IDP.2A (one op adds the entering row and
//            removes the leaving row), exchanged through shared memory, horizontal sliding sums with IADD3;
//            the numerator N*S_Ip - S_I*S_p is evaluated modulo 2^32, exact because |N^2 cov| < 2^31 for r <= 9.
//   stage 2 (fp32):  a, b per pixel; vertical running sums V_a, V_b; exchange; horizontal sliding sums; q.
//            b is formed against the centred guide (I - 128) so the stage-2 sums are ~2x smaller.
//   WTA:     warp min over the 32 disparities of a run (REDUX on the sortable bit pattern, ballot for the
//            lowest d among equals), one 64-bit atomicMin per pixel into the packed-min plane.
#pragma once
#include "gsm_common.cuh"
#include "gsm_sad.cuh"

namespace gsm {

// guide statistic planes (float/int32, same padded geometry as the u8 planes; zero outside the image)
constexpr int GF_STAT_PLANES = 7;
enum { ST_N = 0, ST_SI = 1, ST_INVDEN = 2, ST_CMEAN = 3, ST_INVN = 4, ST_IC = 5, ST_COEF = 6 };
constexpr float GF_CENTRE = 128.0f;

// N, S_I, 1/(N*S_II - S_I^2 + eps*N^2), S_I/N - 128, 1/N, I - 128 for every image pixel.
constexpr int GS_T = 32;
__global__ void __launch_bounds__(GS_T * 8)
gf_stats_kernel(const u8* __restrict__ Ip, float* __restrict__ stats, PlaneGeom pg, int R, float eps) {
  extern __shared__ int gs_smem[];
  const int tw = GS_T + 2 * R;
  u8* tile = reinterpret_cast<u8*>(gs_smem);                       // [tw][tw] (padded to 4)
  int* hsI = gs_smem + (tw * tw + 3) / 4;                          // [tw][GS_T]
  int* hsII = hsI + tw * GS_T;
  const int f = blockIdx.z;
  const int bx = blockIdx.x * GS_T, by = blockIdx.y * GS_T;
  const u8* src = Ip + (size_t)f * pg.plane_stride + (size_t)(PADV + by - R) * pg.pitch + pg.xoff + bx - R;
  const int tid = threadIdx.y * GS_T + threadIdx.x;
  for (int i = tid; i < tw * tw; i += GS_T * 8) {
    const int ty = i / tw, tx = i - ty * tw;
    // rows beyond the bottom pad can only be reached by tiles hanging below the image; clamp the read
    const int prow = min(PADV + by - R + ty, pg.plane_rows - 1) - (PADV + by - R);
    tile[i] = src[(size_t)prow * pg.pitch + tx];
  }
  __syncthreads();
  for (int i = tid; i < tw * GS_T; i += GS_T * 8) {
    const int ty = i / GS_T, tx = i - ty * GS_T;
    int s = 0, s2 = 0;
    for (int j = 0; j <= 2 * R; ++j) {
      const int v = tile[ty * tw + tx + j];
      s += v;
      s2 += v * v;
    }
    hsI[i] = s;
    hsII[i] = s2;
  }
  __syncthreads();
  const size_t plane_elems = pg.plane_stride;  // elements per statistic plane
  float* base = stats + (size_t)f * GF_STAT_PLANES * plane_elems;
  for (int ry = threadIdx.y; ry < GS_T; ry += 8) {
    const int x = bx + threadIdx.x, y = by + ry;
    if (x >= pg.W || y >= pg.H) continue;
    int SI = 0, SII = 0;
    for (int j = 0; j <= 2 * R; ++j) {
      SI += hsI[(ry + j) * GS_T + threadIdx.x];
      SII += hsII[(ry + j) * GS_T + threadIdx.x];
    }
    const int nx = min(pg.W - 1, x + R) - max(0, x - R) + 1;
    const int ny = min(pg.H - 1, y + R) - max(0, y - R) + 1;
    const int N = nx * ny;
    const long long den = (long long)N * SII - (long long)SI * SI;
    const double dden = (double)den + (double)eps * (double)N * (double)N;
    const size_t o = (size_t)(PADV + y) * pg.pitch + pg.xoff + x;
    reinterpret_cast<int*>(base + ST_N * plane_elems)[o] = N;
    reinterpret_cast<int*>(base + ST_SI * plane_elems)[o] = SI;
    base[ST_INVDEN * plane_elems + o] = (float)(1.0 / dden);
    base[ST_CMEAN * plane_elems + o] = (float)((double)SI / N - (double)GF_CENTRE);
    base[ST_INVN * plane_elems + o] = 1.0f / (float)N;
    base[ST_IC * plane_elems + o] = (float)tile[(ry + R) * tw + threadIdx.x + R] - GF_CENTRE;
  }
}

// IDP.2A coefficient plane: COEF[t][x] = I[t+R][x] - 65536 * I[t-R-1][x]  (lo16 = +I entering row t+R,
// hi16 = -I leaving row t-R-1), for t in [-R, H+R].  The trail pipeline reads the same plane 2R+1 rows up.
__global__ void gf_coef_kernel(const u8* __restrict__ Ip, float* __restrict__ stats, PlaneGeom pg, int R) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = (int)blockIdx.y - R;
  const int f = blockIdx.z;
  if (x >= pg.W) return;
  const u8* p = Ip + (size_t)f * pg.plane_stride + pg.xoff + x;
  const int in = p[(size_t)(PADV + t + R) * pg.pitch];
  const int out = p[(size_t)(PADV + t - R - 1) * pg.pitch];
  int* coef = reinterpret_cast<int*>(stats + ((size_t)f * GF_STAT_PLANES + ST_COEF) * pg.plane_stride);
  coef[(size_t)(PADV + t) * pg.pitch + pg.xoff + x] = in - 65536 * out;
}

template <int K>
__device__ __forceinline__ void load_i32x(const int* p, int (&v)[K]) {
#pragma unroll
  for (int w = 0; w < K / 4; ++w) {
    const int4 t = __ldg(reinterpret_cast<const int4*>(p) + w);
    v[4 * w] = t.x; v[4 * w + 1] = t.y; v[4 * w + 2] = t.z; v[4 * w + 3] = t.w;
  }
}
template <int K>
__device__ __forceinline__ void load_f32x(const float* p, float (&v)[K]) {
#pragma unroll
  for (int w = 0; w < K / 4; ++w) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + w);
    v[4 * w] = t.x; v[4 * w + 1] = t.y; v[4 * w + 2] = t.z; v[4 * w + 3] = t.w;
  }
}

// publish K words at buf, then (after the CTA barrier) gather the window [-HL4, K+HL4) around them
template <int K, int HL4>
__device__ __forceinline__ void exch_store(u32* buf, const u32 (&v)[K]) {
#pragma unroll
  for (int w = 0; w < K / 4; ++w)
    reinterpret_cast<uint4*>(buf)[w] = make_uint4(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3]);
}
template <int K, int HL4>
__device__ __forceinline__ void exch_window(const u32* buf, const u32 (&own)[K], u32 (&win)[HL4 + K + HL4]) {
#pragma unroll
  for (int w = 0; w < HL4 / 4; ++w) {
    const uint4 a = reinterpret_cast<const uint4*>(buf - HL4)[w];
    win[4 * w] = a.x; win[4 * w + 1] = a.y; win[4 * w + 2] = a.z; win[4 * w + 3] = a.w;
    const uint4 b = reinterpret_cast<const uint4*>(buf + K)[w];
    win[HL4 + K + 4 * w] = b.x; win[HL4 + K + 4 * w + 1] = b.y;
    win[HL4 + K + 4 * w + 2] = b.z; win[HL4 + K + 4 * w + 3] = b.w;
  }
#pragma unroll
  for (int c = 0; c < K; ++c) win[HL4 + c] = own[c];
}
template <int R, int K, int HL4>
__device__ __forceinline__ void slide_i32(const u32 (&win)[HL4 + K + HL4], int (&S)[K]) {
  int s = 0;
#pragma unroll
  for (int j = -R; j <= R; ++j) s += (int)win[HL4 + j];
  S[0] = s;
#pragma unroll
  for (int c = 1; c < K; ++c) {
    s += (int)win[HL4 + c + R] - (int)win[HL4 + c - R - 1];
    S[c] = s;
  }
}
template <int R, int K, int HL4>
__device__ __forceinline__ void slide_f32(const u32 (&win)[HL4 + K + HL4], float (&S)[K]) {
  float s = 0.f;
#pragma unroll
  for (int j = -R; j <= R; ++j) s += __uint_as_float(win[HL4 + j]);
  S[0] = s;
#pragma unroll
  for (int c = 1; c < K; ++c) {
    s = (s + __uint_as_float(win[HL4 + c + R])) - __uint_as_float(win[HL4 + c - R - 1]);
    S[c] = s;
  }
}

template <int K>
__device__ __forceinline__ void ad_row(const u8* gbase, const u32* obase, u32 osel, size_t ro, u32 (&p)[K / 4]) {
  u32 gw[K / 4], ow[K / 4];
  load_aligned<K>(gbase + ro, gw);
  load_unaligned<K>(reinterpret_cast<const u32*>(reinterpret_cast<const u8*>(obase) + ro), osel, ow);
#pragma unroll
  for (int w = 0; w < K / 4; ++w) p[w] = __vabsdiffu4(gw[w], ow[w]);
}

// a, b of one row from the exact stage-1 sums and the guide statistics of that row
template <int K>
__device__ __forceinline__ void ab_row(const float* srow, size_t plane_elems, const int (&Sp)[K], const int (&SIp)[K],
                                       float (&a)[K], float (&b)[K]) {
  int N[K], SI[K];
  float invden[K], cmean[K], invn[K];
  load_i32x<K>(reinterpret_cast<const int*>(srow + ST_N * plane_elems), N);
  load_i32x<K>(reinterpret_cast<const int*>(srow + ST_SI * plane_elems), SI);
  load_f32x<K>(srow + ST_INVDEN * plane_elems, invden);
  load_f32x<K>(srow + ST_CMEAN * plane_elems, cmean);
  load_f32x<K>(srow + ST_INVN * plane_elems, invn);
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const int num = N[c] * SIp[c] - SI[c] * Sp[c];  // exact modulo 2^32, true value fits int32 for r <= 9
    a[c] = (float)num * invden[c];
    b[c] = fmaf(-a[c], cmean[c], (float)Sp[c] * invn[c]);  // mean_p - a * (mean_I - 128)
  }
}

template <int R, int K, bool EXPORT>
__global__ void __launch_bounds__(384, 1)
gf_wta_kernel(const u8* __restrict__ Gp, const u8* __restrict__ Op, const float* __restrict__ stats,
              i64* __restrict__ keys, FusedGeom g) {
  constexpr int HL4 = (R + 3) / 4 * 4;
  constexpr int KW = K / 4;
  extern __shared__ __align__(16) u32 smem[];

  const int lane = threadIdx.x;
  const int run = threadIdx.y;
  const int runs = blockDim.y;
  const int strip = blockIdx.x;
  const int d = g.d_begin + blockIdx.y * WARP + lane;
  const int frame = blockIdx.z / g.bands;
  const int band = blockIdx.z - frame * g.bands;
  const int H = g.pg.H, W = g.pg.W, pitch = g.pg.pitch;
  const int yb0 = band * g.band_rows;
  const int yb1 = min(H, yb0 + g.band_rows);
  if (yb0 >= H) return;

  const int pitchw = exch_pitch_words(runs, K, HL4);
  const int planew = WARP * pitchw;
  for (int i = threadIdx.y * WARP + threadIdx.x; i < 6 * planew; i += runs * WARP) smem[i] = 0u;
  __syncthreads();
  u32* xb = smem + (size_t)lane * pitchw + HL4 + run * K;  // this thread's slot in exchange plane 0

  const int x0 = strip * g.TW - g.hl + run * K;
  const int dd = min(d, MAX_DISP - 1);
  const int osh = (g.view == 0) ? -dd : dd;
  const size_t plane_elems = g.pg.plane_stride;
  const size_t org = (size_t)PADV * pitch + g.pg.xoff + x0;  // (row 0, column x0) inside a padded plane
  const u8* gbase = Gp + (size_t)frame * g.pg.plane_stride + org;
  const u8* obase_b = Op + (size_t)frame * g.pg.plane_stride + org + osh;
  const u32 omis = (u32)(reinterpret_cast<uintptr_t>(obase_b) & 3u);
  const u32* obase = reinterpret_cast<const u32*>(obase_b - omis);
  const u32 osel = 0x3210u + 0x1111u * omis;
  const float* sbase = stats + (size_t)frame * GF_STAT_PLANES * plane_elems + org;
  const int* coefbase = reinterpret_cast<const int*>(sbase + ST_COEF * plane_elems);

  u32 mask[KW];
  bool full = true;
#pragma unroll
  for (int w = 0; w < KW; ++w) {
    u32 m = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int x = x0 + 4 * w + b;
      const bool ok = (x < W) && (g.view == 0 ? (x >= dd) : (x >= 0));
      m |= ok ? (0xffu << (8 * b)) : 0u;
    }
    mask[w] = m;
    full = full && (m == 0xffffffffu);
  }
  const bool need_mask = __any_sync(0xffffffffu, !full);

  const int out0 = strip * g.TW;
  const int c_lo = max(0, out0 - x0);
  int c_hi = min(K - 1, min(out0 + g.TW, W) - 1 - x0);
  if (d >= g.d_end) c_hi = -1;
  const bool all_valid = __all_sync(0xffffffffu, c_lo == 0 && c_hi == K - 1);

  int Vp_l[K], VIp_l[K], Vp_t[K], VIp_t[K];
  float VA[K], VB[K];
#pragma unroll
  for (int c = 0; c < K; ++c) { Vp_l[c] = VIp_l[c] = Vp_t[c] = VIp_t[c] = 0; VA[c] = VB[c] = 0.f; }

  const int r0 = yb0 - 2 * R;  // first image row whose AD may enter a stage-1 window of this band
  const int a0 = yb0 - R;      // first row whose (a, b) may enter a stage-2 window of this band
  constexpr int COEF_PM = (int)0xFFFF0001;  // lo16 = +1, hi16 = -1
  const int row_lo = -PADV, row_hi = H + PADV - 1;

  for (int t = yb0 - 3 * R; t < yb1 + R; ++t) {
    // ---------------- stage 1, vertical: rows t+R (enters lead), t-R-1 (lead -> trail), t-3R-2 (leaves trail)
    const int t2 = t - 2 * R - 1;  // row of (a, b) recomputed by the trail pipeline
    u32 pn[KW], pm[KW], po[KW];
    ad_row<K>(gbase, obase, osel, (size_t)((long long)(t + R) * pitch), pn);
    if (t - R - 1 >= r0) ad_row<K>(gbase, obase, osel, (size_t)((long long)(t - R - 1) * pitch), pm);
    else {
#pragma unroll
      for (int w = 0; w < KW; ++w) pm[w] = 0u;
    }
    if (t - 3 * R - 2 >= r0) ad_row<K>(gbase, obase, osel, (size_t)((long long)(t - 3 * R - 2) * pitch), po);
    else {
#pragma unroll
      for (int w = 0; w < KW; ++w) po[w] = 0u;
    }
    if (need_mask) {
#pragma unroll
      for (int w = 0; w < KW; ++w) { pn[w] &= mask[w]; pm[w] &= mask[w]; po[w] &= mask[w]; }
    }
    {
      int cA[K], cB[K];
      load_i32x<K>(coefbase + (long long)max(row_lo, min(row_hi, t)) * pitch, cA);
      load_i32x<K>(coefbase + (long long)max(row_lo, min(row_hi, t2)) * pitch, cB);
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const u32 sel = (c & 3) | ((4 + (c & 3)) << 4);
        const u32 nm = __byte_perm(pn[c / 4], pm[c / 4], sel);  // {p_enter, p_leave, x, x}
        const u32 mo = __byte_perm(pm[c / 4], po[c / 4], sel);
        Vp_l[c] = dp2a_lo_su(COEF_PM, nm, Vp_l[c]);
        VIp_l[c] = dp2a_lo_su(cA[c], nm, VIp_l[c]);
        Vp_t[c] = dp2a_lo_su(COEF_PM, mo, Vp_t[c]);
        VIp_t[c] = dp2a_lo_su(cB[c], mo, VIp_t[c]);
      }
    }
    // ---------------- stage 1, horizontal
    exch_store<K, HL4>(xb + 0 * planew, reinterpret_cast<u32(&)[K]>(Vp_l));
    exch_store<K, HL4>(xb + 1 * planew, reinterpret_cast<u32(&)[K]>(VIp_l));
    exch_store<K, HL4>(xb + 2 * planew, reinterpret_cast<u32(&)[K]>(Vp_t));
    exch_store<K, HL4>(xb + 3 * planew, reinterpret_cast<u32(&)[K]>(VIp_t));
    __syncthreads();
    float al[K], bl[K];
    if (t >= a0) {
      int Sp[K], SIp[K];
      u32 win[HL4 + K + HL4];
      exch_window<K, HL4>(xb + 0 * planew, reinterpret_cast<u32(&)[K]>(Vp_l), win);
      slide_i32<R, K, HL4>(win, Sp);
      exch_window<K, HL4>(xb + 1 * planew, reinterpret_cast<u32(&)[K]>(VIp_l), win);
      slide_i32<R, K, HL4>(win, SIp);
      ab_row<K>(sbase + (long long)t * pitch, plane_elems, Sp, SIp, al, bl);
    } else {
#pragma unroll
      for (int c = 0; c < K; ++c) al[c] = bl[c] = 0.f;
    }
    if (t2 >= a0) {
      int Sp[K], SIp[K];
      float at[K], bt[K];
      u32 win[HL4 + K + HL4];
      exch_window<K, HL4>(xb + 2 * planew, reinterpret_cast<u32(&)[K]>(Vp_t), win);
      slide_i32<R, K, HL4>(win, Sp);
      exch_window<K, HL4>(xb + 3 * planew, reinterpret_cast<u32(&)[K]>(VIp_t), win);
      slide_i32<R, K, HL4>(win, SIp);
      ab_row<K>(sbase + (long long)t2 * pitch, plane_elems, Sp, SIp, at, bt);
#pragma unroll
      for (int c = 0; c < K; ++c) { al[c] -= at[c]; bl[c] -= bt[c]; }
    }
    // ---------------- stage 2, vertical
#pragma unroll
    for (int c = 0; c < K; ++c) { VA[c] += al[c]; VB[c] += bl[c]; }

    const int y = t - R;  // output row
    if (y >= yb0) {
      exch_store<K, HL4>(xb + 4 * planew, reinterpret_cast<u32(&)[K]>(VA));
      exch_store<K, HL4>(xb + 5 * planew, reinterpret_cast<u32(&)[K]>(VB));
    }
    __syncthreads();
    if (y < yb0) continue;

    // ---------------- stage 2, horizontal + q + WTA
    float A[K], B[K];
    {
      u32 win[HL4 + K + HL4];
      exch_window<K, HL4>(xb + 4 * planew, reinterpret_cast<u32(&)[K]>(VA), win);
      slide_f32<R, K, HL4>(win, A);
      exch_window<K, HL4>(xb + 5 * planew, reinterpret_cast<u32(&)[K]>(VB), win);
      slide_f32<R, K, HL4>(win, B);
    }
    float ic[K], invn[K];
    load_f32x<K>(sbase + ST_IC * plane_elems + (long long)y * pitch, ic);
    load_f32x<K>(sbase + ST_INVN * plane_elems + (long long)y * pitch, invn);
    u32 key[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const float q = fmaf(A[c], ic[c], B[c]) * invn[c];
      key[c] = (u32)sortable_i32(q) ^ 0x80000000u;  // unsigned order == float order
    }
    if constexpr (EXPORT) {
      const int de = d - g.export_d0;
      if (de >= 0 && de < g.export_nd && d < g.d_end) {
        float* out = reinterpret_cast<float*>(g.export_ptr) + ((size_t)de * H + y) * W;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const int x = x0 + c;
          if (x >= out0 && x < min(out0 + g.TW, W)) out[x] = unsortable_f32((int)(key[c] ^ 0x80000000u));
        }
      }
    }
    if (!all_valid) {
#pragma unroll
      for (int c = 0; c < K; ++c)
        if (c < c_lo || c > c_hi) key[c] = 0xffffffffu;
    }
    u32 mine = 0xffffffffu;
    int mine_d = 0;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const u32 m = __reduce_min_sync(0xffffffffu, key[c]);
      const u32 who = __ballot_sync(0xffffffffu, key[c] == m);  // strict '<': the lowest d among equal costs wins
      if (lane == c) { mine = m; mine_d = d - lane + (__ffs(who) - 1); }
    }
    if (lane < K && mine != 0xffffffffu) {
      const i64 k64 = (i64)(((unsigned long long)(mine ^ 0x80000000u) << 32) | (u32)mine_d);
      atomicMin(keys + ((size_t)frame * H + y) * W + x0 + lane, k64);
    }
  }
}

}  // namespace gsm
