// gsm_gf3.cuh -- fused AD -> guided-filter aggregation -> WTA kernel (GSM_MODE_GF).
//
// Nothing of the D x H x W volume is materialised.  A CTA owns (strip of columns) x (32 disparities) and marches down
// the rows; thread = run of 16 columns x one disparity.  Because box(a), box(b) need a and b on the 2r+1 rows around
// the output row, two instances of the exact integer stage-1 pipeline run 2r+1 rows apart ("lead" adds a row of
// (a,b) to the stage-2 running sums, "trail" recomputes the row that leaves) -- ~30% more arithmetic instead of a
// (2r+1)-row ring of float (a,b) rows, which would not fit on chip for more than ~8 disparities per CTA.
//
// Stage 1 is HORIZONTAL-FIRST and entirely inside the thread: it builds the horizontal window sums of p and I*p of
// its 16 columns directly from the staged image bytes of the 16+2r columns around them (AD is 4 pixels per
// VABSDIFF4, so the redundant halo pixels are nearly free; the slide is one PRMT + two IDP.2A per column and row,
// with the (+I_in, -I_out) coefficient word precomputed per pixel), and accumulates them vertically:
//      S(t) += H(t+r) - H(t-r-1)           (lead)        S'(t-2r-1) += H(t-r-1) - H(t-3r-2)      (trail)
// The box sums S_p, S_Ip come out exact without any exchange; the numerator N*S_Ip - S_I*S_p is evaluated modulo
// 2^32, exact because |N^2 cov| < 2^31 for r <= 9.  (An earlier layout summed vertically first and exchanged four
// integer planes through shared memory: 6 exchange planes, two barriers per row, strip halo 2r -- see git history.)
// What remains in shared memory is the stage-2 exchange of (V_A, V_B) -- double buffered, ONE barrier per row -- and
// the asynchronous input stage (cp.async.bulk + mbarrier, one step ahead).  WTA: lane index in the 5 low bits of the
// sortable key, REDUX.MIN over the warp's disparities, one 64-bit atomicMin per pixel into the packed-min plane.
// gsm_gf.cuh describes the stage-2 numerics (local centres).
#pragma once
#include "gsm_gf.cuh"

namespace gsm {

// IDP.2A coefficient plane of the horizontal slide: HC[y][x] = I[y][x+R] - 65536 * I[y][x-R-1]
// (lo16 = +I of the column entering the window of x, hi16 = -I of the column leaving it).  Stored in ST_COEF by
// gf_prepass_kernel (gsm_gf.cuh).

// Stage of one march step (TWt strip columns):
//   G[3][TWt+32] u8      guide rows t+R, t-R-1, t-3R-2, columns [xs-16, xs+TWt+16)
//   O[3][TWt+64] u8      other-image rows, shifted window covering the CTA's LPR disparities and the +-12 halo
//   HC[3][TWt] i32       slide coefficients of the same three rows
//   ST[2][5][TWt]        N, S_I, 1/den, mean_I-128, 1/N at rows t (lead) and t-2R-1 (trail)
//   ICY[TWt], INVNY[TWt] I-128 and 1/N at the output row t-R;  CEN[runs] local centres of the output row
struct Gf3Stage {
  int TWt, GW, OW, CENB;
  int off_O, off_HC, off_ST, off_ICY, off_INVNY, off_CEN, bytes;
  __host__ __device__ constexpr Gf3Stage(int twt, int runs)
      : TWt(twt), GW(twt + 32), OW(twt + 64), CENB(4 * ((runs + 3) / 4 * 4)), off_O(3 * (twt + 32)),
        off_HC(3 * (twt + 32) + 3 * (twt + 64)), off_ST(3 * (twt + 32) + 3 * (twt + 64) + 12 * twt),
        off_ICY(3 * (twt + 32) + 3 * (twt + 64) + 52 * twt), off_INVNY(3 * (twt + 32) + 3 * (twt + 64) + 56 * twt),
        off_CEN(3 * (twt + 32) + 3 * (twt + 64) + 60 * twt),
        bytes(3 * (twt + 32) + 3 * (twt + 64) + 60 * twt + 4 * ((runs + 3) / 4 * 4)) {}  // == sum of the bulk copies
};

// Input stages in flight: the copies of step t + GF3_NST - 1 are issued during step t (right after its barrier), so
// with 3 stages the bulk copies have a whole march step to land.  (With 2 stages they had only the stage-2 + WTA
// part of a step: ncu showed 11.5% of all warp samples waiting on the stage mbarrier.)
#ifndef GSM_GF_STAGES
#define GSM_GF_STAGES 3
#endif
constexpr int GF3_NST = GSM_GF_STAGES;
#ifndef GSM_GF_INIT2
#define GSM_GF_INIT2 1
#endif
#ifndef GSM_GF_MINB
#define GSM_GF_MINB 1  // CTAs per SM the register allocation is sized for
#endif
static_assert(GF3_NST >= 2 && GF3_NST <= 4, "2..4 input stages");

__host__ __device__ inline size_t gf3_smem_bytes(int runs, int K, int HL4, int LPR) {
  // barriers + centres | GF3_NST input stages | 2 (double buffer) x 2 (V_A, V_B) exchange planes
  return 512 + GF3_NST * (size_t)Gf3Stage(runs * K, runs).bytes +
         4 * (size_t)LPR * exch_pitch_words(runs, K, HL4) * sizeof(u32);
}

// window of a thread per staged row: columns x0-12 .. x0+K+11, i.e. K+24 bytes = WW = 10 words

// AD bytes of the thread's window of one staged row.  growk = staged guide row + run*K (byte 0 is column x0-16).
template <int K>
__device__ __forceinline__ void gf3_ad_window(const u8* growk, const u8* orow, int ooff, u32 (&g)[(K + 24) / 4],
                                              u32 (&p)[(K + 24) / 4]) {
  constexpr int WW = (K + 24) / 4;
  static_assert(K == 16, "16-byte aligned runs: three 128-bit loads");
  const uint4 a = reinterpret_cast<const uint4*>(growk)[0];
  const uint4 b = reinterpret_cast<const uint4*>(growk)[1];
  const uint4 c = reinterpret_cast<const uint4*>(growk)[2];
  g[0] = a.y; g[1] = a.z; g[2] = a.w; g[3] = b.x; g[4] = b.y; g[5] = b.z; g[6] = b.w; g[7] = c.x; g[8] = c.y; g[9] = c.z;
  u32 ow[WW];
  lds_unaligned<K + 24>(orow, ooff, ow);
#pragma unroll
  for (int i = 0; i < WW; ++i) p[i] = __vabsdiffu4(g[i], ow[i]);
}

// byte mask (0xff per valid byte) of window word i: column inside the image and, for the left view, x >= d
__device__ __forceinline__ u32 gf3_word_mask(int xw, int W, int lo) {
  u32 m = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) m |= (xw + b < W && xw + b >= lo) ? (0xffu << (8 * b)) : 0u;
  return m;
}

// mask of the bytes of window word i (bytes 4i..4i+3) that lie in [LO, HI]
template <int LO, int HI>
__host__ __device__ constexpr u32 range_mask(int i) {
  u32 m = 0;
  for (int b = 0; b < 4; ++b)
    if (4 * i + b >= LO && 4 * i + b <= HI) m |= 0xffu << (8 * b);
  return m;
}

// window sums of p and I*p over window bytes [12-R, 12+R] (the window of the thread's column 0)
template <int R, int WW>
__device__ __forceinline__ void gf3_init_sums(const u32 (&g)[WW], const u32 (&p)[WW], int& hp, int& hip) {
  constexpr int LO = 12 - R, HI = 12 + R;
#if GSM_GF_INIT2
  // two accumulators per sum: the dp4a chains are half as long (exact integers: the order does not matter)
  u32 sp[2] = {0, 0}, sip[2] = {0, 0};
#pragma unroll
  for (int i = 0; i < WW; ++i) {
    const u32 m = range_mask<LO, HI>(i);
    if (m != 0) {
      sp[i & 1] = __dp4a(p[i], m & 0x01010101u, sp[i & 1]);
      sip[i & 1] = __dp4a(g[i] & m, p[i], sip[i & 1]);
    }
  }
  hp = (int)(sp[0] + sp[1]);
  hip = (int)(sip[0] + sip[1]);
#else
  u32 sp = 0, sip = 0;
#pragma unroll
  for (int i = 0; i < WW; ++i) {
    const u32 m = range_mask<LO, HI>(i);
    if (m != 0) {
      sp = __dp4a(p[i], m & 0x01010101u, sp);
      sip = __dp4a(g[i] & m, p[i], sip);
    }
  }
  hp = (int)sp;
  hip = (int)sip;
#endif
}

#ifdef GSM_GF_PROFILE
__device__ unsigned long long g_gf_prof[16][9];  // [run][phase] cycles, summed over all CTAs of a launch
#endif

template <int R, int K, int RUNS, int LPR, bool EXPORT>
__global__ void __launch_bounds__(RUNS * LPR, GSM_GF_MINB)
gf3_wta_kernel(const u8* __restrict__ Gp, const u8* __restrict__ Op, const float* __restrict__ stats,
               i64* __restrict__ keys, FusedGeom g) {
  static_assert(K == 16 && R <= 12 && R >= 1 && R < K, "16-column runs, halo of at most 12 columns");
  constexpr int WW = (K + 24) / 4;
  static_assert(LPR == 32 || LPR == 16, "lanes per run");
  constexpr int HL4 = (R + 3) / 4 * 4;
  extern __shared__ __align__(128) u8 smem_raw[];

  const int lane = threadIdx.x & (LPR - 1);
  const int run = threadIdx.y * (WARP / LPR) + threadIdx.x / LPR;
  constexpr int runs = RUNS;
  const int strip = blockIdx.x;
  const int d0 = g.d_begin + blockIdx.y * LPR;
  const int d = d0 + lane;
  const int frame = blockIdx.z / g.bands;
  const int band = blockIdx.z - frame * g.bands;
  int H = g.pg.H, W = g.pg.W;
  size_t koff = (size_t)frame * H * W;
  if (g.ft) { const FrameDesc fd = g.ft[frame]; H = fd.H; W = fd.W; koff = (size_t)fd.off; }
  keys += koff;  // this frame's packed-min plane
  const int pitch = g.pg.pitch;
  const int yb0 = band * g.band_rows;
  const int yb1 = min(H, yb0 + g.band_rows);
  if (yb0 >= H || strip * g.TW >= W) return;

  constexpr int TWt = runs * K;
  constexpr Gf3Stage sg(TWt, runs);
  constexpr int pitchw = exch_pitch_words(runs, K, HL4);
  constexpr int planew = LPR * pitchw;
  float* ccs = reinterpret_cast<float*>(smem_raw + 64);  // [2][<=48] per-run centres, double buffered
  u8* stage_base = smem_raw + 512;
  u32* exch = reinterpret_cast<u32*>(stage_base + GF3_NST * sg.bytes);  // [2 buffers][V_A, V_B][LPR][pitchw]
  const u32 bar0 = smem_u32(smem_raw);
  const bool producer = (threadIdx.x == 0 && threadIdx.y == 0);

  for (int i = threadIdx.y * WARP + threadIdx.x; i < 4 * planew; i += runs * LPR) exch[i] = 0u;
  if (producer) {
#pragma unroll
    for (int i = 0; i < GF3_NST; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  u32* xb = exch + (size_t)lane * pitchw + HL4 + run * K;

  const int xs = strip * g.TW - g.hl;
  const int x0 = xs + run * K;
  const size_t plane_elems = g.pg.plane_stride;
  const int row_lo = -PADV, row_hi = H + PADV - 1;

  const size_t org = (size_t)PADV * pitch + g.pg.xoff + xs;
  const u8* gsrc = Gp + (size_t)frame * g.pg.plane_stride + org - 16;
  const int ostart = g.pg.xoff + xs - 12 + (g.view == 0 ? -(d0 + LPR - 1) : d0);
  const int oalign = ostart & 15;
  const u8* osrc = Op + (size_t)frame * g.pg.plane_stride + (size_t)PADV * pitch + (ostart - oalign);
  const float* ssrc = stats + (size_t)frame * GF_STAT_PLANES * plane_elems + org;
  constexpr int CENW = (RUNS + 3) / 4 * 4;  // centres of one strip: RUNS floats padded to 16 bytes
  const float* csrc = stats + ((size_t)frame * GF_STAT_PLANES + ST_CEN) * plane_elems + (size_t)PADV * pitch +
                      (size_t)strip * CENW;
  const int ooff = oalign + run * K + (g.view == 0 ? (LPR - 1 - lane) : lane);

  auto issue = [&](int t, int s) {
    const u32 bar = bar0 + 8 * s;
    const u32 dst = smem_u32(stage_base + (size_t)s * sg.bytes);
    mbar_expect_tx(bar, (u32)sg.bytes);
    const int rows3[3] = {t + R, t - R - 1, t - 3 * R - 2};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const long long ro = (long long)max(row_lo, min(row_hi, rows3[i])) * pitch;
      bulk_g2s(dst + i * sg.GW, gsrc + ro, sg.GW, bar);
      bulk_g2s(dst + sg.off_O + i * sg.OW, osrc + ro, sg.OW, bar);
      bulk_g2s(dst + sg.off_HC + i * 4 * TWt, ssrc + ST_COEF * plane_elems + ro, 4 * TWt, bar);
    }
    const int rows2[2] = {t, t - 2 * R - 1};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long ro = (long long)max(row_lo, min(row_hi, rows2[i])) * pitch;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        bulk_g2s(dst + sg.off_ST + (i * 5 + k) * 4 * TWt, ssrc + (size_t)k * plane_elems + ro, 4 * TWt, bar);
    }
    const long long ry = (long long)max(row_lo, min(row_hi, t - R)) * pitch;
    bulk_g2s(dst + sg.off_ICY, ssrc + ST_IC * plane_elems + ry, 4 * TWt, bar);
    bulk_g2s(dst + sg.off_INVNY, ssrc + ST_INVN * plane_elems + ry, 4 * TWt, bar);
    bulk_g2s(dst + sg.off_CEN, csrc + ry, sg.CENB, bar);
  };

  // window bytes that may contribute: inside the image and (left view) x >= d  (BlockMatching.cpp:147-149)
  const int dd = min(d, MAX_DISP - 1);
  const int col_lo = g.view == 0 ? dd : 0;
  const bool full = (x0 - 12 >= col_lo) && (x0 + K + 11 < W);
  const bool need_mask = __any_sync(0xffffffffu, !full);

  const int out0 = strip * g.TW;
  const int c_lo = max(0, out0 - x0);
  int c_hi = min(K - 1, min(out0 + g.TW, W) - 1 - x0);
  if (d >= g.d_end) c_hi = -1;
  const bool all_valid = __all_sync(0xffffffffu, c_lo == 0 && c_hi == K - 1);
  // warp-uniform stage skipping: (a, b) is needed on strip columns [hl-R, hl+TW+R) inside the image (+-R)
  const bool need_out = __any_sync(0xffffffffu, min(K - 1, min(out0 + g.TW, W) - 1 - x0) >= c_lo);
  const bool need_ab = __any_sync(
      0xffffffffu, (run * K < g.hl + g.TW + R) && (run * K + K > g.hl - R) && (x0 < W + R) && (x0 + K > -R));

  int Sp_l[K], SIp_l[K], Sp_t[K], SIp_t[K];
  float VA[K], VB[K];
#pragma unroll
  for (int c = 0; c < K; ++c) { Sp_l[c] = SIp_l[c] = Sp_t[c] = SIp_t[c] = 0; VA[c] = VB[c] = 0.f; }
  float cc = 0.f;

  const int r0 = yb0 - 2 * R;
  const int a0 = yb0 - R;
  constexpr int COEF_PM = (int)0xFFFF0001;
  const int t_begin = yb0 - 3 * R, t_end = yb1 + R;

  if (producer) {
#pragma unroll
    for (int i = 0; i < GF3_NST - 1; ++i)
      if (t_begin + i < t_end) issue(t_begin + i, i);
  }

  // ================ part A of row `it` (input stage s): stage 1, (a, b), publish (V_A, V_B)
#ifdef GSM_GF_PROFILE  // per-phase clock() shares of a march step (tools/phase_prof.py); not part of the product build
  unsigned pacc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned plast = (unsigned)clock();
#define GSM_PH(i) { asm volatile("" ::: "memory"); const unsigned now_ = (unsigned)clock(); pacc[i] += now_ - plast; plast = now_; }
#else
#define GSM_PH(i)
#endif
  auto part_a = [&](int it, int s, u32 sphase) {
    const int t = t_begin + it;
    mbar_wait(bar0 + 8 * s, sphase);
    const u8* stg = stage_base + (size_t)s * sg.bytes;
    const int t2 = t - 2 * R - 1;

    GSM_PH(0)  // stage mbarrier wait
    if (need_ab) {
      // ---------------- stage 1: horizontal window sums of the three rows, folded into the vertical sums
      u32 pn[WW], pm[WW], po[WW];
      int hp_n, hip_n, hp_m = 0, hip_m = 0, hp_o = 0, hip_o = 0;
      const bool has_m = t - R - 1 >= r0, has_o = t - 3 * R - 2 >= r0;
      {
        u32 gn[WW];
        gf3_ad_window<K>(stg + run * K, stg + sg.off_O, ooff, gn, pn);
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < WW; ++i) pn[i] &= gf3_word_mask(x0 - 12 + 4 * i, W, col_lo);
        }
        gf3_init_sums<R, WW>(gn, pn, hp_n, hip_n);
      }
      if (has_m) {
        u32 gm[WW];
        gf3_ad_window<K>(stg + sg.GW + run * K, stg + sg.off_O + sg.OW, ooff, gm, pm);
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < WW; ++i) pm[i] &= gf3_word_mask(x0 - 12 + 4 * i, W, col_lo);
        }
        gf3_init_sums<R, WW>(gm, pm, hp_m, hip_m);
      } else {
#pragma unroll
        for (int i = 0; i < WW; ++i) pm[i] = 0u;
      }
      if (has_o) {
        u32 go[WW];
        gf3_ad_window<K>(stg + 2 * sg.GW + run * K, stg + sg.off_O + 2 * sg.OW, ooff, go, po);
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < WW; ++i) po[i] &= gf3_word_mask(x0 - 12 + 4 * i, W, col_lo);
        }
        gf3_init_sums<R, WW>(go, po, hp_o, hip_o);
      } else {
#pragma unroll
        for (int i = 0; i < WW; ++i) po[i] = 0u;
      }
      GSM_PH(1)  // AD windows + initial sums
      const int* hcn = reinterpret_cast<const int*>(stg + sg.off_HC) + run * K;
      const int* hcm = hcn + TWt;
      const int* hco = hcm + TWt;
#pragma unroll
      for (int g4 = 0; g4 < K; g4 += 4) {
        const int4 cn4 = *reinterpret_cast<const int4*>(hcn + g4);
        const int4 cm4 = *reinterpret_cast<const int4*>(hcm + g4);
        const int4 co4 = *reinterpret_cast<const int4*>(hco + g4);
        const int cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w}, cm[4] = {cm4.x, cm4.y, cm4.z, cm4.w},
                  co[4] = {co4.x, co4.y, co4.z, co4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g4 + j;
          if (c > 0) {
            // window byte entering: 12 + c + R, leaving: 12 + c - R - 1
            const int bi = 12 + c + R, bo = 12 + c - R - 1;
            const u32 sel = (u32)(bi & 3) | ((4u + (u32)(bo & 3)) << 4);
            const u32 qn = __byte_perm(pn[bi >> 2], pn[bo >> 2], sel);  // {p_in, p_out, x, x}
            const u32 qm = __byte_perm(pm[bi >> 2], pm[bo >> 2], sel);
            const u32 qo = __byte_perm(po[bi >> 2], po[bo >> 2], sel);
            hp_n = dp2a_lo_su(COEF_PM, qn, hp_n);
            hip_n = dp2a_lo_su(cn[j], qn, hip_n);
            hp_m = dp2a_lo_su(COEF_PM, qm, hp_m);
            hip_m = dp2a_lo_su(cm[j], qm, hip_m);
            hp_o = dp2a_lo_su(COEF_PM, qo, hp_o);
            hip_o = dp2a_lo_su(co[j], qo, hip_o);
          }
          Sp_l[c] += hp_n - hp_m;
          SIp_l[c] += hip_n - hip_m;
          Sp_t[c] += hp_m - hp_o;
          SIp_t[c] += hip_m - hip_o;
        }
      }

      GSM_PH(2)  // horizontal slides
      // ---------------- (a, b) of the lead and trail rows folded into the stage-2 vertical sums
      {
        const float target = reinterpret_cast<const float*>(stg + sg.off_CEN)[run];
        const float dc = target - cc;
        if (fabsf(dc) > GF_RECENTRE) {
#pragma unroll
          for (int c = 0; c < K; ++c) VB[c] = fmaf(dc, VA[c], VB[c]);
          cc = target;
        }
      }
      const float* st_l = reinterpret_cast<const float*>(stg + sg.off_ST) + run * K;
      const bool has_lead = t >= a0, has_trail = t2 >= a0;
#pragma unroll
      for (int g4 = 0; g4 < K; g4 += 4) {
        float da[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          if (which == 0 ? has_lead : has_trail) {
            const float* st = st_l + which * 5 * TWt;
            const int4 N = *reinterpret_cast<const int4*>(st + ST_N * TWt + g4);
            const int4 SI = *reinterpret_cast<const int4*>(st + ST_SI * TWt + g4);
            const float4 invden = *reinterpret_cast<const float4*>(st + ST_INVDEN * TWt + g4);
            const float4 cmean = *reinterpret_cast<const float4*>(st + ST_CMEAN * TWt + g4);
            const float4 invn = *reinterpret_cast<const float4*>(st + ST_INVN * TWt + g4);
            const int Nn[4] = {N.x, N.y, N.z, N.w}, SIi[4] = {SI.x, SI.y, SI.z, SI.w};
            const float idn[4] = {invden.x, invden.y, invden.z, invden.w}, cmv[4] = {cmean.x, cmean.y, cmean.z, cmean.w},
                        inn[4] = {invn.x, invn.y, invn.z, invn.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = g4 + j;
              const int sp = which == 0 ? Sp_l[c] : Sp_t[c];
              const int sip = which == 0 ? SIp_l[c] : SIp_t[c];
              const int num = Nn[j] * sip - SIi[j] * sp;  // exact modulo 2^32; true value fits int32 for r <= 9
              const float a = (float)num * idn[j];
              const float b = fmaf(-a, cmv[j] - cc, (float)sp * inn[j]);
              // lead row enters, trail row leaves: form the small difference first, then ONE rounding at the
              // magnitude of the running sum
              if (which == 0) { da[j] = a; db[j] = b; } else { da[j] -= a; db[j] -= b; }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { VA[g4 + j] += da[j]; VB[g4 + j] += db[j]; }
      }
    }

    GSM_PH(3)  // (a, b)
    const int y = t - R;
    u32* xbuf = xb + (size_t)(it & 1) * 2 * planew;  // double-buffered (V_A, V_B) planes: one barrier per row
    float* ccbuf = ccs + (it & 1) * 48;
    if (need_ab && y >= yb0) {
      exch_store<K, HL4>(xbuf, reinterpret_cast<u32(&)[K]>(VA));
      exch_store<K, HL4>(xbuf + planew, reinterpret_cast<u32(&)[K]>(VB));
      if (lane == 0) ccbuf[run] = cc;
    }
    GSM_PH(4)  // exchange stores
  };

  // ================ part B of row `it` (after the barrier that publishes it): stage 2 horizontal, q, WTA
  auto part_b = [&](int it, int s) {
    const int y = t_begin + it - R;
    if (y < yb0 || !need_out) return;
    const u8* stg = stage_base + (size_t)s * sg.bytes;
    const u32* xbuf = xb + (size_t)(it & 1) * 2 * planew;
    const float* ccbuf = ccs + (it & 1) * 48;
    float A[K], B[K];
    {
      u32 winA[HL4 + K + HL4], winB[HL4 + K + HL4];
      exch_window<K, HL4>(xbuf, reinterpret_cast<u32(&)[K]>(VA), winA);
      exch_window<K, HL4>(xbuf + planew, reinterpret_cast<u32(&)[K]>(VB), winB);
      const float dl = run > 0 ? cc - ccbuf[run - 1] : 0.f;
      const float dr = run + 1 < runs ? cc - ccbuf[run + 1] : 0.f;
      slide_ab<R, K, HL4>(winA, winB, dl, dr, A, B);
    }
    GSM_PH(6)  // exchange loads + stage-2 slide
    // The WTA compares N(x)*q_d(x): N > 0 does not depend on d, so the argmin is that of q (the packed-min plane
    // therefore carries the un-normalised cost); only the exported slices are divided by N.
    const float* icy = reinterpret_cast<const float*>(stg + sg.off_ICY) + run * K;
    float qn[K];
#pragma unroll
    for (int g4 = 0; g4 < K; g4 += 4) {
      const float4 ic = *reinterpret_cast<const float4*>(icy + g4);
      qn[g4 + 0] = fmaf(A[g4 + 0], ic.x - cc, B[g4 + 0]);
      qn[g4 + 1] = fmaf(A[g4 + 1], ic.y - cc, B[g4 + 1]);
      qn[g4 + 2] = fmaf(A[g4 + 2], ic.z - cc, B[g4 + 2]);
      qn[g4 + 3] = fmaf(A[g4 + 3], ic.w - cc, B[g4 + 3]);
    }
    if constexpr (EXPORT) {
      const float* iny = reinterpret_cast<const float*>(stg + sg.off_INVNY) + run * K;
      const int de = d - g.export_d0;
      if (de >= 0 && de < g.export_nd && d < g.d_end) {
        float* out = reinterpret_cast<float*>(g.export_ptr) + ((size_t)de * H + y) * W;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const int x = x0 + c;
          if (x >= out0 && x < min(out0 + g.TW, W)) out[x] = qn[c] * iny[c];
        }
      }
    }
    GSM_PH(7)  // q
    int key[K];
#pragma unroll
    for (int c = 0; c < K; ++c) key[c] = sortable_i32(qn[c]);
#pragma unroll
    for (int c = 0; c < K; ++c) key[c] = (key[c] & ~31) | lane;
    if (!all_valid) {
#pragma unroll
      for (int c = 0; c < K; ++c)
        if (c < c_lo || c > c_hi) key[c] = 0x7fffffff;
    }
    int mine = 0x7fffffff;
    if (LPR == 32) {
      // lane c keeps the minimum of column c: a 4-level select tree on the (loop-invariant) lane bits instead of a
      // compare + select per column
      int m[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) m[c] = c < K ? __reduce_min_sync(0xffffffffu, key[c < K ? c : 0]) : 0x7fffffff;
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = (lane & 1) ? m[2 * i + 1] : m[2 * i];
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = (lane & 2) ? m[2 * i + 1] : m[2 * i];
#pragma unroll
      for (int i = 0; i < 2; ++i) m[i] = (lane & 4) ? m[2 * i + 1] : m[2 * i];
      mine = (lane & 8) ? m[1] : m[0];
    } else {
      const bool upper = (threadIdx.x & 16) != 0;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int m0 = __reduce_min_sync(0xffffffffu, upper ? 0x7fffffff : key[c]);
        const int m1 = __reduce_min_sync(0xffffffffu, upper ? key[c] : 0x7fffffff);
        if (lane == c) mine = upper ? m1 : m0;
      }
    }
    if (lane < K && mine != 0x7fffffff) {
      const i64 k64 = (i64)(((unsigned long long)(u32)(mine & ~31) << 32) | (u32)(d0 + (mine & 31)));
      atomicMin(keys + (size_t)y * W + x0 + lane, k64);
    }
    GSM_PH(8)  // keys + WTA
  };

  const int T = t_end - t_begin;
  int s = 0;       // stage of row it: it % GF3_NST
  u32 sphase = 0;  // its mbarrier parity: (it / GF3_NST) & 1
  for (int it = 0; it < T; ++it) {
    part_a(it, s, sphase);
    __syncthreads();
    // every thread has now finished row it-1 completely: its stage is refilled for row it + GF3_NST - 1
    if (producer && it + GF3_NST - 1 < T) issue(t_begin + it + GF3_NST - 1, s == 0 ? GF3_NST - 1 : s - 1);
    GSM_PH(5)  // row barrier (+ bulk-copy issue in the producer warp)
    part_b(it, s);
    if (++s == GF3_NST) { s = 0; sphase ^= 1u; }
  }
#ifdef GSM_GF_PROFILE
  if (lane == 0)
    for (int i = 0; i < 9; ++i) atomicAdd(&g_gf_prof[run][i], (unsigned long long)pacc[i]);
#endif
#undef GSM_PH
}

}  // namespace gsm
