// gsm_gf5.cuh -- fused AD -> guided-filter aggregation -> WTA kernel (GSM_MODE_GF), "ring" version.
//
// Same decomposition as gsm_gf3.cuh (CTA = strip of columns x 32 disparities marching down the rows, thread = run of
// 16 columns x one disparity, stage 1 exact and horizontal-first inside the thread, stage-2 exchange of (V_A, V_B)
// through double-buffered shared-memory planes with one barrier per row, REDUX WTA into the packed-min plane), but the
// row of (a, b) that LEAVES the stage-2 vertical window is not recomputed by a second ("trail") stage-1 pipeline
// 2r+1 rows behind the first: every thread keeps the (a, b) values of its own columns for the last 2r+1 rows in a
// ring in global memory and reads back the row it wrote 2r+1 steps earlier.
//   * The ring of a CTA (2r+1 rows x 192 columns x 32 disparities x 8 B = 933 KB at r = 9) is addressed by the SM id
//     (one CTA per SM), so the 148 rings (138 MB) are reused by successive CTAs and stay hot in the 126 MB L2: the
//     accesses carry a fractional evict_last policy so that most of the ring is protected from the streaming planes
//     and the remainder streams through HBM instead of thrashing the whole ring.
//   * A thread reads and writes only its own entries (warp-contiguous 1 KB per instruction): no synchronisation, and
//     the value subtracted is bit-identical to the value added 2r+1 rows earlier, so the running sums do not drift.
//   * b was formed against the run's centre of THAT row; the centres of the last 2r+1 rows are kept in shared memory
//     and the leaving b is re-based by the exact identity b'(c2) = b'(c1) + (c2 - c1) a.
// Removes the third AD window / slide, the second pair of vertical sums, half of the (a, b) arithmetic and the second
// set of statistic-plane loads of gsm_gf3.cuh (about a quarter of its instructions and 32 registers per thread).
#pragma once
#include "gsm_gf3.cuh"

namespace gsm {

// Stage of one march step (TWt strip columns):
//   G[2][TWt+32] u8      guide rows t+R, t-R-1, columns [xs-16, xs+TWt+16)
//   O[2][TWt+64] u8      other-image rows, shifted window covering the CTA's LPR disparities and the +-12 halo
//   HC[2][TWt] i32       slide coefficients of the same two rows
//   ST[5][TWt]           N, S_I, 1/den, mean_I-128, 1/N at row t
//   ICY[TWt], INVNY[TWt] I-128 and 1/N at the output row t-R;  CEN[runs] local centres of the output row
struct Gf5Stage {
  int TWt, GW, OW, CENB;
  int off_O, off_HC, off_ST, off_ICY, off_INVNY, off_CEN, bytes;
  __host__ __device__ constexpr Gf5Stage(int twt, int runs)
      : TWt(twt), GW(twt + 32), OW(twt + 64), CENB(4 * ((runs + 3) / 4 * 4)), off_O(2 * (twt + 32)),
        off_HC(2 * (twt + 32) + 2 * (twt + 64)), off_ST(2 * (twt + 32) + 2 * (twt + 64) + 8 * twt),
        off_ICY(2 * (twt + 32) + 2 * (twt + 64) + 28 * twt), off_INVNY(2 * (twt + 32) + 2 * (twt + 64) + 32 * twt),
        off_CEN(2 * (twt + 32) + 2 * (twt + 64) + 36 * twt),
        bytes(2 * (twt + 32) + 2 * (twt + 64) + 36 * twt + 4 * ((runs + 3) / 4 * 4)) {}  // == sum of the bulk copies
};

constexpr int GF5_HDR = 2048;  // mbarriers | centres of the two exchange buffers | centres of the last 2R+1 rows
__host__ __device__ inline size_t gf5_smem_bytes(int runs, int K, int HL4, int LPR) {
  const size_t b = GF5_HDR + GF3_NST * (size_t)Gf5Stage(runs * K, runs).bytes +
                   4 * (size_t)LPR * exch_pitch_words(runs, K, HL4) * sizeof(u32);
  return b < 120 * 1024 ? 120 * 1024 : b;  // more than half an SM: one CTA per SM, so the SM id names the CTA's ring
}
// floats of one ring row of a CTA: (a, b) x strip columns x disparities
__host__ __device__ constexpr size_t gf5_ring_row_floats(int runs, int K, int LPR) { return (size_t)runs * K * LPR * 2; }

#ifndef GSM_GF_RING_STREAM_ROWS
#define GSM_GF_RING_STREAM_ROWS 0
#endif
#ifndef GSM_GF_RING_HINT
#define GSM_GF_RING_HINT 1  // 1: per-instruction fractional evict_last policy; 0: plain accesses
#endif
__device__ __forceinline__ float4 ld_hint(const float4* p, unsigned long long pol) {
#if !GSM_GF_RING_HINT
  (void)pol;
  return *p;
#endif
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_hint(float4* p, const float4& v, unsigned long long pol) {
#if !GSM_GF_RING_HINT
  (void)pol;
  *p = v;
  return;
#endif
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}

template <int R, int K, int RUNS, int LPR, bool EXPORT>
__global__ void __launch_bounds__(RUNS * LPR, 1)
gf5_wta_kernel(const u8* __restrict__ Gp, const u8* __restrict__ Op, const float* __restrict__ stats,
               i64* __restrict__ keys, float* __restrict__ ring, FusedGeom g) {
  static_assert(K == 16 && R <= 12 && R >= 1 && R < K, "16-column runs, halo of at most 12 columns");
  static_assert(LPR == 32, "one warp = the 32 disparities of a run");
  constexpr int WW = (K + 24) / 4;
  constexpr int HL4 = (R + 3) / 4 * 4;
  constexpr int RD = 2 * R + 1;  // ring depth: rows of (a, b) inside the vertical window
  extern __shared__ __align__(128) u8 smem_raw[];

  const int lane = threadIdx.x;
  const int run = threadIdx.y;
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
  constexpr Gf5Stage sg(TWt, runs);
  constexpr int pitchw = exch_pitch_words(runs, K, HL4);
  constexpr int planew = LPR * pitchw;
  float* ccs = reinterpret_cast<float*>(smem_raw + 64);        // [2][48] per-run centres, double buffered
  float* ccring = reinterpret_cast<float*>(smem_raw + 512);    // [RD][16] centre of each of the last RD rows (own run)
  static_assert(512 + RD * RUNS * 4 <= GF5_HDR && RUNS <= 16, "header layout");
  u8* stage_base = smem_raw + GF5_HDR;
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

  // this CTA's ring: one per SM; this thread's entries: [row slot][4-column group][disparity lane][a4 | b4]
  u32 smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  unsigned long long pol_keep = 0, pol_stream = 0;
#if GSM_GF_RING_HINT
  // GSM_GF_RING_STREAM_ROWS of the 2R+1 row slots bypass the L2 working set (evict_first), the others are kept
  // (evict_last): the cached part of the 148 rings is sized to fit the L2, the rest streams through HBM
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
#endif
  constexpr size_t ROW4 = (size_t)(TWt / 4) * LPR * 2;  // float4 per ring row
  // [row slot][a | b][4-column group][disparity lane]: every load / store instruction of the warp covers 512 contiguous bytes
  float4* rbase = reinterpret_cast<float4*>(ring) + (size_t)smid * RD * ROW4 + (size_t)(run * (K / 4)) * LPR + lane;
  constexpr size_t RB = ROW4 / 2;  // offset of the b half of a ring row

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
    const int rows2[2] = {t + R, t - R - 1};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long ro = (long long)max(row_lo, min(row_hi, rows2[i])) * pitch;
      bulk_g2s(dst + i * sg.GW, gsrc + ro, sg.GW, bar);
      bulk_g2s(dst + sg.off_O + i * sg.OW, osrc + ro, sg.OW, bar);
      bulk_g2s(dst + sg.off_HC + i * 4 * TWt, ssrc + ST_COEF * plane_elems + ro, 4 * TWt, bar);
    }
    {
      const long long ro = (long long)max(row_lo, min(row_hi, t)) * pitch;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        bulk_g2s(dst + sg.off_ST + k * 4 * TWt, ssrc + (size_t)k * plane_elems + ro, 4 * TWt, bar);
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

  int Sp[K], SIp[K];
  float VA[K], VB[K];
#pragma unroll
  for (int c = 0; c < K; ++c) { Sp[c] = SIp[c] = 0; VA[c] = VB[c] = 0.f; }
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

  // ================ part A of row `it` (input stage s, ring slot rs): stage 1, (a, b), publish (V_A, V_B)
  auto part_a = [&](int it, int s, u32 sphase, int rs) {
    const int t = t_begin + it;
    const int t2 = t - RD;  // the row that leaves the vertical window: its (a, b) are in ring slot rs
    const bool has_lead = t >= a0, has_trail = t2 >= a0;
    float4* rrow = rbase + (size_t)rs * ROW4;
    const unsigned long long pol = rs < GSM_GF_RING_STREAM_ROWS ? pol_stream : pol_keep;
    float4 ra[K / 4], rb[K / 4];
    if (need_ab && has_trail) {  // issued first: an L2 round trip hides behind stage 1
#pragma unroll
      for (int q = 0; q < K / 4; ++q) {
        ra[q] = ld_hint(rrow + (size_t)q * LPR, pol);
        rb[q] = ld_hint(rrow + RB + (size_t)q * LPR, pol);
      }
    }
    mbar_wait(bar0 + 8 * s, sphase);
    const u8* stg = stage_base + (size_t)s * sg.bytes;

    if (need_ab) {
      // ---------------- stage 1: horizontal window sums of the two rows, folded into the vertical sums
      u32 pn[WW], pm[WW];
      int hp_n, hip_n, hp_m = 0, hip_m = 0;
      const bool has_m = t - R - 1 >= r0;
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
      const int* hcn = reinterpret_cast<const int*>(stg + sg.off_HC) + run * K;
      const int* hcm = hcn + TWt;
#pragma unroll
      for (int g4 = 0; g4 < K; g4 += 4) {
        const int4 cn4 = *reinterpret_cast<const int4*>(hcn + g4);
        const int4 cm4 = *reinterpret_cast<const int4*>(hcm + g4);
        const int cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w}, cm[4] = {cm4.x, cm4.y, cm4.z, cm4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g4 + j;
          if (c > 0) {
            // window byte entering: 12 + c + R, leaving: 12 + c - R - 1
            const int bi = 12 + c + R, bo = 12 + c - R - 1;
            const u32 sel = (u32)(bi & 3) | ((4u + (u32)(bo & 3)) << 4);
            const u32 qn = __byte_perm(pn[bi >> 2], pn[bo >> 2], sel);  // {p_in, p_out, x, x}
            const u32 qm = __byte_perm(pm[bi >> 2], pm[bo >> 2], sel);
            hp_n = dp2a_lo_su(COEF_PM, qn, hp_n);
            hip_n = dp2a_lo_su(cn[j], qn, hip_n);
            hp_m = dp2a_lo_su(COEF_PM, qm, hp_m);
            hip_m = dp2a_lo_su(cm[j], qm, hip_m);
          }
          Sp[c] += hp_n - hp_m;
          SIp[c] += hip_n - hip_m;
        }
      }

      // ---------------- (a, b) of the entering row; the leaving row comes back from the ring
      {
        const float target = reinterpret_cast<const float*>(stg + sg.off_CEN)[run];
        const float dc = target - cc;
        if (fabsf(dc) > GF_RECENTRE) {
#pragma unroll
          for (int c = 0; c < K; ++c) VB[c] = fmaf(dc, VA[c], VB[c]);
          cc = target;
        }
      }
      // centre the leaving row's b was formed against -> re-base it to the current centre
      const float dct = has_trail ? cc - ccring[rs * 16 + run] : 0.f;
      const float* st_l = reinterpret_cast<const float*>(stg + sg.off_ST) + run * K;
#pragma unroll
      for (int g4 = 0; g4 < K; g4 += 4) {
        float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (has_lead) {
          const int4 N = *reinterpret_cast<const int4*>(st_l + ST_N * TWt + g4);
          const int4 SI = *reinterpret_cast<const int4*>(st_l + ST_SI * TWt + g4);
          const float4 invden = *reinterpret_cast<const float4*>(st_l + ST_INVDEN * TWt + g4);
          const float4 cmean = *reinterpret_cast<const float4*>(st_l + ST_CMEAN * TWt + g4);
          const float4 invn = *reinterpret_cast<const float4*>(st_l + ST_INVN * TWt + g4);
          const int Nn[4] = {N.x, N.y, N.z, N.w}, SIi[4] = {SI.x, SI.y, SI.z, SI.w};
          const float idn[4] = {invden.x, invden.y, invden.z, invden.w}, cmv[4] = {cmean.x, cmean.y, cmean.z, cmean.w},
                      inn[4] = {invn.x, invn.y, invn.z, invn.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = g4 + j;
            const int num = Nn[j] * SIp[c] - SIi[j] * Sp[c];  // exact modulo 2^32; true value fits int32 for r <= 9
            a4[j] = (float)num * idn[j];
            b4[j] = fmaf(-a4[j], cmv[j] - cc, (float)Sp[c] * inn[j]);
          }
          st_hint(rrow + (size_t)(g4 / 4) * LPR, make_float4(a4[0], a4[1], a4[2], a4[3]), pol);
          st_hint(rrow + RB + (size_t)(g4 / 4) * LPR, make_float4(b4[0], b4[1], b4[2], b4[3]), pol);
        }
        if (has_trail) {
          const float ta[4] = {ra[g4 / 4].x, ra[g4 / 4].y, ra[g4 / 4].z, ra[g4 / 4].w};
          const float tb[4] = {rb[g4 / 4].x, rb[g4 / 4].y, rb[g4 / 4].z, rb[g4 / 4].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // entering minus leaving: the small difference first, then ONE rounding at the magnitude of the sum
            a4[j] -= ta[j];
            b4[j] -= fmaf(dct, ta[j], tb[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { VA[g4 + j] += a4[j]; VB[g4 + j] += b4[j]; }
      }
      __syncwarp();  // every lane has read the leaving row's centre from this slot
      if (has_lead && lane == 0) ccring[rs * 16 + run] = cc;
    }

    const int y = t - R;
    u32* xbuf = xb + (size_t)(it & 1) * 2 * planew;  // double-buffered (V_A, V_B) planes: one barrier per row
    float* ccbuf = ccs + (it & 1) * 48;
    if (need_ab && y >= yb0) {
      exch_store<K, HL4>(xbuf, reinterpret_cast<u32(&)[K]>(VA));
      exch_store<K, HL4>(xbuf + planew, reinterpret_cast<u32(&)[K]>(VB));
      if (lane == 0) ccbuf[run] = cc;
    }
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
    // lane c keeps the minimum of column c: a 4-level select tree on the (loop-invariant) lane bits instead of a
    // compare + select per column
    int m[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) m[c] = __reduce_min_sync(0xffffffffu, key[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = (lane & 1) ? m[2 * i + 1] : m[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = (lane & 2) ? m[2 * i + 1] : m[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) m[i] = (lane & 4) ? m[2 * i + 1] : m[2 * i];
    const int mine = (lane & 8) ? m[1] : m[0];
    if (lane < K && mine != 0x7fffffff) {
      const i64 k64 = (i64)(((unsigned long long)(u32)(mine & ~31) << 32) | (u32)(d0 + (mine & 31)));
      atomicMin(keys + (size_t)y * W + x0 + lane, k64);
    }
  };

  const int T = t_end - t_begin;
  int s = 0;       // stage of row it: it % GF3_NST
  u32 sphase = 0;  // its mbarrier parity: (it / GF3_NST) & 1
  int rs = 0;      // ring slot of row it: it % RD
  for (int it = 0; it < T; ++it) {
    part_a(it, s, sphase, rs);
    __syncthreads();
    // every thread has now finished row it-1 completely: its stage is refilled for row it + GF3_NST - 1
    if (producer && it + GF3_NST - 1 < T) issue(t_begin + it + GF3_NST - 1, s == 0 ? GF3_NST - 1 : s - 1);
    part_b(it, s);
    if (++s == GF3_NST) { s = 0; sphase ^= 1u; }
    if (++rs == RD) rs = 0;
  }
}

}  // namespace gsm
