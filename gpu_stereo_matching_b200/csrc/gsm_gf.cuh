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
//            b is formed against a LOCAL centre: b' = mean_p - a*(mean_I - c_run), c_run = rounded local mean of the
//            guide over the thread's 16-column run, re-centred (VB += dc*VA, exact identity) when it drifts by more
//            than GF_RECENTRE grey levels.  q = (A*(I - c_run) + B')/N is independent of c in exact arithmetic, but
//            with c near the local intensities the two terms no longer cancel, which is what fp32 needs to stay
//            within 1e-4 of the float64 oracle.  Halo columns owned by neighbour runs (different centres) are
//            converted through the partial sums A_L, A_R of the window: B' += (c_own - c_nbr) * A_{L|R}.
//   WTA:     warp min over the 32 disparities of a run (REDUX on the sortable bit pattern, ballot for the
//            lowest d among equals), one 64-bit atomicMin per pixel into the packed-min plane.
#pragma once
#include "gsm_common.cuh"
#include "gsm_sad.cuh"

namespace gsm {

// guide statistic planes (float/int32, same padded geometry as the u8 planes; zero outside the image)
constexpr int GF_STAT_PLANES = 8;
enum { ST_N = 0, ST_SI = 1, ST_INVDEN = 2, ST_CMEAN = 3, ST_INVN = 4, ST_IC = 5, ST_COEF = 6, ST_CEN = 7 };
constexpr float GF_CENTRE = 128.0f;
constexpr float GF_RECENTRE = 8.0f;
#ifndef GSM_GF_FRESH
#define GSM_GF_FRESH 0  // 1: add-only shadow accumulators swapped in every 2R+1 rows (bounds stage-2 drift; ~15% slower)
#endif

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

// Local centre plane: for every image row and every 16-column block of the padded grid (the blocks the fused
// kernel's runs are aligned to), the rounded mean of (mean_I - 128) over the block's in-image columns.  Stored 4x
// replicated (one float per 4 columns) so that a strip's centres are a 16-byte aligned, contiguous run.
__global__ void gf_centre_kernel(float* __restrict__ stats, PlaneGeom pg) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;  // 16-column block of the padded row
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (blk * 16 >= pg.pitch) return;
  float* base = stats + (size_t)f * GF_STAT_PLANES * pg.plane_stride;
  const float* cm = base + ST_CMEAN * pg.plane_stride + (size_t)(PADV + y) * pg.pitch + blk * 16;
  float sum = 0.f;
  int n = 0;
  for (int i = 0; i < 16; ++i) {
    const int x = blk * 16 + i - pg.xoff;
    if (x >= 0 && x < pg.W) { sum += cm[i]; ++n; }
  }
  const float c = n ? rintf(sum / (float)n) : 0.f;
  float* cen = base + ST_CEN * pg.plane_stride + (size_t)(PADV + y) * pg.pitch + blk * 4;
  cen[0] = cen[1] = cen[2] = cen[3] = c;
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
// sum of win[LO..HI] (inclusive, window-relative indices) with three independent partial sums (short dependency chain)
template <int LO, int HI, int N>
__device__ __forceinline__ int wsum_i(const u32 (&win)[N], int base) {
  int p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
  for (int j = LO; j <= HI; ++j) {
    const int v = (int)win[base + j];
    if ((j - LO) % 3 == 0) p0 += v; else if ((j - LO) % 3 == 1) p1 += v; else p2 += v;
  }
  return p0 + p1 + p2;
}
template <int LO, int HI, int N>
__device__ __forceinline__ float wsum_f(const u32 (&win)[N], int base) {
  float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
  for (int j = LO; j <= HI; ++j) {
    const float v = __uint_as_float(win[base + j]);
    if ((j - LO) % 3 == 0) p0 += v; else if ((j - LO) % 3 == 1) p1 += v; else p2 += v;
  }
  return (p0 + p1) + p2;
}

// Horizontal sliding window sums over the thread's K = 16 columns as TWO independent chains (columns 0..7 and
// 8..15): with only ~3 warps per scheduler the dependent add chain of a single slide is what the issue slots wait
// on.  The two starting sums share their overlapping middle part.
template <int R, int K, int HL4>
__device__ __forceinline__ void slide_i32(const u32 (&win)[HL4 + K + HL4], int (&S)[K]) {
  static_assert(K == 16, "two chains of 8");
  constexpr int H = K / 2;
  int sA, sB;
  if constexpr (2 * R + 1 > H) {
    const int mid = wsum_i<H - R, R>(win, HL4);                // columns common to both start windows
    sA = mid + wsum_i<-R, H - R - 1>(win, HL4);
    sB = mid + wsum_i<R + 1, H + R>(win, HL4);
  } else {
    sA = wsum_i<-R, R>(win, HL4);
    sB = wsum_i<H - R, H + R>(win, HL4);
  }
  S[0] = sA;
  S[H] = sB;
#pragma unroll
  for (int c = 1; c < H; ++c) {
    sA += (int)win[HL4 + c + R] - (int)win[HL4 + c - R - 1];
    sB += (int)win[HL4 + H + c + R] - (int)win[HL4 + H + c - R - 1];
    S[c] = sA;
    S[H + c] = sB;
  }
}

// ---- asynchronous row staging (cp.async.bulk == TMA 1-D, completion on an mbarrier) -------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  u32 done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

// Shared-memory stage holding every global input of ONE march step for the CTA's strip (TWt columns):
//   G[3][TWt] u8        guide rows   t+R, t-R-1, t-3R-2
//   O[3][TWt+64] u8     other-image rows, shifted window covering the CTA's 32 disparities
//   COEF[2][TWt] i32    IDP.2A coefficient rows t, t-2R-1
//   ST[2][5][TWt]       N, S_I, 1/den, mean_I-128, 1/N at rows t (lead) and t-2R-1 (trail)
//   ICY[TWt], INVNY[TWt] f32   I-128 and 1/N at the output row t-R
//   CEN[TWt/4] f32     local centre (mean_I - 128, rounded) of each 16-column run at the output row, 4x replicated
struct GfStage {
  int TWt, OW;
  int off_O, off_COEF, off_ST, off_ICY, off_INVNY, off_CEN, bytes;
  __host__ __device__ constexpr explicit GfStage(int twt)
      : TWt(twt), OW(twt + 64), off_O(3 * twt), off_COEF(3 * twt + 3 * (twt + 64)),
        off_ST(3 * twt + 3 * (twt + 64) + 8 * twt), off_ICY(3 * twt + 3 * (twt + 64) + 48 * twt),
        off_INVNY(3 * twt + 3 * (twt + 64) + 52 * twt), off_CEN(3 * twt + 3 * (twt + 64) + 56 * twt),
        bytes(3 * twt + 3 * (twt + 64) + 57 * twt) {}
};

__host__ __device__ inline size_t gf_smem_bytes(int runs, int K, int HL4, int LPR) {
  return 256 + 2 * (size_t)GfStage(runs * K).bytes + 6 * (size_t)LPR * exch_pitch_words(runs, K, HL4) * sizeof(u32);
}

// K bytes at byte offset `off` (any alignment) of a shared-memory row
template <int K>
__device__ __forceinline__ void lds_unaligned(const u8* row, int off, u32 (&w)[K / 4]) {
  const u32* pa = reinterpret_cast<const u32*>(row + (off & ~3));
  const u32 sel = 0x3210u + 0x1111u * (u32)(off & 3);
  u32 t[K / 4 + 1];
#pragma unroll
  for (int i = 0; i <= K / 4; ++i) t[i] = pa[i];
#pragma unroll
  for (int i = 0; i < K / 4; ++i) w[i] = __byte_perm(t[i], t[i + 1], sel);
}

// Stage-2 horizontal pass.  winA / winB: [left halo HL4 | own K | right halo HL4].  The A window sum is kept as
// three partial sums (columns owned by the left neighbour, by this run, by the right neighbour) because the
// neighbours' B' sums are relative to THEIR centres: B'(x) = sum(winB) + dl * A_L(x) + dr * A_R(x).
// Two independent chains (columns 0..7 and 8..15), see slide_i32.
template <int R, int K, int HL4>
__device__ __forceinline__ void slide_ab(const u32 (&winA)[HL4 + K + HL4], const u32 (&winB)[HL4 + K + HL4], float dl,
                                         float dr, float (&A)[K], float (&B)[K]) {
  static_assert(R < K && K == 16, "window must not reach beyond the adjacent runs");
  constexpr int H = K / 2;
  // chain starting at column c0: window [c0-R, c0+R]; parts: left halo (< 0), own [0, K), right halo (>= K)
  float aL[2], aO[2], aR[2], b[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c0 = h * H;
    float l = 0.f, o = 0.f, r = 0.f;
#pragma unroll
    for (int j = -R; j <= R; ++j) {
      const int idx = c0 + j;
      const float va = __uint_as_float(winA[HL4 + idx]);
      if (idx < 0) l += va; else if (idx < K) o += va; else r += va;
    }
    aL[h] = l; aO[h] = o; aR[h] = r;
  }
  if constexpr (2 * R + 1 > H) {
    const float mid = wsum_f<H - R, R>(winB, HL4);
    b[0] = mid + wsum_f<-R, H - R - 1>(winB, HL4);
    b[1] = mid + wsum_f<R + 1, H + R>(winB, HL4);
  } else {
    b[0] = wsum_f<-R, R>(winB, HL4);
    b[1] = wsum_f<H - R, H + R>(winB, HL4);
  }
#pragma unroll
  for (int c = 0; c < H; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = h * H + c;
      if (c > 0) {
        const int in = col + R, out = col - R - 1;
        const float vin = __uint_as_float(winA[HL4 + in]), vout = __uint_as_float(winA[HL4 + out]);
        if (in < K) aO[h] += vin; else aR[h] += vin;
        if (out < 0) aL[h] -= vout; else aO[h] -= vout;
        b[h] = (b[h] + __uint_as_float(winB[HL4 + in])) - __uint_as_float(winB[HL4 + out]);
      }
      A[col] = (aL[h] + aO[h]) + aR[h];
      B[col] = fmaf(dr, aR[h], fmaf(dl, aL[h], b[h]));
    }
  }
}

template <int K>
__device__ __forceinline__ void ad_row_s(const u8* grow, const u8* orow, int ooff, u32 (&p)[K / 4]) {
  u32 ow[K / 4];
  lds_unaligned<K>(orow, ooff, ow);
#pragma unroll
  for (int w = 0; w < K / 4; w += 4) {
    const uint4 gq = *reinterpret_cast<const uint4*>(grow + 4 * w);
    p[w] = __vabsdiffu4(gq.x, ow[w]);
    p[w + 1] = __vabsdiffu4(gq.y, ow[w + 1]);
    p[w + 2] = __vabsdiffu4(gq.z, ow[w + 2]);
    p[w + 3] = __vabsdiffu4(gq.w, ow[w + 3]);
  }
}

// One (a, b) row: slide the two exact stage-1 sums across the thread's K columns and fold
// SIGN * (a, b) into the stage-2 vertical running sums.  st = staged statistics of that row for this thread.
template <int R, int K, int HL4, int SIGN>
__device__ __forceinline__ void fold_ab(const u32* xbP, const u32* xbI, const int (&Vp)[K], const int (&VIp)[K],
                                        const float* st, int TWt, float cc, float (&VA)[K], float (&VB)[K],
                                        float (&VAf)[K], float (&VBf)[K]) {
  int Sp[K], SIp[K];
  {
    u32 win[HL4 + K + HL4];
    exch_window<K, HL4>(xbP, reinterpret_cast<const u32(&)[K]>(Vp), win);
    slide_i32<R, K, HL4>(win, Sp);
  }
  {
    u32 win[HL4 + K + HL4];
    exch_window<K, HL4>(xbI, reinterpret_cast<const u32(&)[K]>(VIp), win);
    slide_i32<R, K, HL4>(win, SIp);
  }
#pragma unroll
  for (int g4 = 0; g4 < K; g4 += 4) {
    const int4 N = *reinterpret_cast<const int4*>(st + ST_N * TWt + g4);
    const int4 SI = *reinterpret_cast<const int4*>(st + ST_SI * TWt + g4);
    const float4 invden = *reinterpret_cast<const float4*>(st + ST_INVDEN * TWt + g4);
    const float4 cmean = *reinterpret_cast<const float4*>(st + ST_CMEAN * TWt + g4);
    const float4 invn = *reinterpret_cast<const float4*>(st + ST_INVN * TWt + g4);
    const int Nn[4] = {N.x, N.y, N.z, N.w}, SIi[4] = {SI.x, SI.y, SI.z, SI.w};
    const float idn[4] = {invden.x, invden.y, invden.z, invden.w}, cm[4] = {cmean.x, cmean.y, cmean.z, cmean.w},
                inn[4] = {invn.x, invn.y, invn.z, invn.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = g4 + j;
      const int num = Nn[j] * SIp[c] - SIi[j] * Sp[c];  // exact modulo 2^32; true value fits int32 for r <= 9
      const float a = (float)num * idn[j];
      const float b = fmaf(-a, cm[j] - cc, (float)Sp[c] * inn[j]);  // mean_p - a * (mean_I - centre)
      if (SIGN > 0) {
        VA[c] += a; VB[c] += b;
        if (GSM_GF_FRESH) { VAf[c] += a; VBf[c] += b; }
      } else {
        VA[c] -= a; VB[c] -= b;
      }
    }
  }
}

template <int R, int K, int RUNS, int LPR, bool EXPORT>
__global__ void __launch_bounds__(RUNS * LPR, (RUNS * LPR <= 192) ? 2 : 1)
gf_wta_kernel(const u8* __restrict__ Gp, const u8* __restrict__ Op, const float* __restrict__ stats,
              i64* __restrict__ keys, FusedGeom g) {
  constexpr int HL4 = (R + 3) / 4 * 4;
  constexpr int KW = K / 4;
  extern __shared__ __align__(128) u8 smem_raw[];

  // LPR lanes (= disparities) per run: 32 -> one run per warp; 16 -> two runs per warp, i.e. a strip twice as wide
  // for the same thread count (less halo overhead) at the price of staging the rows for half as many disparities.
  static_assert(LPR == 32 || LPR == 16, "lanes per run");
  const int lane = threadIdx.x & (LPR - 1);
  const int run = threadIdx.y * (WARP / LPR) + threadIdx.x / LPR;
  constexpr int runs = RUNS;  // blockDim = (32, RUNS*LPR/32): all shared-memory offsets are immediates
  const int strip = blockIdx.x;
  const int d0 = g.d_begin + blockIdx.y * LPR;
  const int d = d0 + lane;
  const int frame = blockIdx.z / g.bands;
  const int band = blockIdx.z - frame * g.bands;
  const int H = g.pg.H, W = g.pg.W, pitch = g.pg.pitch;
  const int yb0 = band * g.band_rows;
  const int yb1 = min(H, yb0 + g.band_rows);
  if (yb0 >= H) return;

  constexpr int TWt = runs * K;
  constexpr GfStage sg(TWt);
  constexpr int pitchw = exch_pitch_words(runs, K, HL4);
  constexpr int planew = LPR * pitchw;
  float* ccs = reinterpret_cast<float*>(smem_raw + 64);  // per-run centres of the current step (<= 48 runs)
  u8* stage_base = smem_raw + 256;
  u32* exch = reinterpret_cast<u32*>(stage_base + 2 * sg.bytes);
  const u32 bar0 = smem_u32(smem_raw);  // two 8-byte mbarriers at the start of shared memory
  const bool producer = (threadIdx.x == 0 && threadIdx.y == 0);

  for (int i = threadIdx.y * WARP + threadIdx.x; i < 6 * planew; i += runs * LPR) exch[i] = 0u;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  u32* xb = exch + (size_t)lane * pitchw + HL4 + run * K;  // this thread's slot in exchange plane 0

  const int xs = strip * g.TW - g.hl;  // image column of the strip's first (halo) column
  const int x0 = xs + run * K;
  const size_t plane_elems = g.pg.plane_stride;
  const int row_lo = -PADV, row_hi = H + PADV - 1;

  // producer-side source addresses (row 0 of the padded planes, strip column xs)
  const size_t org = (size_t)PADV * pitch + g.pg.xoff + xs;
  const u8* gsrc = Gp + (size_t)frame * g.pg.plane_stride + org;
  const int ostart = g.pg.xoff + xs + (g.view == 0 ? -(d0 + LPR - 1) : d0);  // byte column of the staged O window
  const int oalign = ostart & 15;
  const u8* osrc = Op + (size_t)frame * g.pg.plane_stride + (size_t)PADV * pitch + (ostart - oalign);
  const float* ssrc = stats + (size_t)frame * GF_STAT_PLANES * plane_elems + org;
  const float* csrc = stats + ((size_t)frame * GF_STAT_PLANES + ST_CEN) * plane_elems + (size_t)PADV * pitch +
                      (g.pg.xoff + xs) / 4;  // 4x replicated centre plane: one float per 4 columns
  // consumer-side byte offset of this thread's first pixel inside a staged O row
  const int ooff = oalign + run * K + (g.view == 0 ? (LPR - 1 - lane) : lane);

  // One elected thread issues the 20 bulk copies of a step (uniform-datapath address arithmetic; spreading them
  // over the warps was measured: same speed, +6 instructions per DE of per-lane address arithmetic).
  auto issue = [&](int t, int s) {
    const u32 bar = bar0 + 8 * s;
    const u32 dst = smem_u32(stage_base + (size_t)s * sg.bytes);
    mbar_expect_tx(bar, (u32)sg.bytes);
    const int rows3[3] = {t + R, t - R - 1, t - 3 * R - 2};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const long long ro = (long long)max(row_lo, min(row_hi, rows3[i])) * pitch;
      bulk_g2s(dst + i * TWt, gsrc + ro, TWt, bar);
      bulk_g2s(dst + sg.off_O + i * sg.OW, osrc + ro, sg.OW, bar);
    }
    const int rows2[2] = {t, t - 2 * R - 1};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long ro = (long long)max(row_lo, min(row_hi, rows2[i])) * pitch;
      bulk_g2s(dst + sg.off_COEF + i * 4 * TWt, ssrc + ST_COEF * plane_elems + ro, 4 * TWt, bar);
#pragma unroll
      for (int k = 0; k < 5; ++k)
        bulk_g2s(dst + sg.off_ST + (i * 5 + k) * 4 * TWt, ssrc + (size_t)k * plane_elems + ro, 4 * TWt, bar);
    }
    const long long ry = (long long)max(row_lo, min(row_hi, t - R)) * pitch;
    bulk_g2s(dst + sg.off_ICY, ssrc + ST_IC * plane_elems + ry, 4 * TWt, bar);
    bulk_g2s(dst + sg.off_INVNY, ssrc + ST_INVN * plane_elems + ry, 4 * TWt, bar);
    bulk_g2s(dst + sg.off_CEN, csrc + ry, TWt, bar);
  };

  const int dd = min(d, MAX_DISP - 1);
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
  // Runs that only supply stage-1 halo columns skip the later stages (warp-uniform): (a, b) is needed on strip
  // columns [hl-R, hl+TW+R) inside the image (+R), outputs on [hl, hl+TW) inside the image.
  // Control flow must stay warp-uniform (__syncthreads inside the loop), so with two runs per warp the warp runs a
  // stage when either of its runs needs it.
  const bool need_out =
      __any_sync(0xffffffffu, min(K - 1, min(out0 + g.TW, W) - 1 - x0) >= c_lo);
  const bool need_ab = __any_sync(
      0xffffffffu, (run * K < g.hl + g.TW + R) && (run * K + K > g.hl - R) && (x0 < W + R) && (x0 + K > -R));

  int Vp_l[K], VIp_l[K], Vp_t[K], VIp_t[K];
  // Stage-2 vertical sums.  VA/VB are add/subtract running sums; VAf/VBf only ever add and are swapped in every
  // 2R+1 rows, when they hold exactly the current window: rounding drift is bounded by 2(2R+1) additions
  // instead of growing with the image height.
  float VA[K], VB[K], VAf[K], VBf[K];
#pragma unroll
  for (int c = 0; c < K; ++c) {
    Vp_l[c] = VIp_l[c] = Vp_t[c] = VIp_t[c] = 0;
    VA[c] = VB[c] = VAf[c] = VBf[c] = 0.f;
  }
  int fresh_cnt = 0;
  float cc = 0.f;  // current centre of this run, relative to 128 (warp-uniform)

  const int r0 = yb0 - 2 * R;  // first image row whose AD may enter a stage-1 window of this band
  const int a0 = yb0 - R;      // first row whose (a, b) may enter a stage-2 window of this band
  constexpr int COEF_PM = (int)0xFFFF0001;  // lo16 = +1, hi16 = -1
  const int t_begin = yb0 - 3 * R, t_end = yb1 + R;

  if (producer) issue(t_begin, 0);

  for (int t = t_begin; t < t_end; ++t) {
    const int it = t - t_begin;
    const int s = it & 1;
    mbar_wait(bar0 + 8 * s, (u32)((it >> 1) & 1));
    const u8* stg = stage_base + (size_t)s * sg.bytes;
    const int t2 = t - 2 * R - 1;  // row of (a, b) recomputed by the trail pipeline

    // ---------------- stage 1, vertical: rows t+R (enters lead), t-R-1 (lead -> trail), t-3R-2 (leaves trail)
    {
      u32 pn[KW], pm[KW], po[KW];
      ad_row_s<K>(stg + run * K, stg + sg.off_O, ooff, pn);
      if (t - R - 1 >= r0) ad_row_s<K>(stg + TWt + run * K, stg + sg.off_O + sg.OW, ooff, pm);
      else {
#pragma unroll
        for (int w = 0; w < KW; ++w) pm[w] = 0u;
      }
      if (t - 3 * R - 2 >= r0) ad_row_s<K>(stg + 2 * TWt + run * K, stg + sg.off_O + 2 * sg.OW, ooff, po);
      else {
#pragma unroll
        for (int w = 0; w < KW; ++w) po[w] = 0u;
      }
      if (need_mask) {
#pragma unroll
        for (int w = 0; w < KW; ++w) { pn[w] &= mask[w]; pm[w] &= mask[w]; po[w] &= mask[w]; }
      }
      const int* cA = reinterpret_cast<const int*>(stg + sg.off_COEF) + run * K;
      const int* cB = cA + TWt;
#pragma unroll
      for (int g4 = 0; g4 < K; g4 += 4) {
        const int4 a4 = *reinterpret_cast<const int4*>(cA + g4);
        const int4 b4 = *reinterpret_cast<const int4*>(cB + g4);
        const int ca[4] = {a4.x, a4.y, a4.z, a4.w}, cb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g4 + j;
          const u32 sel = j | ((4 + j) << 4);
          const u32 nm = __byte_perm(pn[c / 4], pm[c / 4], sel);  // {p_enter, p_leave, x, x}
          const u32 mo = __byte_perm(pm[c / 4], po[c / 4], sel);
          Vp_l[c] = dp2a_lo_su(COEF_PM, nm, Vp_l[c]);
          VIp_l[c] = dp2a_lo_su(ca[j], nm, VIp_l[c]);
          Vp_t[c] = dp2a_lo_su(COEF_PM, mo, Vp_t[c]);
          VIp_t[c] = dp2a_lo_su(cb[j], mo, VIp_t[c]);
        }
      }
    }
    // ---------------- stage 1, horizontal
    exch_store<K, HL4>(xb + 0 * planew, reinterpret_cast<u32(&)[K]>(Vp_l));
    exch_store<K, HL4>(xb + 1 * planew, reinterpret_cast<u32(&)[K]>(VIp_l));
    exch_store<K, HL4>(xb + 2 * planew, reinterpret_cast<u32(&)[K]>(Vp_t));
    exch_store<K, HL4>(xb + 3 * planew, reinterpret_cast<u32(&)[K]>(VIp_t));
    __syncthreads();
    // Prefetch the next step's rows into the other stage.  Its previous contents (step t-1) are read up to the end
    // of that step (I-128, 1/N of the output row), so the earliest safe point is after this barrier, which every
    // thread reaches only once it has finished step t-1.
    if (producer && t + 1 < t_end) issue(t + 1, s ^ 1);
    const float* st_l = reinterpret_cast<const float*>(stg + sg.off_ST) + run * K;
    {
      // follow the local intensity level: B'(c + dc) = B'(c) + dc * A exactly, applied to the running sums
      const float target = reinterpret_cast<const float*>(stg + sg.off_CEN)[run * (K / 4)];
      const float dc = target - cc;
      if (fabsf(dc) > GF_RECENTRE) {
#pragma unroll
        for (int c = 0; c < K; ++c) {
          VB[c] = fmaf(dc, VA[c], VB[c]);
          if (GSM_GF_FRESH) VBf[c] = fmaf(dc, VAf[c], VBf[c]);
        }
        cc = target;
      }
    }
    if (need_ab && t >= a0)
      fold_ab<R, K, HL4, +1>(xb + 0 * planew, xb + 1 * planew, Vp_l, VIp_l, st_l, TWt, cc, VA, VB, VAf, VBf);
    if (need_ab && t2 >= a0)
      fold_ab<R, K, HL4, -1>(xb + 2 * planew, xb + 3 * planew, Vp_t, VIp_t, st_l + 5 * TWt, TWt, cc, VA, VB, VAf, VBf);
    if (GSM_GF_FRESH && t >= a0 && ++fresh_cnt == 2 * R + 1) {
      fresh_cnt = 0;
#pragma unroll
      for (int c = 0; c < K; ++c) { VA[c] = VAf[c]; VB[c] = VBf[c]; VAf[c] = 0.f; VBf[c] = 0.f; }
    }

    const int y = t - R;  // output row
    if (need_ab && y >= yb0) {
      exch_store<K, HL4>(xb + 4 * planew, reinterpret_cast<u32(&)[K]>(VA));
      exch_store<K, HL4>(xb + 5 * planew, reinterpret_cast<u32(&)[K]>(VB));
      if (lane == 0) ccs[run] = cc;
    }
    __syncthreads();
    if (y < yb0 || !need_out) continue;

    // ---------------- stage 2, horizontal + q + WTA
    float A[K], B[K];
    {
      u32 winA[HL4 + K + HL4], winB[HL4 + K + HL4];
      exch_window<K, HL4>(xb + 4 * planew, reinterpret_cast<u32(&)[K]>(VA), winA);
      exch_window<K, HL4>(xb + 5 * planew, reinterpret_cast<u32(&)[K]>(VB), winB);
      const float dl = run > 0 ? cc - ccs[run - 1] : 0.f;          // outer halos of the strip are zero pad
      const float dr = run + 1 < runs ? cc - ccs[run + 1] : 0.f;
      slide_ab<R, K, HL4>(winA, winB, dl, dr, A, B);
    }
    const float* icy = reinterpret_cast<const float*>(stg + sg.off_ICY) + run * K;
    const float* iny = reinterpret_cast<const float*>(stg + sg.off_INVNY) + run * K;
    int key[K];  // signed order == float order
#pragma unroll
    for (int g4 = 0; g4 < K; g4 += 4) {
      const float4 ic = *reinterpret_cast<const float4*>(icy + g4);
      const float4 in = *reinterpret_cast<const float4*>(iny + g4);
      key[g4 + 0] = sortable_i32(fmaf(A[g4 + 0], ic.x - cc, B[g4 + 0]) * in.x);
      key[g4 + 1] = sortable_i32(fmaf(A[g4 + 1], ic.y - cc, B[g4 + 1]) * in.y);
      key[g4 + 2] = sortable_i32(fmaf(A[g4 + 2], ic.z - cc, B[g4 + 2]) * in.z);
      key[g4 + 3] = sortable_i32(fmaf(A[g4 + 3], ic.w - cc, B[g4 + 3]) * in.w);
    }
    if constexpr (EXPORT) {
      const int de = d - g.export_d0;
      if (de >= 0 && de < g.export_nd && d < g.d_end) {
        float* out = reinterpret_cast<float*>(g.export_ptr) + ((size_t)de * H + y) * W;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const int x = x0 + c;
          if (x >= out0 && x < min(out0 + g.TW, W)) out[x] = unsortable_f32(key[c]);
        }
      }
    }
    // WTA over the warp's 32 disparities: the lane index rides in the 5 low bits of the sortable key, so one
    // REDUX gives both the minimum and (lowest-d-first) its owner; costs closer than 2^-18 relative count as ties.
#pragma unroll
    for (int c = 0; c < K; ++c) key[c] = (key[c] & ~31) | lane;
    if (!all_valid) {
#pragma unroll
      for (int c = 0; c < K; ++c)
        if (c < c_lo || c > c_hi) key[c] = 0x7fffffff;
    }
    int mine = 0x7fffffff;
    if (LPR == 32) {
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int m = __reduce_min_sync(0xffffffffu, key[c]);
        if (lane == c) mine = m;
      }
    } else {
      // two runs per warp: two full-warp REDUX per column, each half contributing the neutral element to the other's
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
      atomicMin(keys + ((size_t)frame * H + y) * W + x0 + lane, k64);
    }
  }
}

}  // namespace gsm
