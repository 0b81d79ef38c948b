// gsm_gf.cuh -- guided-filter mode (GSM_MODE_GF): guide-statistics pre-pass and the device helpers shared by the
// fused kernel in gsm_gf3.cuh.  The reference has no guided filter; the arithmetic follows "GF-v1"
// (oracle/stereo_oracle.c gf_slice) and the WTA follows STMatching/StereoHelper.cpp:131-154.
//
// Per disparity d, guide I, p = AD_d, clipped (2r+1)^2 window sums S_x = box(x), N = window pixel count:
//     a = (N*S_Ip - S_I*S_p) / (N*S_II - S_I^2 + eps*N^2)      b = (S_p - a*S_I) / N
//     q = (box(a)*I + box(b)) / N
// Numerics of stage 2 (fp32): b is formed against a LOCAL centre, b' = mean_p - a*(mean_I - c_run), c_run = rounded
// local mean of the guide over the thread's 16-column run, re-centred (V_B += dc*V_A, an exact identity) when it
// drifts by more than GF_RECENTRE grey levels.  q = (A*(I - c_run) + B')/N does not depend on c in exact arithmetic,
// but with c near the local intensities the two terms no longer cancel, which is what fp32 needs to stay within 1e-4
// of the float64 oracle.  Halo columns owned by neighbour runs (different centres) are converted through the partial
// sums A_L, A_R of the window: B' += (c_own - c_nbr) * A_{L|R}  (slide_ab below).
#pragma once
#include "gsm_common.cuh"
#include "gsm_sad.cuh"  // exch_pitch_words

namespace gsm {

// guide statistic planes (float/int32, same padded geometry as the u8 planes; zero outside the image)
constexpr int GF_STAT_PLANES = 8;
enum { ST_N = 0, ST_SI = 1, ST_INVDEN = 2, ST_CMEAN = 3, ST_INVN = 4, ST_IC = 5, ST_COEF = 6, ST_CEN = 7 };
constexpr float GF_CENTRE = 128.0f;
constexpr float GF_RECENTRE = 8.0f;

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

// Local centre plane: for every image row, every strip of the launch and every run of K columns of that strip, the
// rounded mean of (mean_I - 128) over the run's in-image columns.  Layout [row][strip * CENW + run] inside a plane of
// the usual row pitch (CENW = runs rounded up to 4), so a strip's centres are one 16-byte aligned bulk copy.
__global__ void gf_centre_kernel(float* __restrict__ stats, PlaneGeom pg, int TW, int hl, int K, int runs, int strips) {
  const int cenw = (runs + 3) / 4 * 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // strip * cenw + run
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (i >= strips * cenw) return;
  const int strip = i / cenw, run = i - strip * cenw;
  float* base = stats + (size_t)f * GF_STAT_PLANES * pg.plane_stride;
  float c = 0.f;
  if (run < runs) {
    const int xa = strip * TW - hl + run * K;
    const float* cm = base + ST_CMEAN * pg.plane_stride + (size_t)(PADV + y) * pg.pitch + pg.xoff;
    float sum = 0.f;
    int n = 0;
    for (int j = 0; j < K; ++j) {
      const int x = xa + j;
      if (x >= 0 && x < pg.W) { sum += cm[x]; ++n; }
    }
    c = n ? rintf(sum / (float)n) : 0.f;
  }
  base[ST_CEN * pg.plane_stride + (size_t)(PADV + y) * pg.pitch + i] = c;
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
__device__ __forceinline__ float wsum_f(const u32 (&win)[N], int base) {
  float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
  for (int j = LO; j <= HI; ++j) {
    const float v = __uint_as_float(win[base + j]);
    if ((j - LO) % 3 == 0) p0 += v; else if ((j - LO) % 3 == 1) p1 += v; else p2 += v;
  }
  return (p0 + p1) + p2;
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
template <int R, int K, int HL4>
__device__ __forceinline__ void slide_ab(const u32 (&winA)[HL4 + K + HL4], const u32 (&winB)[HL4 + K + HL4], float dl,
                                         float dr, float (&A)[K], float (&B)[K]) {
  static_assert(R < K, "window must not reach beyond the adjacent runs");
  // One sliding chain over the K columns.  a: whole window sum of V_A; aL / aR: the part of it owned by the left /
  // right neighbour run (aL only shrinks, aR only grows as the window moves right); b: window sum of V_B.
  float aL = 0.f, aO = 0.f, b = wsum_f<-R, R>(winB, HL4);
#pragma unroll
  for (int j = -R; j <= R; ++j) {
    const float va = __uint_as_float(winA[HL4 + j]);
    if (j < 0) aL += va; else aO += va;
  }
  float a = aL + aO, aR = 0.f;
  A[0] = a;
  B[0] = fmaf(dl, aL, b);
#pragma unroll
  for (int c = 1; c < K; ++c) {
    const int in = c + R, out = c - R - 1;
    const float vin = __uint_as_float(winA[HL4 + in]), vout = __uint_as_float(winA[HL4 + out]);
    a += vin - vout;  // difference first: one rounding at the magnitude of the window sum instead of two
    if (in >= K) aR += vin;
    if (out < 0) aL -= vout;
    b += __uint_as_float(winB[HL4 + in]) - __uint_as_float(winB[HL4 + out]);
    A[c] = a;
    B[c] = fmaf(dr, aR, fmaf(dl, aL, b));
  }
}

}  // namespace gsm
