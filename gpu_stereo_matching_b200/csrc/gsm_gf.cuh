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

// ---- fused guide pre-pass -----------------------------------------------------------------------------------
// One kernel writes every disparity-independent plane of a frame: N, S_I, 1/(N*S_II - S_I^2 + eps*N^2),
// mean_I - 128, 1/N, I - 128 (image pixels), the horizontal-slide coefficient word of gsm_gf3.cuh (image columns
// and R + 1 margin columns) and the per-run local centres; it also initialises the frame's packed-min plane.
// A block owns PP_TX columns (PP_HALO more on each side feed the window sums) and marches down PP_ROWS rows:
// thread = 4 adjacent columns (one 32-bit load per image row, one 128-bit store per plane and row), vertical running
// sums of I and I^2 in registers, horizontal window sums from per-warp prefix sums (in-thread prefix + one shuffle scan
// per 4 columns + shared memory), one barrier per row.  The 16-column runs of all strips lie on one global grid (TW
// is a multiple of 16, every strip starts hl columns left of a multiple of TW), so a block's columns start on a run
// boundary and a run's centre is written to the (strip, run) slots of the one or two strips that contain it.
constexpr int PP_C = 4, PP_THREADS = 96, PP_COLS = PP_THREADS * PP_C, PP_HALO = 16, PP_TX = PP_COLS - 2 * PP_HALO,
              PP_ROWS = 32;  // rows per block of a large launch; small launches use fewer (more blocks)
static_assert(PP_TX % 16 == 0 && PP_HALO % PP_C == 0, "tiles start on run boundaries");
__global__ void __launch_bounds__(PP_THREADS)
gf_prepass_kernel(const u8* __restrict__ Ip, float* __restrict__ stats, PlaneGeom pg, int R, float eps, int TW, int hl,
                  int runs, int strips, i64* __restrict__ keys, i64 key_init, int rows_per_block,
                  const FrameDesc* __restrict__ ft) {
  __shared__ __align__(16) int PW1[2][PP_COLS], PW2[2][PP_COLS], PIX[2][PP_COLS];
  __shared__ __align__(16) float CM[2][PP_COLS];
  const int tid = threadIdx.x, lane = tid & 31;
  const int f = blockIdx.z;
  const int c0 = PP_C * tid;                                                   // first column of the thread in the tile
  const int x = -hl - PP_HALO + (int)blockIdx.x * PP_TX + c0 - PP_HALO;        // its image column
  int H = pg.H, W = pg.W;
  size_t koff = (size_t)f * H * W;
  if (ft) { const FrameDesc fd = ft[f]; H = fd.H; W = fd.W; koff = (size_t)fd.off; }
  const int yb = blockIdx.y * rows_per_block, ye = min(H, yb + rows_per_block);
  if (yb >= H || x - c0 >= W + R + 1 + PP_HALO) return;  // (mixed-size batches) block outside this frame
  const size_t plane_elems = pg.plane_stride;
  const u8* col = Ip + (size_t)f * pg.plane_stride + (size_t)PADV * pg.pitch + pg.xoff + x;  // 4-byte aligned
  float* base = stats + (size_t)f * GF_STAT_PLANES * plane_elems;
  const bool outp = c0 >= PP_HALO && c0 < PP_HALO + PP_TX;
  const bool anyimg = outp && x + PP_C > 0 && x < W;
  const int cenw = (runs + 3) / 4 * 4, rps = TW / 16;
  const bool leader = outp && (c0 & 15) == 0;
  const int gb = (x + hl) >> 4;  // global run index of a leader's run (x + hl is a multiple of 16 there)
  int nx[PP_C];
#pragma unroll
  for (int k = 0; k < PP_C; ++k) nx[k] = min(W - 1, x + k + R) - max(0, x + k - R) + 1;

  auto write_centre = [&](int y) {  // centre of row y of the leader's run, from the row's CM buffer
    if (!leader || gb < 0) return;
    const float* cm = CM[y & 1] + c0;
    float sum = 0.f;
    int n = 0;
    for (int j = 0; j < 16; ++j)
      if (x + j >= 0 && x + j < W) { sum += cm[j]; ++n; }
    const float c = n ? rintf(sum / (float)n) : 0.f;
    float* cen = base + ST_CEN * plane_elems + (size_t)(PADV + y) * pg.pitch;
    const int s0 = gb / rps;
    for (int s = s0; s >= s0 - 1; --s) {
      const int run = gb - s * rps;
      if (s >= 0 && s < strips && run < runs) cen[s * cenw + run] = c;
    }
  };
  auto row4 = [&](int yy) { return *reinterpret_cast<const u32*>(col + (long long)yy * pg.pitch); };

  int v1[PP_C], v2[PP_C];
#pragma unroll
  for (int k = 0; k < PP_C; ++k) v1[k] = v2[k] = 0;
  for (int yy = yb - R; yy < yb + R; ++yy) {
    const u32 p4 = row4(yy);
#pragma unroll
    for (int k = 0; k < PP_C; ++k) {
      const int p = (p4 >> (8 * k)) & 0xff;
      v1[k] += p;
      v2[k] += p * p;
    }
  }
  // the three pixel groups of a row (entering, centre, leaving) are fetched one row ahead
  u32 n_in = row4(yb + R), n_pc = row4(yb), n_out = row4(yb - R);
  for (int y = yb; y < ye; ++y) {
    const int b = y & 1;
    const u32 pin = n_in, pc = n_pc, pout = n_out;
    n_in = row4(y + 1 + R);  // rows up to H + R: inside the bottom pad
    n_pc = row4(y + 1);
    n_out = row4(y + 1 - R);
    int a1[PP_C], a2[PP_C];  // in-thread inclusive prefix of the column sums over rows y-R .. y+R
#pragma unroll
    for (int k = 0; k < PP_C; ++k) {
      const int p = (pin >> (8 * k)) & 0xff;
      v1[k] += p;
      v2[k] += p * p;
      a1[k] = v1[k] + (k ? a1[k - 1] : 0);
      a2[k] = v2[k] + (k ? a2[k - 1] : 0);
    }
    int s1 = a1[PP_C - 1], s2 = a2[PP_C - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t1 = __shfl_up_sync(0xffffffffu, s1, o), t2 = __shfl_up_sync(0xffffffffu, s2, o);
      if (lane >= o) { s1 += t1; s2 += t2; }
    }
    const int o1 = s1 - a1[PP_C - 1], o2 = s2 - a2[PP_C - 1];  // exclusive warp offsets
    *reinterpret_cast<int4*>(&PW1[b][c0]) = make_int4(o1 + a1[0], o1 + a1[1], o1 + a1[2], o1 + a1[3]);
    *reinterpret_cast<int4*>(&PW2[b][c0]) = make_int4(o2 + a2[0], o2 + a2[1], o2 + a2[2], o2 + a2[3]);
    *reinterpret_cast<int4*>(&PIX[b][c0]) =
        make_int4((int)(pc & 0xff), (int)((pc >> 8) & 0xff), (int)((pc >> 16) & 0xff), (int)(pc >> 24));
    __syncthreads();
    if (y > yb) write_centre(y - 1);
    if (outp) {
      const int ny = min(H - 1, y + R) - max(0, y - R) + 1;
      int Nn[PP_C], SIv[PP_C], coef[PP_C];
      float invden[PP_C], cmv[PP_C], invn[PP_C], ic[PP_C];
#pragma unroll
      for (int k = 0; k < PP_C; ++k) {
        const int ci = c0 + k, lo = ci - R - 1, hi = ci + R;  // window = (lo, hi]; it spans at most two warps
        int SI = PW1[b][hi] - PW1[b][lo], SII = PW2[b][hi] - PW2[b][lo];
        if ((lo >> 7) != (hi >> 7)) { SI += PW1[b][lo | 127]; SII += PW2[b][lo | 127]; }
        coef[k] = PIX[b][ci + R] - 65536 * PIX[b][ci - R - 1];
        const bool in = x + k >= 0 && x + k < W;
        const int N = nx[k] * ny;
        // N^2 var and the mean offset are formed exactly in integers and rounded once
        const long long den = (long long)N * SII - (long long)SI * SI;
        const float dden = (float)den + eps * (float)(N * N);
        Nn[k] = in ? N : 0;
        SIv[k] = in ? SI : 0;
        invden[k] = in ? 1.0f / dden : 0.f;
        cmv[k] = in ? (float)(SI - (int)GF_CENTRE * N) / (float)N : 0.f;
        invn[k] = in ? 1.0f / (float)N : 0.f;
        ic[k] = in ? (float)((pc >> (8 * k)) & 0xff) - GF_CENTRE : 0.f;
        if (in) keys[koff + (size_t)y * W + x + k] = key_init;  // packed-min plane: +inf cost, d = 0
      }
      const size_t o = (size_t)(PADV + y) * pg.pitch + pg.xoff + x;  // multiple of 4: 16-byte aligned stores
      *reinterpret_cast<int4*>(base + ST_COEF * plane_elems + o) = make_int4(coef[0], coef[1], coef[2], coef[3]);
      if (anyimg) {
        *reinterpret_cast<int4*>(base + ST_N * plane_elems + o) = make_int4(Nn[0], Nn[1], Nn[2], Nn[3]);
        *reinterpret_cast<int4*>(base + ST_SI * plane_elems + o) = make_int4(SIv[0], SIv[1], SIv[2], SIv[3]);
        *reinterpret_cast<float4*>(base + ST_INVDEN * plane_elems + o) = make_float4(invden[0], invden[1], invden[2], invden[3]);
        *reinterpret_cast<float4*>(base + ST_CMEAN * plane_elems + o) = make_float4(cmv[0], cmv[1], cmv[2], cmv[3]);
        *reinterpret_cast<float4*>(base + ST_INVN * plane_elems + o) = make_float4(invn[0], invn[1], invn[2], invn[3]);
        *reinterpret_cast<float4*>(base + ST_IC * plane_elems + o) = make_float4(ic[0], ic[1], ic[2], ic[3]);
      }
      *reinterpret_cast<float4*>(&CM[b][c0]) = make_float4(cmv[0], cmv[1], cmv[2], cmv[3]);
    }
#pragma unroll
    for (int k = 0; k < PP_C; ++k) {
      const int p = (pout >> (8 * k)) & 0xff;
      v1[k] -= p;
      v2[k] -= p * p;
    }
  }
  __syncthreads();
  if (ye > yb) write_centre(ye - 1);
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

// Stage-2 horizontal pass.  winA / winB: [left halo HL4 | own K | right halo HL4].  The A window sum is kept as
// three partial sums (columns owned by the left neighbour, by this run, by the right neighbour) because the
// neighbours' B' sums are relative to THEIR centres: B'(x) = sum(winB) + dl * A_L(x) + dr * A_R(x).
template <int R, int K, int HL4>
__device__ __forceinline__ void slide_ab(const u32 (&winA)[HL4 + K + HL4], const u32 (&winB)[HL4 + K + HL4], float dl,
                                         float dr, float (&A)[K], float (&B)[K]) {
  static_assert(R < K, "window must not reach beyond the adjacent runs");
  // One sliding chain over the K columns.  a: whole window sum of V_A; aL / aR: the part of it owned by the left /
  // right neighbour run (aL only shrinks, aR only grows as the window moves right); b: window sum of V_B.
  float aL = 0.f, aO = 0.f;
  // the run's own columns first (registers), the neighbours' halo second (shared-memory loads still in flight); the two
  // partial sums are separate accumulators, so the order changes nothing numerically
#pragma unroll
  for (int j = 0; j <= R; ++j) aO += __uint_as_float(winA[HL4 + j]);
#pragma unroll
  for (int j = -R; j < 0; ++j) aL += __uint_as_float(winA[HL4 + j]);
  float b = wsum_f<-R, R>(winB, HL4);
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
