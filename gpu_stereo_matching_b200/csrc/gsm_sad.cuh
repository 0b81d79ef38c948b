// gsm_sad.cuh -- fused AD -> clipped-window SAD -> WTA kernel (GSM_MODE_SAD).
//
// Replaces kernalPreCal_V2 + kernalFindCorr (reference: BlockMatching/Device.cu:19-64) and is bit-exact to
// getDisp (BlockMatching/BlockMatching.cpp:111-189).  The D x H x W difference volume the reference writes
// to HBM (Device.cu:187-194) is never materialised: a CTA owns a strip of columns x 32 disparities and
// marches down the rows keeping only running window sums in registers.
//
//   thread  = (run of K consecutive columns) x (one disparity);  warp = 32 disparities of one run
//   per row : AD of the entering and leaving image rows (VABSDIFF4, 4 pixels/op)
//             -> vertical running sums V[K] (IDP.2A: V += 256*p_new - 256*p_old, one op per column)
//             -> V rows exchanged through shared memory (halo of R columns from the neighbour runs)
//             -> horizontal sliding sums, seeded with d so that the sum IS the packed key (SAD<<8)|d
//             -> warp min over the 32 disparities (REDUX) -> one atomicMin per pixel on the packed plane
// Algorithmic HBM traffic: 2 B/pixel read + 8 B/pixel packed-min plane traffic (L2 resident).
#pragma once
#include "gsm_common.cuh"

namespace gsm {

// (pitch/4) odd: the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups
__host__ __device__ constexpr int exch_pitch_words(int runs, int K, int HL4) {
  return (HL4 + runs * K + HL4) + (((((HL4 + runs * K + HL4) / 4) & 1) == 0) ? 4 : 0);
}

template <int R, int K, bool EXPORT>
__global__ void __launch_bounds__(512, 2)  // two CTAs of 16 warps per SM: at most 64 registers
sad_wta_kernel(const u8* __restrict__ Gp, const u8* __restrict__ Op, i64* __restrict__ keys, FusedGeom g) {
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
  for (int i = threadIdx.y * WARP + threadIdx.x; i < 2 * WARP * pitchw; i += runs * WARP) smem[i] = 0u;
  __syncthreads();

  const int x0 = strip * g.TW - g.hl + run * K;  // image column of this thread's first pixel
  const int dd = min(d, MAX_DISP - 1);           // lanes past d_end compute in-bounds garbage, never submitted
  const int osh = (g.view == 0) ? -dd : dd;       // "other" image is sampled at x - d (left) or x + d (right view)

  const u8* gbase = Gp + (size_t)frame * g.pg.plane_stride + (size_t)PADV * pitch + g.pg.xoff + x0;
  const u8* obase_b = Op + (size_t)frame * g.pg.plane_stride + (size_t)PADV * pitch + g.pg.xoff + x0 + osh;
  const u32 omis = (u32)(reinterpret_cast<uintptr_t>(obase_b) & 3u);
  const u32* obase = reinterpret_cast<const u32*>(obase_b - omis);
  const u32 osel = 0x3210u + 0x1111u * omis;

  // byte masks: a pixel contributes iff it is inside the image and (left view) x >= d:
  // BlockMatching.cpp:147-149 leaves dif_ at its memset 0 for c = x - d < 0.
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

  // output validity of column c of this thread: inside the strip's output window, inside the image,
  // d in range, and the reference's search cut-off `col + d > cols -> break` (BlockMatching.cpp:166).
  const int out0 = strip * g.TW;
  const int c_lo = max(0, out0 - x0);
  int c_hi = min(K - 1, min(out0 + g.TW, W) - 1 - x0);
  c_hi = min(c_hi, W - d - x0);
  if (d >= g.d_end) c_hi = -1;
  const bool all_valid = __all_sync(0xffffffffu, c_lo == 0 && c_hi == K - 1);

  int V[K];
#pragma unroll
  for (int c = 0; c < K; ++c) V[c] = 0;

  const int r0 = yb0 - R;  // first image row that may enter a window of this band
  constexpr int COEF = (int)0xFF000100;  // lo16 = +256, hi16 = -256: V holds 256 * (vertical sum)

  // running row pointers (one add per row instead of a 64-bit multiply per load)
  const u8* g_new = gbase + (long long)(yb0 - R) * pitch;                   // row y + R
  const u8* o_new = reinterpret_cast<const u8*>(obase) + (long long)(yb0 - R) * pitch;
  const u8* g_old = gbase + (long long)(yb0 - 3 * R - 1) * pitch;           // row y - R - 1
  const u8* o_old = reinterpret_cast<const u8*>(obase) + (long long)(yb0 - 3 * R - 1) * pitch;
  for (int y = yb0 - 2 * R; y < yb1; ++y, g_new += pitch, o_new += pitch, g_old += pitch, o_old += pitch) {
    u32 pn[KW], po[KW];
    {
      u32 gw[KW], ow[KW];
      load_aligned<K>(g_new, gw);
      load_unaligned<K>(reinterpret_cast<const u32*>(o_new), osel, ow);
#pragma unroll
      for (int w = 0; w < KW; ++w) pn[w] = __vabsdiffu4(gw[w], ow[w]);
    }
    if (y - R - 1 >= r0) {
      u32 gw[KW], ow[KW];
      load_aligned<K>(g_old, gw);
      load_unaligned<K>(reinterpret_cast<const u32*>(o_old), osel, ow);
#pragma unroll
      for (int w = 0; w < KW; ++w) po[w] = __vabsdiffu4(gw[w], ow[w]);
    } else {
#pragma unroll
      for (int w = 0; w < KW; ++w) po[w] = 0u;
    }
    if (need_mask) {
#pragma unroll
      for (int w = 0; w < KW; ++w) { pn[w] &= mask[w]; po[w] &= mask[w]; }
    }
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const u32 pair = __byte_perm(pn[c / 4], po[c / 4], (c & 3) | ((4 + (c & 3)) << 4));  // {p_new, p_old, x, x}
      V[c] = dp2a_lo_su(COEF, pair, V[c]);
    }
    if (y < yb0) continue;

    // ---- horizontal pass: publish V, read the neighbours' halo, slide ----
    u32* buf = smem + (size_t)(y & 1) * WARP * pitchw + (size_t)lane * pitchw + HL4 + run * K;
#pragma unroll
    for (int w = 0; w < KW; ++w)
      reinterpret_cast<uint4*>(buf)[w] = make_uint4(V[4 * w], V[4 * w + 1], V[4 * w + 2], V[4 * w + 3]);
    __syncthreads();
    int win[HL4 + K + HL4];
#pragma unroll
    for (int w = 0; w < HL4 / 4; ++w) {
      const uint4 a = reinterpret_cast<const uint4*>(buf - HL4)[w];
      win[4 * w] = a.x; win[4 * w + 1] = a.y; win[4 * w + 2] = a.z; win[4 * w + 3] = a.w;
      const uint4 b = reinterpret_cast<const uint4*>(buf + K)[w];
      win[HL4 + K + 4 * w] = b.x; win[HL4 + K + 4 * w + 1] = b.y;
      win[HL4 + K + 4 * w + 2] = b.z; win[HL4 + K + 4 * w + 3] = b.w;
    }
#pragma unroll
    for (int c = 0; c < K; ++c) win[HL4 + c] = V[c];

    u32 key[K];
    int S = d;  // seeding with d makes the window sum the packed key (SAD << 8) | d
#pragma unroll
    for (int j = -R; j <= R; ++j) S += win[HL4 + j];
    key[0] = (u32)S;
#pragma unroll
    for (int c = 1; c < K; ++c) {
      S += win[HL4 + c + R] - win[HL4 + c - R - 1];
      key[c] = (u32)S;
    }

    if constexpr (EXPORT) {
      const int de = d - g.export_d0;
      if (de >= 0 && de < g.export_nd && d < g.d_end) {
        int* out = reinterpret_cast<int*>(g.export_ptr) + ((size_t)de * H + y) * W;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const int x = x0 + c;
          if (x >= out0 && x < min(out0 + g.TW, W)) out[x] = (int)(key[c] >> 8);
        }
      }
    }

    if (!all_valid) {
#pragma unroll
      for (int c = 0; c < K; ++c)
        if (c < c_lo || c > c_hi) key[c] = 0xffffffffu;
    }
    // lane c keeps the minimum of column c: select tree on the lane bits (loop-invariant predicates)
    static_assert(K == 16, "select tree below assumes 16 columns per thread");
    u32 m[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) m[c] = __reduce_min_sync(0xffffffffu, key[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = (lane & 1) ? m[2 * i + 1] : m[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = (lane & 2) ? m[2 * i + 1] : m[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) m[i] = (lane & 4) ? m[2 * i + 1] : m[2 * i];
    const u32 mine = (lane & 8) ? m[1] : m[0];
    if (lane < K && mine != 0xffffffffu)
      atomicMin(keys + ((size_t)frame * H + y) * W + x0 + lane, (i64)mine);
  }
}

}  // namespace gsm
