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
// The four image rows of a step (guide and other image, entering and leaving row) are staged in shared memory two
// steps ahead by one elected thread with cp.async.bulk + mbarrier (three stages), like the guided-filter kernel:
// with direct LDGs 27% of the warp samples waited on the long scoreboard (profiles/r01_sad_*).
// Algorithmic HBM traffic: 2 B/pixel read + 8 B/pixel packed-min plane traffic (L2 resident).
#pragma once
#include "gsm_common.cuh"

namespace gsm {

// (pitch/4) odd: the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups
__host__ __device__ constexpr int exch_pitch_words(int runs, int K, int HL4) {
  return (HL4 + runs * K + HL4) + (((((HL4 + runs * K + HL4) / 4) & 1) == 0) ? 4 : 0);
}

// Input stage of one march step: G[2][TWt] guide rows y+R and y-R-1, O[2][TWt+64] the other image's rows, window
// covering the CTA's 32 disparities.  SAD_NST stages, copies issued SAD_NST-1 steps ahead.
constexpr int SAD_NST = 3, SAD_HDR = 64;
__host__ __device__ constexpr int sad_stage_bytes(int twt) { return 2 * twt + 2 * (twt + 64); }
__host__ __device__ inline size_t sad_smem_bytes(int runs, int K, int HL4) {
  return SAD_HDR + SAD_NST * (size_t)sad_stage_bytes(runs * K) + 2 * (size_t)WARP * exch_pitch_words(runs, K, HL4) * sizeof(u32);
}

template <int R, int K, bool EXPORT>
__global__ void __launch_bounds__(512, 2)  // two CTAs of 16 warps per SM: at most 64 registers
sad_wta_kernel(const u8* __restrict__ Gp, const u8* __restrict__ Op, i64* __restrict__ keys, FusedGeom g) {
  constexpr int HL4 = (R + 3) / 4 * 4;
  constexpr int KW = K / 4;
  extern __shared__ __align__(128) u8 smem_raw[];

  const int lane = threadIdx.x;
  const int run = threadIdx.y;
  const int runs = blockDim.y;
  const int strip = blockIdx.x;
  const int d = g.d_begin + blockIdx.y * WARP + lane;
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

  const int pitchw = exch_pitch_words(runs, K, HL4);
  const int TWt = runs * K;
  const int GW = TWt, OW = TWt + 64, stage_bytes = sad_stage_bytes(TWt);
  u8* stage_base = smem_raw + SAD_HDR;
  u32* smem = reinterpret_cast<u32*>(stage_base + SAD_NST * stage_bytes);  // exchange planes [2][WARP][pitchw]
  const u32 bar0 = smem_u32(smem_raw);
  const bool producer = (threadIdx.x == 0 && threadIdx.y == 0);
  for (int i = threadIdx.y * WARP + threadIdx.x; i < 2 * WARP * pitchw; i += runs * WARP) smem[i] = 0u;
  if (producer) {
#pragma unroll
    for (int i = 0; i < SAD_NST; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int xs = strip * g.TW - g.hl;
  const int x0 = xs + run * K;            // image column of this thread's first pixel
  const int d0 = g.d_begin + blockIdx.y * WARP;
  const int dd = min(d, MAX_DISP - 1);    // lanes past d_end compute in-bounds garbage, never submitted

  // "other" image is sampled at x - d (left) or x + d (right view): staged window of the CTA's 32 disparities
  const u8* gsrc = Gp + (size_t)frame * g.pg.plane_stride + (size_t)PADV * pitch + g.pg.xoff + xs;
  const int ostart = g.pg.xoff + xs + (g.view == 0 ? -(min(d0, MAX_DISP - 1) + WARP - 1) : min(d0, MAX_DISP - 1));
  const int oalign = ostart & 15;
  const u8* osrc = Op + (size_t)frame * g.pg.plane_stride + (size_t)PADV * pitch + (ostart - oalign);
  const int ooff = oalign + run * K + (g.view == 0 ? (min(d0, MAX_DISP - 1) + WARP - 1 - dd) : dd - min(d0, MAX_DISP - 1));
  const int row_lo = -PADV, row_hi = H + PADV - 1;
  auto issue = [&](int y, int s) {  // rows y + R (entering) and y - R - 1 (leaving) of march step y
    const u32 bar = bar0 + 8 * s;
    const u32 dst = smem_u32(stage_base + (size_t)s * stage_bytes);
    mbar_expect_tx(bar, (u32)stage_bytes);
    const int rows2[2] = {y + R, y - R - 1};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long ro = (long long)max(row_lo, min(row_hi, rows2[i])) * pitch;
      bulk_g2s(dst + i * GW, gsrc + ro, GW, bar);
      bulk_g2s(dst + 2 * GW + i * OW, osrc + ro, OW, bar);
    }
  };

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

  const int y_begin = yb0 - 2 * R;
  if (producer) {
#pragma unroll
    for (int i = 0; i < SAD_NST - 1; ++i)
      if (y_begin + i < yb1) issue(y_begin + i, i);
  }
  int s = 0;       // stage of step y
  u32 sphase = 0;  // its mbarrier parity
  for (int y = y_begin; y < yb1; ++y) {
    mbar_wait(bar0 + 8 * s, sphase);
    const u8* stg = stage_base + (size_t)s * stage_bytes;
    u32 pn[KW], po[KW];
    {
      u32 gw[KW], ow[KW];
      const uint4 gv = *reinterpret_cast<const uint4*>(stg + run * K);
      gw[0] = gv.x; gw[1] = gv.y; gw[2] = gv.z; gw[3] = gv.w;
      lds_unaligned<K>(stg + 2 * GW, ooff, ow);
#pragma unroll
      for (int w = 0; w < KW; ++w) pn[w] = __vabsdiffu4(gw[w], ow[w]);
    }
    if (y - R - 1 >= r0) {
      u32 gw[KW], ow[KW];
      const uint4 gv = *reinterpret_cast<const uint4*>(stg + GW + run * K);
      gw[0] = gv.x; gw[1] = gv.y; gw[2] = gv.z; gw[3] = gv.w;
      lds_unaligned<K>(stg + 2 * GW + OW, ooff, ow);
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
    // ---- horizontal pass: publish V, read the neighbours' halo, slide ----
    u32* buf = smem + (size_t)(y & 1) * WARP * pitchw + (size_t)lane * pitchw + HL4 + run * K;
    if (y >= yb0) {
#pragma unroll
      for (int w = 0; w < KW; ++w)
        reinterpret_cast<uint4*>(buf)[w] = make_uint4(V[4 * w], V[4 * w + 1], V[4 * w + 2], V[4 * w + 3]);
    }
    __syncthreads();
    // every thread has finished step y-1 completely: its stage is refilled for step y + SAD_NST - 1
    if (producer && y + SAD_NST - 1 < yb1) issue(y + SAD_NST - 1, s == 0 ? SAD_NST - 1 : s - 1);
    if (++s == SAD_NST) { s = 0; sphase ^= 1u; }
    if (y < yb0) continue;
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
      atomicMin(keys + (size_t)y * W + x0 + lane, (i64)mine);
  }
}

}  // namespace gsm
