// gsm_util.cuh -- the small per-pixel kernels around the fused aggregation kernels:
// plane packing, packed-min plane init / finalize, LR check, median, debug exports.
#pragma once
#include "gsm_common.cuh"

namespace gsm {

// round-to-nearest-even + saturate to u8: the reference's float2uchar (Device.cu:145-150)
__device__ __forceinline__ u8 float2uchar_rni_sat(float a) {
  u32 res;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(res) : "f"(a));
  return (u8)res;
}

// BilinearInterpolation + float2uchar (Device.cu:145-167) of src at (row = my, col = mx).  Products and sums are
// rounded separately (__fmul_rn/__fadd_rn: no FMA contraction) so the result equals the reference's CPU twin
// (CPU_BilinearInterpolation, Utility.cpp:248-264) bit for bit.
__device__ __forceinline__ u8 remap_sample(const u8* __restrict__ src, int rows, int cols, float mx, float my) {
  const float x = my, y = mx;  // the reference calls the interpolator as (src, ycoo, xcoo): x is the row coordinate
  const int x1 = (int)floorf(x), y1 = (int)floorf(y), x2 = x1 + 1, y2 = y1 + 1;
  float result = 0.f;
  if (!(x1 < 0 || x2 >= rows || y1 < 0 || y2 >= cols)) {
    const size_t b = (size_t)x1 * cols + y1;
    const float Q11 = src[b], Q12 = src[b + 1], Q21 = src[b + cols], Q22 = src[b + cols + 1];
    const float wx2 = __fsub_rn((float)x2, x), wx1 = __fsub_rn(x, (float)x1);
    const float left = __fadd_rn(__fmul_rn(wx2, Q11), __fmul_rn(wx1, Q21));
    const float right = __fadd_rn(__fmul_rn(wx2, Q12), __fmul_rn(wx1, Q22));
    result = __fadd_rn(__fmul_rn(__fsub_rn((float)y2, y), left), __fmul_rn(__fsub_rn(y, (float)y1), right));
  }
  return float2uchar_rni_sat(result);
}

// tight [n][H][W] u8  ->  padded plane (see gsm_common.cuh).  fill 0: zero pad everywhere;
// fill 1: columns >= W of image rows replicate src[y][W-1] (right-view "other" image,
// STMatching/StereoHelper.cpp:170-176: x+d >= W falls back to the last valid disparity).
// mapx/mapy (optional, tight float [H][W]): rectification fused into the packer -- the plane receives
// remap(src) (raw frames in, disparity out) instead of a copy of src.
__global__ void pack_plane_kernel(const u8* __restrict__ src, u8* __restrict__ dst, PlaneGeom pg, int fill,
                                  const float* __restrict__ mapx, const float* __restrict__ mapy,
                                  const FrameDesc* __restrict__ ft) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;  // 4-byte group within a padded row
  const int prow = blockIdx.y;
  const int frame = blockIdx.z;
  if (q * 4 >= pg.pitch) return;
  const int y = prow - PADV;
  int H = pg.H, W = pg.W;
  size_t off = (size_t)frame * H * W;
  if (ft) { const FrameDesc fd = ft[frame]; H = fd.H; W = fd.W; off = (size_t)fd.off; }
  u32 v = 0;
  if (y >= 0 && y < H) {
    const u8* s = src + off + (size_t)y * W;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int x = q * 4 + b - pg.xoff;
      u32 px = 0;
      if (fill == 1 && x >= W) x = W - 1;
      if (x >= 0 && x < W) {
        if (mapx) {
          const size_t mi = (size_t)y * W + x;
          px = remap_sample(s - (size_t)y * W, H, W, mapx[mi], mapy[mi]);
        } else {
          px = s[x];
        }
      }
      v |= px << (8 * b);
    }
  }
  reinterpret_cast<u32*>(dst + (size_t)frame * pg.plane_stride + (size_t)prow * pg.pitch)[q] = v;
}

__global__ void fill_keys_kernel(i64* __restrict__ keys, size_t n, i64 v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = v;
}

// packed (cost, d) word -> u8 disparity.  Both key formats keep d in the low byte.
__global__ void finalize_keys_kernel(const i64* __restrict__ keys, u8* __restrict__ disp, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) disp[i] = (u8)(keys[i] & 0xff);
}

// Disparity-split combine over peer memory (NVLink P2P loads / stores).  Every rank holds a packed-min
// plane for its own disparity range; rank r reduces pixels [begin, end) -- its 1/world slice -- over ALL ranks'
// planes with a signed 64-bit min (order of (cost, d), ties to the lowest d), turns the winners into disparities
// and stores the slice into every rank's map: reduce-scatter, finalize and all-gather of the u8 result in one
// pass that moves 8 B/pixel/rank in and 1 B/pixel/rank out instead of an all-reduce of the 8-byte planes.
constexpr int P2P_MAX_RANKS = 16;
struct PeerPlanes {
  const i64* keys[P2P_MAX_RANKS];
  u8* disp[P2P_MAX_RANKS];
};
// A warp reduces chunks of 512 pixels: every load instruction of the warp covers 512 contiguous bytes of ONE plane
// (whole 128-byte lines over NVLink instead of 32 scattered 16-byte pieces), the next peer's eight vectors are in
// flight while the current peer's are reduced, the 512 winners are transposed through 512 bytes of shared memory so
// that every lane stores 16 contiguous bytes per map.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
reduce_keys_p2p_kernel(PeerPlanes pp, int world, int rank, size_t begin, size_t end) {
  __shared__ __align__(16) u8 tr[WARPS][512];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const size_t nwarps = (size_t)gridDim.x * WARPS;
  const size_t nchunks = (end - begin) / 512;
  for (size_t ch = (size_t)blockIdx.x * WARPS + wib; ch < nchunks; ch += nwarps) {
    const size_t base = begin + ch * 512;
    longlong2 m[8], nx[8];
    const longlong2* own = reinterpret_cast<const longlong2*>(pp.keys[rank] + base) + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = own[32 * j];
    if (world > 1) {
      const longlong2* src = reinterpret_cast<const longlong2*>(pp.keys[(rank + 1) % world] + base) + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) nx[j] = src[32 * j];
    }
    for (int w = 1; w < world; ++w) {  // every rank starts on a different peer
      longlong2 cur[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nx[j];
      if (w + 1 < world) {
        const longlong2* src = reinterpret_cast<const longlong2*>(pp.keys[(rank + w + 1) % world] + base) + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) nx[j] = src[32 * j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m[j].x = min(m[j].x, cur[j].x);
        m[j].y = min(m[j].y, cur[j].y);
      }
    }
    // lane l holds pixels 64 j + 2 l, + 1: transpose to 16 contiguous pixels per lane
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<unsigned short*>(&tr[wib][64 * j + 2 * lane]) =
          (unsigned short)((u32)(m[j].x & 0xff) | ((u32)(m[j].y & 0xff) << 8));
    __syncwarp();
    const uint4 ov = *reinterpret_cast<const uint4*>(&tr[wib][16 * lane]);
    __syncwarp();
    for (int w = 0; w < world; ++w) *reinterpret_cast<uint4*>(pp.disp[(rank + w) % world] + base + 16 * lane) = ov;
  }
  // tail of a slice whose length is not a multiple of 512
  if (blockIdx.x == 0) {
    for (size_t k = begin + nchunks * 512 + threadIdx.x; k < end; k += blockDim.x) {
      i64 mk = pp.keys[rank][k];
      for (int w = 1; w < world; ++w) mk = min(mk, pp.keys[(rank + w) % world][k]);
      for (int w = 0; w < world; ++w) pp.disp[w][k] = (u8)(mk & 0xff);
    }
  }
}

// STMatching/StereoDisparity.cpp:136-147.  out_disp (optional) = DL with occluded pixels zeroed.
__global__ void lr_check_kernel(const u8* __restrict__ DL, const u8* __restrict__ DR, u8* __restrict__ occ,
                                u8* __restrict__ mask, u8* __restrict__ out_disp, int H, int W, int n,
                                const FrameDesc* __restrict__ ft) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  size_t off = (size_t)f * H * W;
  if (ft) { const FrameDesc fd = ft[f]; H = fd.H; W = fd.W; off = (size_t)fd.off; }
  if (x >= W || y >= H) return;
  const size_t i = off + (size_t)y * W + x;
  const int d = DL[i];
  u8 o;
  if (x - d >= 0) {
    const int dc = DR[i - d];
    o = (u8)(d == 0 || abs(d - dc) > 1);
  } else {
    o = 1;
  }
  if (occ) occ[i] = o;
  if (mask) mask[i] = (u8)!o;
  if (out_disp) out_disp[i] = o ? 0 : (u8)d;
}

// (2m+1)^2 median, replicate border: smallest v whose cumulative count exceeds t = 2m^2+2m
// (STMatching/ctmf.c:281,288-294,319-325).  One pixel per thread, binary search on the value,
// neighbourhood served from a shared-memory tile.
constexpr int MED_TX = 32, MED_TY = 16, MED_MAXR = 7;
__global__ void __launch_bounds__(MED_TX* MED_TY)
median_kernel(const u8* __restrict__ src, u8* __restrict__ dst, int H, int W, int m, const FrameDesc* __restrict__ ft) {
  __shared__ u8 tile[(MED_TY + 2 * MED_MAXR) * (MED_TX + 2 * MED_MAXR)];
  const int f = blockIdx.z;
  size_t off = (size_t)f * H * W;
  if (ft) { const FrameDesc fd = ft[f]; H = fd.H; W = fd.W; off = (size_t)fd.off; }
  const int bx = blockIdx.x * MED_TX, by = blockIdx.y * MED_TY;
  if (bx >= W || by >= H) return;  // (mixed-size batches) tile outside this frame
  const int tw = MED_TX + 2 * m, th = MED_TY + 2 * m;
  const u8* s = src + off;
  for (int i = threadIdx.y * MED_TX + threadIdx.x; i < tw * th; i += MED_TX * MED_TY) {
    const int ty = i / tw, tx = i - ty * tw;
    const int yy = min(H - 1, max(0, by + ty - m));
    const int xx = min(W - 1, max(0, bx + tx - m));
    tile[ty * tw + tx] = s[(size_t)yy * W + xx];
  }
  __syncthreads();
  const int x = bx + threadIdx.x, y = by + threadIdx.y;
  if (x >= W || y >= H) return;
  const int t = 2 * m * m + 2 * m;
  int lo = 0, hi = 255;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    int cnt = 0;
    for (int dy = 0; dy <= 2 * m; ++dy) {
      const u8* row = tile + (threadIdx.y + dy) * tw + threadIdx.x;
      for (int dx = 0; dx <= 2 * m; ++dx) cnt += (row[dx] <= mid);
    }
    if (cnt > t) hi = mid; else lo = mid + 1;
  }
  dst[off + (size_t)y * W + x] = (u8)lo;
}

// PreCal, BlockMatching/BlockMatching.cpp:89-109 (== kernalPreCal_V2, Device.cu:19-32): debug export only,
// the production path never materialises this volume.
__global__ void ad_volume_kernel(const u8* __restrict__ L, const u8* __restrict__ R, u8* __restrict__ out, int H,
                                 int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int d = blockIdx.z;
  if (x >= W) return;
  const size_t i = (size_t)y * W + x;
  u8 v = 0;
  if (x - d >= 0) v = (u8)abs((int)L[i] - (int)R[i - d]);
  out[(size_t)d * H * W + i] = v;
}

// int32 SAD slices [nd][H][W] -> getAllSAD layout u8 [H*W][D] (BlockMatching.cpp:244-258)
__global__ void all_sad_pack_kernel(const int* __restrict__ slices, u8* __restrict__ out, int H, int W, int D,
                                    int d0, int nd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int k = blockIdx.z;
  if (x >= W || k >= nd) return;
  const int d = d0 + k;
  const size_t i = (size_t)y * W + x;
  out[i * D + d] = (x + d > W) ? (u8)255 : (u8)slices[(size_t)k * H * W + i];
}

// kernalRemap (Device.cu:127-134): one image through one pair of maps
__global__ void remap_kernel(const u8* __restrict__ src, const float* __restrict__ mapx, const float* __restrict__ mapy,
                             u8* __restrict__ dst, int rows, int cols) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= cols) return;
  const size_t i = (size_t)row * cols + col;
  dst[i] = remap_sample(src, rows, cols, mapx[i], mapy[i]);
}

// kernalCvtColor (Device.cu:136-143) / cvtColor_cpu (Utility.cpp:289-298).  The two reference functions do not
// round alike: the host function is compiled without contraction (three products, two sums, truncation), while nvcc
// contracts the kernel's expression into FMUL(.587 c1), FFMA(.299 c0), FFMA(.114 c2) before cvt.rni.sat -- checked
// on the reference's own kernel compiled for sm_100a (profiles/r01_reference_gpu_on_b200.txt).
__global__ void cvtcolor_kernel(const u8* __restrict__ src3, u8* __restrict__ dst, size_t n, int truncate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c0 = src3[3 * i], c1 = src3[3 * i + 1], c2 = src3[3 * i + 2];
  if (truncate) {
    const float sum = __fadd_rn(__fadd_rn(__fmul_rn(.299f, c0), __fmul_rn(.587f, c1)), __fmul_rn(.114f, c2));
    dst[i] = (u8)sum;
  } else {
    const float sum = __fmaf_rn(.114f, c2, __fmaf_rn(.299f, c0, __fmul_rn(.587f, c1)));
    dst[i] = float2uchar_rni_sat(sum);
  }
}

// depth = f*B / d for a rectified rig (Q matrix of stereoRectify, Utility.cpp:228-234); 0 where d == 0
__global__ void depth_kernel(const u8* __restrict__ disp, float* __restrict__ depth, size_t n, float fB) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = disp[i];
  depth[i] = d ? __fdiv_rn(fB, (float)d) : 0.f;
}

// number of SM ids the device hands out (%nsmid): sizes the per-SM rings of gsm_gf5.cuh
__global__ void nsmid_kernel(u32* out) {
  u32 n;
  asm("mov.u32 %0, %%nsmid;" : "=r"(n));
  *out = n;
}

// FFMA + IADD3 issue-peak probe (roofline denominator for the ALU-bound fused kernels)
__global__ void __launch_bounds__(256) alu_peak_kernel(u32* out, int iters) {
  float f[8];
  u32 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = 1.0f + threadIdx.x * 1e-3f + i; a[i] = threadIdx.x * (i + 3) + 1u; }
  const float fb = 1.000001f;
  const u32 c = blockIdx.x + 7u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = (i + 1) & 7;
      f[i] = fmaf(f[i], fb, f[j]);
      a[i] = a[i] + a[j] + c;
    }
  }
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace gsm
