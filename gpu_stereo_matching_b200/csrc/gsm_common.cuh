// gsm_common.cuh -- geometry, padded-plane layout and small device helpers shared by all kernels.
//
// HBM layout ("padded plane"): every u8 image the fused kernels read lives in a plane of
//   plane_rows = H + 2*PADV rows x pitch bytes,  image pixel (y, x) at  (PADV + y) * pitch + xoff + x.
// The pad is zero (or right-replicated for the right-view "other" image), pitch and xoff are chosen so
// that every thread's run of K columns starts 16-byte aligned: rows move as 16-byte aligned bulk copies into shared
// memory and are read from there with 128-bit loads.
// Zero pad rows/cols make out-of-image absolute differences vanish without per-row predicates.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gsm {

typedef uint8_t u8;
typedef uint32_t u32;
typedef long long i64;

constexpr int PADV = 24;        // zero rows above and below the image
constexpr int PADL_BASE = 320;  // >= max left halo (24) + max disparity (256) + word over-read
constexpr int PADR = 544;       // >= strip width (<=256) + max disparity (256) + slack
constexpr int MAX_DISP = 256;
constexpr int WARP = 32;

struct PlaneGeom {
  int H, W;            // image size
  int pitch;           // bytes per padded row (multiple of 16)
  int xoff;            // byte offset of image column 0 inside a padded row
  int plane_rows;      // H + 2*PADV
  size_t plane_stride; // bytes per frame
};

// Per-frame geometry of a mixed-size batch (gsm_stereo_batch_v): frame f is H x W pixels at pixel offset `off` of the
// tight image / packed-min / disparity arrays, inside a padded plane slot laid out for the LARGEST frame of the batch.
// A null table means every frame is pg.H x pg.W at offset f * H * W.
struct FrameDesc {
  int H, W;
  long long off;
};

// One launch of a fused aggregation+WTA kernel.
struct FusedGeom {
  PlaneGeom pg;
  const FrameDesc* ft;  // per-frame sizes (device), or nullptr
  int D;          // total disparities (for validity)
  int d_begin;    // first disparity evaluated by blockIdx.y == 0
  int d_end;      // one past the last disparity evaluated
  int TW;         // output columns per strip (multiple of 16)
  int hl;         // left halo of a strip (multiple of 4, >= stage halo)
  int runs;       // == blockDim.y
  int bands;      // row bands per frame
  int band_rows;  // rows per band
  int view;       // 0: left (guide L, other R at x-d, zero for x<d); 1: right (guide R, other Lrep at x+d)
  // debug export of aggregated cost slices (nullptr = off)
  void* export_ptr;
  int export_d0, export_nd;
};

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// float -> int32 whose SIGNED order equals the float order (no NaNs on this path)
__device__ __forceinline__ int sortable_i32(float f) {
  int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}

// acc + a.lo16 * b.byte0 + a.hi16 * b.byte1   (IDP.2A, signed 16-bit x unsigned 8-bit)
__device__ __forceinline__ int dp2a_lo_su(int a16x2, u32 b8x4, int acc) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b8x4), "r"(acc));
  return d;
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

}  // namespace gsm
