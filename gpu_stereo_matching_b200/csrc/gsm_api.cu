// gsm_api.cu -- host side of libgsm.so: context, launch plans and the C ABI declared in include/gsm.h.
// Replaces the host "proxy" functions of the reference (BlockMatching/Device.cu:173-367).
#include "../../include/gsm.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gsm_common.cuh"
#include "gsm_gf.cuh"
#include "gsm_gf3.cuh"
#include "gsm_sad.cuh"
#include "gsm_st.cuh"
#include "gsm_st_host.hpp"
#include <cub/device/device_radix_sort.cuh>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include "gsm_util.cuh"

using namespace gsm;

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(GSM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

struct gsm_ctx {
  int device = 0;
  int max_rows = 0, max_cols = 0, max_disp = 0, max_batch = 0;
  cudaStream_t stream = nullptr;                 // compute
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // host path: uploads / downloads overlap the kernels
  cudaEvent_t ev_h2d[2] = {}, ev_in_free[2] = {}, ev_done[2] = {}, ev_d2h[2] = {};
  int slot_frames = 1;                             // frames per host-path slot (two slots)
  float* rect_maps = nullptr;                      // [mapx_L][mapy_L][mapx_R][mapy_R], tight rows x cols each
  int rect_rows = 0, rect_cols = 0;
  long long chunk_seq = 0;                         // host-path chunks submitted so far (slot = chunk_seq & 1)
  // device buffers, each sized for max_batch frames
  u8 *tightL = nullptr, *tightR = nullptr;                       // uploads on the host path
  u8 *planeL = nullptr, *planeR = nullptr, *planeLrep = nullptr;  // padded planes
  float* stats[2] = {nullptr, nullptr};                           // GF guide statistics (left / right guide)
  i64 *keysL = nullptr, *keysR = nullptr;                         // packed-min planes
  u8 *dispA = nullptr, *dispB = nullptr, *dispC = nullptr, *dispD = nullptr, *maskD = nullptr;
  u8* dispOut = nullptr;                                          // final map of the host path before D2H
  struct StatGeom { int rows = -1, cols = -1, xoff = -1, n = -1; } stat_geom[2];  // geometry the statistic planes were zeroed for
  // segment-tree stereo (gsm_st.cuh): one device arena, grown on demand
  void* st_buf = nullptr;
  cudaStream_t st_streams[8] = {};      // gsm_segment_tree_stereo_batch: one per device arena (created on first use)
  struct StWorker;                      // host work space of one tree-builder thread (gsm_segment_tree_stereo_batch)
  std::vector<StWorker*> st_workers;
  void* st_pin = nullptr;  // pinned host mirror of the arena's tree block + the weight / disparity read-back
  size_t st_pin_bytes = 0;
  size_t st_bytes = 0;
  int st_ring_smem[3] = {-1, -1, -1};  // dynamic shared memory st_filter_ring_kernel<16 / 8 / 4> may use, -1 = not asked yet
  float* ring = nullptr;                                          // gf5_wta_kernel: per-SM rings of (a, b) rows (gsm_gf5.cuh)
  int ring_slots = 0;                                             // == %nsmid of the device
  size_t ring_bytes = 0, l2_bytes = 0;
  unsigned attr_gf5[2] = {0, 0};
  cudaEvent_t fused_wait = nullptr;                                // one-shot: the next fused kernel waits for it (gsm_partial_keys_device_ex)
  FrameDesc* ft_dev = nullptr;                                    // per-frame sizes of a mixed-size batch (max_batch entries)
  unsigned attr_sad[2] = {0, 0}, attr_gf[2] = {0, 0};             // radii whose kernels already carry the smem attribute
  void* export_buf = nullptr;
  size_t export_bytes = 0;
  u32* peak_buf = nullptr;
  size_t plane_bytes_per_frame = 0;
  long long launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> ev;  // (start, stop) pairs of the fused kernels of the last device call
  size_t ev_used = 0;
};
struct gsm_ctx::StWorker {
  gsm_st::Tree t;
  gsm_st::detail::Work k;
  cudaStream_t s = nullptr;  // the builder's own stream and device scratch (records phase)
  void* dev = nullptr;
  size_t dev_bytes = 0, sort_temp = 0, scratch_n = 0;
};

// strip halo (columns) a fused kernel needs on each side of its output columns
static int stage_halo_of(int /*mode*/, int radius) { return radius; }  // GF: stage 1 needs no exchanged halo

// CTA shape of the fused kernels (compile-time in the kernels, mirrored here for the plane geometry)
#ifndef GSM_SAD_RUNS
#define GSM_SAD_RUNS 16
#endif
#ifndef GSM_GF_K
#define GSM_GF_K 16
#endif
#ifndef GSM_GF_RUNS
#define GSM_GF_RUNS 12
#endif
#ifndef GSM_GF_LPR
#define GSM_GF_LPR 32
#endif
// GSM_GF_RING: which guided-filter kernel a launch uses.
//   0: always gf3_wta_kernel (recomputes the (a, b) row that leaves the vertical window);
//   2 (product): gf5_wta_kernel (gsm_gf5.cuh: reads that row back from a per-SM ring in global memory) for the radii
//      whose rings stay in the L2 -- 148 rings x (2r+1) rows x 48 KB within 55 % of the L2: r <= 4 on a B200, where it is
//      a tenth faster -- and gf3_wta_kernel for the larger radii, where the rings overflow the L2 and the ring kernel is
//      HBM-bound (profiles/experiments_r02/ab_results.txt, calls 5-9);
//   1 (experiments): gf5_wta_kernel for every radius.
#ifndef GSM_GF_RING
#define GSM_GF_RING 2
#endif
#if GSM_GF_RING == 1
#define GSM_GF_RING_MAXR 9
#elif GSM_GF_RING == 2
#define GSM_GF_RING_MAXR 5  // gf5 is instantiated up to this radius; the L2 test below decides per device
#else
#define GSM_GF_RING_MAXR 0
#endif
#if GSM_GF_RING
#include "gsm_gf5.cuh"
#endif
static int strip_columns(int mode) { return mode == GSM_MODE_SAD ? GSM_SAD_RUNS * 16 : GSM_GF_RUNS * GSM_GF_K; }
// left halo of a strip: a multiple of 4 >= the stage halo; a whole 16-column run when that costs no output column
// (then the first run, like the last, only feeds its neighbours and skips the output stage)
static int strip_left_halo(int TWt, int stage_halo) {
  const int hl = round_up(stage_halo, 4);
  return ((TWt - 16 - stage_halo) / 16 == (TWt - hl - stage_halo) / 16) ? 16 : hl;
}

static int max_pitch(int cols) { return round_up(PADL_BASE + 16 + cols + PADR, 16); }

static PlaneGeom make_plane_geom(int rows, int cols, int hl) {
  PlaneGeom pg;
  pg.H = rows;
  pg.W = cols;
  pg.pitch = max_pitch(cols);
  pg.xoff = PADL_BASE + (hl % 16);  // makes (xoff - hl) a multiple of 16: every run starts 16-byte aligned
  pg.plane_rows = rows + 2 * PADV;
  pg.plane_stride = (size_t)pg.pitch * pg.plane_rows;
  return pg;
}

// ------------------------------------------------------------------------------------------------
extern "C" const char* gsm_last_error(void) { return g_err; }
extern "C" const char* gsm_version(void) { return "gsm-b200 0.1 (sm_100a)"; }

extern "C" void gsm_destroy(gsm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  void* bufs[] = {c->tightL, c->tightR, c->planeL, c->planeR, c->planeLrep, c->stats[0], c->stats[1], c->keysL,
                  c->keysR,  c->dispA,  c->dispB,  c->dispC,  c->dispD,     c->maskD,    c->export_buf, c->peak_buf, c->dispOut, c->rect_maps,
                  c->ft_dev, c->ring, c->st_buf};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (c->st_pin) cudaFreeHost(c->st_pin);
  for (gsm_ctx::StWorker* w : c->st_workers) {
    if (w->s) cudaStreamDestroy(w->s);
    if (w->dev) cudaFree(w->dev);
    delete w;
  }
  for (cudaStream_t st : c->st_streams)
    if (st) cudaStreamDestroy(st);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
    if (c->ev_in_free[i]) cudaEventDestroy(c->ev_in_free[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
  }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int gsm_create(gsm_ctx** out, int device, int max_rows, int max_cols, int max_disp, int max_batch) {
  if (!out) return fail(GSM_ERR_INVALID, "gsm_create: out is null");
  *out = nullptr;
  if (max_rows < 1 || max_cols < 1 || max_disp < 1 || max_disp > MAX_DISP || max_batch < 1)
    return fail(GSM_ERR_INVALID, "gsm_create: bad capacity rows=%d cols=%d disp=%d batch=%d", max_rows, max_cols,
                max_disp, max_batch);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GSM_ERR_CUDA, "gsm_create: no CUDA device (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(GSM_ERR_INVALID, "gsm_create: device %d of %d", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(GSM_ERR_UNSUPPORTED, "gsm_create: device %s is sm_%d%d; this library is built for sm_100a only",
                prop.name, prop.major, prop.minor);
  gsm_ctx* c = new gsm_ctx();
  c->device = device;
  c->max_rows = max_rows;
  c->max_cols = max_cols;
  c->max_disp = max_disp;
  c->max_batch = max_batch;
  const size_t px = (size_t)max_rows * max_cols * max_batch;
  c->plane_bytes_per_frame = (size_t)max_pitch(max_cols) * (max_rows + 2 * PADV);
  const size_t plane = c->plane_bytes_per_frame * max_batch;
  cudaError_t st = cudaSuccess;
  auto A = [&](void** p, size_t bytes) {
    if (st == cudaSuccess) st = cudaMalloc(p, bytes);
    // zero-fill on the context's own (non-blocking) stream: a legacy-stream cudaMemset would not be ordered with it
    if (st == cudaSuccess) st = cudaMemsetAsync(*p, 0, bytes, c->stream);
  };
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) st = cudaErrorUnknown;
  if (cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess) st = cudaErrorUnknown;
  if (cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess) st = cudaErrorUnknown;
  for (int i = 0; i < 2 && st == cudaSuccess; ++i) {
    st = cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming);
    if (st == cudaSuccess) st = cudaEventCreateWithFlags(&c->ev_in_free[i], cudaEventDisableTiming);
    if (st == cudaSuccess) st = cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming);
    if (st == cudaSuccess) st = cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming);
  }
  c->slot_frames = (max_batch + 1) / 2;
  const size_t slot_px = (size_t)max_rows * max_cols * c->slot_frames * 2;  // two host-path slots
  A((void**)&c->tightL, slot_px);
  A((void**)&c->tightR, slot_px);
  A((void**)&c->planeL, plane);
  A((void**)&c->planeR, plane);
  A((void**)&c->planeLrep, plane);
  A((void**)&c->stats[0], plane * sizeof(float) * GF_STAT_PLANES);
  A((void**)&c->stats[1], plane * sizeof(float) * GF_STAT_PLANES);
  A((void**)&c->keysL, px * sizeof(i64));
  A((void**)&c->keysR, px * sizeof(i64));
  A((void**)&c->dispA, px);
  A((void**)&c->dispB, px);
  A((void**)&c->dispC, px);
  A((void**)&c->dispD, px);
  A((void**)&c->maskD, slot_px);
  A((void**)&c->dispOut, slot_px);
  A((void**)&c->peak_buf, (size_t)prop.multiProcessorCount * 8 * 256 * sizeof(u32));
  A((void**)&c->ft_dev, (size_t)max_batch * sizeof(FrameDesc));
#if GSM_GF_RING
  {
    // one ring of (a, b) rows per SM id (gsm_gf5.cuh): ask the device how many SM ids it hands out
    u32 nsm = 0;
    if (st == cudaSuccess) {
      nsmid_kernel<<<1, 1, 0, c->stream>>>(c->peak_buf);
      st = cudaMemcpyAsync(&nsm, c->peak_buf, sizeof(u32), cudaMemcpyDeviceToHost, c->stream);
      if (st == cudaSuccess) st = cudaStreamSynchronize(c->stream);
      if (st == cudaSuccess) st = cudaMemsetAsync(c->peak_buf, 0, sizeof(u32), c->stream);
    }
    c->ring_slots = (int)std::max<u32>(nsm, (u32)prop.multiProcessorCount);
    c->l2_bytes = (size_t)prop.l2CacheSize;
    c->ring_bytes = (size_t)c->ring_slots * (2 * GSM_GF_RING_MAXR + 1) * gf5_ring_row_floats(GSM_GF_RUNS, GSM_GF_K, GSM_GF_LPR) * sizeof(float);
    A((void**)&c->ring, c->ring_bytes);
  }
#endif
  if (st == cudaSuccess) st = cudaStreamSynchronize(c->stream);  // every buffer is zero before the first call
  if (st != cudaSuccess) {
    int rc = fail(GSM_ERR_CUDA, "gsm_create: allocation failed: %s", cudaGetErrorString(st));
    gsm_destroy(c);
    return rc;
  }
  *out = c;
  return GSM_OK;
}

extern "C" long long gsm_launch_count(const gsm_ctx* c) { return c ? c->launches : 0; }
extern "C" int gsm_set_kernel_timing(gsm_ctx* c, int enabled) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  c->timing = enabled != 0;
  return GSM_OK;
}
extern "C" float gsm_last_kernel_ms(gsm_ctx* c) {
  if (!c || c->ev_used == 0) return -1.f;
  float total = 0.f;
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    if (cudaEventSynchronize(c->ev[i + 1]) != cudaSuccess) return -1.f;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]) != cudaSuccess) return -1.f;
    total += ms;
  }
  return total;
}
extern "C" int gsm_sync(gsm_ctx* c) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->s_h2d));
  CK(cudaStreamSynchronize(c->s_d2h));
  return GSM_OK;
}

// ------------------------------------------------------------------------------------------------
static int check_params(const gsm_ctx* c, const gsm_params* p, int n, int rows, int cols, int* d_begin,
                        int* d_end, float* eps) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  if (!p) return fail(GSM_ERR_INVALID, "null params");
  if (rows < 1 || cols < 1 || n < 1) return fail(GSM_ERR_INVALID, "bad shape n=%d rows=%d cols=%d", n, rows, cols);
  if (rows > c->max_rows || cols > c->max_cols || p->num_disp > c->max_disp ||
      (size_t)rows * cols > (size_t)c->max_rows * c->max_cols)
    return fail(GSM_ERR_CAPACITY, "request %dx%d D=%d exceeds context capacity %dx%d D=%d", rows, cols, p->num_disp,
                c->max_rows, c->max_cols, c->max_disp);
  if (p->num_disp < 1 || p->num_disp > MAX_DISP) return fail(GSM_ERR_INVALID, "num_disp %d not in 1..256", p->num_disp);
  if (p->mode == GSM_MODE_SAD) {
    if (p->radius < 1 || p->radius > 12) return fail(GSM_ERR_INVALID, "SAD radius %d not in 1..12", p->radius);
    if (p->lr_check) return fail(GSM_ERR_INVALID, "lr_check needs mode GF (the reference defines no right-view SAD)");
  } else if (p->mode == GSM_MODE_GF) {
    if (p->radius < 1 || p->radius > 9) return fail(GSM_ERR_INVALID, "GF radius %d not in 1..9", p->radius);
  } else {
    return fail(GSM_ERR_INVALID, "mode %d", p->mode);
  }
  if (p->median_radius < 0 || p->median_radius > MED_MAXR)
    return fail(GSM_ERR_INVALID, "median_radius %d not in 0..%d", p->median_radius, MED_MAXR);
  int b = p->d_begin, e = p->d_end;
  if (b == 0 && e == 0) e = p->num_disp;
  if (b < 0 || e > p->num_disp || b >= e) return fail(GSM_ERR_INVALID, "d range [%d,%d) not inside [0,%d)", b, e, p->num_disp);
  if (p->rectify && (!c->rect_maps || c->rect_rows != rows || c->rect_cols != cols))
    return fail(GSM_ERR_INVALID, "rectify: no maps set for %dx%d (gsm_set_rectification)", rows, cols);
  *d_begin = b;
  *d_end = e;
  *eps = p->eps > 0.f ? p->eps : 6.5025f;
  return GSM_OK;
}

// launch plan of a fused kernel
struct Plan {
  FusedGeom g;
  dim3 grid, block;
  size_t smem;
};

static Plan make_plan(const gsm_params* p, int n, int rows, int cols, int d_begin, int d_end, int view, int K,
                      int runs, int stage_halo, int exch_planes, int HL4, int lpr = WARP) {
  Plan pl;
  const int TWt = runs * K;
  const int hl = strip_left_halo(TWt, stage_halo);
  int TW = (TWt - hl - stage_halo) / 16 * 16;
  pl.g.pg = make_plane_geom(rows, cols, hl);
  pl.g.D = p->num_disp;
  pl.g.d_begin = d_begin;
  pl.g.d_end = d_end;
  pl.g.TW = TW;
  pl.g.hl = hl;
  pl.g.runs = runs;
  int bands = p->row_bands;
  const int strips = (cols + TW - 1) / TW;
  const int dchunks = (d_end - d_begin + lpr - 1) / lpr;
  if (bands <= 0) {
    // automatic: minimise (waves of 148 CTAs) x (rows marched per CTA incl. the band's warm-up rows)
    const int warm = (p->mode == GSM_MODE_SAD ? 2 : 4) * p->radius;
    long long best_cost = -1;
    // GF: bands of at most 768 rows bound the fp32 drift of the stage-2 running sums (measured at 2160 rows in one
    // band: error grows from 2e-5 to 9e-5 top to bottom; with <= 768-row bands it stays below 3e-5)
    const int b_min = p->mode == GSM_MODE_GF ? (rows + 767) / 768 : 1;
    // The choice depends on the FULL disparity count, not on the sub-range [d_begin, d_end) this call evaluates: the
    // fp32 running sums restart at every band, so ranks of a disparity split must all use the band structure of the
    // single-GPU pass for their packed minima to combine into a bit-identical map (dist.py passes an explicit
    // row_bands tuned for the world size to every rank instead).
    const int dchunks_full = (p->num_disp + lpr - 1) / lpr;
    for (int b = b_min; b <= 16 && (rows / b >= 32 || b == b_min); ++b) {
      const long long ctas = (long long)strips * dchunks_full * n * b;
      const long long cost = ((ctas + 147) / 148) * ((rows + b - 1) / b + warm);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; bands = b; }
    }
    if (bands <= 0) bands = 1;
  }
  bands = std::max(1, std::min(bands, rows));
  bands = std::max(1, std::min(bands, 65535 / std::max(1, n)));  // grid.z = n * bands
  pl.g.ft = nullptr;
  pl.g.bands = bands;
  pl.g.band_rows = (rows + bands - 1) / bands;
  pl.g.view = view;
  pl.g.export_ptr = nullptr;
  pl.g.export_d0 = 0;
  pl.g.export_nd = 0;
  pl.grid = dim3(strips, dchunks, n * bands);
  pl.block = dim3(WARP, runs * lpr / WARP, 1);
  pl.smem = (size_t)exch_planes * WARP * exch_pitch_words(runs, K, HL4) * sizeof(u32);
  return pl;
}

static int timing_begin(gsm_ctx* c, cudaStream_t s) {
  if (!c->timing) return GSM_OK;
  while (c->ev.size() < c->ev_used + 2) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    c->ev.push_back(e);
  }
  CK(cudaEventRecord(c->ev[c->ev_used], s));
  return GSM_OK;
}
static int timing_end(gsm_ctx* c, cudaStream_t s) {
  if (!c->timing) return GSM_OK;
  CK(cudaEventRecord(c->ev[c->ev_used + 1], s));
  c->ev_used += 2;
  return GSM_OK;
}

static i64 key_init(const gsm_params* p) {
  if (p->mode == GSM_MODE_SAD) {
    const int w = 2 * p->radius + 1;
    return ((i64)(50 * w * w) << 8);  // BlockMatching.cpp:157-158: min = 50 * dNum, dm = -256 -> (uchar)0
  }
  return (i64)0x7fffffffffffff00LL;  // +inf cost, d = 0
}

#define SAD_CASES(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)

template <bool EXPORT>
static int launch_sad(gsm_ctx* c, const gsm_params* p, int n, int rows, int cols, int d_begin, int d_end, int view,
                      const u8* G, const u8* O, i64* keys, cudaStream_t s, void* export_ptr, int ed0, int end_,
                      const FrameDesc* ft) {
  constexpr int K = 16;
  const int runs = GSM_SAD_RUNS;
  const int R = p->radius;
  const int HL4 = (R + 3) / 4 * 4;
  Plan pl = make_plan(p, n, rows, cols, d_begin, d_end, view, K, runs, R, 2, HL4);
  pl.smem = sad_smem_bytes(runs, K, HL4);
  pl.g.ft = ft;
  pl.g.export_ptr = export_ptr;
  pl.g.export_d0 = ed0;
  pl.g.export_nd = end_;
  switch (R) {
#define X(r)                                                                                                \
  case r: {                                                                                                 \
    auto kfn = sad_wta_kernel<r, K, EXPORT>;                                                                \
    if (!(c->attr_sad[EXPORT] >> r & 1u)) { /* once per context and kernel, not per launch */               \
      CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));             \
      c->attr_sad[EXPORT] |= 1u << r;                                                                       \
    }                                                                                                       \
    kfn<<<pl.grid, pl.block, pl.smem, s>>>(G, O, keys, pl.g);                                               \
    break;                                                                                                  \
  }
    SAD_CASES(X)
#undef X
    default:
      return fail(GSM_ERR_INVALID, "SAD radius %d", R);
  }
  c->launches++;
  CK(cudaGetLastError());
  return GSM_OK;
}

#define GF_KERNEL gf3_wta_kernel
#define GF_CASES(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)

// one radius instantiation of the guided-filter kernel(s)
template <int r, bool EXPORT>
static int launch_gf_r(gsm_ctx* c, const Plan& pl, bool use_ring, const u8* G, const u8* O, const float* stats, i64* keys,
                       cudaStream_t s) {
  constexpr int K = GSM_GF_K, runs = GSM_GF_RUNS, lpr = GSM_GF_LPR;
#if GSM_GF_RING
  if constexpr (r <= GSM_GF_RING_MAXR) {
    if (use_ring) {
      auto kfn = gf5_wta_kernel<r, K, runs, lpr, EXPORT>;
      const size_t smem = gf5_smem_bytes(runs, K, (r + 3) / 4 * 4, lpr);
      if (!(c->attr_gf5[EXPORT] >> r & 1u)) {
        CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->attr_gf5[EXPORT] |= 1u << r;
      }
      kfn<<<pl.grid, pl.block, smem, s>>>(G, O, stats, keys, c->ring, pl.g);
      return GSM_OK;
    }
  }
#endif
  auto kfn = gf3_wta_kernel<r, K, runs, lpr, EXPORT>;
  if (!(c->attr_gf[EXPORT] >> r & 1u)) {  // once per context and kernel, not per launch
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    c->attr_gf[EXPORT] |= 1u << r;
  }
  kfn<<<pl.grid, pl.block, pl.smem, s>>>(G, O, stats, keys, pl.g);
  return GSM_OK;
}

template <bool EXPORT>
static int launch_gf(gsm_ctx* c, const gsm_params* p, int n, int rows, int cols, int d_begin, int d_end, float eps,
                     int view, const u8* G, const u8* O, i64* keys, cudaStream_t s, void* export_ptr, int ed0,
                     int end_, const FrameDesc* ft) {
  constexpr int K = GSM_GF_K;
  // 12 runs of 16 columns x 32 disparities per CTA: a 192-column strip, 160 of them output columns (v3 needs a halo
  // of only r columns).  Measured alternatives at 720p x 128d: 24 runs x 16 disparities 1654 fps, this 1713 fps.
  constexpr int runs = GSM_GF_RUNS, lpr = GSM_GF_LPR;
  const int R = p->radius;
  const int HL4 = (R + 3) / 4 * 4;
  Plan pl = make_plan(p, n, rows, cols, d_begin, d_end, view, K, runs, stage_halo_of(GSM_MODE_GF, R), 6, HL4, lpr);
  pl.smem = gf3_smem_bytes(runs, K, HL4, lpr);
  pl.g.ft = ft;
  pl.g.export_ptr = export_ptr;
  pl.g.export_d0 = ed0;
  pl.g.export_nd = end_;
  const PlaneGeom& pg = pl.g.pg;
  float* stats = c->stats[view];
  // the statistic planes must be zero outside the image: re-zero when the padded geometry changes (a mixed-size
  // batch always does: its frames occupy slots that may have held larger images)
  gsm_ctx::StatGeom& sg = c->stat_geom[view];
  if (ft || sg.rows != rows || sg.cols != cols || sg.xoff != pg.xoff || sg.n < n) {
    CK(cudaMemsetAsync(stats, 0, (size_t)n * GF_STAT_PLANES * pg.plane_stride * sizeof(float), s));
    sg.rows = ft ? -1 : rows;
    sg.cols = cols;
    sg.xoff = pg.xoff;
    sg.n = n;
  }
  {
    static_assert(K == 16, "gf_prepass_kernel assumes one global grid of 16-column runs");
    const int strips = (int)pl.grid.x;
    const int nbx = (cols + R + 1 + pl.g.hl + PP_HALO + PP_TX - 1) / PP_TX;
    // rows per block: 32 when that already gives two blocks per SM, else fewer rows (more, shorter blocks)
    int rpb = PP_ROWS;
    while (rpb > 8 && (long long)nbx * ((rows + rpb - 1) / rpb) * n < 2 * 148) rpb /= 2;
    gf_prepass_kernel<<<dim3(nbx, (rows + rpb - 1) / rpb, n), PP_THREADS, 0, s>>>(
        G, stats, pg, R, eps, pl.g.TW, pl.g.hl, runs, strips, keys, key_init(p), rpb, ft);
    c->launches++;
    CK(cudaGetLastError());
  }
  int rc;
  if (c->fused_wait) {  // the disparity-independent passes above may overlap whatever records this event
    CK(cudaStreamWaitEvent(s, c->fused_wait, 0));
    c->fused_wait = nullptr;
  }
  if ((rc = timing_begin(c, s))) return rc;
  // the ring kernel where the rings of this radius stay in the L2 (or always, in the experiment build)
  bool use_ring = false;
#if GSM_GF_RING
  if (c->ring && R <= GSM_GF_RING_MAXR) {
    const size_t rings = (size_t)c->ring_slots * (2 * R + 1) * gf5_ring_row_floats(runs, K, lpr) * sizeof(float);
    use_ring = GSM_GF_RING == 1 || (double)rings <= 0.55 * (double)c->l2_bytes;
  }
#endif
  switch (R) {
#define X(r)                                                                                       \
  case r: {                                                                                        \
    if ((rc = launch_gf_r<r, EXPORT>(c, pl, use_ring, G, O, stats, keys, s))) return rc;           \
    break;                                                                                         \
  }
    GF_CASES(X)
#undef X
    default:
      return fail(GSM_ERR_INVALID, "GF radius %d", R);
  }
  c->launches++;
  CK(cudaGetLastError());
  if ((rc = timing_end(c, s))) return rc;
  return GSM_OK;
}

static int pack_planes(gsm_ctx* c, const PlaneGeom& pg, int n, const u8* src, u8* dst, int fill, cudaStream_t s,
                       const float* mapx, const float* mapy, const FrameDesc* ft) {
  dim3 block(128);
  dim3 grid((pg.pitch / 4 + 127) / 128, pg.plane_rows, n);
  pack_plane_kernel<<<grid, block, 0, s>>>(src, dst, pg, fill, mapx, mapy, ft);
  c->launches++;
  CK(cudaGetLastError());
  return GSM_OK;
}

static int fill_keys(gsm_ctx* c, i64* keys, size_t npx, i64 v, cudaStream_t s) {
  fill_keys_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(keys, npx, v);
  c->launches++;
  CK(cudaGetLastError());
  return GSM_OK;
}


static int median_launch(gsm_ctx* c, const u8* src, u8* dst, int n, int rows, int cols, int m, cudaStream_t s,
                         const FrameDesc* ft = nullptr) {
  dim3 grid((cols + MED_TX - 1) / MED_TX, (rows + MED_TY - 1) / MED_TY, n);
  median_kernel<<<grid, dim3(MED_TX, MED_TY), 0, s>>>(src, dst, rows, cols, m, ft);
  c->launches++;
  CK(cudaGetLastError());
  return GSM_OK;
}

// Fused aggregation + WTA of one view for a sub-batch that fits the context: tight images -> packed keys.
static int run_view_keys(gsm_ctx* c, const gsm_params* p, int n, int rows, int cols, int d_begin, int d_end,
                         float eps, int view, const u8* Ltight, const u8* Rtight, i64* keys, cudaStream_t s,
                         void* export_ptr = nullptr, int ed0 = 0, int end_ = 0, const FrameDesc* ft = nullptr,
                         size_t npx_total = 0) {
  const size_t npx = ft ? npx_total : (size_t)n * rows * cols;
  const int stage_halo = stage_halo_of(p->mode, p->radius);
  const PlaneGeom pg = make_plane_geom(rows, cols, strip_left_halo(strip_columns(p->mode), stage_halo));
  int rc;
  // guide / other planes.  view 0: guide L, other R (zero pad).  view 1: guide R, other L right-replicated.
  u8* Gp = view == 0 ? c->planeL : c->planeR;
  u8* Op = view == 0 ? c->planeR : c->planeLrep;
  // with p->rectify the planes receive remap(raw frame): left frames through the left maps, right through the right
  const size_t mpx = (size_t)rows * cols;
  const float* mL = p->rectify ? c->rect_maps : nullptr;
  const float* mR = p->rectify ? c->rect_maps + 2 * mpx : nullptr;
  if (view == 0) {
    if ((rc = pack_planes(c, pg, n, Ltight, Gp, 0, s, mL, mL ? mL + mpx : nullptr, ft))) return rc;
    if ((rc = pack_planes(c, pg, n, Rtight, Op, 0, s, mR, mR ? mR + mpx : nullptr, ft))) return rc;
  } else {
    if ((rc = pack_planes(c, pg, n, Rtight, Gp, 0, s, mR, mR ? mR + mpx : nullptr, ft))) return rc;
    if ((rc = pack_planes(c, pg, n, Ltight, Op, 1, s, mL, mL ? mL + mpx : nullptr, ft))) return rc;
  }
  // GF: the guide pre-pass also initialises the packed-min plane (it visits every pixel anyway)
  if (p->mode == GSM_MODE_SAD && (rc = fill_keys(c, keys, npx, key_init(p), s))) return rc;
  if (p->mode == GSM_MODE_SAD) {
    if (c->fused_wait) {
      CK(cudaStreamWaitEvent(s, c->fused_wait, 0));
      c->fused_wait = nullptr;
    }
    if ((rc = timing_begin(c, s))) return rc;
    if (export_ptr)
      rc = launch_sad<true>(c, p, n, rows, cols, d_begin, d_end, view, Gp, Op, keys, s, export_ptr, ed0, end_, ft);
    else
      rc = launch_sad<false>(c, p, n, rows, cols, d_begin, d_end, view, Gp, Op, keys, s, nullptr, 0, 0, ft);
    if (rc) return rc;
    if ((rc = timing_end(c, s))) return rc;
  } else {
    if (export_ptr)
      rc = launch_gf<true>(c, p, n, rows, cols, d_begin, d_end, eps, view, Gp, Op, keys, s, export_ptr, ed0, end_, ft);
    else
      rc = launch_gf<false>(c, p, n, rows, cols, d_begin, d_end, eps, view, Gp, Op, keys, s, nullptr, 0, 0, ft);
    if (rc) return rc;
  }
  return GSM_OK;
}

// u8 maps (left [, right]) -> final disparity [, mask]: median on both views, then the LR check
// (STMatching/StereoDisparity.cpp:119,126,128-147).  dl / dr may be scratch maps of the context or caller maps.
static int post_maps(gsm_ctx* c, const gsm_params* p, int n, int rows, int cols, const u8* dl, const u8* dr,
                     u8* disp_out, u8* mask_out, cudaStream_t s, const FrameDesc* ft, size_t npx) {
  int rc;
  const bool lr = p->lr_check && dr;
  const int m = p->median_radius;
  if (m > 0) {
    u8* dst = lr ? c->dispB : disp_out;
    if ((rc = median_launch(c, dl, dst, n, rows, cols, m, s, ft))) return rc;
    dl = dst;
  }
  if (lr) {
    if (m > 0) {
      if ((rc = median_launch(c, dr, c->dispD, n, rows, cols, m, s, ft))) return rc;
      dr = c->dispD;
    }
    dim3 grid((cols + 255) / 256, rows, n);
    lr_check_kernel<<<grid, 256, 0, s>>>(dl, dr, nullptr, mask_out, disp_out, rows, cols, n, ft);
    c->launches++;
  } else if (dl != disp_out) {
    CK(cudaMemcpyAsync(disp_out, dl, npx, cudaMemcpyDeviceToDevice, s));
  }
  CK(cudaGetLastError());
  return GSM_OK;
}

// keys (left [, right]) -> disparity [, mask]; order follows STMatching/StereoDisparity.cpp:115-147:
// WTA -> median on both views -> LR check.
static int finalize_views(gsm_ctx* c, const gsm_params* p, int n, int rows, int cols, const i64* kL, const i64* kR,
                          u8* disp_out, u8* mask_out, cudaStream_t s, const FrameDesc* ft = nullptr,
                          size_t npx_total = 0) {
  const size_t npx = ft ? npx_total : (size_t)n * rows * cols;
  const unsigned gb = (unsigned)((npx + 255) / 256);
  const bool lr = p->lr_check && kR;
  u8* dl = (p->median_radius > 0 || lr) ? c->dispA : disp_out;
  finalize_keys_kernel<<<gb, 256, 0, s>>>(kL, dl, npx);
  c->launches++;
  if (lr) {
    finalize_keys_kernel<<<gb, 256, 0, s>>>(kR, c->dispC, npx);
    c->launches++;
  }
  CK(cudaGetLastError());
  return post_maps(c, p, n, rows, cols, dl, lr ? c->dispC : nullptr, disp_out, mask_out, s, ft, npx);
}

extern "C" int gsm_stereo_device(gsm_ctx* c, const gsm_params* p, int n, const void* left_dev, const void* right_dev,
                                 void* disparity_dev, void* mask_dev, int rows, int cols, void* stream) {
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, n, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!left_dev || !right_dev || !disparity_dev) return fail(GSM_ERR_INVALID, "null image pointer");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  c->ev_used = 0;
  const size_t fpx = (size_t)rows * cols;
  for (int f0 = 0; f0 < n; f0 += c->max_batch) {
    const int nb = std::min(c->max_batch, n - f0);
    const u8* L = (const u8*)left_dev + f0 * fpx;
    const u8* R = (const u8*)right_dev + f0 * fpx;
    if ((rc = run_view_keys(c, p, nb, rows, cols, d_begin, d_end, eps, 0, L, R, c->keysL, s))) return rc;
    if (p->lr_check)
      if ((rc = run_view_keys(c, p, nb, rows, cols, d_begin, d_end, eps, 1, L, R, c->keysR, s))) return rc;
    if ((rc = finalize_views(c, p, nb, rows, cols, c->keysL, p->lr_check ? c->keysR : nullptr,
                             (u8*)disparity_dev + f0 * fpx, mask_dev ? (u8*)mask_dev + f0 * fpx : nullptr, s)))
      return rc;
  }
  return GSM_OK;
}

extern "C" int gsm_stereo_batch_async(gsm_ctx* c, const gsm_params* p, int n, const uint8_t* left,
                                      const uint8_t* right, uint8_t* disparity, uint8_t* mask, int rows, int cols) {
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, n, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!left || !right || !disparity) return fail(GSM_ERR_INVALID, "null image pointer");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const size_t fpx = (size_t)rows * cols;
  const size_t slot_stride = (size_t)c->max_rows * c->max_cols * c->slot_frames;
  const bool want_mask = mask && p->lr_check;
  c->ev_used = 0;
  // Two slots of input / output buffers: while the kernels of chunk i run on the compute stream, chunk i+1 is
  // uploaded and chunk i-1 downloaded on their own streams (asynchronous when the host buffers are pinned).
  // Chunks of half the batch capacity.  Smaller chunks (quarter size, or a half-size first chunk) shorten the
  // exposed first upload / last download but were measured slower end to end (1455 / 1518 vs 1587 fps at 32 frames
  // of 720p): their launches fill the 148 SMs worse.
  for (int f0 = 0; f0 < n; ++c->chunk_seq) {
    const long long chunk = c->chunk_seq;  // persists across calls: back-to-back submissions keep pipelining
    const int step = c->slot_frames;
    const int nb = std::min(step, n - f0);
    const int slot = (int)(chunk & 1);
    u8* tl = c->tightL + slot * slot_stride;
    u8* tr = c->tightR + slot * slot_stride;
    u8* dres = c->dispOut + slot * slot_stride;
    u8* mres = c->maskD + slot * slot_stride;
    if (chunk >= 2) CK(cudaStreamWaitEvent(c->s_h2d, c->ev_in_free[slot], 0));
    CK(cudaMemcpyAsync(tl, left + f0 * fpx, nb * fpx, cudaMemcpyHostToDevice, c->s_h2d));
    CK(cudaMemcpyAsync(tr, right + f0 * fpx, nb * fpx, cudaMemcpyHostToDevice, c->s_h2d));
    CK(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
    CK(cudaStreamWaitEvent(s, c->ev_h2d[slot], 0));
    if ((rc = run_view_keys(c, p, nb, rows, cols, d_begin, d_end, eps, 0, tl, tr, c->keysL, s))) return rc;
    if (p->lr_check)
      if ((rc = run_view_keys(c, p, nb, rows, cols, d_begin, d_end, eps, 1, tl, tr, c->keysR, s))) return rc;
    CK(cudaEventRecord(c->ev_in_free[slot], s));
    if (chunk >= 2) CK(cudaStreamWaitEvent(s, c->ev_d2h[slot], 0));
    if ((rc = finalize_views(c, p, nb, rows, cols, c->keysL, p->lr_check ? c->keysR : nullptr, dres,
                             want_mask ? mres : nullptr, s)))
      return rc;
    CK(cudaEventRecord(c->ev_done[slot], s));
    CK(cudaStreamWaitEvent(c->s_d2h, c->ev_done[slot], 0));
    CK(cudaMemcpyAsync(disparity + f0 * fpx, dres, nb * fpx, cudaMemcpyDeviceToHost, c->s_d2h));
    if (want_mask) CK(cudaMemcpyAsync(mask + f0 * fpx, mres, nb * fpx, cudaMemcpyDeviceToHost, c->s_d2h));
    CK(cudaEventRecord(c->ev_d2h[slot], c->s_d2h));
    f0 += nb;
  }
  return GSM_OK;
}

extern "C" int gsm_stereo_batch(gsm_ctx* c, const gsm_params* p, int n, const uint8_t* left, const uint8_t* right,
                                uint8_t* disparity, uint8_t* mask, int rows, int cols) {
  const int rc = gsm_stereo_batch_async(c, p, n, left, right, disparity, mask, rows, cols);
  if (rc) return rc;
  return gsm_sync(c);
}

// Mixed-size batch (BASELINE config 2: the nine Middlebury sets come in three sizes): every frame has its own rows x
// cols and its own host buffers, like the separate Mats of Caller.cpp:12-19; the whole batch runs as ONE launch per
// stage over a per-frame geometry table (FrameDesc).  Blocking.
extern "C" int gsm_stereo_batch_v(gsm_ctx* c, const gsm_params* p, int n, const uint8_t* const* left,
                                  const uint8_t* const* right, uint8_t* const* disparity, uint8_t* const* mask,
                                  const int* rows, const int* cols) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  if (!p || !left || !right || !disparity || !rows || !cols || n < 1)
    return fail(GSM_ERR_INVALID, "gsm_stereo_batch_v: null argument or n=%d", n);
  if (p->rectify) return fail(GSM_ERR_INVALID, "gsm_stereo_batch_v: rectify needs frames of the maps' size (gsm_stereo_batch)");
  int max_r = 0, max_c = 0;
  for (int i = 0; i < n; ++i) {
    if (!left[i] || !right[i] || !disparity[i]) return fail(GSM_ERR_INVALID, "gsm_stereo_batch_v: null image pointer (frame %d)", i);
    if (rows[i] < 1 || cols[i] < 1) return fail(GSM_ERR_INVALID, "gsm_stereo_batch_v: frame %d is %dx%d", i, rows[i], cols[i]);
    max_r = std::max(max_r, rows[i]);
    max_c = std::max(max_c, cols[i]);
  }
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, 1, max_r, max_c, &d_begin, &d_end, &eps))) return rc;
  CK(cudaSetDevice(c->device));
  if ((rc = gsm_sync(c))) return rc;  // drain any streaming batches still in flight
  cudaStream_t s = c->stream;
  const bool want_mask = mask && p->lr_check;
  c->ev_used = 0;
  std::vector<FrameDesc> ft;
  for (int f0 = 0; f0 < n; f0 += c->max_batch) {
    const int nb = std::min(c->max_batch, n - f0);
    // geometry of THIS sub-batch: plane slots are laid out for its largest frame
    int br = 0, bc = 0;
    ft.assign(nb, FrameDesc());
    long long off = 0;
    for (int i = 0; i < nb; ++i) {
      ft[i].H = rows[f0 + i];
      ft[i].W = cols[f0 + i];
      ft[i].off = off;
      off += (long long)ft[i].H * ft[i].W;
      br = std::max(br, ft[i].H);
      bc = std::max(bc, ft[i].W);
    }
    const size_t npx = (size_t)off;
    CK(cudaMemcpyAsync(c->ft_dev, ft.data(), nb * sizeof(FrameDesc), cudaMemcpyHostToDevice, s));
    for (int i = 0; i < nb; ++i) {
      const size_t bytes = (size_t)ft[i].H * ft[i].W;
      CK(cudaMemcpyAsync(c->tightL + ft[i].off, left[f0 + i], bytes, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(c->tightR + ft[i].off, right[f0 + i], bytes, cudaMemcpyHostToDevice, s));
    }
    if ((rc = run_view_keys(c, p, nb, br, bc, d_begin, d_end, eps, 0, c->tightL, c->tightR, c->keysL, s, nullptr, 0, 0,
                            c->ft_dev, npx)))
      return rc;
    if (p->lr_check)
      if ((rc = run_view_keys(c, p, nb, br, bc, d_begin, d_end, eps, 1, c->tightL, c->tightR, c->keysR, s, nullptr, 0,
                              0, c->ft_dev, npx)))
        return rc;
    if ((rc = finalize_views(c, p, nb, br, bc, c->keysL, p->lr_check ? c->keysR : nullptr, c->dispOut,
                             want_mask ? c->maskD : nullptr, s, c->ft_dev, npx)))
      return rc;
    for (int i = 0; i < nb; ++i) {
      const size_t bytes = (size_t)ft[i].H * ft[i].W;
      CK(cudaMemcpyAsync(disparity[f0 + i], c->dispOut + ft[i].off, bytes, cudaMemcpyDeviceToHost, s));
      if (want_mask && mask[f0 + i])
        CK(cudaMemcpyAsync(mask[f0 + i], c->maskD + ft[i].off, bytes, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));  // the table and the slots are reused by the next sub-batch
  }
  return GSM_OK;
}

// Same on DEVICE buffers: frame i is rows[i] x cols[i] at pixel offset sum_{j<i} rows[j]*cols[j] of left_dev /
// right_dev / disparity_dev / mask_dev (tight, concatenated).  n <= the context's batch capacity.  Asynchronous on
// `stream`; rows / cols are host arrays and are consumed before the call returns.
extern "C" int gsm_stereo_device_v(gsm_ctx* c, const gsm_params* p, int n, const void* left_dev, const void* right_dev,
                                   void* disparity_dev, void* mask_dev, const int* rows, const int* cols, void* stream) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  if (!p || !left_dev || !right_dev || !disparity_dev || !rows || !cols || n < 1)
    return fail(GSM_ERR_INVALID, "gsm_stereo_device_v: null argument or n=%d", n);
  if (n > c->max_batch) return fail(GSM_ERR_CAPACITY, "gsm_stereo_device_v: %d frames exceed the batch capacity %d", n, c->max_batch);
  if (p->rectify) return fail(GSM_ERR_INVALID, "gsm_stereo_device_v: rectify needs frames of the maps' size");
  std::vector<FrameDesc> ft(n);
  int br = 0, bc = 0;
  long long off = 0;
  for (int i = 0; i < n; ++i) {
    if (rows[i] < 1 || cols[i] < 1) return fail(GSM_ERR_INVALID, "gsm_stereo_device_v: frame %d is %dx%d", i, rows[i], cols[i]);
    ft[i].H = rows[i];
    ft[i].W = cols[i];
    ft[i].off = off;
    off += (long long)rows[i] * cols[i];
    br = std::max(br, rows[i]);
    bc = std::max(bc, cols[i]);
  }
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, 1, br, bc, &d_begin, &d_end, &eps))) return rc;
  CK(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  c->ev_used = 0;
  const size_t npx = (size_t)off;
  CK(cudaMemcpyAsync(c->ft_dev, ft.data(), n * sizeof(FrameDesc), cudaMemcpyHostToDevice, s));  // pageable: staged now
  if ((rc = run_view_keys(c, p, n, br, bc, d_begin, d_end, eps, 0, (const u8*)left_dev, (const u8*)right_dev, c->keysL, s,
                          nullptr, 0, 0, c->ft_dev, npx)))
    return rc;
  if (p->lr_check)
    if ((rc = run_view_keys(c, p, n, br, bc, d_begin, d_end, eps, 1, (const u8*)left_dev, (const u8*)right_dev, c->keysR,
                            s, nullptr, 0, 0, c->ft_dev, npx)))
      return rc;
  return finalize_views(c, p, n, br, bc, c->keysL, p->lr_check ? c->keysR : nullptr, (u8*)disparity_dev, (u8*)mask_dev, s,
                        c->ft_dev, npx);
}

extern "C" int gsm_block_matching(gsm_ctx* c, const uint8_t* left, const uint8_t* right, uint8_t* disparity, int rows,
                                  int cols, int radius, int num_disp) {
  gsm_params p;
  memset(&p, 0, sizeof(p));
  p.mode = GSM_MODE_SAD;
  p.radius = radius;
  p.num_disp = num_disp;
  return gsm_stereo_batch(c, &p, 1, left, right, disparity, nullptr, rows, cols);
}

extern "C" int gsm_partial_keys_device(gsm_ctx* c, const gsm_params* p, int view, const void* left_dev,
                                       const void* right_dev, void* keys_dev, int rows, int cols, void* stream) {
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, 1, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!left_dev || !right_dev || !keys_dev) return fail(GSM_ERR_INVALID, "null pointer");
  if (view != 0 && view != 1) return fail(GSM_ERR_INVALID, "view %d", view);
  if (view == 1 && p->mode != GSM_MODE_GF) return fail(GSM_ERR_INVALID, "right view needs mode GF");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  c->ev_used = 0;
  return run_view_keys(c, p, 1, rows, cols, d_begin, d_end, eps, view, (const u8*)left_dev, (const u8*)right_dev,
                       (i64*)keys_dev, s);
}

// Like gsm_partial_keys_device, but the fused aggregation+WTA kernel additionally waits for `wait_event` (a
// cudaEvent_t recorded on another stream; NULL = none).  The disparity-independent passes in front of it (plane
// packing, guide statistics) do not wait: they overlap whatever the event marks the end of -- dist.DsplitStream uses it
// to run them beside the previous frame's peer-memory combine, which cannot share SMs with the fused kernel (it
// occupies every register of an SM).
extern "C" int gsm_partial_keys_device_ex(gsm_ctx* c, const gsm_params* p, int view, const void* left_dev,
                                          const void* right_dev, void* keys_dev, int rows, int cols, void* stream,
                                          void* wait_event) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  c->fused_wait = (cudaEvent_t)wait_event;
  const int rc = gsm_partial_keys_device(c, p, view, left_dev, right_dev, keys_dev, rows, cols, stream);
  c->fused_wait = nullptr;
  return rc;
}

extern "C" int gsm_finalize_keys_device(gsm_ctx* c, const gsm_params* p, const void* keys_left_dev,
                                        const void* keys_right_dev, void* disparity_dev, void* mask_dev, int rows,
                                        int cols, void* stream) {
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, 1, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!keys_left_dev || !disparity_dev) return fail(GSM_ERR_INVALID, "null pointer");
  if (p->lr_check && !keys_right_dev) return fail(GSM_ERR_INVALID, "lr_check needs right-view keys");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  return finalize_views(c, p, 1, rows, cols, (const i64*)keys_left_dev, (const i64*)keys_right_dev, (u8*)disparity_dev,
                        (u8*)mask_dev, s);
}

// Post-filters on DEVICE maps: median on both views, then the LR check (StereoDisparity.cpp:119,126,128-147) -- what
// gsm_finalize_keys_device does after extracting the disparities, for callers that already hold u8 maps (the
// peer-memory disparity split, whose combine kernel writes the raw WTA maps of both views into every rank).
extern "C" int gsm_postfilter_device(gsm_ctx* c, const gsm_params* p, const void* disp_left_dev,
                                     const void* disp_right_dev, void* disparity_dev, void* mask_dev, int rows, int cols,
                                     void* stream) {
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, 1, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!disp_left_dev || !disparity_dev) return fail(GSM_ERR_INVALID, "null pointer");
  if (p->lr_check && !disp_right_dev) return fail(GSM_ERR_INVALID, "lr_check needs the right-view map");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  return post_maps(c, p, 1, rows, cols, (const u8*)disp_left_dev, p->lr_check ? (const u8*)disp_right_dev : nullptr,
                   (u8*)disparity_dev, (u8*)mask_dev, s, nullptr, (size_t)rows * cols);
}

static int reduce_keys_p2p(gsm_ctx* c, const void* const* key_ptrs, void* const* disp_ptrs, int world, int rank,
                           long long npx, void* stream, int max_blocks) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  if (!key_ptrs || !disp_ptrs || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || npx < 1)
    return fail(GSM_ERR_INVALID, "gsm_reduce_keys_p2p: world=%d rank=%d npx=%lld", world, rank, npx);
  PeerPlanes pp;
  memset(&pp, 0, sizeof(pp));
  for (int w = 0; w < world; ++w) {
    if (!key_ptrs[w] || !disp_ptrs[w]) return fail(GSM_ERR_INVALID, "gsm_reduce_keys_p2p: null plane of rank %d", w);
    if (((uintptr_t)key_ptrs[w] | (uintptr_t)disp_ptrs[w]) & 15)
      return fail(GSM_ERR_INVALID, "gsm_reduce_keys_p2p: planes must be 16-byte aligned");
    pp.keys[w] = (const i64*)key_ptrs[w];
    pp.disp[w] = (u8*)disp_ptrs[w];
  }
  // slice boundaries on multiples of 16 pixels (128-byte runs of keys, 16-byte stores of disparities)
  auto cut = [&](int r) { return r >= world ? (size_t)npx : (size_t)((long long)npx * r / world) / 16 * 16; };
  const size_t begin = cut(rank), end = cut(rank + 1);
  CK(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (end > begin) {
    const size_t chunks = (end - begin + 511) / 512;  // one warp per 512-pixel chunk
    if (max_blocks > 0) {
      // confined to a few SMs (one 16-warp block each): runs BESIDE a fused kernel that leaves that many SMs idle
      reduce_keys_p2p_kernel<16><<<(unsigned)std::max(1, std::min<int>(max_blocks, (int)((chunks + 15) / 16))), 512, 0, s>>>(
          pp, world, rank, begin, end);
    } else {
      const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((chunks + 7) / 8, 148 * 8));
      reduce_keys_p2p_kernel<8><<<blocks, 256, 0, s>>>(pp, world, rank, begin, end);
    }
    c->launches++;
    CK(cudaGetLastError());
  }
  return GSM_OK;
}

extern "C" int gsm_reduce_keys_p2p(gsm_ctx* c, const void* const* key_ptrs, void* const* disp_ptrs, int world,
                                   int rank, long long npx, void* stream) {
  return reduce_keys_p2p(c, key_ptrs, disp_ptrs, world, rank, npx, stream, 0);
}

// Same, confined to at most max_blocks thread blocks of 512 threads (one per SM).  The fused aggregation kernel owns
// every register of the SMs it runs on; when its grid leaves a few SMs idle (config 5 at 8 ranks: 24 strips x 1 chunk
// x 6 bands = 144 CTAs on 148 SMs) the combine of the previous frame fits on those and runs beside it -- it is bound by
// NVLink latency x bytes in flight, not by SMs.
extern "C" int gsm_reduce_keys_p2p_ex(gsm_ctx* c, const void* const* key_ptrs, void* const* disp_ptrs, int world,
                                      int rank, long long npx, void* stream, int max_blocks) {
  return reduce_keys_p2p(c, key_ptrs, disp_ptrs, world, rank, npx, stream, max_blocks);
}

// ---- cost-stage exports ------------------------------------------------------------------------
static int ensure_export(gsm_ctx* c, size_t bytes) {
  if (c->export_bytes >= bytes) return GSM_OK;
  if (c->export_buf) cudaFree(c->export_buf);
  c->export_buf = nullptr;
  c->export_bytes = 0;
  CK(cudaMalloc(&c->export_buf, bytes));
  c->export_bytes = bytes;
  return GSM_OK;
}

extern "C" int gsm_ad_volume(gsm_ctx* c, const uint8_t* left, const uint8_t* right, uint8_t* volume, int rows,
                             int cols, int num_disp) {
  gsm_params p;
  memset(&p, 0, sizeof(p));
  p.mode = GSM_MODE_SAD;
  p.radius = 1;
  p.num_disp = num_disp;
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, &p, 1, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!left || !right || !volume) return fail(GSM_ERR_INVALID, "null pointer");
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t fpx = (size_t)rows * cols;
  if ((rc = ensure_export(c, fpx * num_disp))) return rc;
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(c->tightL, left, fpx, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->tightR, right, fpx, cudaMemcpyHostToDevice, s));
  ad_volume_kernel<<<dim3((cols + 255) / 256, rows, num_disp), 256, 0, s>>>(c->tightL, c->tightR, (u8*)c->export_buf,
                                                                              rows, cols);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(volume, c->export_buf, fpx * num_disp, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_cost_slices(gsm_ctx* c, const gsm_params* p, int view, const uint8_t* left, const uint8_t* right,
                               int d0, int nd, void* out, int rows, int cols) {
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, p, 1, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!left || !right || !out) return fail(GSM_ERR_INVALID, "null pointer");
  if (d0 < 0 || nd < 1 || d0 + nd > p->num_disp) return fail(GSM_ERR_INVALID, "slice range [%d,%d)", d0, d0 + nd);
  if (view != 0 && view != 1) return fail(GSM_ERR_INVALID, "view %d", view);
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t fpx = (size_t)rows * cols;
  const size_t bytes = fpx * nd * 4;
  if ((rc = ensure_export(c, bytes))) return rc;
  cudaStream_t s = c->stream;
  CK(cudaMemsetAsync(c->export_buf, 0, bytes, s));
  CK(cudaMemcpyAsync(c->tightL, left, fpx, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->tightR, right, fpx, cudaMemcpyHostToDevice, s));
  // evaluate only the 32-disparity chunks that cover [d0, d0+nd)
  const int cb = d0 / WARP * WARP;
  const int ce = std::min(p->num_disp, round_up(d0 + nd, WARP));
  if ((rc = run_view_keys(c, p, 1, rows, cols, cb, ce, eps, view, c->tightL, c->tightR, c->keysL, s, c->export_buf, d0, nd)))
    return rc;
  CK(cudaMemcpyAsync(out, c->export_buf, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_all_sad(gsm_ctx* c, const uint8_t* left, const uint8_t* right, uint8_t* out, int rows, int cols,
                           int radius, int num_disp) {
  gsm_params p;
  memset(&p, 0, sizeof(p));
  p.mode = GSM_MODE_SAD;
  p.radius = radius;
  p.num_disp = num_disp;
  int d_begin, d_end, rc;
  float eps;
  if ((rc = check_params(c, &p, 1, rows, cols, &d_begin, &d_end, &eps))) return rc;
  if (!left || !right || !out) return fail(GSM_ERR_INVALID, "null pointer");
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t fpx = (size_t)rows * cols;
  const size_t slice_bytes = fpx * num_disp * 4;
  if ((rc = ensure_export(c, slice_bytes + fpx * num_disp))) return rc;
  cudaStream_t s = c->stream;
  u8* packed = (u8*)c->export_buf + slice_bytes;
  CK(cudaMemsetAsync(c->export_buf, 0, slice_bytes, s));
  CK(cudaMemcpyAsync(c->tightL, left, fpx, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->tightR, right, fpx, cudaMemcpyHostToDevice, s));
  if ((rc = run_view_keys(c, &p, 1, rows, cols, 0, num_disp, eps, 0, c->tightL, c->tightR, c->keysL, s, c->export_buf, 0,
                          num_disp)))
    return rc;
  all_sad_pack_kernel<<<dim3((cols + 255) / 256, rows, num_disp), 256, 0, s>>>((const int*)c->export_buf, packed, rows,
                                                                                 cols, num_disp, 0, num_disp);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, packed, fpx * num_disp, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

// ---- post-filters as stand-alone calls -----------------------------------------------------------
extern "C" int gsm_median(gsm_ctx* c, const uint8_t* src, uint8_t* dst, int rows, int cols, int radius) {
  if (!c || !src || !dst) return fail(GSM_ERR_INVALID, "null pointer");
  if (rows < 1 || cols < 1 || rows > c->max_rows || cols > c->max_cols) return fail(GSM_ERR_CAPACITY, "shape %dx%d", rows, cols);
  if (radius < 0 || radius > MED_MAXR) return fail(GSM_ERR_INVALID, "median radius %d", radius);
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t fpx = (size_t)rows * cols;
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(c->dispA, src, fpx, cudaMemcpyHostToDevice, s));
  int rc;
  if ((rc = median_launch(c, c->dispA, c->dispB, 1, rows, cols, radius, s))) return rc;
  CK(cudaMemcpyAsync(dst, c->dispB, fpx, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_lr_check(gsm_ctx* c, const uint8_t* dl, const uint8_t* dr, uint8_t* occ, uint8_t* mask, int rows,
                            int cols) {
  if (!c || !dl || !dr) return fail(GSM_ERR_INVALID, "null pointer");
  if (rows < 1 || cols < 1 || rows > c->max_rows || cols > c->max_cols) return fail(GSM_ERR_CAPACITY, "shape %dx%d", rows, cols);
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t fpx = (size_t)rows * cols;
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(c->dispA, dl, fpx, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->dispB, dr, fpx, cudaMemcpyHostToDevice, s));
  lr_check_kernel<<<dim3((cols + 255) / 256, rows, 1), 256, 0, s>>>(c->dispA, c->dispB, c->dispC, c->maskD, nullptr, rows,
                                                                      cols, 1, nullptr);
  c->launches++;
  CK(cudaGetLastError());
  if (occ) CK(cudaMemcpyAsync(occ, c->dispC, fpx, cudaMemcpyDeviceToHost, s));
  if (mask) CK(cudaMemcpyAsync(mask, c->maskD, fpx, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_remap(gsm_ctx* c, const uint8_t* src, const float* mapx, const float* mapy, uint8_t* dst, int rows,
                         int cols) {
  if (!c || !src || !mapx || !mapy || !dst) return fail(GSM_ERR_INVALID, "null pointer");
  if (rows < 1 || cols < 1) return fail(GSM_ERR_INVALID, "bad shape %dx%d", rows, cols);
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t n = (size_t)rows * cols;
  int rc;
  if ((rc = ensure_export(c, n * 10))) return rc;  // [mapx f32][mapy f32][src u8][dst u8]
  float* dmx = (float*)c->export_buf;
  float* dmy = dmx + n;
  u8* dsrc = (u8*)(dmy + n);
  u8* ddst = dsrc + n;
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(dmx, mapx, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dmy, mapy, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dsrc, src, n, cudaMemcpyHostToDevice, s));
  remap_kernel<<<dim3((cols + 255) / 256, rows), 256, 0, s>>>(dsrc, dmx, dmy, ddst, rows, cols);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, ddst, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_cvtcolor(gsm_ctx* c, const uint8_t* src3, uint8_t* dst, int rows, int cols, int truncate) {
  if (!c || !src3 || !dst) return fail(GSM_ERR_INVALID, "null pointer");
  if (rows < 1 || cols < 1) return fail(GSM_ERR_INVALID, "bad shape %dx%d", rows, cols);
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t n = (size_t)rows * cols;
  int rc;
  if ((rc = ensure_export(c, n * 4))) return rc;
  u8* dsrc = (u8*)c->export_buf;
  u8* ddst = dsrc + 3 * n;
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(dsrc, src3, 3 * n, cudaMemcpyHostToDevice, s));
  cvtcolor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dsrc, ddst, n, truncate);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, ddst, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_disparity_to_depth(gsm_ctx* c, const uint8_t* disparity, float* depth, int rows, int cols, float fB) {
  if (!c || !disparity || !depth) return fail(GSM_ERR_INVALID, "null pointer");
  if (rows < 1 || cols < 1) return fail(GSM_ERR_INVALID, "bad shape %dx%d", rows, cols);
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;  // drain any streaming batches still in flight
  const size_t n = (size_t)rows * cols;
  int rc;
  if ((rc = ensure_export(c, n * 5))) return rc;  // [depth f32][disp u8]
  float* dd = (float*)c->export_buf;
  u8* ds = (u8*)(dd + n);
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(ds, disparity, n, cudaMemcpyHostToDevice, s));
  depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ds, dd, n, fB);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(depth, dd, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_set_rectification(gsm_ctx* c, const float* mxl, const float* myl, const float* mxr, const float* myr,
                                     int rows, int cols) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  if (int rc_sync = gsm_sync(c)) return rc_sync;
  if (c->rect_maps) {
    cudaFree(c->rect_maps);
    c->rect_maps = nullptr;
    c->rect_rows = c->rect_cols = 0;
  }
  if (!mxl && !myl && !mxr && !myr) return GSM_OK;
  if (!mxl || !myl || !mxr || !myr) return fail(GSM_ERR_INVALID, "all four maps or none");
  if (rows < 1 || cols < 1 || rows > c->max_rows || cols > c->max_cols) return fail(GSM_ERR_CAPACITY, "maps %dx%d", rows, cols);
  const size_t n = (size_t)rows * cols;
  CK(cudaMalloc((void**)&c->rect_maps, 4 * n * sizeof(float)));
  const float* srcs[4] = {mxl, myl, mxr, myr};
  for (int i = 0; i < 4; ++i)
    CK(cudaMemcpyAsync(c->rect_maps + i * n, srcs[i], n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // the host maps may be freed by the caller; kernels run on other streams too
  c->rect_rows = rows;
  c->rect_cols = cols;
  return GSM_OK;
}


// ---- segment-tree stereo (kernels in gsm_st.cuh, tree in gsm_st_host.hpp) --------------------------
namespace {
// The ordered tree as the host's breadth-first pass leaves it, one contiguous block with the same layout in the pinned
// host slot and on the device: ONE copy, then st_pack_kernel forms the words the filter reads.
struct StTreeBlock {
  size_t o_order, o_level, o_father, o_child0, o_fdist, o_nchild, o_wtab, bytes;
  explicit StTreeBlock(size_t n) {
    size_t o = 0;
    auto take = [&](size_t b) { const size_t at = o; o += (b + 255) / 256 * 256; return at; };
    o_order = take(4 * n); o_level = take(4 * (n + 2)); o_father = take(4 * n); o_child0 = take(4 * n);
    o_fdist = take(n); o_nchild = take(n); o_wtab = take(4 * 256);
    bytes = o;
  }
};
struct StArena {  // carve-up of gsm_ctx::st_buf for one rows x cols x D problem
  u8 *L3, *R3, *med, *wr, *wu, *disp, *disp2, *dispL, *dispR, *mask;
  float *gL, *gR, *buf, *fin, *vol, *wrf, *wuf;
  char* tree;  // StTreeBlock
  int *order, *level_off, *pos;
  int2 *up, *down;
  size_t bytes;
  StArena(void* base, size_t n, int D, bool with_vol) {
    size_t o = 0;
    auto take = [&](size_t b) { void* p = base ? (char*)base + o : nullptr; o += (b + 255) / 256 * 256; return p; };
    L3 = (u8*)take(3 * n); R3 = (u8*)take(3 * n); med = (u8*)take(3 * n);
    wr = (u8*)take(n); wu = (u8*)take(n); disp = (u8*)take(n); disp2 = (u8*)take(n);
    dispL = (u8*)take(n); dispR = (u8*)take(n); mask = (u8*)take(n);
    gL = (float*)take(4 * n); gR = (float*)take(4 * n);
    wrf = (float*)take(4 * n); wuf = (float*)take(4 * n);
    const StTreeBlock tb(n);
    tree = (char*)take(tb.bytes);
    order = (int*)(tree + tb.o_order); level_off = (int*)(tree + tb.o_level);
    up = (int2*)take(8 * n); down = (int2*)take(8 * n); pos = (int*)take(4 * n);
    buf = (float*)take(4 * n * D); fin = (float*)take(4 * n * D);
    vol = with_vol ? (float*)take(4 * n * D) : nullptr;
    bytes = o;
  }
};
// Pinned slot k of gsm_ctx::st_pin: [StTreeBlock mirror][weight read-back: two planes of n floats (or of n bytes)]
// [sorted edge codes, <= 2n words][sorted edge weights, <= 2n floats][kept-edge flags, n bytes][per-pixel records, 8 n bytes]
// [batches: the frame's two images, 3n bytes each]
struct StPinSlot {
  char *tree, *w, *code, *ws, *flags, *rec, *img;
  static size_t stride(size_t n) {
    return (StTreeBlock(n).bytes + 8 * n * 3 + (n + 255) / 256 * 256 + 8 * n + 2 * ((3 * n + 255) / 256 * 256) + 255) / 256 * 256;
  }
  StPinSlot(const gsm_ctx* c, size_t n, int k) {
    tree = (char*)c->st_pin + k * stride(n);
    w = tree + StTreeBlock(n).bytes;
    code = w + 8 * n;
    ws = code + 8 * n;
    flags = ws + 8 * n;
    rec = flags + (n + 255) / 256 * 256;
    img = rec + 8 * n;
  }
};
int st_reserve(gsm_ctx* c, size_t n, int D, bool with_vol, int pin_slots = 1, int arenas = 1) {
  const size_t pin = pin_slots * StPinSlot::stride(n);
  if (c->st_pin_bytes < pin) {
    if (c->st_pin) cudaFreeHost(c->st_pin);
    c->st_pin = nullptr;
    c->st_pin_bytes = 0;
    CK(cudaHostAlloc(&c->st_pin, pin, cudaHostAllocDefault));
    c->st_pin_bytes = pin;
  }
  const size_t need = arenas * StArena(nullptr, n, D, with_vol).bytes;
  if (c->st_bytes >= need) return GSM_OK;
  if (c->st_buf) cudaFree(c->st_buf);
  c->st_buf = nullptr;
  c->st_bytes = 0;
  CK(cudaMalloc(&c->st_buf, need));
  c->st_bytes = need;
  return GSM_OK;
}
int st_check(const gsm_ctx* c, int rows, int cols, int D) {
  if (!c) return fail(GSM_ERR_INVALID, "null ctx");
  // the reference's 3x3 median (ctmf.c:211-212) asserts on images below 3x3
  if (rows < 3 || cols < 3) return fail(GSM_ERR_INVALID, "segment-tree stereo needs at least 3x3 pixels (got %dx%d)", rows, cols);
  if (D < 1 || D > MAX_DISP) return fail(GSM_ERR_INVALID, "num_disp %d not in 1..256", D);
  return GSM_OK;
}
// A tree builder: host work space + its own stream and device scratch for the data-parallel middle phase.  One per
// thread that builds trees (the calling thread uses builder 0).
int st_worker(gsm_ctx* c, int k, size_t n, gsm_ctx::StWorker** out) {
  while ((int)c->st_workers.size() <= k) c->st_workers.push_back(new gsm_ctx::StWorker());
  gsm_ctx::StWorker* w = c->st_workers[k];
  if (!w->s) {  // high priority: a builder's small kernels must not queue behind the tree filters of other frames
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&w->s, cudaStreamNonBlocking, hi));
  }
  // weights (two float planes at most) | flags | records | edge keys in/out, codes in/out, sorted weights (<= 2n words each) | sort temp
  if (w->scratch_n != n) {  // (the size query walks the sort's dispatch: not something to repeat per call and builder)
    size_t temp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, temp, (const u32*)nullptr, (u32*)nullptr, (const u32*)nullptr, (u32*)nullptr,
                                       (int)(2 * n), 0, 32));
    temp = (temp + 255) / 256 * 256;
    const size_t need = 7 * ((8 * n + 255) / 256 * 256) + (n + 255) / 256 * 256 + temp + 2 * ((3 * n + 255) / 256 * 256);
    if (w->dev_bytes < need) {
      if (w->dev) cudaFree(w->dev);
      w->dev = nullptr;
      w->dev_bytes = 0;
      w->scratch_n = 0;
      CK(cudaMalloc(&w->dev, need));
      w->dev_bytes = need;
    }
    w->sort_temp = temp;
    w->scratch_n = n;
  }
  *out = w;
  return GSM_OK;
}
// ---- the tree of one view, in three phases ---------------------------------------------------------------------------
// phase 1 (GPU, asynchronous on s): 3x3 median of the image, edge weights, read-back into pw.  With disp / mask (device
// u8 maps) the weights are CColorDepthWeight's (SegmentTree.cpp:197-218, float), else CColorWeight's (u8).
int st_weights_async(gsm_ctx* c, const StArena& a, const u8* img3, int rows, int cols, char* pw, cudaStream_t s,
                     const u8* disp = nullptr, const u8* mask = nullptr, int level = 0) {
  const size_t n = (size_t)rows * cols;
  const dim3 blk(128), grd((cols + 127) / 128, rows);
  st_median3_kernel<<<grd, blk, 0, s>>>(img3, a.med, rows, cols);
  if (disp) {
    st_edge_weight_depth_kernel<<<grd, blk, 0, s>>>(a.med, disp, mask, (float)level, a.wrf, a.wuf, rows, cols);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pw, a.wrf, 4 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(pw + 4 * n, a.wuf, 4 * n, cudaMemcpyDeviceToHost, s));
  } else {
    st_edge_weight_kernel<<<grd, blk, 0, s>>>(a.med, a.wr, a.wu, rows, cols);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pw, a.wr, n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(pw + n, a.wu, n, cudaMemcpyDeviceToHost, s));
  }
  c->launches += 2;
  return GSM_OK;
}
// GSM_ST_TIMING=1: wall-clock milliseconds of the builder's phases on stderr (dev aid; the GPU phases include their syncs)
struct StPhaseClock {
  bool on;
  std::chrono::steady_clock::time_point t0;
  char line[256];
  int len = 0;
  StPhaseClock() : on(getenv("GSM_ST_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) { line[0] = 0; }
  void mark(const char* name) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    len += snprintf(line + len, sizeof(line) - len, " %s %.3f", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
  void print(const char* what) {
    if (on) fprintf(stderr, "[gsm st] %s:%s ms\n", what, line);
  }
};
// phase 2 (safe to run concurrently on different builders / pinned slots; the calling thread must have the context's
// device current): the ordered tree from the weights in pin.w, left in the pinned mirror of the arena's tree block.
//   GPU (the builder's stream): edges enumerated and sorted by weight      (stable radix sort)
//   host: the two Kruskal passes -> kept-edge flags                        (sequential by definition)
//   GPU: per-pixel records from flags + weights                            (data parallel: st_records_kernel)
//   host: breadth-first ordering                                           (sequential, one record per node)
// Fills dt except the device pointers.
int st_build(gsm_ctx::StWorker& w, const StPinSlot& pin, bool float_weights, int rows, int cols, float sigma, float tau,
             StTree* dt, long long* launches, const u8* left_host = nullptr, const u8* right_host = nullptr) {
  const size_t n = (size_t)rows * cols;
  const size_t wb = float_weights ? 4 : 1;
  const int m = (cols - 1) * rows + (rows - 1) * cols;  // edges of the grid
  StPhaseClock clk;
  const size_t r8 = (8 * n + 255) / 256 * 256, r1 = (n + 255) / 256 * 256;  // (the carve-up st_worker sized)
  char* dw = (char*)w.dev;
  u8* dflags = (u8*)(dw + r8);
  unsigned long long* drec = (unsigned long long*)(dw + r8 + r1);
  u32* key_in = (u32*)(dw + 2 * r8 + r1);
  u32* key_out = (u32*)(dw + 3 * r8 + r1);
  u32* code_in = (u32*)(dw + 4 * r8 + r1);
  u32* code_out = (u32*)(dw + 5 * r8 + r1);
  float* dws = (float*)(dw + 6 * r8 + r1);
  void* temp = dw + 7 * r8 + r1;
  const dim3 blk(128), grd((cols + 127) / 128, rows);
  if (left_host) {
    // batches: the builder stages its frame itself (the copies out of pageable memory run on all builder threads) and
    // computes the edge weights on its own stream -- they never leave the device
    const size_t r3 = (3 * n + 255) / 256 * 256;
    u8* dimg = (u8*)temp + w.sort_temp;
    u8* dmed = dimg + r3;
    memcpy(pin.img, left_host, 3 * n);
    memcpy(pin.img + r3, right_host, 3 * n);
    CK(cudaMemcpyAsync(dimg, pin.img, 3 * n, cudaMemcpyHostToDevice, w.s));
    st_median3_kernel<<<grd, blk, 0, w.s>>>(dimg, dmed, rows, cols);
    st_edge_weight_kernel<<<grd, blk, 0, w.s>>>(dmed, (u8*)dw, (u8*)dw + n, rows, cols);
    *launches += 2;
  } else {
    CK(cudaMemcpyAsync(dw, pin.w, 2 * n * wb, cudaMemcpyHostToDevice, w.s));
  }
  // ---- the edges in the reference's sorted order (GPU: enumerate in (b, a) order, stable radix sort by weight)
  size_t temp_bytes = w.sort_temp;
  if (float_weights) {
    st_enumerate_edges_kernel<float><<<grd, blk, 0, w.s>>>((const float*)dw, (const float*)dw + n, key_in, code_in, rows, cols);
    CK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const u32*)key_in, key_out, (const u32*)code_in, code_out, m, 0, 32, w.s));
    st_edge_weights_from_keys_kernel<float><<<(m + 255) / 256, 256, 0, w.s>>>(key_out, dws, m);
  } else {
    st_enumerate_edges_kernel<u8><<<grd, blk, 0, w.s>>>((const u8*)dw, (const u8*)dw + n, key_in, code_in, rows, cols);
    CK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const u32*)key_in, key_out, (const u32*)code_in, code_out, m, 0, 8, w.s));
    st_edge_weights_from_keys_kernel<u8><<<(m + 255) / 256, 256, 0, w.s>>>(key_out, dws, m);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(pin.code, code_out, 4 * (size_t)m, cudaMemcpyDeviceToHost, w.s));
  CK(cudaMemcpyAsync(pin.ws, dws, 4 * (size_t)m, cudaMemcpyDeviceToHost, w.s));
  CK(cudaStreamSynchronize(w.s));
  *launches += 3;  // + the sort's own kernels
  clk.mark("weights+sort(gpu)");
  const uint32_t* code = (const uint32_t*)pin.code;
  const float* ws = (const float*)pin.ws;
  // ---- the two Kruskal passes (host: sequential by definition) -> kept-edge flags
  gsm_st::detail::kruskal(w.k, code, ws, rows, cols, m, tau);
  memcpy(pin.flags, w.k.flags.data(), n);
  clk.mark("kruskal(host)");
  // ---- per-pixel records (GPU)
  CK(cudaMemcpyAsync(dflags, pin.flags, n, cudaMemcpyHostToDevice, w.s));
  if (float_weights)
    st_records_kernel<float><<<grd, blk, 0, w.s>>>(dflags, (const float*)dw, (const float*)dw + n, drec, rows, cols,
                                                   /*CColorDepthWeight::GetScale*/ 255.0f);
  else
    st_records_kernel<u8><<<grd, blk, 0, w.s>>>(dflags, (const u8*)dw, (const u8*)dw + n, drec, rows, cols,
                                                /*CColorWeight::GetScale*/ 1.0f);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(pin.rec, drec, 8 * n, cudaMemcpyDeviceToHost, w.s));
  CK(cudaStreamSynchronize(w.s));
  ++*launches;
  clk.mark("records(gpu)");
  gsm_st::Tree& t = w.t;
  gsm_st::detail::bfs((const uint64_t*)pin.rec, rows, cols, t);
  clk.mark("bfs(host)");
  gsm_st::detail::release_if_large(w.k, rows * cols);
  const StTreeBlock tb(n);
  memcpy(pin.tree + tb.o_order, t.order.data(), 4 * n);
  memcpy(pin.tree + tb.o_level, t.level_off.data(), 4 * t.level_off.size());
  memcpy(pin.tree + tb.o_father, t.father.data(), 4 * n);
  memcpy(pin.tree + tb.o_child0, t.child0.data(), 4 * n);
  memcpy(pin.tree + tb.o_fdist, t.fdist.data(), n);
  memcpy(pin.tree + tb.o_nchild, t.nchild.data(), n);
  float table[256];
  gsm_st::weight_table(sigma, table);
  memcpy(pin.tree + tb.o_wtab, table, sizeof(table));
  dt->levels = (int)t.level_off.size() - 1;
  dt->n = (int)n;
  dt->max_width = 0;
  for (size_t l = 0; l + 1 < t.level_off.size(); ++l) dt->max_width = std::max(dt->max_width, t.level_off[l + 1] - t.level_off[l]);
  clk.mark("stage");
  clk.print("tree builder");
  return GSM_OK;
}
// phase 3 (asynchronous on s; pin.tree stays untouched until s has passed it): the tree block to the device as one copy,
// then the words the filter kernels read
int st_upload(gsm_ctx* c, const StArena& a, const StPinSlot& pin, StTree* dt, cudaStream_t s) {
  const size_t n = (size_t)dt->n;
  const StTreeBlock tb(n);
  CK(cudaMemcpyAsync(a.tree, pin.tree, tb.bytes, cudaMemcpyHostToDevice, s));
  st_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a.order, (const int*)(a.tree + tb.o_father), (const int*)(a.tree + tb.o_child0),
                                                           (const u8*)(a.tree + tb.o_fdist), (const u8*)(a.tree + tb.o_nchild),
                                                           (const int*)(a.tree + tb.o_wtab), a.up, a.down, a.pos, (int)n);
  c->launches++;
  CK(cudaGetLastError());
  dt->up = a.up; dt->down = a.down; dt->level_off = a.level_off;
  return GSM_OK;
}
// all three for one image on the calling thread: image (device, interleaved 3-channel) -> the ordered tree on the device;
// the host form of the tree is left in builder 0 (c->st_workers[0]->t)
int st_tree(gsm_ctx* c, const StArena& a, const u8* img3, int rows, int cols, float sigma, float tau, StTree* dt,
            cudaStream_t s, const u8* disp = nullptr, const u8* mask = nullptr, int level = 0) {
  const size_t n = (size_t)rows * cols;
  const StPinSlot pin(c, n, 0);
  gsm_ctx::StWorker* w;
  int rc;
  if ((rc = st_worker(c, 0, n, &w))) return rc;
  if ((rc = st_weights_async(c, a, img3, rows, cols, pin.w, s, disp, mask, level))) return rc;
  CK(cudaStreamSynchronize(s));  // (also: every earlier upload from this pinned slot has been consumed)
  if ((rc = st_build(*w, pin, disp != nullptr, rows, cols, sigma, tau, dt, &c->launches))) return rc;
  return st_upload(c, a, pin, dt, s);
}
// the tree filter over D channels: the on-chip ring kernel with the deepest ring the widest level allows, else the plain one
template <int RING>
bool st_filter_ring_try(gsm_ctx* c, float* buf, float* fin, const StTree& dt, int D, cudaStream_t s, int slot) {
  if (c->st_ring_smem[slot] < 0) {  // once per context: how much dynamic shared memory this instance may use
    int optin = 0;
    cudaFuncAttributes fa;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, st_filter_ring_kernel<RING>) != cudaSuccess)
      return false;
    const int dyn = optin - (int)fa.sharedSizeBytes;
    if (cudaFuncSetAttribute(st_filter_ring_kernel<RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) return false;
    c->st_ring_smem[slot] = dyn;
  }
  const int cap = (dt.max_width + 3) & ~3;
  const size_t smem = (size_t)RING * (cap + ST_PAD) * 16;
  if (smem > (size_t)c->st_ring_smem[slot]) return false;
  st_filter_ring_kernel<RING><<<D, 256, smem, s>>>(buf, fin, dt, cap);
  return true;
}
int st_filter_launch(gsm_ctx* c, float* buf, float* fin, const StTree& dt, int D, cudaStream_t s) {
  // GSM_ST_RING = 16 / 8 / 4 caps the ring depth, 0 forces the plain kernel (tests run every variant); default: deepest that fits
  const char* e = getenv("GSM_ST_RING");
  const int maxr = e ? atoi(e) : 16;
  if (!((maxr >= 16 && st_filter_ring_try<16>(c, buf, fin, dt, D, s, 0)) || (maxr >= 8 && st_filter_ring_try<8>(c, buf, fin, dt, D, s, 1)) ||
        (maxr >= 4 && st_filter_ring_try<4>(c, buf, fin, dt, D, s, 2))))
    st_filter_kernel<<<D, 256, 0, s>>>(buf, fin, dt);
  c->launches++;
  return GSM_OK;
}
}  // namespace

// The host stage on its own (no GPU involved): edge weights in, ordered tree out -- what gsm_st_filter runs between its
// two GPU stages.  wr[p] = weight of edge (p, p+1), wu[p] = weight of edge (p, p-cols) (entries of edges that do not
// exist are ignored).
extern "C" int gsm_st_build_tree_host(const uint8_t* wr, const uint8_t* wu, int rows, int cols, float tau, int* order,
                                      int* father_id, uint8_t* father_dist, int* levels) {
  if (!wr || !wu || rows < 1 || cols < 1) return fail(GSM_ERR_INVALID, "gsm_st_build_tree_host: bad arguments");
  static thread_local gsm_st::Tree t;
  gsm_st::build_tree(wr, wu, rows, cols, tau > 0.f ? tau : 1200.f, 1.0f, t);
  const size_t n = (size_t)rows * cols;
  if (order) memcpy(order, t.order.data(), 4 * n);
  if (father_id) memcpy(father_id, t.father_id.data(), 4 * n);
  if (father_dist) memcpy(father_dist, t.fdist.data(), n);
  if (levels) *levels = (int)t.level_off.size() - 1;
  return GSM_OK;
}

extern "C" int gsm_st_matching_cost(gsm_ctx* c, const uint8_t* left3, const uint8_t* right3, float* cost, int rows,
                                    int cols, int num_disp) {
  int rc;
  if ((rc = st_check(c, rows, cols, num_disp))) return rc;
  if (!left3 || !right3 || !cost) return fail(GSM_ERR_INVALID, "null pointer");
  CK(cudaSetDevice(c->device));
  if ((rc = gsm_sync(c))) return rc;
  const size_t n = (size_t)rows * cols;
  if ((rc = st_reserve(c, n, num_disp, true))) return rc;
  const StArena a(c->st_buf, n, num_disp, true);
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(a.L3, left3, 3 * n, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(a.R3, right3, 3 * n, cudaMemcpyHostToDevice, s));
  const dim3 blk(128), grd((cols + 127) / 128, rows);
  st_gray_grad_kernel<<<grd, blk, 0, s>>>(a.L3, a.gL, rows, cols);
  st_gray_grad_kernel<<<grd, blk, 0, s>>>(a.R3, a.gR, rows, cols);
  st_cost_kernel<<<grd, blk, 0, s>>>(a.L3, a.R3, a.gL, a.gR, a.vol, nullptr, 0, rows, cols, num_disp);
  c->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(cost, a.vol, 4 * n * num_disp, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_st_filter(gsm_ctx* c, const uint8_t* image3, float* cost, int rows, int cols, int num_disp,
                             float sigma, float tau, int* order, int* father_id, uint8_t* father_dist) {
  int rc;
  if ((rc = st_check(c, rows, cols, num_disp))) return rc;
  if (!image3) return fail(GSM_ERR_INVALID, "null pointer");
  CK(cudaSetDevice(c->device));
  if ((rc = gsm_sync(c))) return rc;
  const size_t n = (size_t)rows * cols;
  if ((rc = st_reserve(c, n, num_disp, true))) return rc;
  const StArena a(c->st_buf, n, num_disp, true);
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(a.L3, image3, 3 * n, cudaMemcpyHostToDevice, s));
  StTree dt;
  if ((rc = st_tree(c, a, a.L3, rows, cols, sigma, tau, &dt, s))) return rc;
  gsm_st::Tree& t = c->st_workers[0]->t;
  if (father_id) gsm_st::detail::father_ids(t);
  if (order) memcpy(order, t.order.data(), 4 * n);
  if (father_id) memcpy(father_id, t.father_id.data(), 4 * n);
  if (father_dist) memcpy(father_dist, t.fdist.data(), n);
  if (cost) {
    CK(cudaMemcpyAsync(a.vol, cost, 4 * n * num_disp, cudaMemcpyHostToDevice, s));
    const dim3 g2((unsigned)((n + 255) / 256), num_disp);
    st_permute_kernel<<<g2, 256, 0, s>>>(a.vol, a.order, a.buf, (int)n, num_disp);
    if ((rc = st_filter_launch(c, a.buf, a.fin, dt, num_disp, s))) return rc;
    st_unpermute_kernel<<<g2, 256, 0, s>>>(a.fin, a.order, a.vol, (int)n, num_disp);
    c->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cost, a.vol, 4 * n * num_disp, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  return GSM_OK;
}

extern "C" int gsm_segment_tree_stereo(gsm_ctx* c, const gsm_st_params* p, const uint8_t* left3, const uint8_t* right3,
                                       uint8_t* disparity, int rows, int cols) {
  if (!p) return fail(GSM_ERR_INVALID, "null params");
  int rc;
  if ((rc = st_check(c, rows, cols, p->num_disp))) return rc;
  if (!left3 || !right3 || !disparity) return fail(GSM_ERR_INVALID, "null pointer");
  if (p->median_radius < 0 || p->median_radius > MED_MAXR) return fail(GSM_ERR_INVALID, "median_radius %d", p->median_radius);
  if (p->scale < 1) return fail(GSM_ERR_INVALID, "scale %d", p->scale);
  StPhaseClock clk;
  CK(cudaSetDevice(c->device));
  if ((rc = gsm_sync(c))) return rc;
  const size_t n = (size_t)rows * cols;
  const int D = p->num_disp;
  if ((rc = st_reserve(c, n, D, false, p->refined ? 2 : 1))) return rc;
  const StArena a(c->st_buf, n, D, false);
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(a.L3, left3, 3 * n, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(a.R3, right3, 3 * n, cudaMemcpyHostToDevice, s));
  const dim3 blk(128), grd((cols + 127) / 128, rows);
  st_gray_grad_kernel<<<grd, blk, 0, s>>>(a.L3, a.gL, rows, cols);
  st_gray_grad_kernel<<<grd, blk, 0, s>>>(a.R3, a.gR, rows, cols);
  c->launches += 2;
  StTree dt;
  const float tau = p->tau > 0.f ? p->tau : 1200.f;
  const unsigned gn = (unsigned)((n + 255) / 256);
  // with the tree of a view on the device (dt): its cost, aggregated over the tree, winner (+ median) -> a.disp / a.disp2
  auto view_gpu = [&](int right, u8** out) -> int {
    // the cost kernel writes straight into the [D][BFS position] layout the filter works on
    if (right) st_cost_right_kernel<<<grd, blk, 0, s>>>(a.L3, a.R3, a.gL, a.gR, a.buf, a.pos, n, rows, cols, D);
    else st_cost_kernel<<<grd, blk, 0, s>>>(a.L3, a.R3, a.gL, a.gR, a.buf, a.pos, n, rows, cols, D);
    int r;
    if ((r = st_filter_launch(c, a.buf, a.fin, dt, D, s))) return r;
    st_wta_kernel<<<gn, 256, 0, s>>>(a.fin, a.order, a.disp, (int)n, D);
    c->launches += 2;
    *out = a.disp;
    if (p->median_radius > 0) {
      if ((r = median_launch(c, a.disp, a.disp2, 1, rows, cols, p->median_radius, s))) return r;
      *out = a.disp2;
    }
    CK(cudaGetLastError());
    return GSM_OK;
  };
  auto view = [&](const u8* img, int right, float sigma, const u8* tdisp, const u8* tmask, u8** out) -> int {
    int r;
    clk.mark("upload+launches");
    if ((r = st_tree(c, a, img, rows, cols, sigma, tau, &dt, s, tdisp, tmask, D))) return r;
    clk.mark("tree(weights, builder, upload)");
    return view_gpu(right, out);
  };
  u8* out = nullptr;
  if (!p->refined) {
    if ((rc = view(a.L3, 0, p->sigma, nullptr, nullptr, &out))) return rc;
  } else {
    // stereo_disparity_iteration (StereoDisparity.cpp:92-160): both views with SIGMA_ONE (Toolkit.h:34), L-R check
    // (:128-147), then a second left pass over the tree of CColorDepthWeight(left image, left disparity, mask)
    if (D > cols) return fail(GSM_ERR_INVALID, "refined segment-tree stereo needs num_disp <= cols (StereoHelper.cpp:162-177)");
    const float SIGMA_ONE = 0.08f;
    // the trees of the two views are independent and host-bound: the right one is built on a second thread while this
    // one builds the left one; the GPU takes the left view as soon as its tree is up
    const StPinSlot pinL(c, n, 0), pinR(c, n, 1);
    gsm_ctx::StWorker *wL, *wR;
    if ((rc = st_worker(c, 0, n, &wL)) || (rc = st_worker(c, 1, n, &wR))) return rc;
    if ((rc = st_weights_async(c, a, a.L3, rows, cols, pinL.w, s))) return rc;
    if ((rc = st_weights_async(c, a, a.R3, rows, cols, pinR.w, s))) return rc;
    CK(cudaStreamSynchronize(s));
    StTree dtR;
    int rcR = GSM_OK;
    std::string errR;
    long long launchesR = 0;
    const int dev = c->device;
    auto build_right = [&] {
      if (cudaSetDevice(dev) != cudaSuccess) { rcR = GSM_ERR_CUDA; errR = "cudaSetDevice failed on the builder thread"; return; }
      rcR = st_build(*wR, pinR, false, rows, cols, SIGMA_ONE, tau, &dtR, &launchesR);
      if (rcR) errR = g_err;  // (thread-local text)
    };
    std::thread right_builder;
    try {
      right_builder = std::thread(build_right);
    } catch (...) {  // no thread to be had: build it here, after the left one
    }
    rc = st_build(*wL, pinL, false, rows, cols, SIGMA_ONE, tau, &dt, &c->launches);
    if (!rc) rc = st_upload(c, a, pinL, &dt, s);
    if (!rc) rc = view_gpu(0, &out);
    if (!rc && cudaMemcpyAsync(a.dispL, out, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = fail(GSM_ERR_CUDA, "copy of the left disparity");
    if (right_builder.joinable()) right_builder.join();  // (before any return: the thread references this frame)
    else build_right();
    c->launches += launchesR;
    if (rc) return rc;
    if (rcR) return fail(rcR, "tree builder, right view: %s", errR.c_str());
    dt = dtR;
    if ((rc = st_upload(c, a, pinR, &dt, s))) return rc;
    if ((rc = view_gpu(1, &out))) return rc;
    CK(cudaMemcpyAsync(a.dispR, out, n, cudaMemcpyDeviceToDevice, s));
    lr_check_kernel<<<dim3((cols + 255) / 256, rows, 1), 256, 0, s>>>(a.dispL, a.dispR, nullptr, a.mask, nullptr, rows, cols, 1,
                                                                        nullptr);
    c->launches++;
    if ((rc = view(a.L3, 0, p->sigma, a.dispL, a.mask, &out))) return rc;
  }
  if (p->scale != 1) {
    st_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, n, p->scale);
    c->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(disparity, out, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  clk.mark("cost+filter+wta+median+download");
  clk.print("gsm_segment_tree_stereo");
  return GSM_OK;
}

// A batch of pairs (frames of one size): the host-bound part of the pipeline -- one tree per frame -- is data parallel
// across frames, so the trees are built concurrently on host threads while the GPU takes each frame as soon as its tree
// is packed.  Per frame the stages and their results are those of gsm_segment_tree_stereo (ST-1).
extern "C" int gsm_segment_tree_stereo_batch(gsm_ctx* c, const gsm_st_params* p, const uint8_t* left3, const uint8_t* right3,
                                             uint8_t* disparity, int nframes, int rows, int cols, int host_threads) {
  if (!p) return fail(GSM_ERR_INVALID, "null params");
  int rc;
  if ((rc = st_check(c, rows, cols, p->num_disp))) return rc;
  if (!left3 || !right3 || !disparity) return fail(GSM_ERR_INVALID, "null pointer");
  if (nframes < 1) return fail(GSM_ERR_INVALID, "nframes %d", nframes);
  if (p->refined) return fail(GSM_ERR_UNSUPPORTED, "gsm_segment_tree_stereo_batch runs ST-1 (refined = 0); call gsm_segment_tree_stereo per pair for ST-2");
  if (p->median_radius < 0 || p->median_radius > MED_MAXR) return fail(GSM_ERR_INVALID, "median_radius %d", p->median_radius);
  if (p->scale < 1) return fail(GSM_ERR_INVALID, "scale %d", p->scale);
  CK(cudaSetDevice(c->device));
  if ((rc = gsm_sync(c))) return rc;
  const size_t n = (size_t)rows * cols;
  const int D = p->num_disp;
  const float tau = p->tau > 0.f ? p->tau : 1200.f;
  int T = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  T = std::max(1, std::min(std::min(T, 64), nframes));
  // frames in flight: two per builder thread, within 512 MB of pinned host memory (a slot mirrors one frame's tree),
  // and chunks of equal size (a short last chunk costs a whole round of tree builds)
  const size_t slot_bytes = StPinSlot::stride(n);
  const int kmax = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(2 * (size_t)T, (size_t)nframes), std::max<size_t>(2, ((size_t)512 << 20) / slot_bytes)));
  const int chunks = (nframes + kmax - 1) / kmax;
  const int K = (nframes + chunks - 1) / chunks;
  // Several device arenas, each on its own stream: a tree filter is one CTA per disparity and latency-bound (a quarter
  // of its SM's issue slots), so the filters of several frames run side by side -- on different SMs while there are
  // free ones, co-resident after that.
  const size_t arena_bytes = StArena(nullptr, n, D, false).bytes;
  const int NA = (int)std::max<size_t>(1, std::min<size_t>(4, ((size_t)16 << 30) / arena_bytes));  // within 16 GB; 8 arenas measured no faster
  if ((rc = st_reserve(c, n, D, false, K, NA))) return rc;
  for (int k = 0; k < T; ++k) {
    gsm_ctx::StWorker* unused;
    if ((rc = st_worker(c, k, n, &unused))) return rc;
  }
  std::vector<StArena> arena;
  for (int k = 0; k < NA; ++k) {
    arena.emplace_back((char*)c->st_buf + k * arena_bytes, n, D, false);
    if (!c->st_streams[k]) CK(cudaStreamCreateWithFlags(&c->st_streams[k], cudaStreamNonBlocking));
  }
  cudaStream_t* streams = c->st_streams;
  const StArena& a = arena[0];
  cudaStream_t s = streams[0];
  const dim3 blk(128), grd((cols + 127) / 128, rows);
  const unsigned gn = (unsigned)((n + 255) / 256);

  for (int f0 = 0; f0 < nframes; f0 += K) {
    const int kc = std::min(K, nframes - f0);
    // ---- builder threads: frames in order off a shared counter; each stages its frame in its pinned slot, computes the
    // edge weights on its own stream and builds the tree (st_build)
    std::vector<StTree> dts(kc);
    std::vector<char> done(kc, 0);
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> next{0};
    std::vector<int> brc(kc, GSM_OK);
    std::vector<std::string> berr(kc);  // (the error text is thread-local: carried over by hand)
    std::vector<long long> blaunches(kc, 0);
    const int dev = c->device;
    auto builder = [&](int w) {
      const bool dev_ok = cudaSetDevice(dev) == cudaSuccess;
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= kc) break;
        const StPinSlot pin(c, n, i);
        brc[i] = dev_ok ? st_build(*c->st_workers[w], pin, false, rows, cols, p->sigma, tau, &dts[i], &blaunches[i],
                                   left3 + (size_t)(f0 + i) * 3 * n, right3 + (size_t)(f0 + i) * 3 * n)
                        : GSM_ERR_CUDA;
        if (brc[i]) berr[i] = dev_ok ? g_err : "cudaSetDevice failed on a builder thread";
        {
          std::lock_guard<std::mutex> lk(mu);
          done[i] = 1;
        }
        cv.notify_all();
      }
    };
    std::vector<std::thread> threads;
    for (int w = 0; w < std::min(T, kc); ++w) {
      try {
        threads.emplace_back(builder, w);
      } catch (...) {
        break;  // fewer threads than asked for
      }
    }
    if (threads.empty()) builder(0);  // none at all: build here, then run the GPU stages
    // ---- the calling thread: each frame to the GPU as soon as its tree is packed (asynchronous copies and launches
    // only); the disparity comes back through the frame's slot
    auto frame_gpu = [&](int i) -> int {
      const StPinSlot pin(c, n, i);
      const StArena& a = arena[i % NA];
      cudaStream_t s = streams[i % NA];
      if (brc[i]) return fail(brc[i], "tree builder, frame %d: %s", f0 + i, berr[i].c_str());
      CK(cudaMemcpyAsync(a.L3, pin.img, 3 * n, cudaMemcpyHostToDevice, s));  // staged by the builder
      CK(cudaMemcpyAsync(a.R3, pin.img + (3 * n + 255) / 256 * 256, 3 * n, cudaMemcpyHostToDevice, s));
      st_gray_grad_kernel<<<grd, blk, 0, s>>>(a.L3, a.gL, rows, cols);
      st_gray_grad_kernel<<<grd, blk, 0, s>>>(a.R3, a.gR, rows, cols);
      int r;
      if ((r = st_upload(c, a, pin, &dts[i], s))) return r;
      st_cost_kernel<<<grd, blk, 0, s>>>(a.L3, a.R3, a.gL, a.gR, a.buf, a.pos, n, rows, cols, D);
      if ((r = st_filter_launch(c, a.buf, a.fin, dts[i], D, s))) return r;
      st_wta_kernel<<<gn, 256, 0, s>>>(a.fin, a.order, a.disp, (int)n, D);
      c->launches += 4;
      u8* out = a.disp;
      if (p->median_radius > 0) {
        if ((r = median_launch(c, a.disp, a.disp2, 1, rows, cols, p->median_radius, s))) return r;
        out = a.disp2;
      }
      if (p->scale != 1) {
        st_scale_kernel<<<gn, 256, 0, s>>>(out, n, p->scale);
        c->launches++;
      }
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(pin.w, out, n, cudaMemcpyDeviceToHost, s));  // the weights in pin.w have been consumed by the builder
      return GSM_OK;
    };
    rc = GSM_OK;
    for (int i = 0; i < kc && !rc; ++i) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done[i] != 0; });
      }
      rc = frame_gpu(i);
    }
    next.store(kc);  // (on an error: no further frames are started)
    for (std::thread& th : threads) th.join();
    for (int i = 0; i < kc; ++i) c->launches += blaunches[i];
    if (rc) {
      for (int k = 0; k < NA; ++k) cudaStreamSynchronize(streams[k]);
      return rc;
    }
    for (int k = 0; k < NA; ++k) CK(cudaStreamSynchronize(streams[k]));
    for (int i = 0; i < kc; ++i) memcpy(disparity + (size_t)(f0 + i) * n, StPinSlot(c, n, i).w, n);
  }
  return GSM_OK;
}

extern "C" int gsm_measure_alu_peak(gsm_ctx* c, double* lane_ops_per_s) {
  if (!c || !lane_ops_per_s) return fail(GSM_ERR_INVALID, "null pointer");
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  const int grid = prop.multiProcessorCount * 8, iters = 8192;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0, c->stream));
    alu_peak_kernel<<<grid, 256, 0, c->stream>>>(c->peak_buf, iters);
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *lane_ops_per_s = (double)grid * 256 * iters * 8 * 2 / (best * 1e-3);
  return GSM_OK;
}

#ifdef GSM_GF_PROFILE
// dev builds only (make EXTRA=-DGSM_GF_PROFILE): read / reset the per-phase cycle counters of gf3_wta_kernel
extern "C" int gsm_debug_gf_prof(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, gsm::g_gf_prof, sizeof(unsigned long long) * 16 * 9);
  if (reset) {
    unsigned long long z[16 * 9] = {0};
    cudaMemcpyToSymbol(gsm::g_gf_prof, z, sizeof(z));
  }
  return 0;
}
#endif
