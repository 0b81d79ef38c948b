/*
 * gsm.h -- C ABI of libgsm.so: the B200-native (sm_100a) BlockMatching hot path.
 *
 * Drop-in boundary for ningw42/GPU_Stereo_Matching's GPU "proxy" functions
 * (reference: BlockMatching/Device.cuh:50-52, implemented in BlockMatching/Device.cu:173-367)
 * and the CPU functions they are checked against (BlockMatching/BlockMatching.h:8-15).
 * Plain pointers and sizes only; no OpenCV, no torch types.  include/gsm_compat.hpp layers the
 * reference's exact C++ signatures (cv::Mat&) on top of this file.
 *
 * Conventions
 *   - images: uint8, row-major, contiguous, rows x cols  (reference: raw .data memcpy of
 *     rows*cols bytes, Device.cu:213-214).
 *   - radius  == the reference's `SADWindowSize` / `SAD` argument: window = 2*radius+1
 *     (Device.cu:181, BlockMatching.cpp:119).
 *   - num_disp == the reference's `searchRange`: disparities d in [0, num_disp), <= 256
 *     (u8 output, BlockMatching.cpp:184).
 *   - every entry point returns 0 on success, a negative gsm_status otherwise;
 *     gsm_last_error() returns a thread-local description.  (The reference ignores every
 *     CUDA status; this ABI never does.)
 *   - a gsm_ctx owns all device memory and streams for ONE GPU and is used by
 *     one host thread at a time.  Nothing is allocated per call once the context is warm
 *     (reference: 6 cudaMalloc per call, never freed, Device.cu:187-194).
 *   - There is NO CPU fallback: without a CUDA device every call fails with GSM_ERR_CUDA.
 */
#ifndef GSM_H
#define GSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsm_ctx gsm_ctx;

typedef enum {
  GSM_OK = 0,
  GSM_ERR_INVALID = -1,  /* bad argument (size, radius, num_disp, null pointer) */
  GSM_ERR_CUDA = -2,     /* CUDA runtime error (message in gsm_last_error) */
  GSM_ERR_CAPACITY = -3, /* request exceeds what the context was created for */
  GSM_ERR_UNSUPPORTED = -4
} gsm_status;

typedef enum {
  GSM_MODE_SAD = 0, /* clipped-window SAD + WTA with the reference quirks: bit-exact to getDisp,
                       BlockMatching.cpp:111-189 / kernalFindCorr, Device.cu:34-64 */
  GSM_MODE_GF = 1   /* guided-filter aggregation (north_star; no reference implementation: parity unpinned) + WTA
                       over all d, lowest d on ties (STMatching/StereoHelper.cpp:131-154).  The WTA compares the
                       fp32 cost with its 5 low mantissa bits cleared (they carry the lane index inside the warp
                       reduction): costs closer than 2^-18 relative count as ties and resolve to the lowest d,
                       where the reference's strict '<' on exact floats would keep the smaller one.  That is two
                       orders of magnitude below the 1e-4 tolerance of the costs themselves. */
} gsm_mode;

/* Parameters of one stereo pass.  Zero-initialise, then set what you need. */
typedef struct gsm_params {
  int mode;          /* gsm_mode */
  int radius;        /* window radius r: SAD 1..12, GF 1..9 */
  int num_disp;      /* D, 1..256 */
  float eps;         /* GF regulariser on the 0..255 intensity scale; <= 0 selects 6.5025 (= 1e-4 * 255^2) */
  int lr_check;      /* GF only: also aggregate the right view (STMatching/StereoHelper.cpp:156-180) and
                        apply the left-right consistency check (STMatching/StereoDisparity.cpp:128-147):
                        occluded pixels -> disparity 0, mask = !occ */
  int median_radius; /* > 0: (2m+1)^2 median, replicate border (STMatching/ctmf.c:378-433, applied like
                        StereoDisparity.cpp:119,126 to both views before the LR check) */
  int row_bands;     /* >= 1: split every frame into this many independent row bands (more CTAs for
                        single-frame latency); 0 = automatic */
  int d_begin;       /* disparity sub-range [d_begin, d_end) evaluated by this call; 0/0 = all.  Used by the */
  int d_end;         /* multi-GPU disparity split: partial results combine through the packed-min plane. */
  int rectify;       /* 1: left/right are RAW (unrectified) frames; they are rectified on the device through the maps
                        given to gsm_set_rectification, fused into the plane packer. */
} gsm_params;

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Replaces the per-call cudaMalloc block of blockMatching_gpu (Device.cu:184-194). */
int gsm_create(gsm_ctx** out, int device, int max_rows, int max_cols, int max_disp, int max_batch);
void gsm_destroy(gsm_ctx* ctx);
const char* gsm_last_error(void);
const char* gsm_version(void);

/* ---- the reference entry point -------------------------------------------------------------- */
/* == blockMatching_gpu(h_left, h_right, h_disparity, SADWindowSize, searchRange), Device.cuh:50 /
 *    Device.cu:173-301, and bit-exact to getDisp, BlockMatching.cpp:111-189.  HOST pointers;
 *    blocking (upload, kernels, download), like the reference. */
int gsm_block_matching(gsm_ctx* ctx, const uint8_t* left, const uint8_t* right, uint8_t* disparity,
                       int rows, int cols, int radius, int num_disp);

/* ---- full path, host buffers (the e2e path) ------------------------------------------------- */
/* n frames of identical size, frame i at left + i*rows*cols.  mask may be NULL (written only with
 * lr_check).  Pipelined internally (uploads, kernels and downloads of consecutive chunks overlap on three streams);
 * fully asynchronous only when the host buffers are page-locked. */
int gsm_stereo_batch(gsm_ctx* ctx, const gsm_params* p, int n, const uint8_t* left, const uint8_t* right,
                     uint8_t* disparity, uint8_t* mask, int rows, int cols);

/* Mixed-size batch: frame i is rows[i] x cols[i] with its own host buffers left[i] / right[i] / disparity[i]
 * (/ mask[i]; mask or mask[i] may be NULL) -- the separate Mats of Caller.cpp:12-19, and BASELINE config 2 (the nine
 * Middlebury sets come in three sizes).  The whole batch runs as ONE launch per stage over a per-frame geometry
 * table; every frame must fit the context (rows[i] <= max_rows, cols[i] <= max_cols).  Blocking.  Results are
 * bit-identical to calling gsm_stereo_batch once per frame with the same row_bands. */
int gsm_stereo_batch_v(gsm_ctx* ctx, const gsm_params* p, int n, const uint8_t* const* left,
                       const uint8_t* const* right, uint8_t* const* disparity, uint8_t* const* mask, const int* rows,
                       const int* cols);

/* Streaming variant: enqueues uploads, kernels and downloads and returns; the results are in `disparity` / `mask`
 * after gsm_sync().  Back-to-back submissions keep the two-slot pipeline full across calls (frame k+1 uploads while
 * frame k computes and frame k-1 downloads) -- what a capture loop (reference: Utility.cpp:198-226) would use.  The
 * host buffers must be page-locked for the copies to be asynchronous and must stay valid until gsm_sync(). */
int gsm_stereo_batch_async(gsm_ctx* ctx, const gsm_params* p, int n, const uint8_t* left, const uint8_t* right,
                           uint8_t* disparity, uint8_t* mask, int rows, int cols);

/* ---- full path, device-resident buffers ----------------------------------------------------- */
/* Same computation on DEVICE pointers (tightly packed n x rows x cols), enqueued on `stream`
 * (a cudaStream_t; NULL = the context's own stream).  Asynchronous: returns after enqueueing. */
int gsm_stereo_device(gsm_ctx* ctx, const gsm_params* p, int n, const void* left_dev, const void* right_dev,
                      void* disparity_dev, void* mask_dev, int rows, int cols, void* stream);
/* Mixed-size batch on DEVICE pointers: frame i is rows[i] x cols[i] at pixel offset sum_{j<i} rows[j]*cols[j] of
 * every buffer (tight, concatenated); n <= the context's batch capacity; rows / cols are HOST arrays. */
int gsm_stereo_device_v(gsm_ctx* ctx, const gsm_params* p, int n, const void* left_dev, const void* right_dev,
                        void* disparity_dev, void* mask_dev, const int* rows, const int* cols, void* stream);
/* Waits for everything enqueued on the context's own streams (compute, upload, download). */
int gsm_sync(gsm_ctx* ctx);

/* ---- multi-GPU disparity split -------------------------------------------------- */
/* Evaluate only d in [p->d_begin, p->d_end) for ONE frame and leave the per-pixel packed minimum in
 * keys_dev (int64[rows*cols], device).  Packed word: SAD  (int64)((SAD << 8) | d);
 * GF  ((int64)sortable_int32(N*q) << 32) | d  (N = window pixel count of the pixel, so the argmin over d is that
 * of q; the 5 low bits of the cost word are cleared) -- both order like (cost, d) under a SIGNED 64-bit min,
 * so ranks combine with ncclAllReduce(ncclInt64, ncclMin) / torch.distributed ReduceOp.MIN and ties
 * resolve to the lowest d exactly like the reference's strict '<' (BlockMatching.cpp:178).
 * view: 0 = left disparity, 1 = right-view disparity (for the LR check). */
int gsm_partial_keys_device(gsm_ctx* ctx, const gsm_params* p, int view, const void* left_dev,
                            const void* right_dev, void* keys_dev, int rows, int cols, void* stream);
/* Same, but the fused kernel also waits for wait_event (a cudaEvent_t recorded on another stream, NULL = none) while
 * the disparity-independent passes in front of it (plane packing, guide statistics) do not: a caller overlaps them
 * with work on another stream that must not share SMs with the fused kernel, e.g. the previous frame's
 * gsm_reduce_keys_p2p (gpu_stereo_matching_b200/dist.py: DsplitStream). */
int gsm_partial_keys_device_ex(gsm_ctx* ctx, const gsm_params* p, int view, const void* left_dev,
                               const void* right_dev, void* keys_dev, int rows, int cols, void* stream,
                               void* wait_event);
/* keys -> u8 disparity (+ optional median / LR check against keys_right_dev, may be NULL). */
int gsm_finalize_keys_device(gsm_ctx* ctx, const gsm_params* p, const void* keys_left_dev,
                             const void* keys_right_dev, void* disparity_dev, void* mask_dev, int rows,
                             int cols, void* stream);

/* u8 maps -> final map: (2m+1)^2 median on both views, then the LR check against disp_right_dev (may be NULL without
 * lr_check) -- the part of gsm_finalize_keys_device after the disparities are extracted
 * (STMatching/StereoDisparity.cpp:119,126,128-147). */
int gsm_postfilter_device(gsm_ctx* ctx, const gsm_params* p, const void* disp_left_dev, const void* disp_right_dev,
                          void* disparity_dev, void* mask_dev, int rows, int cols, void* stream);

/* Combine the ranks' packed-min planes over PEER MEMORY (NVLink P2P) instead of an all-reduce: rank `rank` of `world`
 * reduces its 1/world slice of the npx pixels over all planes (key_ptrs[w] = rank w's int64 plane, mapped into this
 * process, e.g. CUDA IPC / torch symmetric memory), turns it into u8 disparities and stores that slice into every
 * rank's map disp_ptrs[w].  All planes 16-byte aligned.  The caller orders it across ranks: a barrier after every
 * rank's gsm_partial_keys_device and one after this call (gpu_stereo_matching_b200/dist.py: dsplit_stereo_p2p).
 * Replaces nothing in the reference (it is single-GPU); SURVEY 8e. */
int gsm_reduce_keys_p2p(gsm_ctx* ctx, const void* const* key_ptrs, void* const* disp_ptrs, int world, int rank,
                        long long npx, void* stream);

/* Same, confined to at most max_blocks thread blocks (512 threads, one per SM; <= 0: unlimited).  The fused kernel owns
 * every register of the SMs it runs on, so a concurrent combine only makes progress on SMs the fused grid leaves idle:
 * with max_blocks = that number (config 5 at 8 ranks: 148 - 144 = 4) the combine of frame k runs beside the fused
 * kernel of frame k+1 (dist.DsplitStream). */
int gsm_reduce_keys_p2p_ex(gsm_ctx* ctx, const void* const* key_ptrs, void* const* disp_ptrs, int world, int rank,
                           long long npx, void* stream, int max_blocks);

/* ---- cost-stage exports (keep the reference's compareDiff / compareSAD checks possible) ------ */
/* AD volume u8 [D][rows][cols] == PreCal, BlockMatching.cpp:89-109 (what compareDiff :263-276 checks). */
int gsm_ad_volume(gsm_ctx* ctx, const uint8_t* left, const uint8_t* right, uint8_t* volume, int rows, int cols,
                  int num_disp);
/* Aggregated cost slices for d in [d0, d0+nd): SAD mode -> int32 un-truncated SAD [nd][rows][cols];
 * GF mode -> float32 q_d [nd][rows][cols].  Produced by the SAME fused kernel as the WTA path. view as above. */
int gsm_cost_slices(gsm_ctx* ctx, const gsm_params* p, int view, const uint8_t* left, const uint8_t* right,
                    int d0, int nd, void* out, int rows, int cols);
/* getAllSAD layout: u8 [rows*cols][D], truncated, 255 where col+d > cols (BlockMatching.cpp:191-261). */
int gsm_all_sad(gsm_ctx* ctx, const uint8_t* left, const uint8_t* right, uint8_t* out, int rows, int cols,
                int radius, int num_disp);

/* ---- post-filters as stand-alone calls ------------------------------------------------------ */
int gsm_median(gsm_ctx* ctx, const uint8_t* src, uint8_t* dst, int rows, int cols, int radius);
int gsm_lr_check(gsm_ctx* ctx, const uint8_t* disp_left, const uint8_t* disp_right, uint8_t* occ,
                 uint8_t* mask, int rows, int cols);

/* ---- SURVEY 8(f) "next" rows: the other two proxy functions of Device.cuh:50-52 --------------------- */
/* == kernalRemap / CPU_Remap (Device.cu:127-167, Utility.cpp:236-264): dst(r,c) = bilinear sample of src at
 * (row = mapy(r,c), col = mapx(r,c)), 0 when the 2x2 footprint leaves the image, round-nearest-even + saturate.
 * Host pointers; maps are float32 rows x cols (initUndistortRectifyMap CV_32FC1, Utility.cpp:232-233). */
int gsm_remap(gsm_ctx* ctx, const uint8_t* src, const float* mapx, const float* mapy, uint8_t* dst, int rows,
              int cols);
/* == kernalCvtColor (Device.cu:136-143): gray = .299*ch0 + .587*ch1 + .114*ch2 of interleaved 3-channel u8,
 * rounded to nearest-even (truncate = 0, the GPU kernel) or truncated (truncate = 1, cvtColor_cpu,
 * Utility.cpp:289-298). */
int gsm_cvtcolor(gsm_ctx* ctx, const uint8_t* src3, uint8_t* dst, int rows, int cols, int truncate);

/* depth(r,c) = fB / disparity(r,c), float32; 0 where the disparity is 0 (occluded / no match).  fB = focal length in
 * pixels x baseline: the depth the Q matrix of stereoRectify gives for a rectified rig (Utility.cpp:228-234).  Host
 * pointers.  SURVEY 8(f) row 3. */
int gsm_disparity_to_depth(gsm_ctx* ctx, const uint8_t* disparity, float* depth, int rows, int cols, float fB);

/* Rectification maps for gsm_params.rectify (float32 rows x cols each, host pointers; copied to the device).  They
 * are what Rectify() builds (initUndistortRectifyMap CV_32FC1, Utility.cpp:228-234); sampling follows gsm_remap.
 * Pass NULL maps to drop them. */
int gsm_set_rectification(gsm_ctx* ctx, const float* mapx_left, const float* mapy_left, const float* mapx_right,
                          const float* mapy_right, int rows, int cols);

/* ---- the segment-tree stereo of the reference's STMatching project ------------------------ */
/* stereo_disparity_normal (STMatching/StereoDisparity.cpp:58-90): colour + gradient matching cost
 * (StereoHelper.cpp:75-129) -> segment-tree aggregation (CColorWeight + BuildSegmentTree + Filter,
 * SegmentTree.cpp:38-195) -> winner-take-all (StereoHelper.cpp:131-154) -> (2m+1)^2 median (Toolkit.cpp:33-48) ->
 * disparity * scale.  Images: interleaved 3-channel u8, rows x cols, channel order as cv::imread gives it (B, G, R);
 * at least 3 x 3 pixels (the reference's median asserts below that).  Host pointers, blocking.  Every stage is bit-exact to the reference compiled without FP contraction.
 * Of the tree construction only what is sequential by definition runs on the host, in O(pixels): the two Kruskal
 * passes (adaptive threshold: the outcome depends on the sorted edge order) and the breadth-first ordering; the edge
 * sort, the per-pixel adjacency, cost, tree filter, WTA and median run on the GPU. */
typedef struct gsm_st_params {
  int num_disp;      /* max_dis_level, 1..256 */
  float sigma;       /* range parameter of the edge-weight table exp(-dist / (255 sigma)) (SegmentTree.cpp:141-146) */
  float tau;         /* constant of the merge threshold tau / size (TAU = 1200, Toolkit.h:33); <= 0 selects 1200 */
  int median_radius; /* 3 in the reference (StereoDisparity.cpp:85); 0 = none */
  int scale;         /* final map = disparity * scale, saturated to 255 (StereoDisparity.cpp:87); >= 1 */
  int refined;       /* 0: stereo_disparity_normal (ST-1, the default of STMatching/main.cpp:52); 1:
                        stereo_disparity_iteration (ST-2, StereoDisparity.cpp:92-160): both views aggregated with
                        sigma 0.08, L-R check (:128-147), then a second pass over the tree of CColorDepthWeight
                        (left image, left disparity, mask; SegmentTree.cpp:197-218); needs num_disp <= cols */
} gsm_st_params;
int gsm_segment_tree_stereo(gsm_ctx* ctx, const gsm_st_params* p, const uint8_t* left3, const uint8_t* right3,
                            uint8_t* disparity, int rows, int cols);
/* nframes pairs of one size in one call (left3 / right3: [nframes][rows][cols][3], disparity: [nframes][rows][cols]).
 * The reference's STMatching processes one pair per process (main.cpp:60); its host-bound stage -- one tree per frame --
 * is independent across frames, so the trees are built concurrently on host_threads host threads (0 = one per
 * hardware thread, at most nframes) while the GPU aggregates each frame as soon as its tree is packed.  ST-1 only
 * (p->refined must be 0).  Every frame's map is identical to gsm_segment_tree_stereo on that pair. */
int gsm_segment_tree_stereo_batch(gsm_ctx* ctx, const gsm_st_params* p, const uint8_t* left3, const uint8_t* right3,
                                  uint8_t* disparity, int nframes, int rows, int cols, int host_threads);
/* Stage exports (parity with the reference stage by stage).  GetMatchingCost: float [rows][cols][num_disp]. */
int gsm_st_matching_cost(gsm_ctx* ctx, const uint8_t* left3, const uint8_t* right3, float* cost, int rows, int cols,
                         int num_disp);
/* CColorWeight + BuildSegmentTree (+ Filter when cost != NULL: float [rows][cols][num_disp], aggregated in place).
 * order / father_id / father_dist (each optional, rows*cols entries): the ordered tree in breadth-first order --
 * pixel id, father's pixel id (0 for the root) and quantised weight of the edge to the father (CSegmentTree::m_tree). */
int gsm_st_filter(gsm_ctx* ctx, const uint8_t* image3, float* cost, int rows, int cols, int num_disp, float sigma,
                  float tau, int* order, int* father_id, uint8_t* father_dist);

/* The host stage of the segment tree on its own (no GPU): edge weights in (wr[p]: edge (p, p+1); wu[p]: edge
 * (p, p-cols); u8 as CColorWeight produces them), ordered tree out (as gsm_st_filter); *levels = depth of the tree. */
int gsm_st_build_tree_host(const uint8_t* wr, const uint8_t* wu, int rows, int cols, float tau, int* order,
                           int* father_id, uint8_t* father_dist, int* levels);

/* ---- introspection for benches -------------------------------------------------------------- */
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
long long gsm_launch_count(const gsm_ctx* ctx);
/* Device time (ms, CUDA events on the launching stream) of the fused aggregation+WTA kernel(s) of the
 * most recent gsm_stereo_device call with timing enabled; < 0 if none. */
int gsm_set_kernel_timing(gsm_ctx* ctx, int enabled);
float gsm_last_kernel_ms(gsm_ctx* ctx);
/* FP32/INT issue-peak microbenchmark (FFMA + IADD3 mix): lane-ops per second on this device. */
int gsm_measure_alu_peak(gsm_ctx* ctx, double* lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GSM_H */
