"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed as plumbing).

Two strategies, nothing else:
  * frame batches  -- frames of a stream are independent: contiguous blocks of frames per rank,
                      NO collective on the data path (shard_frames).
  * disparity split -- one very large frame: rank g evaluates d in [d_begin, d_end) over the full frame and
                      leaves a per-pixel packed (cost, d) int64 word; ONE exchange step, an all-reduce(MIN)
                      over NCCL/NVLink, combines the ranks exactly (lowest d wins ties, like the reference's
                      strict '<', BlockMatching.cpp:178), independent of the number of ranks (dsplit_stereo).
                      dsplit_stereo_p2p does the same exchange WITHOUT a collective library: the planes live in
                      symmetric (peer-mapped) memory and one kernel per rank reduces its slice of all planes over
                      NVLink P2P loads, finalizes it and stores it into every rank's map (gsm_reduce_keys_p2p).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

WARP = 32  # the fused kernels evaluate disparities in chunks of 32 (one warp)


def shard_frames(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of frames for `rank`; blocks differ in size by at most one frame."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_disparities(num_disp: int, world: int, rank: int, align: int = WARP) -> Tuple[int, int]:
    """[d_begin, d_end) for `rank`: whole 32-disparity chunks, as even as possible; may be empty."""
    if world < 1 or not (0 <= rank < world) or num_disp < 1:
        raise ValueError("bad shard request")
    chunks = (num_disp + align - 1) // align
    c0, c1 = shard_frames(chunks, world, rank)
    return min(c0 * align, num_disp), min(c1 * align, num_disp)


def dsplit_row_bands(rows: int, cols: int, num_disp: int, world: int, radius: int = 9, sms: int = 148,
                     strip_out_cols: int = 160) -> int:
    """Row bands for a disparity split over `world` ranks: the value of gsm_params.row_bands EVERY rank must pass.

    The fp32 running sums of the guided-filter kernel restart at every band, so the packed minima of the ranks combine
    into a map that is bit-identical to a single-GPU pass only if all of them (and that pass) use the same bands.  The
    choice mirrors the library's own cost model (csrc/gsm_api.cu make_plan: waves of CTAs x rows marched per CTA incl.
    4r warm-up rows, bands of at most 768 rows) for the number of 32-disparity chunks ONE rank evaluates, so that a
    rank with 1/world of the disparities still fills its GPU.  Any value is correct; this one is fast."""
    if world < 1 or rows < 1 or cols < 1 or num_disp < 1:
        raise ValueError("bad request")
    strips = -(-cols // strip_out_cols)
    chunks_rank = -(-(-(-num_disp // WARP)) // world)
    warm, b_min = 4 * radius, -(-rows // 768)
    best, best_cost = b_min, None
    b = b_min
    while b <= 16 and (rows // b >= 32 or b == b_min):
        cost = -(-(strips * chunks_rank * b) // sms) * (-(-rows // b) + warm)
        if best_cost is None or cost < best_cost:
            best, best_cost = b, cost
        b += 1
    return max(1, min(best, rows))


def dsplit_spare_sms(rows: int, cols: int, num_disp: int, world: int, row_bands: int, sms: int = 148,
                     strip_out_cols: int = 160) -> int:
    """SMs the fused kernel of ONE rank of a disparity split leaves idle when its whole grid (strips x 32-disparity
    chunks x row bands, one CTA per SM) fits the GPU in a single wave; 0 otherwise.  DsplitStream runs the peer-memory
    combine on them."""
    strips = -(-cols // strip_out_cols)
    chunks_rank = -(-(-(-num_disp // WARP)) // world)
    ctas = strips * chunks_rank * max(1, row_bands)
    return sms - ctas if 0 < sms - ctas <= 16 else 0


def key_init(mode: int, radius: int) -> int:
    """Initial packed word of the min plane (must be identical on every rank)."""
    if mode == 0:  # SAD: acceptance threshold 50*(2r+1)^2, d = 0  (BlockMatching.cpp:157-158)
        w = 2 * radius + 1
        return (50 * w * w) << 8
    return 0x7FFFFFFFFFFFFF00  # GF: +inf, d = 0


def torch_stream_handle(stream=None) -> int:
    """cudaStream_t for libgsm: torch's current stream; the legacy default stream is passed as
    cudaStreamLegacy (1) because 0 means "the context's own stream" in the C ABI."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream or 1


def dsplit_stereo(partial_keys: Callable, finalize: Callable, keys_left, keys_right, params, world: int, rank: int,
                  group=None, all_reduce: Optional[Callable] = None):
    """Disparity-split evaluation of ONE frame.

    partial_keys(view, d_begin, d_end, keys_tensor) fills keys_tensor (int64 [rows*cols]) with this rank's
    packed minima (gsm_partial_keys_device on a GPU rank); finalize(keys_left, keys_right|None) turns reduced
    keys into the disparity map (gsm_finalize_keys_device).  all_reduce defaults to torch.distributed's MIN.
    Ranks whose range is empty contribute the init word only.
    """
    import torch.distributed as dist
    if all_reduce is None:
        def all_reduce(t):
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    d0, d1 = shard_disparities(params.num_disp, world, rank)
    views = (0, 1) if params.lr_check else (0,)
    for view, keys in zip(views, (keys_left, keys_right)):
        if d1 > d0:
            partial_keys(view, d0, d1, keys)
        else:
            keys.fill_(key_init(params.mode, params.radius))
        if world > 1:
            all_reduce(keys)
    return finalize(keys_left, keys_right if params.lr_check else None)


class PeerPlanes:
    """Per-rank packed-min planes and raw WTA maps in symmetric (peer-mapped) device memory, for dsplit_stereo_p2p.

    Allocated with torch.distributed._symmetric_memory (CUDA IPC over NVLink / NVSwitch): after the rendezvous every
    rank holds device pointers to every other rank's planes.  `views` = 2 keeps a second plane / map for the right view
    (LR check).  `slots` = 2 double-buffers planes and maps so that consecutive frames need ONE cross-rank barrier each
    instead of two (see dsplit_stereo_p2p).  Raises if peer memory is unavailable -- callers fall back to
    dsplit_stereo (NCCL all-reduce).
    """

    def __init__(self, npx: int, group=None, views: int = 1, slots: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        grp = group if group is not None else dist.group.WORLD
        self.npx = int(npx)
        self.views, self.slots = int(views), int(slots)
        self.stride = (self.npx + 15) // 16 * 16  # pixels per (slot, view) plane: keeps every plane 16-byte aligned
        dev = torch.device("cuda", torch.cuda.current_device())
        total = self.slots * self.views * self.stride
        self.keys = symm.empty(total, dtype=torch.int64, device=dev)
        self.disp = symm.empty(total, dtype=torch.uint8, device=dev)
        self.keys_h = symm.rendezvous(self.keys, grp)
        self.disp_h = symm.rendezvous(self.disp, grp)
        self.world, self.rank = self.keys_h.world_size, self.keys_h.rank
        self._kbase = [int(p) for p in self.keys_h.buffer_ptrs]
        self._dbase = [int(p) for p in self.disp_h.buffer_ptrs]
        self.frame = 0  # frames submitted so far (slot = frame % slots)
        # results of the optional post-filters (local, not peer-mapped)
        self.out = torch.empty(self.npx, dtype=torch.uint8, device=dev)
        self.mask = torch.empty(self.npx, dtype=torch.uint8, device=dev)

    def _index(self, slot: int, view: int) -> int:
        return (slot * self.views + view) * self.stride

    def keys_view(self, slot: int, view: int):
        o = self._index(slot, view)
        return self.keys[o:o + self.npx]

    def disp_view(self, slot: int, view: int):
        o = self._index(slot, view)
        return self.disp[o:o + self.npx]

    def key_ptrs(self, slot: int, view: int):
        return [b + 8 * self._index(slot, view) for b in self._kbase]

    def disp_ptrs(self, slot: int, view: int):
        return [b + self._index(slot, view) for b in self._dbase]

    def nvlink_bytes_per_frame(self, views: int = 1) -> int:
        """Bytes this rank moves over NVLink per frame: P2P loads of its 1/world slice of the other ranks' int64
        planes + P2P stores of its u8 slice into the other ranks' maps."""
        sl = self.npx / self.world
        return int(views * (self.world - 1) * sl * (8 + 1))


def dsplit_stereo_p2p(ctx, partial_keys: Callable, planes: PeerPlanes, params, stream_handle: int = 0,
                      rows: Optional[int] = None, cols: Optional[int] = None, final_barrier: bool = True):
    """Disparity-split evaluation of ONE frame with the combine fused over peer memory.

    partial_keys(view, d_begin, d_end, keys_tensor) as in dsplit_stereo; every rank must evaluate its range with the
    SAME gsm_params.row_bands (dsplit_row_bands) for the result to equal the single-GPU map bit for bit.  Every rank
    ends with the full map: the raw left-view WTA map in planes.disp_view(slot, 0) or, with params.lr_check /
    params.median_radius (needs rows, cols), the post-filtered map in planes.out (+ planes.mask), computed by every
    rank on its own copy (gsm_postfilter_device; no further communication).  Returns the result tensor.

    Must run with torch's current stream == the stream behind stream_handle: the cross-rank barriers (symmetric-memory
    signal pads) are enqueued on the current stream.
      barrier A (always): every rank's planes of this frame are complete before anybody reads them over NVLink.  Because
                 each rank enqueues its combine kernel of frame k BEFORE its partial keys of frame k+1, this barrier
                 also proves that every combine of the previous frame has finished: with two plane / map slots
                 (PeerPlanes(slots=2)) nothing written for frame k+1 can race with frame k.
      barrier B (final_barrier=True): every slice of THIS frame has landed in every rank's map.  A stream of frames sets
                 final_barrier=False on all but the last frame and reads frame k's map after barrier A of frame k+1.
    """
    lr = bool(params.lr_check)
    if lr and planes.views < 2:
        raise ValueError("lr_check needs PeerPlanes(views=2)")
    post = lr or params.median_radius > 0
    if post and (rows is None or cols is None or rows * cols != planes.npx):
        raise ValueError("post-filters need rows, cols with rows*cols == planes.npx")
    if planes.slots < 2 and not final_barrier:
        raise ValueError("final_barrier=False needs PeerPlanes(slots=2)")
    slot = planes.frame % planes.slots
    planes.frame += 1
    d0, d1 = shard_disparities(params.num_disp, planes.world, planes.rank)
    views = (0, 1) if lr else (0,)
    for view in views:
        k = planes.keys_view(slot, view)
        if d1 > d0:
            partial_keys(view, d0, d1, k)
        else:
            k.fill_(key_init(params.mode, params.radius))
    planes.keys_h.barrier(channel=0)
    for view in views:
        ctx.reduce_keys_p2p(planes.key_ptrs(slot, view), planes.disp_ptrs(slot, view), planes.rank, planes.npx,
                            stream_handle)
    if final_barrier or post:
        planes.keys_h.barrier(channel=1)
    if not post:
        return planes.disp_view(slot, 0)
    ctx.postfilter_device(planes.disp_view(slot, 0).data_ptr(), planes.disp_view(slot, 1).data_ptr() if lr else 0,
                          planes.out.data_ptr(), planes.mask.data_ptr() if lr else 0, rows, cols, params, stream_handle)
    return planes.out


class DsplitStream:
    """A STREAM of frames through the peer-memory disparity split with the combine off the critical path.

    Two CUDA streams per rank: the compute stream runs gsm_partial_keys_device of frame k+1 while the combine stream
    runs the cross-rank barrier and gsm_reduce_keys_p2p of frame k.  Planes and maps rotate through three slots
    (PeerPlanes(slots=3)), so the only cross-stream waits are two frames old and never block in steady state:
      * compute, before frame k:  "the combine stream of THIS rank has passed the barrier of frame k-2" -- every rank
        enqueues its combine of frame k-3 before that barrier, so nobody still reads the slot frame k overwrites;
        the FUSED kernel of frame k also waits for this rank's combine of frame k-1 (partial_keys is called with that
        event: partial_keys(view, d_begin, d_end, keys, wait_event)), because the fused kernel fills every SM's
        register file and the two would serialise each other; plane packing and the guide statistics run beside it;
      * combine, frame k: this rank's planes of frame k are complete -> barrier (everybody's are) -> reduce my 1/N
        slice of all planes over NVLink, finalize, store it into every rank's map.
    The map of frame k is complete on every rank once the barrier of frame k+1 (or flush()) has passed on the combine
    stream.  Left view only, no post-filters (dsplit_stereo_p2p handles LR check / median per frame).
    """

    def __init__(self, ctx, partial_keys: Callable, planes: PeerPlanes, params, compute_stream, combine_stream,
                 timing: bool = False, spare_sms: int = 0):
        import torch
        self.timing, self._marks = bool(timing), []
        # SMs the fused kernel's grid leaves idle (0 = none / unknown).  With spare SMs the combine is confined to them
        # and runs BESIDE the next frame's fused kernel; without, the fused kernel waits for the combine (they cannot
        # share an SM: the fused kernel owns its whole register file) and only the passes in front of it overlap.
        self.spare_sms = int(spare_sms)
        if planes.slots < 3:
            raise ValueError("DsplitStream needs PeerPlanes(slots=3)")
        if params.lr_check or params.median_radius:
            raise ValueError("DsplitStream combines the left view only")
        self.ctx, self.partial, self.planes, self.params = ctx, partial_keys, planes, params
        self.s_main, self.s_side = compute_stream, combine_stream
        self.h_side = torch_stream_handle(combine_stream)
        self.k = 0
        self._torch = torch
        self._passed = {}  # frame -> event recorded on the combine stream right after that frame's barrier
        self._combined = None  # event recorded after the most recent combine kernel

    def submit(self):
        """Enqueue one frame (partial_keys reads whatever input buffers it was bound to); returns the frame index."""
        torch, pl, k = self._torch, self.planes, self.k
        slot = k % pl.slots
        d0, d1 = shard_disparities(self.params.num_disp, pl.world, pl.rank)
        with torch.cuda.stream(self.s_main):
            ev = self._passed.pop(k - 2, None)
            if ev is not None:
                self.s_main.wait_event(ev)
            keys = pl.keys_view(slot, 0)
            if d1 > d0:
                # the fused kernel owns every register of the SMs it runs on, so it cannot share them with the
                # previous frame's combine: it waits for that combine, the passes in front of it run beside it
                prev = self._combined if self.spare_sms <= 0 else None
                self.partial(0, d0, d1, keys, prev.cuda_event if prev is not None else 0)
            else:
                keys.fill_(key_init(self.params.mode, self.params.radius))
            done = torch.cuda.Event(enable_timing=self.timing)
            done.record(self.s_main)
        with torch.cuda.stream(self.s_side):
            self.s_side.wait_event(done)
            pl.keys_h.barrier(channel=0)
            passed = torch.cuda.Event(enable_timing=self.timing)
            passed.record(self.s_side)
            self._passed[k] = passed
            self.ctx.reduce_keys_p2p(pl.key_ptrs(slot, 0), pl.disp_ptrs(slot, 0), pl.rank, pl.npx, self.h_side,
                                     max_blocks=self.spare_sms)
            self._combined = torch.cuda.Event(enable_timing=self.timing)
            self._combined.record(self.s_side)
        if self.timing:
            self._marks.append((done, passed, self._combined))
        self.k += 1
        return k

    def flush(self):
        """Close the stream: after this (on the compute stream) every submitted frame's map is complete on every rank."""
        torch, pl = self._torch, self.planes
        with torch.cuda.stream(self.s_side):
            pl.keys_h.barrier(channel=1)
            ev = torch.cuda.Event()
            ev.record(self.s_side)
        self.s_main.wait_event(ev)

    def stats(self, skip: int = 3):
        """(mean ms this rank's combine stream waited in the cross-rank barrier, mean ms of the combine kernel) over the
        frames submitted with timing=True, the first `skip` excluded; call after the streams are synchronised."""
        m = self._marks[skip:]
        if not m:
            return None, None
        wait = sum(a.elapsed_time(b) for a, b, _ in m) / len(m)
        red = sum(b.elapsed_time(c) for _, b, c in m) / len(m)
        return wait, red

    def result(self, k: int):
        """u8 map of frame k (valid after the barrier of frame k+1 or flush(), until frame k+3 is submitted)."""
        return self.planes.disp_view(k % self.planes.slots, 0)
