#!/usr/bin/env python
"""bench.py -- headline benchmark of the BlockMatching hot path on B200.

Contract (see the task description): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
  step      = one pass of the fused hot path (guided-filter stereo: AD -> GF aggregation -> WTA) over one batch
              of synthetic rectified 1280x720 pairs, 128 disparities, r = 9  (BASELINE config 3, the shape the
              metric "MDE/s & fps at 720p x 128d" is quoted on).
  value     = whole-job MDE/s (rows*cols*D*frames / s / 1e6) with the frames already resident in HBM.
  e2e       = the same metric through the host-buffer C-ABI call gsm_stereo_batch (pinned host memory,
              H2D + kernels + D2H inside the timed region).
  N > 1     = frame batches sharded across ranks, no collective on the data path ("weak" scaling: every rank
              processes `--frames` frames per step).
  --impl reference = the reference's own CPU BlockMatching (getDisp, compiled unmodified into oracle/_ref)
              timed on the host cores on a bounded sample of the same frames.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, D, R_GF = 720, 1280, 128, 9
WORKLOAD = "config3: rectified 1280x720 synthetic stereo stream, 128 disparities, guided filter r=9, no LR"
GF_OPS_PER_DE = 28   # algorithmic lane-ops per pixel*disparity, guided-filter mode
SAD_OPS_PER_DE = 7   # SAD mode (AD 1 + box sum 4 + WTA 2)
HBM_BYTES_PER_PX = 3  # read L + R, write disparity


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (NVML, 10 ms period; nvidia-smi as a fallback)."""

    _REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._idx = gpu_index
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = int(get(self._h))
            for bit, name in self._REASONS.items():
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self._idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for nme, v in zip(names, f[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(nme)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "source": "nvml" if self._nvml is not None else "nvidia-smi", "reasons": sorted(self.reasons)}


def _reference_arm(args, rank, world):
    """CPU baseline: the reference's own getDisp on the box's host cores (bounded sample, all cores)."""
    if rank != 0:
        return
    from oracle import oracle as O
    from gpu_stereo_matching_b200 import data as gdata
    kind = "reference" if O.have_ref() else "port"
    cores = os.cpu_count() or 1
    band_rows, r_ref = 24, 5
    L, R, _ = gdata.synthetic_pair(H, W, 1234)
    bands = [(L[i * band_rows:(i + 1) * band_rows].copy(), R[i * band_rows:(i + 1) * band_rows].copy())
             for i in range(min(cores, H // band_rows))]
    fn = (lambda a, b: O.ref().ref_getDisp(a, b, band_rows, W, r_ref, D, np.empty_like(a))) if kind == "reference" \
        else (lambda a, b: O.sad_wta(a, b, r_ref, D, direct=True))

    def step():
        ths = [threading.Thread(target=fn, args=b) for b in bands]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    de = len(bands) * band_rows * W * D
    with O.quiet_stdout():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    v = de / dt / 1e6
    sample = (f"{len(bands)} bands of {band_rows}x{W} px x {D} d of a config-3 frame, one band per thread; "
              f"reference getDisp (SAD r={r_ref}, Caller.cpp:19 -- the reference has no guided filter)")
    # the guided-filter aggregation itself has no reference implementation: our float64 oracle port (OpenMP, all cores)
    # on a crop of the same frame, so the GF-vs-GF ratio can be read from the same line
    gf_rows = 120
    a, b = L[:gf_rows].copy(), R[:gf_rows].copy()
    t0 = time.perf_counter()
    O.gf_wta(a, b, R_GF, D)
    gdt = time.perf_counter() - t0
    gf_port = {"value": gf_rows * W * D / gdt / 1e6, "unit": "MDE/s", "cores": cores, "kind": "port",
               "sample": f"oracle GF r={R_GF} float64 (OpenMP) on a {gf_rows}x{W} px x {D} d crop of a config-3 frame"}
    _emit(({
        "impl": "reference", "metric": "MDE/s", "value": v, "unit": "MDE/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32", "data": "synthetic", "fps": v * 1e6 / (H * W * D),
        "config": {"workload": WORKLOAD, "reference_arm": sample},
        "cpu_baseline": {"value": v, "unit": "MDE/s", "cores": len(bands), "kind": kind, "sample": sample,
                         "gf_port": gf_port},
        "e2e": {"value": v, "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _cpu_baseline(frames_L, frames_R):
    """Bounded CPU sample for the product arm's JSON line (rank 0, N=1): the reference getDisp (1 core, as
    shipped) and our GF oracle port (all cores) on crops of the first frame."""
    from oracle import oracle as O
    out = {}
    L, R = frames_L[0], frames_R[0]
    rows = 120
    a, b = L[:rows].copy(), R[:rows].copy()
    kind = "reference" if O.have_ref() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        O.ref_getDisp(a, b, 5, D)
    else:
        O.sad_wta(a, b, 5, D, direct=True)
    dt = time.perf_counter() - t0
    out.update({"value": rows * W * D / dt / 1e6, "unit": "MDE/s", "cores": 1, "kind": kind,
                "sample": f"reference getDisp (SAD r=5, single-threaded as shipped) on a {rows}x{W} px x {D} d crop of frame 0"})
    rows = 240
    a, b = L[:rows].copy(), R[:rows].copy()
    t0 = time.perf_counter()
    O.gf_wta(a, b, R_GF, D)
    dt = time.perf_counter() - t0
    out["gf_port"] = {"value": rows * W * D / dt / 1e6, "unit": "MDE/s", "cores": os.cpu_count(), "kind": "port",
                      "sample": f"oracle GF r={R_GF} float64 (OpenMP) on a {rows}x{W} px x {D} d crop of frame 0"}
    return out


def _stream_frames(gdata, n, seed0):
    """n distinct config-3 frames (rectified_stream, seeds seed0 + i), cached in /tmp between runs."""
    path = os.path.join("/tmp", f"gsm_bench_c3_{H}x{W}_{seed0}_{n}.npz")
    try:
        z = np.load(path)
        if z["L"].shape == (n, H, W):
            return z["L"], z["R"], bool(z["rectified"])
    except Exception:
        pass
    L, R, rectified = gdata.rectified_stream(n, seed0=seed0)
    try:
        np.savez(path + f".{os.getpid()}.tmp.npz", L=L, R=R, rectified=np.array(rectified))
        os.replace(path + f".{os.getpid()}.tmp.npz", path)
    except Exception:
        pass
    return L, R, rectified


def _time_device_steps(torch, stream, flush, fn, steps):
    """Mean ms of `steps` runs of fn() on `stream`, CUDA events on that stream, L2 flushed between runs."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in evs:
            flush.fill_(1)
            a.record(stream)
            fn()
            b.record(stream)
        stream.synchronize()
    return float(sum(a.elapsed_time(b) for a, b in evs) / steps)


def _sad_record(torch, ctx, g, stream, sh, flush, Ld, Rd, Dd, Lh, Rh, Dh, n, steps, alu_peak):
    """Like-for-like with the reference arm (getDisp is SAD r=5): the bit-exact SAD kernel on the same frames --
    device-resident value, end-to-end value through the host-buffer call, roofline at 7 lane-ops/DE."""
    import time as _t
    r = 5
    p = g.make_params("sad", r, D)
    fn = lambda: ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, H, W, p, sh)
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
    stream.synchronize()
    ctx.set_kernel_timing(True)
    kms = []
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for a, b in evs:
            flush.fill_(1)
            a.record(stream)
            fn()
            b.record(stream)
            stream.synchronize()
            kms.append(ctx.last_kernel_ms())
    ctx.set_kernel_timing(False)
    ms = float(sum(a.elapsed_time(b) for a, b in evs) / steps)
    kernel_ms = float(np.mean(kms))
    Ln, Rn, Dn = Lh.numpy(), Rh.numpy(), Dh.numpy()
    for _ in range(2):
        ctx.stereo_batch_async(Ln, Rn, p, out=Dn)
    ctx.sync()
    t0 = _t.perf_counter()
    for _ in range(steps):
        ctx.stereo_batch_async(Ln, Rn, p, out=Dn)
    ctx.sync()
    e2e_ms = (_t.perf_counter() - t0) / steps * 1e3
    return {"ms": ms, "e2e_ms": e2e_ms, "kernel_ms": kernel_ms, "radius": r}


def _segment_tree_record(ctx, gdata, with_reference):
    """SURVEY 8f row 4 beside the headline: the reference's second pipeline (STMatching stereo_disparity_normal) through
    gsm_segment_tree_stereo from HOST buffers (wall clock: upload, GPU stages, host tree build, read-back), and -- the
    cpu_baseline leg -- the reference's own code (oracle/_ref/libsegref.so, one core) on the same pair."""
    import time
    h, w, D = 370, 463, 64  # the Middlebury third-size format of the reference's data sets
    L, R = gdata.synthetic_color_pair(h, w, 7, dmax=D - 8)
    for _ in range(2):
        disp = ctx.segment_tree_stereo(L, R, D, scale=1)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx.segment_tree_stereo(L, R, D, scale=1)
    ms = (time.perf_counter() - t0) / reps * 1e3
    rec = {"what": "STMatching stereo_disparity_normal (cost, segment tree, tree filter, WTA, 7x7 median) from host buffers",
           "workload": f"synthetic colour pair {w}x{h} x {D} d", "ms_per_pair": ms, "pairs_per_s": 1e3 / ms,
           "value": h * w * D / ms / 1e3, "unit": "MDE/s"}
    nb = 32  # a batch: the per-frame trees are built concurrently on the host threads, two frames share the GPU
    Lb, Rb = np.stack([L] * nb), np.stack([R] * nb)
    ctx.segment_tree_stereo_batch(Lb, Rb, D)
    t0 = time.perf_counter()
    bd = ctx.segment_tree_stereo_batch(Lb, Rb, D)
    bms = (time.perf_counter() - t0) * 1e3 / nb
    rec["batch"] = {"pairs_per_call": nb, "ms_per_pair": bms, "pairs_per_s": 1e3 / bms, "host_threads": os.cpu_count(),
                    "identical_to_single_calls": bool(all(np.array_equal(bd[i], disp) for i in range(nb)))}
    if with_reference:
        from oracle import oracle as O
        if O.have_segref():
            with O.quiet_stdout():
                ref = O.ref_st_routine(L, R, D, 1, 0.1)
                t0 = time.perf_counter()
                for _ in range(2):
                    O.ref_st_routine(L, R, D, 1, 0.1)
                rms = (time.perf_counter() - t0) / 2 * 1e3
            rec["cpu_reference"] = {"ms_per_pair": rms, "kind": "reference", "cores": 1,
                                    "identical_to_product": bool(np.array_equal(ref, disp))}
    return rec


def _dsplit_record(args, torch, dist, g, gdata, rank, world, local_rank):
    """BASELINE config 5 in front of the driver (N > 1): one 3840x2160 pair, 256 disparities, GF r=9, split by disparity
    range over the ranks; packed (cost, d) planes combined over NVLink peer memory (gsm_reduce_keys_p2p), one cross-rank
    barrier per frame in a stream of frames.  Speed-up is against the same process's single-GPU pass with ITS best
    (automatic) row bands; the map is checked bit for bit against the single-GPU pass with the SAME row bands and
    against the NCCL all-reduce combine."""
    from gpu_stereo_matching_b200.dist import (DsplitStream, PeerPlanes, dsplit_row_bands, dsplit_spare_sms,
                                              dsplit_stereo, torch_stream_handle)
    h, w, d = 2160, 3840, 256
    path = "/tmp/gsm_bench_c5_pair.npz"
    try:
        z = np.load(path)
        L, R = z["L"], z["R"]
    except Exception:
        L, R, _ = gdata.synthetic_pair(h, w, 3000, dmax=250)
        if rank == 0:
            try:
                np.savez(path + ".tmp.npz", L=L, R=R)
                os.replace(path + ".tmp.npz", path)
            except Exception:
                pass
    Ld, Rd = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    ctx = g.StereoContext(h, w, d, 1, device=local_rank)
    stream = torch.cuda.Stream()
    sh = torch_stream_handle(stream)
    bands = dsplit_row_bands(h, w, d, world)
    p = g.make_params("gf", 9, d, row_bands=bands)
    steps = max(args.steps, 10)

    def partial(view, d0, d1, keys, wait_event=0):
        ctx.partial_keys_device(Ld.data_ptr(), Rd.data_ptr(), keys.data_ptr(), h, w,
                                g.make_params("gf", 9, d, row_bands=bands, d_begin=d0, d_end=d1), view, sh, wait_event)

    rec = {"workload": "config5: one synthetic 3840x2160 pair, 256 disparities, GF r=9, split by disparity range",
           "row_bands": bands, "steps": steps}
    planes, pipe, combine = None, None, "nccl all-reduce(MIN) of the int64 packed (cost,d) plane"
    try:
        planes = PeerPlanes(h * w, views=1, slots=3)
        spare = dsplit_spare_sms(h, w, d, world, bands, torch.cuda.get_device_properties(local_rank).multi_processor_count)
        rec["spare_sms_running_the_combine"] = spare
        pipe = DsplitStream(ctx, partial, planes, p, stream, torch.cuda.Stream(), timing=True, spare_sms=spare)
        combine = ("peer memory over NVLink: reduce-scatter + finalize + all-gather of the u8 map in one kernel "
                   "(gsm_reduce_keys_p2p) on a second stream, overlapping the next frame's kernels; one cross-rank "
                   "barrier per frame (dist.DsplitStream)")
    except Exception as e:  # noqa: BLE001
        rec["peer_memory_unavailable"] = f"{type(e).__name__}: {e}"
    kl = torch.empty(h * w, dtype=torch.int64, device="cuda")
    Dn = torch.empty(h * w, dtype=torch.uint8, device="cuda")

    def nccl_step():
        dsplit_stereo(partial, lambda a, b: ctx.finalize_keys_device(a.data_ptr(), 0, Dn.data_ptr(), 0, h, w, p, sh),
                      kl, None, p, world, rank)

    def run(nsteps):
        """a stream of nsteps frames; returns the last frame's map (complete on every rank when the stream ends)"""
        if pipe is not None:
            last = None
            for _ in range(nsteps):
                last = pipe.submit()
            pipe.flush()
            return pipe.result(last)
        for _ in range(nsteps):
            nccl_step()
        return Dn

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        run(3)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        res = run(steps)
        e1.record(stream)
    sync_all()
    ms_n = e0.elapsed_time(e1) / steps
    map_n = res.clone()
    bar_ms, red_ms = pipe.stats(skip=6) if pipe is not None else (None, None)  # the timed frames only
    # the same GPU alone: full disparity range, automatic bands (its best) and the split's bands (for bit-identity)
    Dd1 = torch.empty(h * w, dtype=torch.uint8, device="cuda")
    single = {}
    for name, b in (("auto", 0), ("same_bands", bands)):
        p1 = g.make_params("gf", 9, d, row_bands=b)
        fn = lambda: ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd1.data_ptr(), 0, 1, h, w, p1, sh)
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
            a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(5):
                fn()
            bb.record(stream)
        stream.synchronize()
        single[name] = a.elapsed_time(bb) / 5
        if name == "auto":
            diff_auto = int((Dd1 != map_n).sum().item())
    identical = int(torch.equal(Dd1, map_n))
    # the NCCL all-reduce combine must give the same map
    with torch.cuda.stream(stream):
        nccl_step()
    stream.synchronize()
    same_nccl = int(torch.equal(Dn, map_n))
    t = torch.tensor([ms_n, single["auto"], single["same_bands"], bar_ms or 0.0, red_ms or 0.0], dtype=torch.float64,
                     device="cuda")
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    bar_max, red_max = float(t[3].item()), float(t[4].item())
    bar_min, red_min = float(tmin[3].item()), float(tmin[4].item())
    t = t[:3]
    flags = torch.tensor([identical, same_nccl], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ms_n, ms1_auto, ms1_same = [float(x) for x in t.tolist()]
    de = h * w * d
    rec.update({
        "n_gpus": world, "ms_per_step": ms_n, "value": de / (ms_n * 1e-3) / 1e6, "unit": "MDE/s", "scaling": "strong",
        "single_gpu_ms_auto_bands": ms1_auto, "single_gpu_ms_same_bands": ms1_same,
        "speedup_vs_single_gpu": ms1_auto / ms_n, "efficiency": ms1_auto / ms_n / world,
        "combine": combine,
        "combine_stream_ms": {"barrier_wait_min_max_over_ranks": [bar_min, bar_max],
                              "reduce_kernel_min_max_over_ranks": [red_min, red_max],
                              "note": "CUDA events on the combine stream: time between this rank's planes being complete and "
                                      "the cross-rank barrier releasing (= waiting for the slowest rank), and the "
                                      "gsm_reduce_keys_p2p kernel (NVLink P2P loads of 7/8 of the slice, 16-byte P2P stores)"},
        "map_identical_to_single_gpu_same_bands_on_all_ranks": bool(int(flags[0].item())),
        "pixels_differing_from_single_gpu_auto_bands": diff_auto,
        "combine_check": ("peer-memory map == all-reduce map on all ranks" if int(flags[1].item()) else
                          "MISMATCH between the peer-memory and the all-reduce combine"),
        "nvlink_bytes_per_rank_per_step": planes.nvlink_bytes_per_frame() if planes is not None else int(2 * h * w * 8 * (world - 1) / world),
        "result_checksum": int(map_n.to(torch.int64).sum().item()),
    })
    ctx.close()
    return rec


def _extra_config(args, rank, world, local_rank):
    """BASELINE configs 4 and 5 (not the contract line; same JSON shape, `config.workload` says which)."""
    import torch
    import torch.distributed as dist
    import gpu_stereo_matching_b200 as g
    from gpu_stereo_matching_b200 import data as gdata
    from gpu_stereo_matching_b200.dist import dsplit_stereo, shard_frames, torch_stream_handle
    stream = torch.cuda.Stream()
    sh = torch_stream_handle(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == "c2":
        # all nine Middlebury sets (tests/golden/middlebury_gray.npz), D = 64, GF r = 9 with L-R check: three image
        # sizes, ONE mixed-size batch call (gsm_stereo_device_v: one launch per stage over a per-frame geometry table)
        fx = np.load(os.path.join(ROOT, "tests", "golden", "middlebury_gray.npz"))
        sets = ["Art", "Books", "Computer", "Dolls", "Drumsticks", "Dwarves", "Laundry", "Moebius", "Reindeer"]
        p = g.make_params("gf", 9, 64, lr_check=True)
        ctx = g.StereoContext(370, 463, 64, 9, device=local_rank)
        shapes = [fx[n_ + "_L"].shape for n_ in sets]
        cat = lambda key: torch.from_numpy(np.concatenate([fx[n_ + key].reshape(-1) for n_ in sets])).cuda()
        Ld, Rd = cat("_L"), cat("_R")
        Dd, Md = torch.empty_like(Ld), torch.empty_like(Ld)
        step = lambda: ctx.stereo_device_v(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), Md.data_ptr(), shapes, p, sh)
        h, w, d = 370, 463, 64
        de_step, scaling = int(Ld.numel()) * 64 * world, "weak"
        workload = ("config2: the nine Middlebury third-size sets (463/447/443 x 370), 64 disparities, GF r=9 with L-R check, "
                    "ONE mixed-size batch call")
        par = f"replicated x{world}" if world > 1 else "1 GPU"
    elif args.config == "c4":
        h, w, d, n = 1080, 1920, 192, min(args.frames, 16)
        p = g.make_params("gf", 9, d, lr_check=True, median_radius=3)
        f0, _ = shard_frames(n * world, world, rank)
        Lu, Ru = gdata.synthetic_batch(2, h, w, 2000 + f0, dmax=180)
        Ld = torch.from_numpy(np.tile(Lu, ((n + 1) // 2, 1, 1))[:n]).cuda()
        Rd = torch.from_numpy(np.tile(Ru, ((n + 1) // 2, 1, 1))[:n]).cuda()
        Dd, Md = torch.empty_like(Ld), torch.empty_like(Ld)
        ctx = g.StereoContext(h, w, d, n, device=local_rank)
        step = lambda: ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), Md.data_ptr(), n, h, w, p, sh)
        de_step, scaling = n * h * w * d * world, "weak"
        workload = f"config4: {n} synthetic 1920x1080 pairs per GPU, 192 disparities, GF r=9, LR check + 7x7 median"
        par = f"frame-batch x{world} (no collective)"
    else:
        h, w, d = 2160, 3840, 256
        from gpu_stereo_matching_b200.dist import dsplit_row_bands
        c5_bands = dsplit_row_bands(h, w, d, world)  # every rank uses the same bands: the map is N-independent
        p = g.make_params("gf", 9, d, row_bands=c5_bands)
        L, R, _ = gdata.synthetic_pair(h, w, 3000, dmax=250)
        Ld, Rd = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
        Dd = torch.empty_like(Ld)
        kl = torch.empty(h * w, dtype=torch.int64, device="cuda")
        ctx = g.StereoContext(h, w, d, 1, device=local_rank)

        def partial(view, d0, d1, keys):
            pp = g.make_params("gf", 9, d, row_bands=c5_bands, d_begin=d0, d_end=d1)
            ctx.partial_keys_device(Ld.data_ptr(), Rd.data_ptr(), keys.data_ptr(), h, w, pp, view, sh)

        def finalize(a, b):
            ctx.finalize_keys_device(a.data_ptr(), 0, Dd.data_ptr(), 0, h, w, p, sh)

        step = lambda: dsplit_stereo(partial, finalize, kl, None, p, world, rank)
        p2p_planes = None
        par = f"disparity-split x{world}, one all-reduce(MIN) of the int64 packed (cost,d) plane over NCCL/NVLink"
        if world > 1 and args.combine == "p2p":
            # the combine fused over peer memory: reduce-scatter + finalize + all-gather of the u8 map in ONE kernel
            # of P2P loads / stores (gsm_reduce_keys_p2p); falls back to the NCCL all-reduce when symmetric memory
            # cannot be set up on this box
            try:
                from gpu_stereo_matching_b200.dist import PeerPlanes, dsplit_stereo_p2p
                planes = PeerPlanes(h * w, views=1, slots=2)
                step = lambda: dsplit_stereo_p2p(ctx, partial, planes, p, sh)
                p2p_planes = planes
                par = (f"disparity-split x{world}, packed (cost,d) planes combined over NVLink peer memory "
                       "(one reduce-scatter + finalize + all-gather kernel, gsm_reduce_keys_p2p)")
            except Exception as e:  # noqa: BLE001
                par += f" [peer memory unavailable: {type(e).__name__}: {e}]"
        de_step, scaling = h * w * d, "strong"
        workload = "config5: one synthetic 3840x2160 pair, 256 disparities, GF r=9, split by disparity range"
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    cfg_extra = {}
    if args.config == "c5" and p2p_planes is not None:
        # outside the timed region: the peer-memory combine must give the map of the NCCL all-reduce path on every rank
        with torch.cuda.stream(stream):
            dsplit_stereo(partial, finalize, kl, None, p, world, rank)
        stream.synchronize()
        last = p2p_planes.disp_view((p2p_planes.frame - 1) % p2p_planes.slots, 0)
        same = torch.tensor([int(torch.equal(last.view(-1), Dd.view(-1)))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        cfg_extra["combine_check"] = ("peer-memory map == all-reduce map on all ranks" if int(same.item()) == 1
                                      else "MISMATCH between the peer-memory and the all-reduce combine")
        Dd.copy_(last.view_as(Dd))
    if rank == 0:
        v = de_step / (ms * 1e-3) / 1e6
        _emit(({"metric": "MDE/s", "value": v, "unit": "MDE/s", "n_gpus": world, "steps": args.steps,
                          "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": scaling,
                          "vs_baseline": None, "dtype": "u8/int32+fp32", "data": "synthetic",
                          "fps": v * 1e6 / (h * w * d) * (1 if args.config == "c5" else 1),
                          "config": {"workload": workload, "parallelism": par, "l2": "inputs + statistic planes exceed L2",
                                     **cfg_extra},
                          "result_checksum": int(Dd.to(torch.int64).sum().item()), "clocks": clk.summary()}))
    ctx.close()


_JSON_FD = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout: keep a private copy of fd 1 for it and point fd 1 at stderr, so that
    library chatter (NCCL prints its version on stdout when NCCL_DEBUG is set; the reference prints phase timings)
    cannot land in front of the line."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dsplit", action="store_true", help="N > 1: skip the config-5 disparity-split record")
    ap.add_argument("--combine", default="p2p", choices=["p2p", "nccl"],
                    help="config c5, N > 1: combine the ranks' planes over peer memory (default) or with an NCCL all-reduce")
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"],
                    help="c3 (default, the contract workload): 720p x128 GF frame batches; c4: 1080p x192 GF+LR+median "
                         "frame batches; c5: one 3840x2160 x256 GF pair split by disparity range across the ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _protect_stdout()
    if args.impl == "reference":
        _reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gpu_stereo_matching_b200 as g
    from gpu_stereo_matching_b200 import data as gdata
    from gpu_stereo_matching_b200.dist import shard_frames, torch_stream_handle

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.config != "c3":
        _extra_config(args, rank, world, local_rank)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    n = args.frames
    # every rank owns its own block of the global stream (frame sharding, no collective on the data path)
    f0, f1 = shard_frames(n * world, world, rank)
    # n DISTINCT synthetic frames per rank (seeds 1234 + global frame index), generated once and cached under /tmp
    # (host work outside the timed region)
    Lu, Ru, rectified = _stream_frames(gdata, n, 1234 + f0)
    Lh = torch.from_numpy(Lu).pin_memory()
    Rh = torch.from_numpy(Ru).pin_memory()
    Dh = torch.empty_like(Lh).pin_memory()
    # capacity of two submissions: the host path double-buffers HALF the capacity per chunk, so a submission of n
    # frames is uploaded, computed (one launch of n frames, like the device-resident step) and downloaded as one chunk
    # while the previous / next submission's copies run on their own streams
    ctx = g.StereoContext(H, W, D, 2 * min(n, 32), device=local_rank)
    p = g.make_params("gf", R_GF, D, row_bands=0)
    Ld, Rd = Lh.cuda(non_blocking=True), Rh.cuda(non_blocking=True)
    Dd = torch.empty_like(Ld)
    stream = torch.cuda.Stream()
    sh = torch_stream_handle(stream)
    de_step = n * H * W * D

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, H, W, p, sh)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- device-resident timing --------------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
    barrier()
    ctx.set_kernel_timing(True)
    l0 = ctx.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kms = []
    with ClockSampler(local_rank) as clk:
        with torch.cuda.stream(stream):
            for a, b in evs:
                flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
                a.record(stream)
                step_device()
                b.record(stream)
                stream.synchronize()
                kms.append(ctx.last_kernel_ms())
        barrier()
    launches = ctx.launch_count - l0
    ctx.set_kernel_timing(False)
    ms = float(sum(a.elapsed_time(b) for a, b in evs) / args.steps)
    kernel_ms = float(np.mean(kms))
    # ---- end-to-end through the host-buffer C-ABI call -----------------------------------------------------
    # gsm_stereo_batch_async per step (pinned host buffers -> H2D -> kernels -> D2H into pinned host memory), as a
    # capture loop would drive it; the timed region ends with gsm_sync(), i.e. when the last result is on the host.
    Ln, Rn, Dn = Lh.numpy(), Rh.numpy(), Dh.numpy()
    for _ in range(2):
        ctx.stereo_batch_async(Ln, Rn, p, out=Dn)
    ctx.sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.stereo_batch_async(Ln, Rn, p, out=Dn)
    ctx.sync()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    barrier()
    checksum = int(Dh.numpy().astype(np.uint64).sum())
    alu_peak = ctx.measure_alu_peak()
    sad = _sad_record(torch, ctx, g, stream, sh, flush, Ld, Rd, Dd, Lh, Rh, Dh, n, args.steps, alu_peak)
    barrier()
    dsplit = None
    if world > 1 and not args.no_dsplit:
        del flush
        dsplit = _dsplit_record(args, torch, dist, g, gdata, rank, world, local_rank)

    # ---- max over ranks ----------------------------------------------------------------------------------
    t = torch.tensor([ms, e2e_ms, kernel_ms, sad["ms"], sad["e2e_ms"], sad["kernel_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms, sad_ms, sad_e2e_ms, sad_kernel_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        peaks, peak_src = _peaks()
        value = de_step * world / (ms * 1e-3) / 1e6
        e2e = de_step * world / (e2e_ms * 1e-3) / 1e6
        de_s_kernel = de_step / (kernel_ms * 1e-3)
        traffic, traffic_note = None, "no ncu capture found under profiles/"
        try:
            with open(os.path.join(ROOT, "profiles", "gf_wta_traffic.json")) as f:
                tr = json.load(f)
            traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["frames_per_launch"] * min(n, 32)
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, scaled by frames per launch, from "
                            + tr["source"] + "; algorithmic bytes are 3 B/px -- the excess is the disparity-independent "
                            "guide-statistic planes (32 B/px, written once per frame, re-read by each 32-disparity chunk) "
                            "and the 8 B/px packed-min plane; DRAM throughput is <1% of peak")
        except Exception:
            pass
        line = {
            "metric": "MDE/s", "value": value, "unit": "MDE/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32+fp32", "data": "synthetic",
            "fps": value * 1e6 / (H * W * D),
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": n, "distinct_frames_per_gpu": n, "rows": H, "cols": W, "num_disp": D,
                       "mode": "gf", "radius": R_GF, "rectified_with_calib_maps": bool(rectified),
                       "parallelism": f"frame-batch x{world} (no collective)", "l2": "flushed between timed steps (256 MiB fill)",
                       "result_checksum": checksum},
            "e2e": {"value": e2e, "unit": "MDE/s", "fps": e2e * 1e6 / (H * W * D), "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * n * H * W, "d2h_bytes_per_step": n * H * W},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {
                "bound": "alu", "kernel": "gf3_wta_kernel<9,16,12,32>",
                "achieved": GF_OPS_PER_DE * de_s_kernel / 1e12, "peak": alu_peak / 1e12, "unit": "Tlane-op/s",
                "frac": GF_OPS_PER_DE * de_s_kernel / alu_peak,
                "peak_source": "FFMA+IADD3 issue-peak microbenchmark run live on this GPU (gsm_measure_alu_peak); "
                               "MEASURED_PEAKS.json has no ALU figure",
                "algorithmic_ops_per_de": GF_OPS_PER_DE, "kernel_ms_per_step": kernel_ms,
                "kernel_share_of_step": kernel_ms / ms,
                "hbm": {"achieved": HBM_BYTES_PER_PX * n * H * W / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": HBM_BYTES_PER_PX * n * H * W / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "peak_source": peak_src},
                "traffic": traffic,
                "traffic_note": traffic_note,
            },
        }
        sad_de_s_kernel = de_step / (sad_kernel_ms * 1e-3)
        line["sad"] = {
            "what": "the reference's own algorithm (getDisp: SAD r=5, bit-exact) on the same frames -- like-for-like "
                    "with the reference arm / cpu_baseline",
            "radius": 5, "value": de_step * world / (sad_ms * 1e-3) / 1e6, "unit": "MDE/s",
            "fps": n * world / (sad_ms * 1e-3), "ms_per_step": sad_ms,
            "e2e": {"value": de_step * world / (sad_e2e_ms * 1e-3) / 1e6, "unit": "MDE/s", "ms_per_step": sad_e2e_ms,
                    "fps": n * world / (sad_e2e_ms * 1e-3)},
            "roofline": {"bound": "alu", "kernel": "sad_wta_kernel<5,16>", "algorithmic_ops_per_de": SAD_OPS_PER_DE,
                         "achieved": SAD_OPS_PER_DE * sad_de_s_kernel / 1e12, "peak": alu_peak / 1e12,
                         "unit": "Tlane-op/s", "frac": SAD_OPS_PER_DE * sad_de_s_kernel / alu_peak,
                         "kernel_ms_per_step": sad_kernel_ms},
        }
        if dsplit is not None:
            line["dsplit"] = dsplit
        if world == 1:
            try:  # a side record: it must never cost the headline line
                line["segment_tree"] = _segment_tree_record(ctx, gdata, with_reference=not args.no_cpu_baseline)
            except Exception as e:  # noqa: BLE001
                line["segment_tree"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            with O.quiet_stdout():
                cb = _cpu_baseline(Lu, Ru)
            line["cpu_baseline"] = cb
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
