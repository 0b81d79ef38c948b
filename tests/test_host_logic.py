"""CPU tests of the host-side partitioning logic, incl. the N>1 path on gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from gpu_stereo_matching_b200 import dist as gdist  # noqa: E402
from gpu_stereo_matching_b200 import make_params  # noqa: E402


def test_shard_frames_partition():
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [gdist.shard_frames(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_disparities_partition():
    for D in (1, 31, 32, 33, 64, 100, 128, 192, 256):
        for world in (1, 2, 4, 8):
            spans = [gdist.shard_disparities(D, world, r) for r in range(world)]
            covered = []
            for a, b in spans:
                assert 0 <= a <= b <= D
                assert a % 32 == 0 or a == D
                covered += list(range(a, b))
            assert covered == list(range(D))


def test_key_init_matches_reference_threshold():
    assert gdist.key_init(0, 5) == (50 * 121) << 8  # BlockMatching.cpp:157
    assert gdist.key_init(1, 9) == 0x7FFFFFFFFFFFFF00


def _sortable(q32):
    b = q32.view(np.int32).astype(np.int64)
    return b ^ ((b >> 31) & 0x7FFFFFFF)


def _oracle_partial_keys(O, L, R, p, view, d0, d1):
    """CPU stand-in for gsm_partial_keys_device built from the ORACLE (tests only)."""
    h, w = L.shape
    keys = np.full((h, w), gdist.key_init(p.mode, p.radius), np.int64)
    xs = np.arange(w)[None, :]
    for d in range(d0, d1):
        if p.mode == 0:
            sad = O.sad_slice(L, R, p.radius, d).astype(np.int64)
            cand = np.where(xs + d <= w, (sad << 8) | d, np.int64(2 ** 62))
        else:
            q = O.gf_cost_slices(L, R, p.radius, d, 1, view=view)[0].astype(np.float32)
            cand = (_sortable(q) << 32) | d
        keys = np.minimum(keys, cand)
    return keys


def _dsplit_worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "middlebury_gray.npz"))
    L, R = fx["ArtDemo_L"][40:120, 60:220].copy(), fx["ArtDemo_R"][40:120, 60:220].copy()
    p = make_params(mode, 5 if mode == "sad" else 3, 64, lr_check=(mode == "gf"))
    h, w = L.shape
    kl = torch.empty(h * w, dtype=torch.int64)
    kr = torch.empty(h * w, dtype=torch.int64)

    def partial(view, d0, d1, keys):
        keys.copy_(torch.from_numpy(_oracle_partial_keys(O, L, R, p, view, d0, d1).reshape(-1)))

    def finalize(a, b):
        dl = (a.numpy().reshape(h, w) & 0xFF).astype(np.uint8)
        if b is None:
            return dl
        dr = (b.numpy().reshape(h, w) & 0xFF).astype(np.uint8)
        occ, _ = O.lr_check(dl, dr)
        dl[occ != 0] = 0
        return dl

    disp = gdist.dsplit_stereo(partial, finalize, kl, kr, p, world, rank)
    if rank == 0:
        np.save(out_path, disp)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("mode", ["sad", "gf"])
def test_dsplit_world2_gloo(tmp_path, mode, orc, fx):
    """Disparity split over 2 ranks + all-reduce(MIN) on packed int64 words == single-rank result."""
    import torch.multiprocessing as mp
    out = str(tmp_path / f"disp_{mode}.npy")
    mp.spawn(_dsplit_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    disp = np.load(out)
    L, R = fx["ArtDemo_L"][40:120, 60:220].copy(), fx["ArtDemo_R"][40:120, 60:220].copy()
    if mode == "sad":
        ref = orc.sad_wta(L, R, 5, 64)
    else:
        # float32 keys: compare against the oracle evaluated on the same float32-rounded costs
        ql = orc.gf_cost_slices(L, R, 3, 0, 64, view=0).astype(np.float32)
        qr = orc.gf_cost_slices(L, R, 3, 0, 64, view=1).astype(np.float32)
        dl = np.argmin(ql, axis=0).astype(np.uint8)
        dr = np.argmin(qr, axis=0).astype(np.uint8)
        occ, _ = orc.lr_check(dl, dr)
        ref = dl.copy()
        ref[occ != 0] = 0
    assert np.array_equal(disp, ref)


def _frames_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 7
    a, b = gdist.shard_frames(n, world, rank)
    # no collective on the data path: each rank just reports which frames it owns
    mine = torch.zeros(n, dtype=torch.int64)
    mine[a:b] = rank + 1
    dist.all_reduce(mine)  # bookkeeping only (test-side)
    if rank == 0:
        np.save(out_path, mine.numpy())
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "own.npy")
    mp.spawn(_frames_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    own = np.load(out)
    assert own.tolist() == [1, 1, 1, 1, 2, 2, 2]  # every frame owned exactly once, contiguous blocks


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on host cores only (the reference's getDisp through oracle/_ref, or the port):
    one JSON line with the contract's keys, the same metric / unit / config.workload as the B200 arm."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "MDE/s" and line["unit"] == "MDE/s" and line["value"] > 0
    assert line["config"]["workload"].startswith("config3") and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_dsplit_row_bands_fills_the_gpu():
    """Row bands of a disparity split: every rank passes the same value; enough CTAs (strips x chunks x bands) for 148
    SMs at 8 ranks on the 4K config, the single-GPU choice at world 1, never more bands than rows."""
    assert gdist.dsplit_row_bands(2160, 3840, 256, 1) == 3      # 768-row cap of the fp32 running sums
    b8 = gdist.dsplit_row_bands(2160, 3840, 256, 8)
    assert 24 * 1 * b8 >= 140 and b8 >= 3
    for world in (1, 2, 3, 4, 8):
        for (h, w, d) in ((2160, 3840, 256), (720, 1280, 128), (40, 50, 33), (1, 16, 1)):
            b = gdist.dsplit_row_bands(h, w, d, world)
            assert 1 <= b <= max(1, h)
    with pytest.raises(ValueError):
        gdist.dsplit_row_bands(0, 10, 10, 1)


def test_python_binding_rejects_mismatched_buffers():
    """The C ABI takes raw pointers and ONE rows x cols: the binding must refuse arrays whose shapes, dtypes or layout
    disagree instead of letting a copy run past the end of a host buffer."""
    from gpu_stereo_matching_b200 import api
    a = np.zeros((4, 8, 8), np.uint8)
    assert api._same_shape(3, left=a, right=a.copy(), out=a.copy(), mask_out=None) == (4, 8, 8)
    with pytest.raises(ValueError):
        api._same_shape(3, left=a, right=a[:, :, :7].copy())
    with pytest.raises(ValueError):
        api._same_shape(3, left=a, right=a[:, ::2, :])          # not contiguous
    with pytest.raises(TypeError):
        api._same_shape(3, left=a, right=a.astype(np.int16))
    with pytest.raises(ValueError):
        api._same_shape(3, left=a, out=a[0])                    # wrong rank
