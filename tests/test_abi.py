"""CPU tests: libgsm.so loads, exports every symbol include/gsm.h declares, the compat header compiles."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gpu_stereo_matching_b200 import lib
    L = lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libgsm.so lacks {n}"
    # the ctypes table covers the header exactly
    assert sorted(lib.SYMBOLS) == names


def test_version_and_error_strings():
    from gpu_stereo_matching_b200 import lib
    L = lib.load()
    assert b"sm_100a" in L.gsm_version()
    assert isinstance(L.gsm_last_error(), bytes)


def test_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device gsm_create must fail with a message, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import gpu_stereo_matching_b200 as g
    with pytest.raises(g.GsmError) as ei:
        g.StereoContext(64, 64, 16, 1)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(g.GsmError):
        g.blockMatching_gpu(np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint8), 2, 4)


def test_product_never_imports_oracle():
    """The shipped package and the CUDA sources must not reference oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "gpu_stereo_matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "libref" not in txt, f


def test_params_struct_layout():
    from gpu_stereo_matching_b200 import lib, make_params
    assert ctypes.sizeof(lib.GsmParams) == 10 * 4
    p = make_params("gf", 9, 128, lr_check=True, median_radius=3)
    assert (p.mode, p.radius, p.num_disp, p.lr_check, p.median_radius) == (1, 9, 128, 1, 3)


def test_compat_header_compiles_and_links(tmp_path):
    """The reference-signature wrapper (include/gsm_compat.hpp) builds against the cv::Mat shim and links
    against libgsm.so; with oracle/_ref present it also links the reference's own compareDisp."""
    exe = tmp_path / "caller_dropin"
    cmd = ["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "cpp", "caller_dropin.cpp"), "-o", str(exe),
           "-L", os.path.join(ROOT, "gpu_stereo_matching_b200"), "-lgsm",
           "-Wl,-rpath," + os.path.join(ROOT, "gpu_stereo_matching_b200")]
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref, "libref.so")):
        cmd += ["-DWITH_REF", "-L", ref, "-lref", "-Wl,-rpath," + ref]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert exe.exists()


def test_header_is_valid_c_and_runs(tmp_path):
    """gsm.h compiles as C99; the C program links libgsm.so and either runs the path (GPU box) or reports the
    missing device loudly (CPU box) -- never a silent fallback."""
    exe = tmp_path / "abi_c"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "abi_c.c"), "-o", str(exe),
           "-L", os.path.join(ROOT, "gpu_stereo_matching_b200"), "-lgsm",
           "-Wl,-rpath," + os.path.join(ROOT, "gpu_stereo_matching_b200")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "sm_100a" in run.stdout
    assert ("no device" in run.stdout and "no CPU fallback" in run.stdout) or "ok " in run.stdout


def _read_pgm(path):
    raw = open(path, "rb").read()
    assert raw[:2] == b"P5"
    parts = raw.split(b"\n", 3)
    w, h = [int(x) for x in parts[1].split()]
    return np.frombuffer(parts[3], np.uint8, w * h).reshape(h, w)


def test_caller_decodes_the_demo_pair_like_opencv(tmp_path):
    """Host side of the GUI-free singleFrame (Caller.cpp:12-16: imread + cvtColor BGR2GRAY): the built-in PNG decoder
    and the fixed-point gray conversion reproduce, byte for byte, the grays cv2 made from the same files
    (tests/golden/make_fixtures.py -> middlebury_gray.npz ArtDemo_L / ArtDemo_R).  No GPU involved."""
    exe = os.path.join(ROOT, "gpu_stereo_matching_b200", "gsm_caller")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "gpu_stereo_matching_b200", "csrc")], check=True, capture_output=True)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "middlebury_gray.npz"))
    for name, key in (("view1_.png", "ArtDemo_L"), ("view5_.png", "ArtDemo_R")):
        out = tmp_path / (name + ".pgm")
        subprocess.run([exe, "gray", os.path.join(ROOT, "tests", "golden", "art_demo", name), str(out)], check=True)
        assert np.array_equal(_read_pgm(out), fx[key]), name
    # PNG written by the front-end decodes back to the same bytes
    png = tmp_path / "g.png"
    subprocess.run([exe, "gray", os.path.join(ROOT, "tests", "golden", "art_demo", "view1_.png"), str(png)], check=True)
    pgm = tmp_path / "g.pgm"
    subprocess.run([exe, "gray", str(png), str(pgm)], check=True)
    assert np.array_equal(_read_pgm(pgm), fx["ArtDemo_L"])
