import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The built libraries are git-ignored: on a fresh checkout build them once (nvcc cross-compiles sm_100a without a
    # GPU).  This is test set-up, not a fallback: the package itself never builds or substitutes anything.
    lib = os.path.join(ROOT, "gpu_stereo_matching_b200", "libgsm.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "gpu_stereo_matching_b200", "csrc")], check=False,
                       capture_output=True)


@pytest.fixture(scope="session")
def fx():
    """Gray u8 Middlebury pairs (tests/golden/make_fixtures.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "middlebury_gray.npz"))


@pytest.fixture(scope="session")
def digests():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "ref_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O
    O.lib()  # builds liboracle.so on first use
    return O


@pytest.fixture(scope="session")
def ctx():
    """One libgsm context for the whole GPU session (C ABI through ctypes)."""
    import gpu_stereo_matching_b200 as g
    c = g.StereoContext(1080, 1920, 256, 4)
    yield c
    c.close()


SETS = ["Art", "Books", "Computer", "Dolls", "Drumsticks", "Dwarves", "Laundry", "Moebius", "Reindeer"]
