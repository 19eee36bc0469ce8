"""Shared fixtures. GPU tests are marked `gpu`; everything else runs on a CPU-only box.

The oracle (oracle/) is the checker here and only here; the product under test is
sparse-voxel-octrees_b200/libsvo_b200.so called through its C ABI (pysvo is a ctypes shim).
/root/reference is never read by `-m gpu` tests: the reference's own object code travels as
oracle/_ref/libsvo_ref.so, and tests that need it skip when it is absent.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))

GOLDEN = ROOT / "tests" / "golden"
DRAGON = GOLDEN / "XYZRGB-Dragon.oct"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pysvo():
    import pysvo as mod
    if not mod.LIB_PATH.exists():
        mod.build_library()
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def port():
    from oracle.pyoracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own object code; skipped where neither the .so nor /root/reference exists."""
    from oracle import pyoracle
    try:
        return pyoracle.Ref()
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"oracle/_ref unavailable: {e}")


@pytest.fixture(scope="session")
def dragon_words(pysvo):
    words, center = pysvo.oct_read(DRAGON)
    return words, center


@pytest.fixture(scope="session")
def gpu_dragon(pysvo):
    if pysvo.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible (no CPU fallback exists)")
    tree = pysvo.VoxelOctree(DRAGON)
    yield tree
    tree.close()


CAMERAS = [(0.0, 0.0, 1.0), (20.0, 135.0, 0.5), (0.0, 0.0, 2.5), (-35.0, 250.0, 0.8), (89.0, 10.0, 0.3)]
