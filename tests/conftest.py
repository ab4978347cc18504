"""Test configuration.  `-m "not gpu"` covers the oracle (vs cv2, golden vectors, KATs), host logic and the ABI
surface; `-m gpu` are the parity tests proper: CUDA path (through the C ABI) vs the CPU oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "orb-slam3-rust_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import orb_oracle

    orb_oracle.build()
    return orb_oracle


@pytest.fixture(scope="session")
def cv2mod():
    return pytest.importorskip("cv2")


@pytest.fixture(scope="session")
def camera_dict():
    from vslam_b200 import synth

    return dict(synth.EUROC_CAMERA)


@pytest.fixture(scope="session")
def pairs():
    """Cache of synthetic stereo pairs keyed by (seed, W, H) -> (left, right) numpy uint8."""
    from vslam_b200 import synth

    cache = {}

    def get(seed, w=752, h=480):
        key = (seed, w, h)
        if key not in cache:
            l, r = synth.make_stereo_pair(seed, w, h)
            cache[key] = (l.numpy(), r.numpy())
        return cache[key]

    return get


def assert_struct_equal(a, b, what):
    """Bit-exact comparison of structured arrays with a readable first difference."""
    assert len(a) == len(b), f"{what}: length {len(a)} != {len(b)}"
    if a.tobytes() == b.tobytes():
        return
    for f in a.dtype.names:
        ne = np.flatnonzero(a[f].view(np.uint32) != b[f].view(np.uint32))
        assert len(ne) == 0, f"{what}.{f}: {len(ne)} differ, first at {ne[0]}: {a[f][ne[0]]!r} != {b[f][ne[0]]!r}"
