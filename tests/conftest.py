import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference compiled unmodified (oracle/_ref).  Present wherever it was prebuilt."""
    from oracle.oracle import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libfle_ref.so not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "fle_golden.npz"))


@pytest.fixture(scope="session")
def lib_path():
    import fle_b200  # noqa: F401
    from fle_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return capi.LIB_PATH


@pytest.fixture(scope="session")
def ctx(lib_path):
    """A library context on cuda:0 bound to torch's current stream.  Fails loudly (no skip, no
    fallback) when the extension cannot be loaded on a GPU box."""
    import torch
    import fle_b200
    torch.cuda.init()
    c = fle_b200.FleContext(0, torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


def tail_mask_words(n):
    """uint64 mask array selecting the first n bits (LSB-first words)."""
    nw = (n + 63) // 64
    m = np.full(nw, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    if n & 63:
        m[-1] = np.uint64((1 << (n & 63)) - 1)
    return m


def valid_row_mask_planes(n_in_group):
    """Plane-word mask of the first n_in_group rows of a group (row k = bit 63-k)."""
    if n_in_group >= 64:
        return np.uint64(0xFFFFFFFFFFFFFFFF)
    return np.uint64(((1 << n_in_group) - 1) << (64 - n_in_group))
