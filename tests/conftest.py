import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """libb200ssl.so and the C oracle must exist; build() is cheap when they are up to date."""
    import __graft_entry__ as entry
    lib = os.path.join(ROOT, "semi-supervised_semantic_segmentation_b200", "lib", "libb200ssl.so")
    if not os.path.exists(lib):
        entry.build()
    import oracle
    oracle.build()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def unpack_mask(g):
    shape = tuple(int(x) for x in g["shape"])
    n = int(np.prod(shape))
    return np.unpackbits(g["mask_bits"])[:n].reshape(shape).astype(np.float32)
