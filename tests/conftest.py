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
    """``gpu``-marked tests need a CUDA device AND the built library: skip (not fail) them elsewhere, so a plain
    ``pytest tests`` is green on a CPU box.  On a GPU box a missing library is an error, not a skip: the product has
    no fallback and a silently skipped suite would read as a pass."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def bits(a):
    """fp32 array -> int32 bit patterns (so that -0.0 != +0.0 and NaNs compare)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.int32)
