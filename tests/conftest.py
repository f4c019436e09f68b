import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the CUDA library is built in-tree (git-ignored): make sure it exists / is current before any test imports it
    try:
        from jolideco_b200 import build

        build.build()
    except Exception as exc:  # no nvcc on this machine: tests that need the library will fail loudly themselves
        print(f"[conftest] could not (re)build libjolideco_b200.so: {exc}")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device: skip them (instead of failing at import) on a CPU-only host."""
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); jolideco_b200 has no CPU path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def unpack_datasets(g, prefix="ds"):
    n = int(g[f"{prefix}n"])
    return [
        {k: g[f"{prefix}{i}_{k}"] for k in ["counts", "psf", "exposure", "background"]}
        for i in range(n)
    ]


@pytest.fixture(scope="session")
def golden():
    return load_golden
