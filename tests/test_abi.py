"""CPU-side checks of the C-ABI library: it builds, loads without a GPU, exports every symbol that
include/jolideco_b200.h declares, and fails loudly (no fallback) when no device is present."""
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "jolideco_b200.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jd_[a-z0-9_]+)\s*\(", src)) - {"jd_status"})


def test_library_builds_and_exports_every_declared_symbol():
    from jolideco_b200 import _lib, build

    build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/jolideco_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.jd_abi_version() == 1


def test_no_silent_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jolideco_b200 import _lib, ops

    with pytest.raises(_lib.JolidecoB200Error):
        ops.require_device()
    with pytest.raises(_lib.JolidecoB200Error):
        ops.flux_forward(torch.zeros(4, 4))  # CPU tensor: refused, never computed on the host
