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


@pytest.mark.parametrize("P,K", [(16129, 256), (8160, 256), (65025, 64), (80, 3), (3969, 256), (127, 11), (40000, 40)])
def test_stream_k_plan_covers_every_tile_component_once(P, K):
    """Host arithmetic of the stream-K decomposition (no GPU needed): the chunks partition the linearised
    (tile pair, component) space, every tile's segments fit its partial slots, and the workspace holds them."""
    import ctypes

    from jolideco_b200 import _lib

    lib = _lib.load()
    n, chunk, smax = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.call("jd_gmm_tc_sk_plan", P, K, ctypes.addressof(n), ctypes.addressof(chunk), ctypes.addressof(smax))
    n, chunk, smax = n.value, chunk.value, smax.value
    n_pairs = ((P + 127) // 128 + 1) // 2
    w_tot = n_pairs * K
    assert (n - 1) * chunk < w_tot <= n * chunk  # every CTA pair owns a non-empty chunk, together they cover everything
    if K % 8 == 0:
        assert chunk % 8 == 0  # no segment shorter than 8 components
    covered = 0
    for tp in range(n_pairs):  # the kernel's own slot arithmetic (gmm_fwd_tc_sk_kernel)
        c_first, c_last = (tp * K) // chunk, ((tp + 1) * K - 1) // chunk
        assert c_last - c_first + 1 <= smax and c_last < n
        for c in range(c_first, c_last + 1):
            lo, hi = max(c * chunk, tp * K), min((c + 1) * chunk, (tp + 1) * K, w_tot)
            assert hi > lo
            covered += hi - lo
    assert covered == w_tot
    nbytes = lib.jd_gmm_tc_sk_workspace_bytes(P, K)
    assert nbytes >= 2 * n_pairs * 4 + 3 * 2 * n_pairs * smax * 128 * 4
