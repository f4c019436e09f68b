"""Host check of the per-pixel shift arithmetic shared with the CUDA kernels (jolideco_b200/csrc/jd_shift.cuh): built
with g++ and compared with the oracle's restatement (itself pinned to the imported reference, shift_kat.npz)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import jolideco_oracle as O


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_shift_taps_match_oracle_and_reference(tmp_path):
    so = tmp_path / "libshift_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "jolideco_b200", "csrc"),
                    os.path.join(ROOT, "tests", "shift_host.cpp"), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    lib.shift_forward_host.argtypes = [f32p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
    lib.shift_backward_host.argtypes = [f32p, f32p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, f32p, f64p]
    g = load_golden("shift_kat.npz")
    image = np.ascontiguousarray(g["image"], dtype=np.float32)
    cot = np.ascontiguousarray(g["cot"], dtype=np.float32)
    H, W = image.shape
    ptr = lambda a, t=f32p: a.ctypes.data_as(t)  # noqa: E731
    for i, (sx, sy, scale) in enumerate(list(g["cases"]) + [(2.0, -1.0, 1), (0.0, 0.7, 2)]):
        out, dimage, dshift = np.empty_like(image), np.empty_like(image), np.zeros(2)
        lib.shift_forward_host(ptr(image), sx, sy, int(scale), H, W, ptr(out))
        lib.shift_backward_host(ptr(cot), ptr(image), sx, sy, int(scale), H, W, ptr(dimage), ptr(dshift, f64p))
        ref, d_dy, d_dx = O.shift_image(image.astype(np.float64), sy, sx, int(scale), return_grads=True)
        assert np.abs(out - ref).max() <= 3e-6 * np.abs(ref).max()
        adj = O.shift_image_adjoint(cot.astype(np.float64), sy, sx, int(scale))
        assert np.abs(dimage - adj).max() <= 3e-6 * np.abs(adj).max()
        if i < len(g["cases"]):  # away from whole-pixel kinks: also against the reference's autograd values
            assert np.abs(out - g[f"c{i}_f32_out"]).max() <= 2e-5 * np.abs(ref).max()
            np.testing.assert_allclose(dshift, g[f"c{i}_f64_dshift_xy"], rtol=2e-5)
            np.testing.assert_allclose(dshift, [(cot * d_dx).sum(), (cot * d_dy).sum()], rtol=2e-5)
