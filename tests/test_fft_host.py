"""Host check of the FFT stage code the CUDA kernels run (jolideco_b200/csrc/jd_fft_stages.cuh): the mixed-radix
Stockham stages, compiled with g++ and driven sequentially, against a naive double-precision DFT for every size class
(2^L, 3 * 2^L, 5 * 2^L), forward and inverse; plus the padded-size selection."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fft_stages_against_naive_dft(tmp_path):
    exe = tmp_path / "fft_stages_host"
    src = os.path.join(ROOT, "tests", "fft_stages_host.cpp")
    inc = os.path.join(ROOT, "jolideco_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, src, "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == "OK", res.stdout + res.stderr
