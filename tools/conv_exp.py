"""Timing sweep of the direct convolution kernels against the FFT path (tuning aid):
    python tools/conv_exp.py > gpurun_out/conv_exp.txt"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolideco_b200 import _lib, ops


def timeit(fn, reps=20):
    """GPU time per launch: `reps` launches captured in one CUDA graph (no CPU launch overhead in the timing)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * reps) * 1e3


cases = [(1024, 17), (512, 34), (512, 23), (256, 17), (1024, 23), (2048, 17), (1024, 9)]
variants = [("auto", 1, 0, 0), ("old", 0, 0, 0)] + [(f"tx{tx}s{s}", 1, tx, s) for tx in (8, 16, 28) for s in (1, 2, 4)]
for n, k in cases:
    flux = torch.rand(n, n, device="cuda"); E = torch.rand(n, n, device="cuda") + 0.5
    psf = torch.rand(k, k, device="cuda"); out = torch.empty_like(flux); d = torch.randn(n, n, device="cuda")
    flop = 2.0 * n * n * k * k
    row = []
    for name, v3, tx, s in variants:
        _lib.call("jd_conv_tuning", v3, tx, s)
        try:
            uf = timeit(lambda: ops.conv_forward(flux, E, psf, out=out))
            ub = timeit(lambda: ops.conv_backward(d, E, psf, 1, out=out))
            row.append(f"{name}: {uf:6.1f}/{ub:6.1f} us ({flop / uf / 1e6:5.1f} TF/s)")
        except Exception as exc:
            row.append(f"{name}: failed ({str(exc)[:40]})")
    _lib.call("jd_conv_tuning", 1, 0, 0)
    plan = ops.FFTConvPlan(psf, n, n)
    uf = timeit(lambda: ops.conv_forward_fft(flux, E, plan, out=out))
    ub = timeit(lambda: ops.conv_backward_fft(d, E, plan, 1, out=out))
    row.append(f"fft: {uf:6.1f}/{ub:6.1f} us")
    print(f"n={n} k={k}: " + " | ".join(row), flush=True)
