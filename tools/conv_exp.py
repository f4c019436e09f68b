"""Timing of the direct convolution kernel (tuning aid): JD_CONV_TILE=0|1|2 python tools/conv_exp.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolideco_b200 import ops
for n, k in [(1024, 17), (512, 17), (256, 17), (512, 9), (1024, 23)]:
    flux = torch.rand(n, n, device="cuda"); E = torch.rand(n, n, device="cuda") + 0.5
    psf = torch.rand(k, k, device="cuda"); out = torch.empty_like(flux)
    for _ in range(3):
        ops.conv_forward(flux, E, psf, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.conv_forward(flux, E, psf, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"tile {os.environ.get('JD_CONV_TILE','0')} n={n} k={k}: {us:7.1f} us  {2*n*n*k*k/us/1e6:6.1f} TFLOP/s")
