"""A few launches of the direct convolution (profiling target): python tools/conv_one.py [n] [k]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolideco_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 17
flux = torch.rand(n, n, device="cuda"); E = torch.rand(n, n, device="cuda") + 0.5
psf = torch.rand(k, k, device="cuda"); out = torch.empty_like(flux); d = torch.randn(n, n, device="cuda")
for _ in range(3):
    ops.conv_forward(flux, E, psf, out=out)
    ops.conv_backward(d, E, psf, 1, out=out)
torch.cuda.synchronize()
