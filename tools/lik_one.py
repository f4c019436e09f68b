"""A few launches of the batched likelihood kernels at the north-star shape (8 datasets, 1024^2, 17x17 PSFs) for ncu:
    ncu --set full --import-source on --clock-control none -k regex:lik_kernel -s 2 -c 2 -o gpurun_out/prof_lik python tools/lik_one.py
Also prints CUDA-event timings of back-to-back launches (warm L2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolideco_b200 import _lib, ops, synthetic  # noqa: E402

n, D, k = int(os.environ.get("N", 1024)), int(os.environ.get("D", 8)), int(os.environ.get("K", 17))
dev = "cuda"
rng = np.random.default_rng(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)  # noqa: E731
flux = t(rng.gamma(2.0, size=(n, n)))
ds = [dict(exposure=t(rng.uniform(0.5, 1.5, size=(n, n))), psf=t(synthetic.gaussian_psf(k)),
           background=t(np.full((n, n), 0.5)), counts=t(rng.poisson(3.0, size=(n, n)))) for _ in range(D)]
for _ in range(3):
    ops.likelihood_batched(flux, ds, 1)
torch.cuda.synchronize()
# timing: rebuild nothing, just relaunch through the ops wrapper's table (python overhead excluded by events per call)
res = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.likelihood_batched(flux, ds, 1, want_grad=False)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) * 1e3)
print(f"forward only ({D} x {n}^2, {k}x{k}): {min(res):.1f} us per launch (incl. table upload)")
peak = None
out = torch.zeros(1, device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fl = _lib.load().jd_probe_fp32_fma(4096, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    peak = max(peak or 0, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
print(f"FP32 FMA probe: {peak:.1f} TFLOP/s")
