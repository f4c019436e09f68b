"""Time the prior forward kernels alone (CUDA events, L2 flushed between launches):
    [JD_TC_DEBUG=n] [JD_TCM_CLUSTERS=c] python tools/tcm_exp.py [size ...]
JD_TC_DEBUG knobs of backend 3 (results wrong, timing only): 1 no epilogue TMEM loads, 2 no FP16 products,
4 no TF32 product, 8 one MMA per component, 16 dense (untrimmed) schedule."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jolideco_b200 as J  # noqa: E402
from jolideco_b200 import ops, synthetic  # noqa: E402

dev = "cuda"
sizes = [int(a) for a in sys.argv[1:]] or [512, 1024]
means, cov, w = synthetic.synthetic_gmm(256, seed=7)
packed = J.GaussianMixtureModel.from_numpy(means, cov, w, meta=J.GaussianMixtureModelMeta(stride=4)).packed(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in sizes:
    flux = torch.from_numpy(np.random.default_rng(0).gamma(2.0, size=(n, n)).astype(np.float32)).to(dev)
    P = ((n - 8) // 4 + 1) ** 2
    for backend in [int(b) for b in os.environ.get("JD_EXP_BACKENDS", "3,4,5").split(",")]:
        for _ in range(3):
            ops.gmm_prior_forward(flux, (1, -2), packed, 4, False, backend=backend)
        ts = []
        for i in range(10):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gmm_prior_forward(flux, (1, -2), packed, 4, False, backend=backend)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        print(f"size {n} P {P} backend {backend} dbg {os.environ.get('JD_TC_DEBUG', '0')}: {us:.1f} us "
              f"(incl. ~10 us of host-side wrapper launches), {2.0 * P * 4096 * 256 / us / 1e6:.0f} TFLOP/s useful")
