"""Hand-over timeline of the two-tiles-per-CTA prior forward (backend 4), CTA 0, from the clock64 stamps of the
JD_TCM_TRACE experiment build:
    JD_NVCC_EXTRA=-DJD_TCM_TRACE JD_LIB_TAG=trace python -m jolideco_b200.build
    JD_LIB_PATH=$PWD/jolideco_b200/libjolideco_b200_trace.so [JD_TC_DEBUG=1] python tools/tcm_trace.py [size] [backend]
Events per position: 0 producer past `empty` | 1 issuer 0 past `tempty` | 2 issuer 0 past `full` | 3 issuer 0 done |
4 epilogue 0 past `tfull` | 5 epilogue 0 loads landed | 6 epilogue 0 released the slot | 7 epilogue 1 past `tfull` |
8 issuer 1 past `tempty` | 9 issuer 1 past `full` | 10 issuer 1 done | 11 epilogue 1 released the slot."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jolideco_b200 as J  # noqa: E402
from jolideco_b200 import _lib, ops, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
backend = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nslot = 3 if backend == 5 else 2
dev = "cuda"
means, cov, w = synthetic.synthetic_gmm(256, seed=7)
packed = J.GaussianMixtureModel.from_numpy(means, cov, w, meta=J.GaussianMixtureModelMeta(stride=4)).packed(dev)
flux = torch.from_numpy(np.random.default_rng(0).gamma(2.0, size=(n, n)).astype(np.float32)).to(dev)
for _ in range(3):
    ops.gmm_prior_forward(flux, (1, -2), packed, 4, False, backend=backend)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
npos = 400
buf = np.zeros((npos, 16), dtype=np.int64)
rc = lib.jd_debug_tcm2_trace(buf.ctypes.data_as(ctypes.c_void_p), npos)
assert rc == 0
t = buf.astype(np.float64)
t0 = t[0, 0]
sel = slice(40, 200)  # steady state inside the first segment
names = ["prod>empty", "iss>tempty", "iss>full", "iss done", "ep0>tfull", "ep0 ld", "ep0 rel", "ep1>tfull",
         "mma issued", "stage commit", "reconverged", "ep1 rel"]
print(f"backend {backend}; first positions (clk since the producer's first issue):")
for p in list(range(0, 6)) + list(range(100, 112)):
    print(p, " ".join(f"{names[e]}={t[p, e] - t0:7.0f}" for e in range(12)))
per = np.diff(t[sel, 3]).mean()
print(f"\nsteady state (positions {sel.start}..{sel.stop}): period {per:.0f} clk per position "
      f"({per / 2:.0f} per tile-position)")
d = lambda a, b: (t[sel, a] - t[sel, b]).mean()
print(f"  issuer:   past tempty -> past full {d(2, 1):6.0f} | MMAs issued {d(8, 2):6.0f} | stage commit {d(9, 8):6.0f} | "
      f"slot commit {d(3, 9):6.0f} | reconverged {d(10, 3):6.0f}")
nxt_own = t[sel.start + nslot:sel.stop + nslot, 1]
print(f"            reconverged (pos p) -> same warp past tempty (pos p+{nslot}) {(nxt_own - t[sel, 10]).mean():6.0f}")
print(f"  TMA:      producer issue -> issuer sees full {d(2, 0):6.0f}")
print(f"  tensor:   issuer done -> epilogue 0 sees tfull {d(4, 3):6.0f} | epilogue 1 {d(7, 3):6.0f}")
print(f"  epilogue: tfull -> loads landed {d(5, 4):6.0f} | loads -> slot released {d(6, 5):6.0f} | epilogue 1 total {d(11, 7):6.0f}")
rel = np.maximum(t[sel.start:sel.stop, 6], t[sel.start:sel.stop, 11])
print(f"  slot:     both epilogues released (pos p) -> issuer past tempty (pos p+{nslot}) {(nxt_own - rel).mean():6.0f}")
nst = 10 if backend == 5 else 6
prv = t[sel.start + nst:sel.stop + nst, 0]
print(f"  stage:    issuer done (pos p) -> producer past empty (pos p+{nst}) {(prv - t[sel, 3]).mean():6.0f}")
