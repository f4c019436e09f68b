import os, sys, torch, numpy as np
B = int(os.environ.get("JD_BACKEND", "1"))
sys.path.insert(0, '/root/repo')
from jolideco_b200 import ops, synthetic
from oracle import jolideco_oracle as O
means, cov, w = synthetic.synthetic_gmm(256, seed=7)
g = O.GMM(means, cov, w)
packed = ops.GMMPacked(g.means, g.precisions_cholesky, g.weights, g.pixel_weights, 'cuda')
flux = torch.rand(512, 512, device='cuda') + 0.5
for _ in range(3):
    ops.gmm_prior_forward(flux, (0, 0), packed, backend=B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gmm_prior_forward(flux, (0, 0), packed, backend=B)
e1.record(); torch.cuda.synchronize()
print("backend", B, "JD_TC_DEBUG", os.environ.get('JD_TC_DEBUG', '0'), 'us per launch', e0.elapsed_time(e1) / 20 * 1e3)
