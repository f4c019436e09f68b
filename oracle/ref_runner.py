"""TEST / BASELINE INFRASTRUCTURE — steps the UNMODIFIED reference package (vendored by
`pip install --no-deps --target baseline/_ref /root/reference`, or `/root/reference` itself in the build
container) through its own classes, on the host cores or on a CUDA device.

Used only by `bench.py --impl reference`, by bench.py's `cpu_baseline` / `gpu_baseline` legs and by
`oracle/make_golden.py`.  Nothing under `jolideco_b200/` imports it.

Two step semantics, both spelled with the reference's own objects (no arithmetic of ours on the path):
  * `step(i)`       the reference loop body, `jolideco/core.py:214-229`;
  * `joint_step()`  one Adam step on sum_d L_d - beta * prior, the objective of `TotalLoss.__call__`
                    (`loss.py:257-261`), with the dataset terms kept in the autograd graph (the reference's own
                    `PoissonLoss.evaluate` re-wraps them in a fresh tensor, `loss.py:71`, which has no gradient).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def locate_reference():
    """baseline/_ref (travels to the GPU box) first, then the read-only source tree of the build container."""
    for cand in (os.environ.get("JOLIDECO_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "jolideco")):
            return cand
    return None


def install():
    root = locate_reference()
    if root is None:
        raise ImportError("reference package not found (baseline/_ref or /root/reference)")
    from . import ref_shim

    ref_shim.REFERENCE_ROOT = root
    return ref_shim.install(), root


class ReferenceRun:
    """The reference's components / losses / optimizer for one workload (`jolideco_b200.synthetic.make_workload`)."""

    def __init__(self, workload, device="cpu", marginalize=False, seed=0, lr=0.1, beta=1.0):
        import torch

        _, self.root = install()
        from jolideco.loss import TotalLoss
        from jolideco.models import FluxComponents, SpatialFluxComponent
        from jolideco.priors import GMMPatchPrior, UniformPrior
        from jolideco.priors.patches.gmm import GaussianMixtureModel, GaussianMixtureModelMeta

        self.torch = torch
        self.device = torch.device(device)
        prior = UniformPrior()
        if workload["gmm_arrays"] is not None:
            gmm = GaussianMixtureModel.from_numpy(*workload["gmm_arrays"], meta=GaussianMixtureModelMeta(stride=4))
            gen = torch.Generator(device=self.device).manual_seed(seed)
            prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=marginalize, device=self.device)
        comps = FluxComponents()
        comps["flux"] = SpatialFluxComponent.from_numpy(flux=workload["flux_init"], upsampling_factor=workload["f"],
                                                        prior=prior)
        self.components = comps.to(self.device)
        self.beta = beta
        self.total_loss = TotalLoss.from_datasets_and_components(datasets=workload["datasets"], components=self.components,
                                                                 beta=beta, device=self.device)
        self.optimizer = torch.optim.Adam(params=list(self.components.parameters()), lr=lr)
        self.pairs = list(self.total_loss.poisson_loss.iter_by_dataset)
        self.D = len(self.pairs)

    def step(self, i):
        """core.py:214-229 for dataset i."""
        tl = self.total_loss
        counts, npred_model = self.pairs[i % self.D]
        self.optimizer.zero_grad()
        fluxes = self.components.to_flux_tuple()
        npred = npred_model.evaluate(fluxes=fluxes)
        loss = tl.poisson_loss.loss_function(npred, counts)
        loss_prior = tl.prior_loss(fluxes=fluxes)
        loss_total = loss - self.beta * loss_prior / tl.prior_weight
        loss_total.backward()
        self.optimizer.step()
        return loss_total

    def joint_loss(self):
        """sum_d L_d - beta * prior with the dataset terms in the graph; consumes one cycle-spin draw."""
        tl = self.total_loss
        self.optimizer.zero_grad()
        fluxes = self.components.to_flux_tuple()
        total = 0.0
        for counts, npred_model in self.pairs:
            total = total + tl.poisson_loss.loss_function(npred_model.evaluate(fluxes=fluxes), counts)
        return total - self.beta * tl.prior_loss(fluxes=fluxes)

    def joint_step(self):
        total = self.joint_loss()
        total.backward()
        self.optimizer.step()
        return total

    def peek_shift(self):
        """(row, col) cycle-spin shift the prior will draw next (utils/torch.py:108-116), without consuming it."""
        torch = self.torch
        gen = getattr(self.components["flux"].prior, "generator", None)
        if gen is None:
            return (0, 0)
        g = torch.Generator(device=gen.device)
        g.set_state(gen.get_state())
        sy = int(torch.randint(-2, 3, (1,), generator=g, device=gen.device))
        sx = int(torch.randint(-2, 3, (1,), generator=g, device=gen.device))
        return sy, sx

    def theta_grad(self):
        return self.components["flux"]._flux_upsampled.grad.detach().cpu().numpy()[0, 0]

    def sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def flux_numpy(self):
        return self.components["flux"].flux_upsampled.detach().cpu().numpy()[0, 0]


def time_steps(run, joint, steps, warmup, budget_s):
    """Wall-clock seconds per step over up to `steps` steps (fewer when `budget_s` would be exceeded)."""
    import time

    def one(i):
        run.joint_step() if joint else run.step(i)

    t0 = time.perf_counter()
    w = max(1, warmup)
    for i in range(w):
        one(i)
    run.sync()
    t_est = (time.perf_counter() - t0) / w
    n = int(max(2, min(steps, budget_s / max(t_est, 1e-9))))
    t0 = time.perf_counter()
    for i in range(n):
        one(i)
    run.sync()
    dt = time.perf_counter() - t0
    return n, dt, w


if __name__ == "__main__":  # quick self-test on a tiny workload
    sys.path.insert(0, ROOT)
    from jolideco_b200 import synthetic

    wl = synthetic.make_workload("tiny", seed=3)
    r = ReferenceRun(wl)
    print("reference from", r.root, "joint loss", float(r.joint_step()), "step loss", float(r.step(0)))
    print(np.abs(r.flux_numpy()).mean())
