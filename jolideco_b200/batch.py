"""Batched independent MAP runs (bootstrap / restarts): BASELINE.json configs[4].

The reference has no batch API (a 4-D N=1 flux is enforced, `models/core.py:394-397`, and the prior
normalises by `flux.numel()`, `patches/core.py:246`, which would mix runs).  Independent runs share
nothing, so instead of a leading batch axis every run keeps its own `MapEngine` (normalised per run
by construction) and the runs are interleaved on a pool of CUDA streams: each run's step is one CUDA
graph, graphs of different runs execute concurrently, which is what fills the 148 SMs when a single
256x256 run only has ~32 CTAs of prior work.  Across GPUs the runs are dealt round-robin
(`dist.shard_indices`), no collective.
"""
import copy

import numpy as np
import torch

from . import dist, ops
from .core import MAPDeconvolver, MAPDeconvolverResult
from .loss import TotalLoss
from .models import FluxComponents, SpatialFluxComponent

__all__ = ["run_many", "BatchedRuns"]


class BatchedRuns:
    """Independent runs dealt to this rank, interleaved on a pool of CUDA streams."""

    def __init__(self, jobs, n_epochs, n_streams=8, rank=0, world=1, **deconvolver_kwargs):
        self.mine = dist.shard_indices(len(jobs), rank, world)
        self.n_epochs = n_epochs
        self.deco = deco = MAPDeconvolver(n_epochs=n_epochs, display_progress=False, **deconvolver_kwargs)
        ops.require_device(deco.device)
        self.streams = [torch.cuda.Stream(device=deco.device) for _ in range(max(1, min(n_streams, len(self.mine))))]
        self.runs = []
        self.epoch = 0
        with torch.cuda.device(deco.device):
            for slot, j in enumerate(self.mine):
                job = jobs[j]
                components = job["components"]
                if isinstance(components, SpatialFluxComponent):
                    components = {deco._default_flux_component: components}
                components = FluxComponents(components)
                components_init = copy.deepcopy(components)
                components = components.to(deco.device)
                if not deco._engine_supported(components, None):
                    raise NotImplementedError("run_many: job %d is not a fused-engine configuration" % j)
                total_loss = TotalLoss.from_datasets_and_components(
                    datasets=job["datasets"], datasets_validation=job.get("datasets_validation"),
                    components=components, beta=deco.beta, device=deco.device)
                D = len(job["datasets"])
                # concurrent runs: throughput, not latency - no second stream per run, and the one-tile-per-CTA prior
                # kernel (backend 1, no stream-K): the persistent backend-3 kernel claims every SM for each launch, which
                # serialises the runs (cfg5: 27.6 k instead of 35.7 k iterations/s)
                engine = deco._build_engine(total_loss, components, n_epochs * (D + 1), stream_k=False, overlap=False,
                                            backend=1)
                stream = self.streams[slot % len(self.streams)]
                with torch.cuda.stream(stream):
                    engine.warmup()
                rows = torch.zeros((n_epochs, engine.n_trace), dtype=torch.float64, device=deco.device)
                self.runs.append(dict(index=j, engine=engine, total_loss=total_loss, components=components,
                                      components_init=components_init, D=D, stream=stream, rows=rows))
            torch.cuda.synchronize(deco.device)

    @property
    def steps_per_epoch(self):
        return sum(r["D"] for r in self.runs)

    def run_epochs(self, n, trace=True):
        """Enqueue n epochs of every run (epoch-major: consecutive launches go to different streams)."""
        with torch.cuda.device(self.deco.device):
            for _ in range(n):
                for i in range(max(r["D"] for r in self.runs) if self.runs else 0):
                    for r in self.runs:
                        if i < r["D"]:
                            with torch.cuda.stream(r["stream"]):
                                r["engine"].step(i)
                if trace and self.epoch < self.n_epochs:
                    for r in self.runs:
                        with torch.cuda.stream(r["stream"]):
                            r["engine"].trace_enqueue(r["rows"][self.epoch])
                self.epoch += 1

    def join(self, stream=None):
        """Make `stream` (default: current) wait for every run stream."""
        stream = torch.cuda.current_stream(self.deco.device) if stream is None else stream
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            stream.wait_event(ev)

    def results(self):
        torch.cuda.synchronize(self.deco.device)
        out = {}
        for r in self.runs:
            host = r["rows"][: min(self.epoch, self.n_epochs)].cpu().numpy()
            names = list(r["total_loss"].prior_loss.priors)
            for vals in host:
                ld, lp, lv = r["engine"].trace_decode(vals)
                r["total_loss"].append_trace_values(ld, [lp] * len(names), "", lv if lv else None)
            out[r["index"]] = MAPDeconvolverResult(config=self.deco.to_dict(), components=r["components"],
                                                   components_init=r["components_init"],
                                                   trace_loss=r["total_loss"].trace)
        return out


def run_many(jobs, n_epochs=100, n_streams=8, rank=0, world=1, **deconvolver_kwargs):
    """Run independent deconvolutions concurrently on one GPU.

    jobs : list of dict(datasets=..., components=..., [datasets_validation=...]) as for `MAPDeconvolver.run`.
    Only the jobs dealt to `rank` (of `world`) are run; returns {job index: MAPDeconvolverResult}.
    Every job must be a configuration the fused engine supports (one spatial component, uniform or GMM
    patch prior, Adam) and uses the reference's sequential step semantics.
    """
    batch = BatchedRuns(jobs, n_epochs, n_streams, rank, world, **deconvolver_kwargs)
    batch.run_epochs(n_epochs)
    return batch.results()
