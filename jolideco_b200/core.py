"""MAP deconvolver behind the reference's public API (jolideco/core.py).

`MAPDeconvolver(...).run(datasets, datasets_validation, components, calibrations)` keeps the
reference's signature, error behaviour and result object.  The inner loop (core.py:209-230) runs on
the fused CUDA engine (`engine.MapEngine`) whenever the configuration is the hot path the engine
covers (one spatial flux component, uniform or GMM patch prior, Adam); anything else still runs on
the same kernels through the autograd bindings with `torch.optim`, mirroring the reference loop.
There is no CPU fallback: a non-CUDA device raises.
"""
import copy
import logging
import os
from pathlib import Path

import numpy as np
import torch

from . import ops
from ._lib import JolidecoB200Error
from .engine import DatasetBuffers, MapEngine
from .loss import TotalLoss
from .models import FluxComponents, SpatialFluxComponent
from .norms import IdentityImageNorm
from .priors import GMMPatchPrior, UniformPrior, default_backend

log = logging.getLogger(__name__)

__all__ = ["MAPDeconvolver", "MAPDeconvolverResult"]

OPTIMIZER = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD}


class MAPDeconvolver:
    """Maximum A-Posteriori deconvolver (core.py:46-282); same constructor arguments."""

    _default_flux_component = "flux"
    _default_checkpoint_filename = "checkpoint-epoch-{epoch}.npz"

    def __init__(self, n_epochs=1_000, beta=1, learning_rate=0.1, compute_error=False, stop_early=False,
                 stop_early_n_average=10, device="cuda", display_progress=True, optimizer_type="adam",
                 optimizer_kwargs=None, checkpoint_path=None, use_cuda_graph=True, fused=True, mode="sequential",
                 process_group=None, collective="nccl"):
        self.n_epochs = n_epochs
        self.beta = beta
        self.learning_rate = learning_rate
        self.compute_error = compute_error
        self.stop_early = stop_early
        self.stop_early_n_average = stop_early_n_average
        self.display_progress = display_progress
        if "cuda" not in str(device):
            raise JolidecoB200Error(f"device {device!r}: jolideco_b200 runs on CUDA (sm_100a) only; no CPU fallback")
        self.device = torch.device(device)
        if optimizer_type not in OPTIMIZER:
            raise ValueError(f"Unknown optimizer: {optimizer_type}, must be one of {OPTIMIZER}")
        self.optimizer_type = optimizer_type
        self.optimizer_kwargs = {} if optimizer_kwargs is None else optimizer_kwargs
        self.optimizer_kwargs.setdefault("lr", self.learning_rate)
        if checkpoint_path is not None:
            checkpoint_path = Path(checkpoint_path)
            checkpoint_path.mkdir(exist_ok=True, parents=True)
        self.checkpoint_path = checkpoint_path
        self.use_cuda_graph = use_cuda_graph
        self.fused = fused
        # "sequential": the reference's loop, one Adam step per dataset with the full prior (core.py:214-229).
        # "joint": one Adam step per epoch on sum_d L_d - beta * prior (TotalLoss.__call__, loss.py:257-261);
        #          with torch.distributed initialised the datasets are sharded over the ranks of `process_group`,
        #          the prior over patch-row blocks, and the flux gradient is all-reduced (NCCL over NVLink).
        if mode not in ("sequential", "joint"):
            raise ValueError(f"Unknown mode: {mode}, must be 'sequential' or 'joint'")
        self.mode = mode
        self.process_group = process_group
        self.collective = collective  # "nccl" or "peer" (fused all-reduce + Adam over NVLink peer memory)

    def to_dict(self):
        data = {}
        data.update(self.__dict__)
        data["device"] = str(self.device)
        data["checkpoint_path"] = str(self.checkpoint_path)
        data.pop("optimizer", None)
        data.pop("optimizer_kwargs", None)
        data.pop("process_group", None)
        data.pop("engine", None)
        return data

    def __str__(self):
        return f"{self.__class__.__name__}\n" + "\n".join(f"  {k:24s}: {v}" for k, v in self.to_dict().items())

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _shift_is_zero(cal):
        shift = cal.shift_xy.detach().cpu()
        return bool(torch.all(torch.isclose(shift, torch.zeros_like(shift))))

    @classmethod
    def _calibrations_fusable(cls, calibrations):
        """Calibrations the engine covers: background norm (trained or frozen); shifts at 0 never train in the
        reference (`shift_image_torch` returns early, utils/torch.py:211); psf_scale ~ 1 is the identity.  Non-zero
        (trainable) shifts run in the engine through jd_shift_forward/backward (JD_FUSED_SHIFT=0 sends them to the
        autograd path with grid_sample instead)."""
        fused_shift = os.environ.get("JD_FUSED_SHIFT", "1") == "1"
        for cal in calibrations.values():
            if not cls._shift_is_zero(cal) and not fused_shift:
                return False
            if not bool(torch.isclose(cal.psf_scale.detach().cpu(), torch.tensor(1.0)).all()):
                return False
        return True

    def _engine_supported(self, components, calibrations):
        if not self.fused or self.optimizer_type != "adam":
            return False
        if calibrations and not self._calibrations_fusable(calibrations):
            return False
        if set(self.optimizer_kwargs) - {"lr", "betas", "eps"}:
            return False
        comps = list(components.values())
        if len(comps) != 1 or not isinstance(comps[0], SpatialFluxComponent) or comps[0].frozen:
            return False
        prior = comps[0].prior
        if isinstance(prior, UniformPrior):
            return True
        identity = prior.norm is None or isinstance(prior.norm, IdentityImageNorm)  # other norms: autograd path
        return isinstance(prior, GMMPatchPrior) and identity and prior.gmm.n_features == ops.PD

    def _group(self):
        if self.mode != "joint" or not torch.distributed.is_available() or not torch.distributed.is_initialized():
            return None, 0, 1
        pg = self.process_group or torch.distributed.group.WORLD
        world = torch.distributed.get_world_size(pg)
        if world == 1:
            return None, 0, 1
        return pg, torch.distributed.get_rank(pg), world

    def _build_engine(self, total_loss, components, n_draws, shard=None, stream_k=None, overlap=None, backend=None):
        (name, comp), = components.items()
        theta = comp._flux_upsampled.data[0, 0]
        mask = comp.mask[0, 0].contiguous() if comp.mask is not None else None

        def buffers(poisson_loss):
            out = []
            for ds_name, counts, models in zip(poisson_loss.names_all, poisson_loss.counts_all,
                                               poisson_loss.npred_models_all):
                model = models[name]
                cal = models.calibration
                shifted = cal is not None and not self._shift_is_zero(cal)
                out.append(DatasetBuffers(counts[0, 0].contiguous(), model.exposure[0, 0], model.psf[0, 0],
                                          models.background[0, 0].contiguous(), model.upsampling_factor, name=ds_name,
                                          bkg_log_norm=None if cal is None else cal._background_norm.data,
                                          train_bkg_norm=cal is not None and not cal.frozen,
                                          shift_xy=cal.shift_xy.data.view(-1) if shifted else None,
                                          train_shift=shifted and not cal.frozen))
            return out

        prior_cfg, table = None, None
        prior = comp.prior
        if isinstance(prior, GMMPatchPrior):
            if prior.backend is not None:
                backend = prior.backend
            elif backend is None:
                backend = default_backend()
            prior_cfg = dict(packed=prior.gmm.packed(self.device), stride=prior.stride, marginalize=prior.marginalize,
                             backend=backend)
            table = np.array([prior.draw_shifts() for _ in range(n_draws)], dtype=np.int32).reshape(-1, 2)
        validation = buffers(total_loss.poisson_loss_validation) if total_loss.poisson_loss_validation else []
        kwargs = {}
        if shard is not None:  # datasets sharded over ranks: global trace layout, identical shifts on every rank
            if table is not None:
                tab = torch.from_numpy(table).to(self.device)
                torch.distributed.broadcast(tab, src=torch.distributed.get_global_rank(shard["pg"], 0), group=shard["pg"])
                table = tab.cpu().numpy()
            f = comp.upsampling_factor or 1
            kwargs = dict(process_group=shard["pg"], collective=self.collective, dataset_index=shard["index"],
                          n_datasets_global=shard["n"],
                          validation_index=shard["vindex"], n_validation_global=shard["nv"],
                          counts_shape=(theta.shape[0] // f, theta.shape[1] // f))
        return MapEngine(theta, buffers(total_loss.poisson_loss), prior=prior_cfg, mask=mask,
                         use_log_flux=comp.use_log_flux, beta=self.beta, lr=self.optimizer_kwargs["lr"],
                         betas=self.optimizer_kwargs.get("betas", (0.9, 0.999)),
                         eps=self.optimizer_kwargs.get("eps", 1e-8), shift_table=table,
                         datasets_validation=validation, use_graph=self.use_cuda_graph, stream_k=stream_k, overlap=overlap,
                         **kwargs)

    def _early_stop(self, trace):
        if self.stop_early and len(trace) > self.stop_early_n_average:
            values = trace["datasets-validation-total"]
            return trace[-1]["datasets-validation-total"] > np.mean(values[-self.stop_early_n_average:])
        return False

    def _checkpoint(self, epoch, total_loss, components, calibrations=None, engine=None):
        """Per-epoch checkpoint (core.py:234-243): a `MAPDeconvolverResult` with config, trace so far, components and
        calibrations, written by rank 0 only (every rank holds the same replica).  The engine may train a working
        copy of theta (symmetric memory of the peer collective): it is copied back into the component first."""
        if not self.checkpoint_path:
            return ""
        filename = self._default_checkpoint_filename.format(epoch=epoch)
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()
        if rank == 0:
            if engine is not None:
                engine.sync_theta()
            checkpoint = MAPDeconvolverResult(config=self.to_dict(), trace_loss=total_loss.trace, components=components,
                                              calibrations=calibrations)
            checkpoint.write(self.checkpoint_path / filename, overwrite=True)
        return filename

    # ------------------------------------------------------------------------------------------
    def run(self, datasets, datasets_validation=None, components=None, calibrations=None):
        """Run the MAP deconvolver (core.py:149-282); returns a `MAPDeconvolverResult`."""
        if self.stop_early and datasets_validation is None:
            raise ValueError("Early stopping requires providing test datasets")
        ops.require_device(self.device)
        if isinstance(components, SpatialFluxComponent):
            components = {self._default_flux_component: components}
        components = FluxComponents(components)
        components_init = copy.deepcopy(components)
        calibrations_init = copy.deepcopy(calibrations)
        components = components.to(self.device)
        if calibrations:
            calibrations = calibrations.to(self.device)

        pg, rank, world = self._group()
        shard = None
        names, vnames = list(datasets), list(datasets_validation or {})
        if pg is not None:  # this rank only uploads and evaluates its own datasets
            from . import dist

            shard = dict(pg=pg, index=dist.shard_indices(len(names), rank, world), n=len(names),
                         vindex=dist.shard_indices(len(vnames), rank, world), nv=len(vnames))
            datasets = {names[i]: datasets[names[i]] for i in shard["index"]}
            if datasets_validation:
                datasets_validation = {vnames[i]: datasets_validation[vnames[i]] for i in shard["vindex"]}

        with torch.cuda.device(self.device):
            total_loss = TotalLoss.from_datasets_and_components(
                datasets=datasets, datasets_validation=datasets_validation, components=components,
                calibrations=calibrations, beta=self.beta, device=self.device)
            total_loss.dataset_names = names
            if vnames and total_loss.poisson_loss_validation is None:
                total_loss.poisson_loss_validation = False  # no local validation shard; the trace column still exists
            if self.mode == "joint":
                if not self._engine_supported(components, calibrations):
                    raise NotImplementedError("mode='joint' needs a configuration the fused engine supports")
                self._run_joint(total_loss, components, shard, calibrations)
            elif self._engine_supported(components, calibrations):
                self._run_fused(total_loss, components, len(datasets), calibrations)
            else:
                self._run_autograd(total_loss, components, calibrations)

        if self.compute_error:
            # The reference's estimate is sqrt(1 / (H 1)) with H 1 a vector-Hessian product of TotalLoss.__call__
            # (loss.py:263-300).  PoissonLoss.evaluate re-wraps the dataset losses in a fresh torch.tensor (loss.py:71),
            # which cuts the graph: H 1 is identically 0 and the reference returns inf everywhere (checked with the
            # imported reference, uniform and GMM priors, f = 1 and 2).  Reproduced as is rather than "fixed".
            log.warning("compute_error=True: the reference's Hessian estimate is identically zero (detached dataset "
                        "losses, loss.py:71), flux errors are inf as in the reference")
            components.set_flux_errors({name: torch.full_like(c.flux_upsampled.detach(), float("inf"))
                                        for name, c in components.items()})

        return MAPDeconvolverResult(config=self.to_dict(), components=components, components_init=components_init,
                                    trace_loss=total_loss.trace, calibrations=calibrations,
                                    calibrations_init=calibrations_init, wcs=None)

    def _draw_state(self, components):
        """Generator states of the priors' cycle-spin RNGs before the shift table is pre-drawn."""
        return {name: prior.generator.get_state() for name, prior in components.priors.items()
                if getattr(prior, "generator", None) is not None}

    def _rewind_draws(self, components, states, n_consumed):
        """The engine pre-draws n_epochs x (D + 1) shifts; an early stop consumes fewer.  Put the generators where the
        reference's would be: initial state advanced by exactly the draws that were used (2 randint calls per draw,
        utils/torch.py:108-116), so that a later run seeded from the same generator reproduces the reference."""
        for name, prior in components.priors.items():
            if name in states:
                prior.generator.set_state(states[name])
                for _ in range(n_consumed):
                    prior.draw_shifts()

    def _run_fused(self, total_loss, components, n_datasets, calibrations=None):
        states = self._draw_state(components)
        engine = self._build_engine(total_loss, components, self.n_epochs * (n_datasets + 1))
        self.engine = engine
        engine.warmup()
        prior_names = list(total_loss.prior_loss.priors)
        # Without early stopping nothing on the host depends on the per-epoch trace: the accumulator rows stay
        # on the device and are read back once at the end (the reference syncs with .item() every epoch).  Checkpoints
        # carry the trace so far (core.py:236-241), so they also need it on the host every epoch.
        deferred = not self.stop_early and not self.checkpoint_path
        rows = torch.zeros((self.n_epochs, engine.n_trace), dtype=torch.float64, device=self.device) if deferred else None
        filenames = []
        for epoch in range(self.n_epochs):
            for i in range(n_datasets):
                engine.step(i)
            filename = self._checkpoint(epoch, total_loss, components, calibrations, engine)
            if deferred:
                engine.trace_enqueue(rows[epoch])
                filenames.append(filename)
                continue
            ld, lp, lv = engine.trace_losses()
            total_loss.append_trace_values(ld, [lp] * len(prior_names), filename, lv if lv else None)
            if self._early_stop(total_loss.trace):
                self._rewind_draws(components, states, (epoch + 1) * (n_datasets + 1))
                break
        if deferred:
            host = rows.cpu().numpy()  # one synchronising read-back
            for vals, filename in zip(host, filenames):
                ld, lp, lv = engine.trace_decode(vals)
                total_loss.append_trace_values(ld, [lp] * len(prior_names), filename, lv if lv else None)
        torch.cuda.synchronize(self.device)

    def _run_joint(self, total_loss, components, shard, calibrations=None):
        """One joint Adam step per epoch (+ trace); dataset- and prior-sharded when `shard` is given."""
        states = self._draw_state(components)
        engine = self._build_engine(total_loss, components, self.n_epochs * 2, shard)
        self.engine = engine
        engine.warmup(joint=True)
        prior_names = list(total_loss.prior_loss.priors)
        deferred = not self.stop_early and not self.checkpoint_path
        rows = torch.zeros((self.n_epochs, engine.n_trace), dtype=torch.float64, device=self.device) if deferred else None
        filenames = []
        for epoch in range(self.n_epochs):
            engine.joint_step()
            filename = self._checkpoint(epoch, total_loss, components, calibrations, engine)
            if deferred:
                engine.trace_enqueue(rows[epoch])
                filenames.append(filename)
                continue
            ld, lp, lv = engine.trace_losses()
            total_loss.append_trace_values(ld, [lp] * len(prior_names), filename, lv if lv else None)
            if self._early_stop(total_loss.trace):
                self._rewind_draws(components, states, (epoch + 1) * 2)
                break
        if deferred:
            if engine.world > 1:  # every slot is written by one rank (prior: partial sums): one all-reduce for all epochs
                torch.distributed.all_reduce(rows, group=engine.pg)
            host = rows.cpu().numpy()
            for vals, filename in zip(host, filenames):
                ld, lp, lv = engine.trace_decode(vals)
                total_loss.append_trace_values(ld, [lp] * len(prior_names), filename, lv if lv else None)
        engine.sync_theta()
        torch.cuda.synchronize(self.device)
        engine.release_peer()  # the next run on this process group takes over the symmetric buffers

    def _run_autograd(self, total_loss, components, calibrations):
        """The reference loop verbatim (core.py:197-267) on the autograd bindings of the kernels."""
        parameters = list(components.parameters())
        if calibrations:
            parameters.extend(calibrations.parameters())
        self.optimizer = OPTIMIZER[self.optimizer_type](params=parameters, **self.optimizer_kwargs)
        for epoch in range(self.n_epochs):
            components.train()
            for counts, npred_model in total_loss.poisson_loss.iter_by_dataset:
                self.optimizer.zero_grad()
                fluxes = components.to_flux_tuple()
                npred = npred_model.evaluate(fluxes=fluxes)
                loss = total_loss.poisson_loss.loss_function(npred, counts)
                loss_prior = total_loss.prior_loss(fluxes=fluxes)
                loss_total = loss - self.beta * loss_prior / total_loss.prior_weight
                loss_total.backward()
                self.optimizer.step()
            components.eval()
            filename = self._checkpoint(epoch, total_loss, components, calibrations)
            total_loss.append_trace(fluxes=fluxes, filename=filename)
            if self._early_stop(total_loss.trace):
                break


class MAPDeconvolverResult:
    """MAP deconvolver result (core.py:285-471)."""

    def __init__(self, config, components, trace_loss, components_init=None, calibrations=None,
                 calibrations_init=None, wcs=None):
        self._components = components
        self._components_init = components_init
        self.trace_loss = trace_loss
        self._calibrations = calibrations
        self._calibrations_init = calibrations_init
        self._config = config
        self._wcs = wcs

    @property
    def components(self):
        return self._components

    @property
    def components_init(self):
        return self._components_init

    @property
    def calibrations(self):
        return self._calibrations

    @property
    def calibrations_init(self):
        return self._calibrations_init

    @property
    def flux_total(self):
        return self.components.flux_total_numpy

    @property
    def flux_upsampled_total(self):
        return self.components.flux_upsampled_total_numpy

    @property
    def config(self):
        return self._config

    @property
    def checkpoint_path(self):
        return Path(self.config.get("checkpoint_path", None))

    def read_checkpoint(self, epoch):
        """Checkpoint of `epoch` as a `MAPDeconvolverResult` (core.py:329-343)."""
        return self.__class__.read(self.checkpoint_path / self.trace_loss["filename"][epoch])

    def write(self, filename, overwrite=False, format=None):
        """Write the result (core.py:435-450).  The reference's FITS / ASDF writers are outside the hot path (no
        astropy / asdf here): the container is an `.npz` with the config (JSON), the loss trace, every component's
        upsampled flux (+ error, mask, parameterisation) and the calibration parameters - what `read` needs to hand
        back an equivalent result and what a resumed run needs to start from."""
        import json

        filename = Path(filename)
        if filename.exists() and not overwrite:
            raise IOError(f"{filename} already exists and overwrite is False")
        data = {"config": np.array(json.dumps({k: v for k, v in self.config.items() if _jsonable(v)}))}
        trace = self.trace_loss
        data["trace_colnames"] = np.array(list(trace.colnames))
        for name in trace.colnames:
            col = trace[name] if len(trace) else np.array([])
            data[f"trace__{name}"] = np.asarray(col) if name != "filename" else np.array([str(v) for v in col])
        data["component_names"] = np.array(list(self.components))
        for name, comp in self.components.items():
            data[f"flux_upsampled__{name}"] = comp.flux_upsampled_numpy
            data[f"parameter__{name}"] = comp._flux_upsampled.detach().cpu().numpy()  # log flux when use_log_flux
            data[f"component_meta__{name}"] = np.array([int(comp.upsampling_factor or 1), int(comp.use_log_flux),
                                                        int(comp.frozen)])
            if comp.flux_upsampled_error is not None:
                data[f"flux_upsampled_error__{name}"] = comp.flux_upsampled_error_numpy
            if comp.mask is not None:
                data[f"mask__{name}"] = comp.mask.detach().cpu().numpy()[0, 0]
        if self.calibrations:
            data["calibration_names"] = np.array(list(self.calibrations))
            for name, cal in self.calibrations.items():
                d = cal.to_dict()
                data[f"calibration__{name}"] = np.array([d["shift_x"], d["shift_y"], d["background_norm"], d["psf_scale"],
                                                         float(d.get("frozen", False))], dtype=np.float64)
        with open(filename, "wb") as fh:
            np.savez(fh, **data)

    @classmethod
    def read(cls, filename, format=None):
        """Read a result written by `write` (core.py:452-471).  Priors are not stored (the reference cannot serialise
        a non-registry GMM either, gmm.py:458-471): components come back with a uniform prior."""
        import json

        from .loss import TotalLoss  # noqa: F401  (trace table type)
        from .models import NPredCalibration, NPredCalibrations
        from .table import TraceTable

        with np.load(filename, allow_pickle=False) as f:
            config = json.loads(str(f["config"]))
            names = [str(n) for n in f["trace_colnames"]]
            trace = TraceTable(names=names)
            cols = {n: f[f"trace__{n}"] for n in names}
            for i in range(len(cols[names[0]]) if names else 0):
                trace.add_row({n: (str(cols[n][i]) if n == "filename" else float(cols[n][i])) for n in names})
            components = FluxComponents()
            for name in [str(n) for n in f["component_names"]]:
                factor, use_log, frozen = (int(v) for v in f[f"component_meta__{name}"])
                mask = f[f"mask__{name}"] if f"mask__{name}" in f else None
                theta = torch.from_numpy(f[f"parameter__{name}"])
                comp = SpatialFluxComponent(flux_upsampled=torch.ones_like(theta), use_log_flux=bool(use_log),
                                            mask=None if mask is None else torch.from_numpy(mask[None, None]),
                                            upsampling_factor=factor, frozen=bool(frozen), prior=UniformPrior())
                comp._flux_upsampled.data.copy_(theta)  # the stored parameter itself: lossless under masks
                if f"flux_upsampled_error__{name}" in f:
                    comp._flux_upsampled_error = torch.from_numpy(f[f"flux_upsampled_error__{name}"])
                components[name] = comp
            calibrations = None
            if "calibration_names" in f:
                calibrations = NPredCalibrations()
                for name in [str(n) for n in f["calibration_names"]]:
                    sx, sy, b, ps, frozen = f[f"calibration__{name}"]
                    calibrations[name] = NPredCalibration(shift_x=float(sx), shift_y=float(sy), background_norm=float(b),
                                                          psf_scale=float(ps), frozen=bool(frozen))
        return cls(config=config, components=components, trace_loss=trace, calibrations=calibrations)


def _jsonable(v):
    import json

    try:
        json.dumps(v)
        return True
    except TypeError:
        return False
