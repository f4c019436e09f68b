"""Poisson / prior / total loss behind the reference's class names (jolideco/loss.py)."""
import numpy as np
import torch

from . import functional as F_b200
from .models import NPredModels
from .table import TraceTable

__all__ = ["PoissonLoss", "PriorLoss", "TotalLoss"]


class _PoissonNLL:
    """Callable with the semantics of nn.PoissonNLLLoss(log_input=False, reduction="mean",
    eps=1e-25, full=True) (loss.py:35-37), evaluated by the fused CUDA reduction."""

    eps = 1e-25

    def __call__(self, npred, counts):
        return F_b200.poisson_nll(npred, counts)


class PoissonLoss:
    """Poisson loss per dataset (loss.py:13-133)."""

    def __init__(self, counts_all, npred_models_all, names_all):
        if len(counts_all) != len(npred_models_all):
            raise ValueError("counts_all and npred_models_all must have the same length")
        self.counts_all = counts_all
        self.npred_models_all = npred_models_all
        self.loss_function = _PoissonNLL()
        self.names_all = names_all

    @property
    def weights(self):
        weights = [m.calibration.weight for m in self.npred_models_all if m.calibration is not None]
        return torch.tensor(weights)

    @property
    def n_datasets(self):
        return len(self.counts_all)

    def evaluate(self, fluxes):
        loss_datasets = []
        for counts, npred_model in zip(self.counts_all, self.npred_models_all):
            npred = npred_model.evaluate(fluxes=fluxes)
            loss_datasets.append(self.loss_function(npred, counts))
        return torch.stack(loss_datasets) if loss_datasets and loss_datasets[0].requires_grad else torch.tensor(
            [float(_) for _ in loss_datasets])

    @property
    def iter_by_dataset(self):
        for data in zip(self.counts_all, self.npred_models_all):
            yield data

    @classmethod
    def from_datasets(cls, datasets, components, calibrations=None, device="cuda"):
        npred_models_all, counts_all = [], []
        for name, dataset in datasets.items():
            calibration = calibrations[name] if calibrations else None
            npred_models = NPredModels.from_dataset_numpy(dataset=dataset, components=components,
                                                          calibration=calibration, device=device)
            npred_models_all.append(npred_models.to(device))
            # the engine and the kernels are float32 (the reference's data helpers default to it; integer counts and
            # float64 arrays, which the reference promotes on the fly, are cast once here)
            counts = torch.from_numpy(np.ascontiguousarray(np.asarray(dataset["counts"], dtype=np.float32)[
                np.newaxis, np.newaxis])).to(device)
            counts_all.append(counts)
        return cls(counts_all=counts_all, npred_models_all=npred_models_all, names_all=list(datasets))

    def __call__(self, fluxes):
        losses = self.evaluate(fluxes=fluxes)
        if len(self.weights):
            losses = losses * self.weights.to(losses.device)
        return torch.sum(losses)


class PriorLoss:
    """Prior loss (loss.py:136-168)."""

    def __init__(self, priors):
        self.priors = priors

    def evaluate(self, fluxes):
        return [prior(flux=flux) for flux, prior in zip(fluxes, self.priors.values())]

    def __call__(self, fluxes):
        return sum(self.evaluate(fluxes=fluxes))


class TotalLoss:
    """Total loss with trace (loss.py:171-360)."""

    def __init__(self, poisson_loss, prior_loss, poisson_loss_validation=None, beta=1):
        self.poisson_loss = poisson_loss
        self.poisson_loss_validation = poisson_loss_validation
        self.prior_loss = prior_loss
        self.beta = beta
        self._trace = None
        # names of ALL datasets of the run (differs from poisson_loss.names_all when the datasets are sharded
        # over ranks: each rank only holds its own)
        self.dataset_names = list(poisson_loss.names_all)

    @property
    def trace(self):
        if self._trace is None:
            names = ["total", "datasets-total", "priors-total"]
            names += [f"prior-{name}" for name in self.prior_loss.priors]
            names += [f"dataset-{name}" for name in self.dataset_names]
            if self.poisson_loss_validation or self.poisson_loss_validation is False:
                names += ["datasets-validation-total"]
            names += ["filename"]
            self._trace = TraceTable(names=names, dtype=[float] * (len(names) - 1) + [str])
        return self._trace

    def append_trace_values(self, loss_datasets, loss_priors, filename, loss_datasets_validation=None):
        """Row assembly of loss.py:226-250 from already evaluated (host float) losses."""
        loss_datasets_total = sum(loss_datasets)
        loss_priors_total = self.beta * sum(loss_priors)
        row = {"total": loss_datasets_total - loss_priors_total, "datasets-total": loss_datasets_total,
               "priors-total": -loss_priors_total, "filename": filename}
        for name, value in zip(self.prior_loss.priors, loss_priors):
            row[f"prior-{name}"] = -self.beta * value
        for name, value in zip(self.dataset_names, loss_datasets):
            row[f"dataset-{name}"] = value
        if loss_datasets_validation is not None:
            row["datasets-validation-total"] = sum(loss_datasets_validation)
        self.trace.add_row(row)

    @torch.no_grad()
    def append_trace(self, fluxes, filename):
        loss_datasets = [float(_) for _ in self.poisson_loss.evaluate(fluxes=fluxes)]
        loss_priors = [float(_) for _ in self.prior_loss.evaluate(fluxes=fluxes)]
        validation = None
        if self.poisson_loss_validation:
            validation = [float(_) for _ in self.poisson_loss_validation.evaluate(fluxes=fluxes)]
        self.append_trace_values(loss_datasets, loss_priors, filename, validation)

    @property
    def prior_weight(self):
        return len(self.poisson_loss.counts_all)

    def __call__(self, fluxes):
        loss_datasets = self.poisson_loss.evaluate(fluxes=fluxes)
        loss_priors = self.prior_loss.evaluate(fluxes=fluxes)
        return sum(loss_datasets) - self.beta * sum(loss_priors)

    @classmethod
    def from_datasets_and_components(cls, datasets, components, datasets_validation=None, beta=1, calibrations=None,
                                     device="cuda"):
        poisson_loss = PoissonLoss.from_datasets(datasets=datasets, components=components, device=device,
                                                 calibrations=calibrations)
        poisson_loss_validation = None
        if datasets_validation:
            poisson_loss_validation = PoissonLoss.from_datasets(datasets=datasets_validation, components=components,
                                                                calibrations=calibrations, device=device)
        return cls(poisson_loss=poisson_loss, poisson_loss_validation=poisson_loss_validation,
                   prior_loss=PriorLoss(priors=components.priors), beta=beta)
