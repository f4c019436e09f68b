"""Priors of the MAP hot path behind the reference's class names.

Mirrors `jolideco/priors/core.py` (Prior, Priors, UniformPrior), `jolideco/priors/patches/gmm.py`
(GaussianMixtureModel, GaussianMixtureModelMeta) and `jolideco/priors/patches/core.py`
(GMMPatchPrior); arithmetic runs in the CUDA kernels of `libjolideco_b200.so`.
"""
import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import functional as F_b200
from . import ops
from ._lib import JolidecoB200Error

__all__ = ["Prior", "Priors", "UniformPrior", "GaussianMixtureModel", "GaussianMixtureModelMeta", "GMMPatchPrior",
           "compute_precision_cholesky", "get_pixel_weights"]


class Prior(nn.Module):
    """Prior base class; pickles the generator by state (priors/core.py:23-47)."""

    def __getstate__(self):
        state = self.__dict__.copy()
        generator = state.pop("generator", None)
        if generator:
            state["generator"] = generator.get_state()
            state["generator-device"] = generator.device
        return state

    def __setstate__(self, state):
        generator_state = state.pop("generator", None)
        generator_device = state.pop("generator-device", "cpu")
        if generator_state is not None:
            generator = torch.Generator(device=generator_device)
            generator.set_state(generator_state)
            state["generator"] = generator
        self.__dict__ = state

    def to_dict(self):
        for name, cls in PRIOR_REGISTRY.items():
            if isinstance(self, cls):
                return {"type": name}
        return {}


class Priors(nn.ModuleDict):
    """Dict of multiple priors (priors/core.py:86-107)."""

    def __call__(self, fluxes):
        value = 0
        for idx, prior in enumerate(self.values()):
            value += prior(flux=fluxes[idx])
        return value


class UniformPrior(Prior):
    """Uniform prior: log prior 0 (priors/core.py:110-129)."""

    def __call__(self, flux):
        return torch.tensor(0)


# ---------------------------------------------------------------------------------------------
# setup-time numpy helpers (utils/numpy.py:16-79); float64 like the reference
# ---------------------------------------------------------------------------------------------
def compute_precision_cholesky(covariances):
    """Precision Cholesky factors (chol(Sigma_k)^-1)^T, upper triangular (utils/numpy.py:16-34)."""
    covariances = np.asarray(covariances, dtype=np.float64)
    out = np.empty(covariances.shape)
    eye = np.eye(covariances.shape[1])
    for k, cov in enumerate(covariances):
        try:
            chol = np.linalg.cholesky(cov)
        except np.linalg.LinAlgError:
            raise ValueError(f"Cholesky decomposition failed for {cov}")
        out[k] = np.linalg.solve(chol, eye).T
    return out


def get_pixel_weights(patch_shape, stride):
    """Trapezoid pixel weights for overlapping patches, sum = stride^2 (utils/numpy.py:37-79)."""
    width = int(np.max(patch_shape))
    overlap = width - stride
    half = (width - 1.0) / 2
    x = np.linspace(-half, half, width)
    flat, slope = stride - overlap, 1.0 / overlap
    x2, x3 = min(-flat / 2.0, 0), max(flat / 2.0, 0)
    x1, x4 = x2 - 1.0 / slope, x3 + 1.0 / slope
    values = np.select(
        [np.logical_and(x >= x1, x < x2), np.logical_and(x >= x2, x < x3), np.logical_and(x >= x3, x < x4)],
        [slope * (x - x1), 1, slope * (x4 - x)],
    )
    weights = values * values[:, np.newaxis]
    return weights / weights.sum() * stride**2


@dataclass
class GaussianMixtureModelMeta:
    """GMM meta data (gmm.py:23-61): `stride` selects the pixel weights, `patch_norm` the patch norm."""

    stride: Optional[int] = None
    patch_norm: str = "subtract-mean"


class GaussianMixtureModel(nn.Module):
    """Gaussian mixture model container (gmm.py:64-299).

    Buffers are float32 exactly as the reference keeps them; the kernel constants are derived
    once per device by `packed()`.
    """

    def __init__(self, means, covariances, weights, precisions_cholesky, meta=None):
        super().__init__()
        self.register_buffer("means", means)
        self.register_buffer("covariances", covariances)
        self.register_buffer("weights", weights)
        self.register_buffer("precisions_cholesky", precisions_cholesky)
        self.meta = meta or GaussianMixtureModelMeta()
        self._packed = {}

    @classmethod
    def from_numpy(cls, means, covariances, weights, meta=None):
        precisions_cholesky = compute_precision_cholesky(covariances=covariances)
        return cls(
            means=torch.from_numpy(np.asarray(means).astype(np.float32)),
            covariances=torch.from_numpy(np.asarray(covariances).astype(np.float32)),
            weights=torch.from_numpy(np.asarray(weights).astype(np.float32)),
            precisions_cholesky=torch.from_numpy(precisions_cholesky.astype(np.float32)),
            meta=meta,
        )

    @classmethod
    def from_sklearn_gmm(cls, gmm):
        return cls.from_numpy(means=gmm.means_, covariances=gmm.covariances_, weights=gmm.weights_)

    @property
    def n_components(self):
        return int(self.covariances.shape[0])

    @property
    def n_features(self):
        return int(self.covariances.shape[1])

    @property
    def patch_shape(self):
        npix = int(round(math.sqrt(self.means.shape[-1])))
        return npix, npix

    @property
    def means_numpy(self):
        return self.means.detach().cpu().numpy()

    @property
    def covariances_numpy(self):
        return self.covariances.detach().cpu().numpy()

    @property
    def weights_numpy(self):
        return self.weights.detach().cpu().numpy()

    @property
    def precisions_cholesky_numpy(self):
        return self.precisions_cholesky.detach().cpu().numpy()

    @property
    def log_weights_numpy(self):
        return np.log(self.weights_numpy)

    @property
    def log_det_cholesky_numpy(self):
        return np.log(np.diagonal(self.precisions_cholesky_numpy, axis1=1, axis2=2)).sum(axis=1)

    @property
    def covariance_det(self):
        """Determinant of the first component's covariance (gmm.py:414-417)."""
        return np.linalg.det(self.covariances_numpy[0])

    def is_equal(self, other):
        """Same shapes and close covariances (gmm.py:447-452)."""
        if self.covariances.shape != other.covariances.shape:
            return False
        return bool(np.allclose(self.covariances_numpy, other.covariances_numpy))

    def reduce_to_topk(self, k):
        """The k components with the largest weights (gmm.py:391-412)."""
        idx = np.argsort(self.weights_numpy)[::-1][:k]
        return self.from_numpy(means=self.means_numpy[idx], covariances=self.covariances_numpy[idx],
                               weights=self.weights_numpy[idx], meta=self.meta)

    def estimate_log_prob_numpy(self, x):
        """Host (float64 numpy) evaluation of the per-component log likelihood (gmm.py:242-260) - setup-time
        inspection only; the MAP path uses the CUDA kernels."""
        x = np.asarray(x, dtype=np.float64)
        out = np.empty((x.shape[0], self.n_components))
        for k, (mu, L) in enumerate(zip(self.means_numpy.astype(np.float64), self.precisions_cholesky_numpy.astype(np.float64))):
            y = x @ L - mu @ L
            out[:, k] = np.sum(np.square(y) * self.pixel_weights_numpy, axis=1)
        return (-0.5 * (self.n_features * np.log(2 * np.pi) + out) + self.log_det_cholesky_numpy
                + self.log_weights_numpy)

    @property
    def pixel_weights_numpy(self):
        if self.meta.stride is None:
            weights = np.ones(self.patch_shape)
        else:
            weights = get_pixel_weights(patch_shape=self.patch_shape, stride=self.meta.stride)
        return weights.reshape((1, -1))

    @property
    def pixel_weights(self):
        return torch.from_numpy(self.pixel_weights_numpy.astype(np.float32)).to(self.means.device)

    @property
    def means_precisions_cholesky(self):
        return torch.einsum("ki,kij->kj", self.means, self.precisions_cholesky)

    @property
    def log_det_cholesky(self):
        return torch.log(torch.diagonal(self.precisions_cholesky, dim1=1, dim2=2)).sum(dim=1)

    @property
    def log_weights(self):
        return torch.log(self.weights)

    def packed(self, device):
        """Kernel constants (ops.GMMPacked) on `device`, built once."""
        key = (str(torch.device(device)), self.meta.stride)
        if key not in self._packed:
            self._packed[key] = ops.GMMPacked(
                self.means.detach().cpu().numpy(),
                self.precisions_cholesky.detach().cpu().numpy(),
                self.weights.detach().cpu().numpy(),
                self.pixel_weights_numpy,
                device,
            )
        return self._packed[key]

    def estimate_log_prob(self, x):
        """Per-component log likelihood of feature vectors x (P, D) (gmm.py:262-281)."""
        if not x.is_cuda:
            raise JolidecoB200Error("estimate_log_prob: x must be a CUDA tensor (no CPU path)")
        return ops.gmm_log_prob(x.contiguous().to(torch.float32), self.packed(x.device))


class GMMPatchPrior(Prior):
    """GMM patch prior (priors/patches/core.py:30-246).

    Same constructor as the reference.  The cycle-spin shifts are drawn on the HOST from
    `generator` exactly as `utils/torch.py:108-116` does (two `torch.randint(-2, 3)`, row shift
    first), so a seeded CPU generator reproduces the reference's CPU trajectory; the reference's
    CUDA-generator stream is not reproduced.
    """

    def __init__(self, gmm=None, stride=None, cycle_spin=True, cycle_spin_subpix=False, generator=None, norm=None,
                 patch_norm=None, jitter=False, marginalize=False, device="cpu", backend=None):
        super().__init__()
        if gmm is None:
            raise JolidecoB200Error("GMMPatchPrior: the packaged GMM library is not available offline, pass `gmm=`")
        self.gmm = gmm
        self.stride = gmm.meta.stride if stride is None else stride
        if self.stride is None:
            raise ValueError("GMMPatchPrior: stride is undefined (gmm.meta.stride is None and no stride given)")
        self.cycle_spin = cycle_spin
        if cycle_spin_subpix or jitter:
            raise NotImplementedError("cycle_spin_subpix / jitter are outside the accelerated hot path")
        self.cycle_spin_subpix, self.jitter = False, False
        self.generator = torch.Generator(device="cpu") if generator is None else generator
        self.norm = norm
        patch_norm = gmm.meta.patch_norm if patch_norm is None else patch_norm
        if not (patch_norm == "subtract-mean" or type(patch_norm).__name__ == "SubtractMeanPatchNorm"):
            raise NotImplementedError("only the subtract-mean patch norm is fused into the prior kernel")
        self.patch_norm = "subtract-mean"
        self.marginalize = marginalize
        self.device = torch.device(device)
        self.backend = backend

    @property
    def patch_shape(self):
        return self.gmm.patch_shape

    @property
    def log_like_weight(self):
        return self.stride**2 / (self.patch_shape[0] * self.patch_shape[1])

    @property
    def overlap(self):
        return max(self.patch_shape) - self.stride

    def draw_shifts(self):
        """The two randint draws of `cycle_spin` (utils/torch.py:108-116): (row shift, col shift)."""
        if not self.cycle_spin:
            return 0, 0
        w = self.patch_shape[0] // 4
        sy = int(torch.randint(-w, w + 1, (1,), generator=self.generator, device=self.generator.device))
        sx = int(torch.randint(-w, w + 1, (1,), generator=self.generator, device=self.generator.device))
        return sy, sx

    def __call__(self, flux, mask=None, shift_yx=None):
        if not flux.is_cuda:
            raise JolidecoB200Error("GMMPatchPrior: flux must be a CUDA tensor (no CPU path)")
        if self.norm is not None:
            flux = self.norm(flux)
        if shift_yx is None:
            shift_yx = self.draw_shifts()
        backend = default_backend() if self.backend is None else self.backend
        return F_b200.gmm_patch_prior(flux, self.gmm.packed(flux.device), shift_yx, self.stride, self.marginalize,
                                      backend)


# 3 = tcgen05 split-TF32 forward with the correction products on the FP16 pipe, persistent stream-K (jd_gmm_tcm.cu);
# 4 = the same split, two patch tiles per CTA and staged operand image, per-tile issuer / epilogue / slots (jd_gmm_tcm2.cu);
# 5 = that kernel with the split-FP16 recipe (12 MMAs, 3 accumulator slots per tile, half the operand bytes);
# 1 = tcgen05 3 x TF32 (jd_gmm_tc.cu), 2 = tcgen05 split-FP16 (jd_gmm_tc16.cu), 0 = FP32 CUDA-core check path.
# JD_PRIOR_BACKEND overrides the default (A/B runs).
_DEFAULT_BACKEND = int(os.environ.get("JD_PRIOR_BACKEND", "4"))


def default_backend():
    return _DEFAULT_BACKEND


def set_default_backend(backend):
    """0 = FP32 CUDA-core prior kernel, 1 / 2 / 3 = the tcgen05 kernels (see _DEFAULT_BACKEND)."""
    global _DEFAULT_BACKEND
    _DEFAULT_BACKEND = int(backend)


PRIOR_REGISTRY = {"uniform": UniformPrior, "gmm-patches": GMMPatchPrior}
