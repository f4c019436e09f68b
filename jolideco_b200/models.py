"""Flux components and the NPred forward model behind the reference's class names.

Mirrors `jolideco/models/core.py` (SpatialFluxComponent, FluxComponents) and
`jolideco/models/npred.py` (NPredModel, NPredModels, NPredCalibration, NPredCalibrations).
Setup (bilinear upsampling, PSF normalisation, exposure edge correction) happens once on the host
as in the reference (npred.py:66-115); the per-iteration arithmetic runs in the CUDA kernels.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_b200
from . import ops
from ._lib import JolidecoB200Error
from .priors import Prior, Priors, UniformPrior

__all__ = ["SpatialFluxComponent", "FluxComponents", "NPredModel", "NPredModels", "NPredCalibration",
           "NPredCalibrations"]


class SpatialFluxComponent(nn.Module):
    """Flux component (models/core.py:354-717): log-flux parameter, optional mask, upsampling."""

    is_sparse = False

    def __init__(self, flux_upsampled, flux_upsampled_error=None, mask=None, use_log_flux=True, upsampling_factor=1,
                 prior=None, frozen=False, wcs=None):
        super().__init__()
        if not flux_upsampled.ndim == 4:
            raise ValueError(f"Flux tensor must be four dimensional. Got {flux_upsampled.ndim}")
        if use_log_flux:
            flux_upsampled = torch.log(flux_upsampled)
        self._flux_upsampled = nn.Parameter(flux_upsampled)
        self._flux_upsampled_error = flux_upsampled_error
        if mask is not None and not mask.shape == flux_upsampled.shape:
            raise ValueError(f"Flux and mask need to have the same shape, got {flux_upsampled.shape} and {mask.shape}")
        if mask is not None:
            self.register_buffer("mask", mask.to(torch.uint8))
        else:
            self.mask = None
        self._use_log_flux = use_log_flux
        self.upsampling_factor = int(upsampling_factor)
        self.prior = UniformPrior() if prior is None else prior
        self.frozen = frozen
        self._wcs = wcs

    @classmethod
    def from_numpy(cls, flux, mask=None, **kwargs):
        """Create from a 2-D flux init array, upsampled bilinearly (models/core.py:505-540)."""
        upsampling_factor = kwargs.get("upsampling_factor", None)
        flux = torch.from_numpy(flux[np.newaxis, np.newaxis].astype(np.float32))
        if upsampling_factor:
            flux = F.interpolate(flux, scale_factor=upsampling_factor, mode="bilinear")
        if mask is not None:
            mask = torch.from_numpy(mask[np.newaxis, np.newaxis].astype(bool))
            if upsampling_factor:
                mask = F.interpolate(mask.type(torch.float32), scale_factor=upsampling_factor, mode="bilinear") > 0.5
        return cls(flux_upsampled=flux, mask=mask, **kwargs)

    @classmethod
    def from_flux_init_datasets(cls, datasets, **kwargs):
        fluxes = [d["counts"] / d["exposure"] - d["background"] for d in datasets]
        return cls.from_numpy(flux=np.nanmean(fluxes, axis=0), **kwargs)

    def parameters(self, recurse=True):
        if self.frozen:
            return []
        return super().parameters(recurse)

    @property
    def wcs(self):
        return self._wcs

    @property
    def shape(self):
        return self._flux_upsampled.shape

    @property
    def shape_image(self):
        return self.shape[-2:]

    @property
    def use_log_flux(self):
        return self._use_log_flux

    @property
    def flux_upsampled(self):
        theta = self._flux_upsampled
        if not theta.is_cuda:  # host-side inspection before/after a run: plain torch
            flux = torch.exp(theta) if self.use_log_flux else theta
            return flux * self.mask if self.mask is not None else flux
        return F_b200.flux_from_theta(theta, self.mask, self.use_log_flux)

    @property
    def flux(self):
        flux = self.flux_upsampled
        if self.upsampling_factor:
            flux = F.avg_pool2d(flux, kernel_size=self.upsampling_factor, divisor_override=1)
        return flux

    @property
    def flux_upsampled_error(self):
        return self._flux_upsampled_error

    @property
    def flux_upsampled_error_numpy(self):
        return self.flux_upsampled_error.detach().cpu().numpy()[0, 0]

    @property
    def flux_numpy(self):
        return self.flux.detach().cpu().numpy()[0, 0]

    @property
    def flux_upsampled_numpy(self):
        return self.flux_upsampled.detach().cpu().numpy()[0, 0]

    def to_dict(self, include_data=None):
        data = {"use_log_flux": self.use_log_flux, "upsampling_factor": int(self.upsampling_factor),
                "frozen": self.frozen, "prior": self.prior.to_dict()}
        if include_data == "numpy":
            data["flux_upsampled"] = self.flux_upsampled_numpy
        return data


class FluxComponents(nn.ModuleDict):
    """Flux components (models/core.py:720-933)."""

    def parameters(self):
        parameters = []
        for component in self.values():
            if not component.frozen:
                parameters.extend(component.parameters())
        return parameters

    @property
    def priors(self):
        priors = Priors()
        for name, component in self.items():
            priors[name] = component.prior
        return priors

    @property
    def fluxes_numpy(self):
        return {name: c.flux_numpy for name, c in self.items()}

    @property
    def fluxes_upsampled_numpy(self):
        return self.to_numpy()

    @property
    def flux_upsampled_total(self):
        """Total summed flux as a tensor (models/core.py:746-757)."""
        values = list(self.values())
        flux = torch.zeros_like(values[0].flux_upsampled)
        for component in values:
            flux = flux + component.flux_upsampled
        return flux

    @property
    def flux_upsampled_total_numpy(self):
        return np.sum([flux for flux in self.fluxes_upsampled_numpy.values()], axis=0)

    @property
    def flux_total_numpy(self):
        return np.sum([flux for flux in self.fluxes_numpy.values()], axis=0)

    def to_numpy(self):
        return {name: np.squeeze(c.flux_upsampled.detach().cpu().numpy()) for name, c in self.items()}

    def to_flux_tuple(self):
        return tuple([_.flux_upsampled for _ in self.values()])

    def set_flux_errors(self, flux_errors):
        for name, flux_error in flux_errors.items():
            self[name]._flux_upsampled_error = flux_error

    def to_dict(self, include_data=None):
        return {name: c.to_dict(include_data=include_data) for name, c in self.items()}


def _convolve_fft_host(image, kernel):
    """Setup-time restatement of utils/torch.py:347-370 on the host (used once per dataset for the
    exposure edge correction, npred.py:108-113)."""
    s = [image.shape[-2] + kernel.shape[-2] - 1, image.shape[-1] + kernel.shape[-1] - 1]
    res = torch.fft.irfft2(torch.fft.rfft2(image, s=s) * torch.fft.rfft2(kernel, s=s), s=s)
    y0, x0 = (s[0] - image.shape[-2]) // 2, (s[1] - image.shape[-1]) // 2
    return res[..., y0 : y0 + image.shape[-2], x0 : x0 + image.shape[-1]]


class NPredModel(nn.Module):
    """Predicted counts model of one component (models/npred.py:31-191)."""

    def __init__(self, exposure, psf=None, rmf=None, upsampling_factor=None):
        super().__init__()
        if rmf is not None:
            raise NotImplementedError("rmf (energy redistribution) is outside the accelerated hot path")
        if exposure.ndim != 4 or exposure.shape[0] != 1 or exposure.shape[1] != 1:
            raise NotImplementedError("only single-channel (1,1,H,W) exposures are supported")
        self.register_buffer("exposure", exposure.contiguous())
        self.register_buffer("psf", None if psf is None else psf.contiguous())
        self.rmf = None
        self.upsampling_factor = upsampling_factor

    @property
    def shape_upsampled(self):
        return tuple(self.exposure.shape)

    @property
    def shape(self):
        shape = list(self.shape_upsampled)
        shape[-1] //= self.upsampling_factor
        shape[-2] //= self.upsampling_factor
        return tuple(shape)

    @classmethod
    def from_numpy(cls, exposure, psf, upsampling_factor, correct_exposure_edges=True, device=None):
        """NPredModel.from_numpy (npred.py:66-115): bilinear upsampling of exposure and PSF, PSF / f^2, exposure edge
        correction exposure /= PSF (*) 1.  With a CUDA `device` the arrays are uploaded first and the whole setup
        runs there (the edge correction through the library's own convolution), which removes ~5 ms of host FFTs per
        1024^2 dataset from `MAPDeconvolver.run`; without, on the host as the reference does."""
        dims = (np.newaxis, np.newaxis)
        on_gpu = device is not None and torch.device(device).type == "cuda"
        kwargs = {
            "upsampling_factor": upsampling_factor,
            "exposure": torch.from_numpy(np.ascontiguousarray(np.asarray(exposure, dtype=np.float32)[dims])),
            "psf": torch.from_numpy(np.ascontiguousarray(np.asarray(psf, dtype=np.float32)[dims])),
        }
        for name in ["psf", "exposure"]:
            tensor = kwargs[name]
            if on_gpu:
                tensor = tensor.to(device)
            if upsampling_factor and upsampling_factor != 1:  # scale factor 1 is the identity (npred.py:96 still calls it)
                tensor = F.interpolate(tensor, scale_factor=upsampling_factor, mode="bilinear")
            if name == "psf" and upsampling_factor:
                tensor = tensor / upsampling_factor**2
            kwargs[name] = tensor
        if correct_exposure_edges:
            exposure_t = kwargs["exposure"]
            if on_gpu:
                from . import ops

                psf_t = kwargs["psf"][0, 0].contiguous()
                ones = torch.ones_like(exposure_t[0, 0])
                with torch.cuda.device(device):
                    if psf_t.shape[0] * psf_t.shape[1] >= ops.FFT_MIN_PSF_AREA:
                        weights = ops.conv_forward_fft(ones, ones, ops.FFTConvPlan(psf_t, *ones.shape))
                    else:
                        weights = ops.conv_forward(ones, ones, psf_t)
                weights = weights[None, None]
            else:
                weights = _convolve_fft_host(torch.ones_like(exposure_t), kwargs["psf"])
            kwargs["exposure"] = exposure_t / weights
        return cls(**kwargs)

    @classmethod
    def from_dataset_numpy(cls, dataset, upsampling_factor=None, correct_exposure_edges=True, device=None):
        return cls.from_numpy(exposure=dataset["exposure"], psf=dataset["psf"], upsampling_factor=upsampling_factor,
                              correct_exposure_edges=correct_exposure_edges, device=device)

    def forward(self, flux, psf_scale=None):
        if psf_scale is not None and not bool(torch.isclose(psf_scale.detach().cpu(), torch.tensor(1.0)).all()):
            raise NotImplementedError("psf_scale != 1 is outside the accelerated hot path")
        if not flux.is_cuda:
            raise JolidecoB200Error("NPredModel: flux must be a CUDA tensor (no CPU path)")
        if self.psf is None:
            raise NotImplementedError("NPredModel without a PSF")
        return F_b200.npred_forward(flux, self.exposure, self.psf, self.upsampling_factor or 1)


class NPredModels(nn.ModuleDict):
    """Sum of the per-component NPred models plus background (models/npred.py:194-295)."""

    def __init__(self, background, calibration=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("background", background)
        self.calibration = calibration

    def evaluate_per_component(self, fluxes):
        npreds = {}
        for (name, npred_model), flux in zip(self.items(), fluxes):
            if self.calibration is not None:
                flux = self.calibration(flux=flux, scale=npred_model.upsampling_factor)
                npreds[name] = npred_model(flux=flux, psf_scale=self.calibration.psf_scale)
            else:
                npreds[name] = npred_model(flux=flux)
        if self.calibration is not None:
            npreds["background"] = self.background * self.calibration.background_norm
        else:
            npreds["background"] = self.background
        return npreds

    def evaluate(self, fluxes):
        npreds = self.evaluate_per_component(fluxes=fluxes)
        npred_total = torch.zeros(self.background.shape, device=fluxes[0].device)
        for npred in npreds.values():
            npred_total = npred_total + npred
        return npred_total

    @classmethod
    def from_dataset_numpy(cls, dataset, components, calibration=None, device=None):
        values = []
        for name, component in components.items():
            psf = dataset["psf"]
            if isinstance(psf, dict):
                psf = psf[name]
            npred_model = NPredModel.from_numpy(exposure=dataset["exposure"], psf=psf,
                                                upsampling_factor=component.upsampling_factor, device=device)
            values.append((name, npred_model))
        background = torch.from_numpy(np.ascontiguousarray(np.asarray(dataset["background"], dtype=np.float32)[
            np.newaxis, np.newaxis]))
        return cls(background, calibration, values)


class NPredCalibration(nn.Module):
    """Dataset calibration parameters (models/npred.py:298-402)."""

    def __init__(self, shift_x=0.0, shift_y=0.0, background_norm=1.0, psf_scale=1.0, frozen=False, weight=1.0):
        super().__init__()
        self.shift_xy = nn.Parameter(torch.tensor([[shift_x, shift_y]]))
        self._background_norm = nn.Parameter(torch.log(torch.tensor([background_norm])))
        self.psf_scale = nn.Parameter(torch.tensor([psf_scale]), requires_grad=False)
        self.frozen = frozen
        self.weight = weight

    @property
    def background_norm(self):
        return torch.exp(self._background_norm)

    def parameters(self, recurse=True):
        if self.frozen:
            return []
        return super().parameters(recurse)

    def to_dict(self):
        shift_xy = self.shift_xy.detach().cpu().numpy()
        return {"shift_x": shift_xy[0, 0].item(), "shift_y": shift_xy[0, 1].item(),
                "background_norm": self.background_norm.detach().cpu().numpy().item(),
                "psf_scale": self.psf_scale.detach().cpu().numpy().item(), "frozen": self.frozen,
                "weight": float(self.weight)}

    @classmethod
    def from_dict(cls, data):
        return cls(**data)

    def __call__(self, flux, scale):
        """Sub-pixel shift (utils/torch.py:196-223): identity (no graph) when the shift is ~0."""
        shift_xy = self.shift_xy
        if bool(torch.all(torch.isclose(shift_xy, torch.zeros_like(shift_xy)))):
            return flux
        size = flux.size()
        sc = 2 * scale / torch.tensor([[size[-1]], [size[-2]]], device=flux.device)
        theta = torch.cat([torch.eye(2, device=flux.device), sc * shift_xy.T], dim=1)[None]
        grid = F.affine_grid(theta=theta, size=size, align_corners=False)
        return F.grid_sample(flux, grid=grid, align_corners=False)


class NPredCalibrations(nn.ModuleDict):
    """Calibration components (models/npred.py:405-510)."""

    def parameters(self, recurse=True):
        parameters = []
        for model in self.values():
            if not model.frozen:
                parameters.extend(list(model.parameters()))
        return parameters

    def to_dict(self):
        return {name: model.to_dict() for name, model in self.items()}

    @classmethod
    def from_dict(cls, data):
        return cls([(name, NPredCalibration.from_dict(data=d)) for name, d in data.items()])
