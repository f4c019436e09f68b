"""Image norms applied to the flux before patch extraction (jolideco/utils/norms.py:115-420).

These are elementwise torch expressions with (optionally trainable) scalar parameters; they stay in torch autograd
AROUND the custom prior op (`GMMPatchPrior(norm=...)` applies the norm, then calls the CUDA prior), exactly as SURVEY
8f ranks them - off the accelerated path, which covers the default identity norm.  Same class names, constructor
arguments, `to_dict` / `from_dict` and registry keys as the reference; the classes are generated from one table of
(parameters, forward, inverse) so that each transfer function is stated once.
"""
import math

import numpy as np
import torch

__all__ = ["ImageNorm", "IdentityImageNorm", "ASinhImageNorm", "MaxImageNorm", "FixedMaxImageNorm", "SigmoidImageNorm",
           "ATanImageNorm", "LogImageNorm", "PowerImageNorm", "InverseCDFImageNorm", "NORMS_REGISTRY"]


class ImageNorm(torch.nn.Module):
    """Base class: `frozen` hides the parameters from the optimiser (norms.py:118-128)."""

    _params = ()          # ((name, default, trainable), ...)
    registry_key = None

    def __init__(self, *args, frozen=False, **kwargs):
        super().__init__()
        self.frozen = frozen
        values = dict(zip([p[0] for p in self._params], args))
        values.update(kwargs)
        for name, default, trainable in self._params:
            if name not in values and default is None:
                raise TypeError(f"{type(self).__name__}: missing argument {name!r}")
            value = torch.tensor([float(values.get(name, default))])
            if trainable:
                setattr(self, name, torch.nn.Parameter(value))
            else:
                self.register_buffer(name, value)

    def parameters(self, recurse=True):
        return [] if self.frozen else super().parameters(recurse)

    def to_dict(self):
        data = {"type": self.registry_key} if self.registry_key else {}
        data.update({name: float(getattr(self, name).detach()) for name, _, _ in self._params})
        return data

    @classmethod
    def from_dict(cls, data):
        kwargs = dict(data)
        if "type" in kwargs:
            cls = NORMS_REGISTRY[kwargs.pop("type")]
        return cls(**kwargs)

    def forward(self, image):
        raise NotImplementedError

    def inverse(self, image):
        raise NotImplementedError

    def __call__(self, image):  # the reference defines __call__ directly (no hooks)
        return self.forward(image)

    def evaluate_numpy(self, image):
        return self(torch.from_numpy(np.asarray(image).astype(np.float32))).detach().numpy()

    def inverse_numpy(self, image):
        return self.inverse(torch.from_numpy(np.asarray(image).astype(np.float32))).detach().numpy()

    def __str__(self):
        args = ", ".join(f"{k}={v}" for k, v in self.to_dict().items() if k != "type")
        return f"{type(self).__name__}({args})"


def _make(name, key, params, forward, inverse=None, doc=""):
    body = {"_params": tuple(params), "registry_key": key, "forward": forward, "__doc__": doc}
    if inverse is not None:
        body["inverse"] = inverse
    return type(name, (ImageNorm,), body)


# name, registry key, ((parameter, default, trainable), ...), forward, inverse            reference lines
IdentityImageNorm = _make("IdentityImageNorm", "identity", (), lambda s, x: x, lambda s, y: y,
                          "y = x (norms.py:225-232)")
ASinhImageNorm = _make("ASinhImageNorm", "asinh", (("alpha", 1.0, True), ("beta", 1.0, True)),
                       lambda s, x: torch.asinh(x / s.alpha) / torch.asinh(s.beta / s.alpha),
                       lambda s, y: s.alpha * torch.sinh(y * torch.asinh(s.beta / s.alpha)),
                       "y = asinh(x / alpha) / asinh(beta / alpha) (norms.py:235-257)")
MaxImageNorm = _make("MaxImageNorm", "max", (), lambda s, x: x / x.max(), None, "y = x / max(x) (norms.py:260-272)")
FixedMaxImageNorm = _make("FixedMaxImageNorm", "fixed-max", (("max_value", None, True),),
                          lambda s, x: torch.clip(x / s.max_value, min=0, max=1), lambda s, y: y * s.max_value,
                          "y = clip(x / max_value, 0, 1) (norms.py:275-293)")
SigmoidImageNorm = _make("SigmoidImageNorm", "sigmoid", (("alpha", 1.0, True), ("beta", 1.0, True)),
                         lambda s, x: 1 / (1 + torch.exp(-(x - s.beta / 2.0) / s.alpha)),
                         lambda s, y: s.alpha * torch.log(y / (1.0 - y)) + s.beta / 2.0,
                         "y = 1 / (1 + exp(-(x - beta / 2) / alpha)) (norms.py:296-316)")
ATanImageNorm = _make("ATanImageNorm", "atan", (("alpha", 1.0, True),),
                      lambda s, x: 2 * torch.atan(x / s.alpha) / math.pi, lambda s, y: 0.5 * math.pi * torch.tan(y),
                      "y = 2 atan(x / alpha) / pi; `inverse` as the reference states it (norms.py:319-337)")
LogImageNorm = _make("LogImageNorm", "log", (("alpha", 1.0, True),), lambda s, x: torch.log(x / s.alpha),
                     lambda s, y: s.alpha * torch.exp(y), "y = log(x / alpha) (norms.py:372-390)")
PowerImageNorm = _make("PowerImageNorm", "power", (("alpha", 1.0, True), ("beta", 1.0, False)),
                       lambda s, x: torch.pow(x / s.beta, s.alpha), lambda s, y: s.beta * torch.pow(y, 1 / s.alpha),
                       "y = (x / beta) ** alpha, beta a fixed buffer (norms.py:393-413)")



class InverseCDFImageNorm(ImageNorm):
    """Histogram equalisation: y = cdf(x), piecewise linear between the tabulated points (norms.py:340-369; the
    interpolation is `interp1d_torch`, utils/torch.py:146-169: segment index searchsorted(xp, x) clipped to
    [0, len - 2], values taken at (index - 1, index), linear extrapolation outside)."""

    registry_key = "inverse-cdf"

    def __init__(self, x, cdf):
        super().__init__()
        if x.shape != cdf.shape:
            raise ValueError(f"'x' and 'cdf' must have same shape, got {x.shape} and {cdf.shape}")
        self.x, self.cdf = x, cdf

    @classmethod
    def from_image(cls, image, bins=1000):
        weights, edges = torch.histogram(torch.from_numpy(image), bins=bins)
        cdf = torch.cumsum(weights, 0)
        cdf = (cdf - cdf.min()) / (cdf - cdf.min()).max()
        return cls(x=(edges[1:] + edges[:-1]) / 2, cdf=cdf)

    def forward(self, image):
        hi = torch.clip(torch.searchsorted(self.x, image), 0, len(self.x) - 2)
        x0, x1, y0, y1 = self.x[hi - 1], self.x[hi], self.cdf[hi - 1], self.cdf[hi]
        return torch.lerp(y0, y1, (image - x0) / (x1 - x0))

    def to_dict(self):
        raise NotImplementedError


NORMS_REGISTRY = {cls.registry_key: cls for cls in (InverseCDFImageNorm, MaxImageNorm, FixedMaxImageNorm, SigmoidImageNorm, ATanImageNorm,
                                                    ASinhImageNorm, IdentityImageNorm, LogImageNorm, PowerImageNorm)}
