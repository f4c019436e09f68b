"""Image norms (jolideco_b200/norms.py) against formulas of the reference (utils/norms.py:225-413), values and
autograd gradients on CPU (they are plain torch modules around the CUDA prior op), plus the dict round trip."""
import math

import numpy as np
import pytest
import torch

from jolideco_b200 import norms as N

X = torch.tensor(np.random.default_rng(0).gamma(2.0, size=(1, 1, 6, 7)).astype(np.float32))

CASES = [
    (N.IdentityImageNorm, {}, lambda x, p: x),
    (N.ASinhImageNorm, dict(alpha=0.7, beta=2.5), lambda x, p: torch.asinh(x / p["alpha"]) / math.asinh(p["beta"] / p["alpha"])),
    (N.MaxImageNorm, {}, lambda x, p: x / x.max()),
    (N.FixedMaxImageNorm, dict(max_value=3.0), lambda x, p: torch.clip(x / p["max_value"], 0, 1)),
    (N.SigmoidImageNorm, dict(alpha=0.8, beta=1.5), lambda x, p: torch.sigmoid((x - p["beta"] / 2) / p["alpha"])),
    (N.ATanImageNorm, dict(alpha=1.7), lambda x, p: 2 * torch.atan(x / p["alpha"]) / math.pi),
    (N.LogImageNorm, dict(alpha=0.5), lambda x, p: torch.log(x / p["alpha"])),
    (N.PowerImageNorm, dict(alpha=0.6, beta=2.0), lambda x, p: (x / p["beta"]) ** p["alpha"]),
]


@pytest.mark.parametrize("cls,params,formula", CASES, ids=[c[0].__name__ for c in CASES])
def test_norm_values_gradients_and_dict_round_trip(cls, params, formula):
    norm = cls(**params)
    x = X.clone().requires_grad_(True)
    y = norm(x)
    ref_x = X.clone().requires_grad_(True)
    ref = formula(ref_x, params)
    torch.testing.assert_close(y, ref, rtol=2e-6, atol=1e-7)
    y.sum().backward()
    ref.sum().backward()
    torch.testing.assert_close(x.grad, ref_x.grad, rtol=2e-5, atol=1e-7)
    # trainable parameters receive gradients; `frozen` hides them from the optimiser (norms.py:122-128)
    trainable = [n for n, _, t in cls._params if t]
    assert len(list(norm.parameters())) == len(trainable)
    if trainable and cls is not N.FixedMaxImageNorm:
        assert all(getattr(norm, n).grad is not None for n in trainable)
    assert list(cls(**params, frozen=True).parameters()) == []
    data = norm.to_dict()
    assert data["type"] == cls.registry_key and N.NORMS_REGISTRY[data["type"]] is cls
    again = N.ImageNorm.from_dict(data)
    assert type(again) is cls and again.to_dict() == pytest.approx(data)
    if cls not in (N.MaxImageNorm, N.ATanImageNorm):  # inverse undoes the norm (ATan: as the reference states it)
        inside = X if cls is not N.FixedMaxImageNorm else X.clamp(max=2.9)
        torch.testing.assert_close(norm.inverse(norm(inside)), inside, rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(norm.evaluate_numpy(X.numpy()), y.detach().numpy(), rtol=1e-6)


def test_constructor_errors_and_positional_arguments():
    assert float(N.ASinhImageNorm(0.3, 2.0).alpha) == pytest.approx(0.3)
    with pytest.raises(TypeError):
        N.FixedMaxImageNorm()
    assert "ASinhImageNorm" in str(N.ASinhImageNorm()) and N.PowerImageNorm().beta.requires_grad is False


def test_inverse_cdf_norm():
    """Histogram equalisation (norms.py:340-369 with interp1d_torch, utils/torch.py:146-169); bit-identical to the
    imported reference when checked in the build container."""
    img = np.random.default_rng(2).gamma(2.0, size=(40, 40)).astype(np.float32)
    norm = N.InverseCDFImageNorm.from_image(img, bins=50)
    assert norm.x.shape == norm.cdf.shape == (50,) and float(norm.cdf[0]) == 0.0 and float(norm.cdf[-1]) == 1.0
    x = torch.tensor(img[None, None, :7, :7])
    y = norm(x)
    ref = np.interp(x.numpy(), norm.x.numpy(), norm.cdf.numpy())
    inside = (x.numpy() >= float(norm.x[0])) & (x.numpy() <= float(norm.x[-2]))
    np.testing.assert_allclose(y.numpy()[inside], ref[inside], rtol=1e-5, atol=1e-6)
    assert N.NORMS_REGISTRY["inverse-cdf"] is N.InverseCDFImageNorm and list(norm.parameters()) == []
    with pytest.raises(ValueError):
        N.InverseCDFImageNorm(torch.zeros(3), torch.zeros(4))
