"""TEST INFRASTRUCTURE — torch-CPU port of the reference's MAP loop, op for op.

Where `jolideco_oracle.py` restates the algorithm in closed form (numpy, explicit adjoints), this
module restates the reference's *implementation*: the same ATen op stream on the host (rfft2 x2 +
irfft2 per convolution with the kernel FFT recomputed every call, `avg_pool2d(divisor_override=1)`,
`nn.PoissonNLLLoss`, `roll` + `unfold` x2 + `reshape` + boolean-mask gather + `nanmean`, the Python
loop over mixture components with one `matmul` each, autograd for the backward and
`torch.optim.Adam`).  It is what `bench.py` times as the CPU baseline (`cpu_baseline.kind = "port"`,
and the `--impl reference` arm) because the reference package itself cannot travel to the GPU box.
It is validated against the imported reference's golden runs in `tests/test_oracle_golden.py`.
Only tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline legs may import it.

File:line citations are relative to the reference root.
"""
import numpy as np
import torch
import torch.nn.functional as F


def convolve_fft_torch(image, kernel):
    """utils/torch.py:347-370 (+ _centered :337-344)."""
    s = [image.shape[-2] + kernel.shape[-2] - 1, image.shape[-1] + kernel.shape[-1] - 1]
    image_ft = torch.fft.rfft2(image, s=s)
    kernel_ft = torch.fft.rfft2(kernel, s=s)
    result = torch.fft.irfft2(image_ft * kernel_ft, s=s)
    y0, x0 = (s[0] - image.shape[-2]) // 2, (s[1] - image.shape[-1]) // 2
    return result[..., y0 : y0 + image.shape[-2], x0 : x0 + image.shape[-1]]


class Dataset:
    """NPredModel.from_numpy + NPredModels buffers (models/npred.py:66-115, 263-295)."""

    def __init__(self, dataset, f):
        dims = (np.newaxis, np.newaxis)
        exposure = torch.from_numpy(np.asarray(dataset["exposure"], dtype=np.float32)[dims])
        psf = torch.from_numpy(np.asarray(dataset["psf"], dtype=np.float32)[dims])
        if f:
            exposure = F.interpolate(exposure, scale_factor=f, mode="bilinear")
            psf = F.interpolate(psf, scale_factor=f, mode="bilinear") / f**2
        exposure = exposure / convolve_fft_torch(torch.ones_like(exposure), psf)
        self.exposure, self.psf, self.f = exposure, psf, f
        self.background = torch.from_numpy(np.asarray(dataset["background"], dtype=np.float32)[dims])
        self.counts = torch.from_numpy(np.asarray(dataset["counts"], dtype=np.float32)[dims])

    def npred(self, flux):
        """NPredModel.forward + NPredModels.evaluate (npred.py:160-191, 241-261)."""
        npred = convolve_fft_torch(flux * self.exposure, self.psf)
        if self.f:
            npred = F.avg_pool2d(npred, kernel_size=self.f, divisor_override=1)
        npred = torch.clip(npred, 0, torch.inf)
        total = torch.zeros(self.background.shape)
        total = total + npred
        return total + self.background


class GMM:
    """GaussianMixtureModel buffers + estimate_log_prob (priors/patches/gmm.py:119-149, 217-299)."""

    def __init__(self, means, covariances, weights, pixel_weights):
        from .jolideco_oracle import compute_precision_cholesky

        prec = compute_precision_cholesky(np.asarray(covariances))
        self.means = torch.from_numpy(np.asarray(means).astype(np.float32))
        self.weights = torch.from_numpy(np.asarray(weights).astype(np.float32))
        self.precisions_cholesky = torch.from_numpy(prec.astype(np.float32))
        self.K, self.D = self.means.shape
        self.means_precisions_cholesky = torch.stack(
            [torch.matmul(mu, pc) for mu, pc in zip(self.means, self.precisions_cholesky)])
        self.log_det_cholesky = torch.sum(
            torch.log(self.precisions_cholesky.reshape(self.K, -1)[:, :: self.D + 1]), axis=1)
        self.log_weights = torch.log(self.weights)
        self.pixel_weights = torch.from_numpy(np.asarray(pixel_weights, dtype=np.float32).reshape(1, -1))

    def estimate_log_prob(self, x):
        n_samples, n_features = x.shape
        log_prob = torch.empty((n_samples, self.K))
        for k, (mu_prec, prec_chol) in enumerate(zip(self.means_precisions_cholesky, self.precisions_cholesky)):
            y = torch.matmul(x, prec_chol) - mu_prec
            log_prob[:, k] = torch.sum(torch.square(y) * self.pixel_weights, axis=1)
        two_pi = torch.tensor(2 * np.pi)
        return -0.5 * (n_features * torch.log(two_pi) + log_prob) + self.log_det_cholesky + self.log_weights


def gmm_patch_prior(flux, gmm, shifts, stride=4, marginalize=False, size=8):
    """GMMPatchPrior.__call__ (priors/patches/core.py:189-246), shifts injected."""
    dims = (flux.ndim - 2, flux.ndim - 1)
    rolled = torch.roll(flux, shifts=(int(shifts[0]), int(shifts[1])), dims=dims)
    windows = rolled.unfold(flux.ndim - 2, size, stride).unfold(flux.ndim - 1, size, stride)
    patches = torch.reshape(windows, (-1, size * size))
    selection = torch.all(patches > -1e5, dim=1, keepdims=False)
    patches = patches[selection, :]
    patches = patches - torch.nanmean(patches, dim=1, keepdims=True)
    loglike = gmm.estimate_log_prob(patches)
    values = torch.logsumexp(loglike, dim=1) if marginalize else torch.max(loglike, dim=1).values
    return torch.sum(values) * (stride**2 / (size * size)) / flux.numel()


class MapLoop:
    """MAPDeconvolver.run inner loop (core.py:197-230) + trace (loss.py:212-250)."""

    def __init__(self, flux_init_up, datasets, gmm=None, beta=1.0, lr=0.1, stride=4, marginalize=False):
        flux = torch.from_numpy(np.asarray(flux_init_up, dtype=np.float32)[np.newaxis, np.newaxis])
        self.theta = torch.nn.Parameter(torch.log(flux))
        self.datasets, self.gmm, self.beta = datasets, gmm, beta
        self.stride, self.marginalize = stride, marginalize
        self.optimizer = torch.optim.Adam(params=[self.theta], lr=lr)
        self.loss_function = torch.nn.PoissonNLLLoss(log_input=False, reduction="mean", eps=1e-25, full=True)
        self.fluxes = None

    def step(self, i, shifts=None):
        ds = self.datasets[i]
        self.optimizer.zero_grad()
        flux = torch.exp(self.theta)
        self.fluxes = flux
        loss = self.loss_function(ds.npred(flux), ds.counts)
        if self.gmm is not None:
            loss_prior = gmm_patch_prior(flux, self.gmm, shifts, self.stride, self.marginalize)
        else:
            loss_prior = torch.tensor(0)
        loss_total = loss - self.beta * loss_prior / len(self.datasets)
        loss_total.backward()
        self.optimizer.step()
        return loss_total

    def joint_step(self, shifts=None):
        """One Adam step on TotalLoss.__call__ (loss.py:257-261)."""
        self.optimizer.zero_grad()
        flux = torch.exp(self.theta)
        self.fluxes = flux
        total = sum(self.loss_function(ds.npred(flux), ds.counts) for ds in self.datasets)
        if self.gmm is not None:
            total = total - self.beta * gmm_patch_prior(flux, self.gmm, shifts, self.stride, self.marginalize)
        total.backward()
        self.optimizer.step()
        return total

    @torch.no_grad()
    def trace(self, shifts=None):
        flux = self.fluxes  # stale tuple quirk: core.py:217 / :245
        ld = [self.loss_function(ds.npred(flux), ds.counts).item() for ds in self.datasets]
        lp = 0.0
        if self.gmm is not None:
            lp = gmm_patch_prior(flux, self.gmm, shifts, self.stride, self.marginalize).item()
        return {"total": sum(ld) - self.beta * lp, "datasets": ld, "priors-total": -self.beta * lp}

    def flux_numpy(self):
        return torch.exp(self.theta).detach().numpy()[0, 0]
