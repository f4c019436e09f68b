"""Autograd bindings of the CUDA kernels: lets the reference's own training loop
(jolideco/core.py:214-229: forward, `loss.backward()`, `optimizer.step()`) run unmodified on top of
the C-ABI kernels.  The fused engine (`engine.py`) calls the same kernels without autograd."""
import torch

from . import ops


def _img(t):
    """(1,1,H,W) or (H,W) tensor -> contiguous (H,W) view."""
    if t.ndim == 4:
        if t.shape[0] != 1 or t.shape[1] != 1:
            raise ValueError(f"only single-channel (1,1,H,W) images are supported, got {tuple(t.shape)}")
        t = t[0, 0]
    return t.contiguous()


class _NPredFunction(torch.autograd.Function):
    """npred = clip(sumpool_f(psf (*) (flux * E)), 0, inf)   (models/npred.py:160-191)."""

    @staticmethod
    def forward(ctx, flux, exposure, psf, f):
        shape = flux.shape
        fl, ex, ps = _img(flux.detach()), _img(exposure), _img(psf)
        conv = ops.conv_forward(fl, ex, ps)
        fH, fW = fl.shape
        H, W = fH // f, fW // f
        pool = ops.pool_clip(conv, H, W, f)
        ctx.save_for_backward(ex, ps, pool)
        ctx.f = f
        ctx.in_shape = shape
        out = torch.clamp_min(pool, 0)
        return out.reshape(shape[:-2] + (H, W))

    @staticmethod
    def backward(ctx, grad):
        ex, ps, pool = ctx.saved_tensors
        dpool = (_img(grad) * (pool >= 0)).contiguous()
        dflux = ops.conv_backward(dpool, ex, ps, ctx.f)
        return dflux.reshape(ctx.in_shape), None, None, None


def npred_forward(flux, exposure, psf, f=1):
    return _NPredFunction.apply(flux, exposure, psf, int(f) if f else 1)


class _PoissonNLLFunction(torch.autograd.Function):
    """nn.PoissonNLLLoss(log_input=False, reduction="mean", eps=1e-25, full=True) (loss.py:35-37)."""

    @staticmethod
    def forward(ctx, npred, counts):
        n, c = _img(npred.detach()), _img(counts)
        zeros = torch.zeros_like(n)
        # conv := npred, background := 0, f = 1: pool == npred (clip is the identity for npred >= 0)
        res = ops.poisson_forward_backward(n, zeros, c, f=1, want_grad=True)
        ctx.save_for_backward(res["dpool"])
        ctx.in_shape = npred.shape
        return (res["loss_sum"] / n.numel()).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad):
        (dn,) = ctx.saved_tensors
        return (dn * grad).reshape(ctx.in_shape), None


def poisson_nll(npred, counts):
    return _PoissonNLLFunction.apply(npred, counts)


class _GMMPatchPriorFunction(torch.autograd.Function):
    """sum_p max_k / logsumexp_k log N_k(patch_p) * stride^2/64 / numel  (priors/patches/core.py:227-246)."""

    @staticmethod
    def forward(ctx, flux, packed, shift_yx, stride, marginalize, backend):
        fl = _img(flux.detach())
        fH, fW = fl.shape
        shift = ops.as_shift_tensor(shift_yx, fl.device)
        value, argmax, logp, s = ops.gmm_prior_forward(fl, shift, packed, stride, marginalize, backend=backend)
        c = stride**2 / ops.PD / (fH * fW)
        ctx.save_for_backward(fl, shift, value, argmax, logp if logp is not None else torch.empty(0, device=fl.device))
        ctx.packed, ctx.stride, ctx.marginalize, ctx.c, ctx.in_shape = packed, stride, marginalize, c, flux.shape
        return (s * c).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad):
        fl, shift, value, argmax, logp = ctx.saved_tensors
        fH, fW = fl.shape
        G = ops.gmm_prior_backward(fl, shift, ctx.packed, -ctx.c, ctx.stride, ctx.marginalize, None, argmax,
                                   logp if ctx.marginalize else None, value)
        dflux = ops.patch_fold(G, fH, fW, shift, ctx.stride)
        return (dflux * grad).reshape(ctx.in_shape), None, None, None, None, None


def gmm_patch_prior(flux, packed, shift_yx, stride=4, marginalize=False, backend=0):
    return _GMMPatchPriorFunction.apply(flux, packed, shift_yx, int(stride), bool(marginalize), int(backend))


class _FluxFunction(torch.autograd.Function):
    """flux = exp(theta) [* mask]   (models/core.py:583-594); d flux / d theta = flux."""

    @staticmethod
    def forward(ctx, theta, mask, use_log_flux):
        th = theta.detach().contiguous()
        flux = ops.flux_forward(th.reshape(-1), None if mask is None else mask.reshape(-1), use_log_flux).reshape(
            theta.shape)
        ctx.use_log_flux = use_log_flux
        ctx.save_for_backward(flux if use_log_flux else (mask if mask is not None else torch.empty(0)))
        return flux

    @staticmethod
    def backward(ctx, grad):
        (saved,) = ctx.saved_tensors
        if ctx.use_log_flux:
            return grad * saved, None, None
        return (grad * saved if saved.numel() else grad), None, None


def flux_from_theta(theta, mask=None, use_log_flux=True):
    return _FluxFunction.apply(theta, mask, bool(use_log_flux))
