"""Batched likelihood kernels (jd_likelihood_forward / _backward: PSF convolution with the Poisson statistic fused into
its epilogue, all datasets of a joint iteration per launch) against the oracle, and the joint-step gradient assembly
kernels (jd_adam_joint_step_dev, jd_grad_reduce_local) against their single-purpose counterparts."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import jolideco_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from jolideco_b200 import _lib, ops

DEV = "cuda"


def t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(DEV)


def rel_max(a, ref):
    return np.abs(a - ref).max() / np.abs(ref).max()


def make_datasets(rng, D, H, W, kh, kw, f, zero_rows=False):
    out = []
    for i in range(D):
        psf = rng.uniform(size=(kh, kw)) * np.outer(np.hanning(kh + 2)[1:-1], np.hanning(kw + 2)[1:-1])
        psf = (psf / psf.sum()).astype(np.float32)
        E = rng.uniform(0.5, 1.5, size=(H * f, W * f)).astype(np.float32)
        if zero_rows:  # npred == 0 exactly where exposure and background vanish (log(1e-25) branch)
            E[: 2 * f + kh] = 0
        bkg = rng.uniform(0.1, 1.0, size=(H, W)).astype(np.float32)
        if zero_rows:
            bkg[:2] = 0
        counts = rng.poisson(3.0, size=(H, W)).astype(np.float32)
        out.append(dict(psf=psf, exposure=E, background=bkg, counts=counts))
    return out


CASES = [  # (D, H, W, kh, kw, f)
    (3, 64, 64, 17, 17, 1),     # the north-star tap layout (lead 0, 5 groups, 1 tail tap)
    (2, 37, 45, 5, 4, 1),       # ragged image (W % 4 != 0: scalar staging / epilogue), even PSF
    (2, 130, 70, 9, 29, 1),     # partial tiles, the widest tap row (lead 2 + 29 = 31)
    (1, 72, 100, 6, 6, 1),      # even PSF: asymmetric crop, forward and adjoint lead differ
    (2, 40, 36, 7, 7, 2),       # upsampling 2: sum-pool in registers, replicated adjoint input
    (1, 21, 17, 6, 6, 2),       # upsampling 2, ragged, even PSF
    (2, 96, 128, 1, 1, 1),      # 1 x 1 PSF
    (1, 128, 192, 23, 13, 1),   # non-square PSF
    (1, 96, 80, 34, 34, 2),     # BASELINE configs[1]: 17 x 17 PSF upsampled by 2 (9 / 10 tap groups)
    (2, 100, 120, 7, 37, 1),    # widest f = 1 row (lead 2 + 37 taps -> 10 groups, padded)
]


FFT_CASES = [  # (D, H, W, kh, kw, f): the shared-memory FFT path, every dataset per launch, Poisson fused
    (2, 96, 80, 41, 41, 1),     # radix-5 plans (136 -> 160, 120 -> 128)
    (3, 64, 96, 34, 34, 2),     # upsampling 2: a row pair = a pooling pair; even PSF (asymmetric crop)
    (1, 130, 70, 64, 64, 1),    # BASELINE configs[2] PSF, odd image sizes (last row pair has one row)
    (2, 100, 52, 9, 31, 1),     # non-square PSF
    (2, 33, 47, 20, 20, 2),     # upsampling 2, ragged, even PSF
]


@pytest.mark.parametrize("D,H,W,kh,kw,f,fft", [c + (False,) for c in CASES] + [c + (True,) for c in FFT_CASES])
@pytest.mark.parametrize("with_norm", [False, True])
def test_batched_likelihood_matches_oracle(D, H, W, kh, kw, f, fft, with_norm):
    assert fft or _lib.load().jd_likelihood_supported(kh, kw, f) == 1
    rng = np.random.default_rng(100 * kh + kw + f)
    flux = (rng.gamma(2.0, size=(H * f, W * f)) * np.exp(rng.normal(0, 0.5, size=(H * f, W * f)))).astype(np.float32)
    ds = make_datasets(rng, D, H, W, kh, kw, f)
    logb = [np.float32(rng.normal(0, 0.3)) for _ in ds]
    dev_ds = []
    for d, lb in zip(ds, logb):
        dd = {k: t(v) for k, v in d.items()}
        if with_norm:
            dd["bkg_log_norm"] = t(np.array([lb], dtype=np.float32))
        dev_ds.append(dd)
    res = ops.likelihood_batched(t(flux), dev_ds, f, fft=fft)
    loss = res["loss_sum"].cpu().numpy() / (H * W)
    for i, (d, lb) in enumerate(zip(ds, logb)):
        f64 = {k: v.astype(np.float64) for k, v in d.items()}
        bnorm = np.exp(np.float64(lb)) if with_norm else None
        npred, pool = O.npred_forward(flux.astype(np.float64), f64["exposure"], f64["psf"], f64["background"], f, bnorm,
                                      return_pool=True)
        tol = 4.0 if fft else 1.0  # float32 FFTs of a few hundred points against the float64 oracle
        assert_allclose(loss[i], O.poisson_nll(npred, f64["counts"]), rtol=2e-6 * tol)
        dn = O.poisson_nll_grad(npred, f64["counts"])
        dpool_ref = dn * (pool >= 0)
        got = res["dpool"][i].cpu().numpy()
        # pixels whose pre-clip pool is within rounding of 0 may fall on either side of the clip
        sure = np.abs(pool) > 1e-6 * tol * np.abs(pool).max()
        assert np.abs(got - dpool_ref)[sure].max() <= 5e-6 * tol * np.abs(dpool_ref).max()
        assert_allclose(res["dlogb"][i].item(), (dn * f64["background"] * (bnorm or 1.0)).sum(), rtol=2e-5, atol=1e-9)
        # adjoint of exactly the dpool the forward produced
        ref_b = O.npred_backward(got.astype(np.float64), np.ones_like(pool), flux.astype(np.float64), f64["exposure"],
                                 f64["psf"], f)
        assert rel_max(res["dflux"][i].cpu().numpy(), ref_b) < 5e-6 * tol


def test_npred_exactly_zero_keeps_the_eps_literal():
    """npred == 0 after the clip: loss term -c log(1e-25), gradient (1 - c 1e25) / (H W), finite in float32."""
    rng = np.random.default_rng(3)
    H, W = 48, 64
    flux = rng.gamma(2.0, size=(H, W)).astype(np.float32)
    (d,) = make_datasets(rng, 1, H, W, 5, 5, 1, zero_rows=True)
    res = ops.likelihood_batched(t(flux), [{k: t(v) for k, v in d.items()}], 1)
    f64 = {k: v.astype(np.float64) for k, v in d.items()}
    pool = O.sum_pool(O.convolve_direct(flux.astype(np.float64) * f64["exposure"], f64["psf"]), 1)
    assert (pool[:2] == 0).all()                       # direct form: exact zeros (the FFT oracle leaves 1e-16 noise)
    npred = np.clip(pool, 0, np.inf) + f64["background"]
    assert (npred[:2] == 0).all()
    got = res["dpool"][0].cpu().numpy()
    assert np.isfinite(got).all()
    ref = O.poisson_nll_grad(npred.astype(np.float32), d["counts"]) * (pool >= 0)
    assert_allclose(got[:2], ref[:2], rtol=1e-5)       # (1 - c 1e25) / (H W): finite in float32
    assert_allclose(got[8:], ref[8:], rtol=1e-4, atol=1e-9)
    assert_allclose(res["loss_sum"].item() / (H * W), O.poisson_nll(npred, f64["counts"]), rtol=2e-6)


def test_batched_likelihood_equals_the_separate_kernels():
    """Same numbers (to float rounding) as conv -> Poisson -> conv adjoint, and loss-only tables leave dpool alone."""
    rng = np.random.default_rng(5)
    H = W = 96
    flux = rng.gamma(2.0, size=(H, W)).astype(np.float32)
    ds = make_datasets(rng, 2, H, W, 11, 11, 1)
    dev_ds = [{k: t(v) for k, v in d.items()} for d in ds]
    res = ops.likelihood_batched(t(flux), dev_ds, 1)
    only = ops.likelihood_batched(t(flux), dev_ds, 1, want_grad=False)
    assert only["dpool"] is None
    assert_allclose(only["loss_sum"].cpu().numpy(), res["loss_sum"].cpu().numpy(), rtol=1e-12)
    for i, d in enumerate(dev_ds):
        conv = ops.conv_forward(t(flux), d["exposure"], d["psf"])
        sep = ops.poisson_forward_backward(conv, d["background"], d["counts"], 1)
        assert_allclose(res["loss_sum"][i].item(), sep["loss_sum"].item(), rtol=2e-6)
        assert rel_max(res["dpool"][i].cpu().numpy(), sep["dpool"].cpu().numpy()) < 5e-6
        back = ops.conv_backward(res["dpool"][i].contiguous(), d["exposure"], d["psf"], 1)
        assert rel_max(res["dflux"][i].cpu().numpy(), back.cpu().numpy()) < 5e-6


def test_unsupported_geometries_are_refused():
    lib = _lib.load()
    assert lib.jd_likelihood_supported(34, 34, 2) == 1 and lib.jd_likelihood_supported(17, 17, 3) == 0
    assert lib.jd_likelihood_supported(64, 64, 1) == 0 and lib.jd_likelihood_supported(29, 29, 1) == 1
    with pytest.raises(_lib.JolidecoB200Error):
        ops.likelihood_batched(torch.ones(64, 64, device=DEV), [dict(
            exposure=torch.ones(64, 64, device=DEV), psf=torch.ones(3, 44, device=DEV),
            background=torch.ones(64, 64, device=DEV), counts=torch.ones(64, 64, device=DEV))])


@pytest.mark.parametrize("with_prior", [True, False])
@pytest.mark.parametrize("fH,fW,stride,shift_yx,rows", [
    (72, 88, 4, (1, -2), None),        # per-pixel gather (four pixels per thread, stride-4 shifts)
    (40, 1024, 4, (1, -2), None),      # one image row per CTA pass: the rolled row folded through shared memory
    (40, 1024, 4, (-2, 3), (2, 7)),    # ... on a patch-row block (multi-GPU shard), another shift
    (40, 1024, 4, (0, 0), None),
    (48, 64, 2, (2, 1), None),         # other strides: the general gather
])
def test_joint_update_kernels_match_fold_plus_adam(with_prior, fH, fW, stride, shift_yx, rows):
    """jd_adam_joint_step_dev == sum of parts + jd_patch_fold + jd_adam_step_dev; jd_grad_reduce_local == the gradient."""
    rng = np.random.default_rng(9)
    D = 3
    n = fH * fW
    ny, nx = ops.patch_grid(fH, fW, stride)
    r0, r1 = (0, ny) if rows is None else rows
    theta = t(rng.normal(size=(fH, fW)))
    flux = ops.flux_forward(theta)
    parts = t(rng.normal(size=(D, fH, fW)))
    G = t(rng.normal(size=((r1 - r0) * nx, 64))) if with_prior else None
    shift = torch.tensor(list(shift_yx), dtype=torch.int32, device=DEV)
    scalars = t(np.array([0.1 / (1 - 0.9), np.sqrt(1 - 0.999)], dtype=np.float32))
    s = torch.cuda.current_stream().cuda_stream
    p = lambda x: None if x is None else x.data_ptr()  # noqa: E731
    # reference: explicit sum, fold, Adam
    g = parts.sum(dim=0)
    if with_prior:
        g = g + 0.7 * ops.patch_fold(G, fH, fW, shift, stride, rows=(r0, r1))
    th_ref, m_ref, v_ref = theta.clone(), torch.zeros_like(theta), torch.zeros_like(theta)
    _lib.call("jd_adam_step_dev", p(th_ref), p(m_ref), p(v_ref), p(flux), None, p(g.contiguous()), None, 0.0, 1, n,
              p(scalars), 0.9, 0.999, 1e-8, s)
    th, m, v = theta.clone(), torch.zeros_like(theta), torch.zeros_like(theta)
    _lib.call("jd_adam_joint_step_dev", p(th), p(m), p(v), p(flux), None, p(parts), D, n, p(G), 0.7, 1, fH, fW, p(shift),
              stride, r0, r1, p(scalars), 0.9, 0.999, 1e-8, s)
    assert_allclose(th.cpu().numpy(), th_ref.cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert_allclose(m.cpu().numpy(), m_ref.cpu().numpy(), rtol=1e-5, atol=1e-7)
    out = torch.empty_like(theta)
    _lib.call("jd_grad_reduce_local", p(parts), D, n, p(G), 0.7, fH, fW, p(shift), stride, r0, r1, p(out), s)
    assert_allclose(out.cpu().numpy(), g.cpu().numpy(), rtol=1e-5, atol=1e-6)
    if with_prior:  # the fold itself: same summation order in every variant of the gather -> bit-identical
        only = torch.empty_like(theta)
        _lib.call("jd_grad_reduce_local", p(torch.zeros_like(parts)), 1, n, p(G), 1.0, fH, fW, p(shift), stride, r0, r1,
                  p(only), s)
        assert torch.equal(only, ops.patch_fold(G, fH, fW, shift, stride, rows=(r0, r1)))
    # in place on parts[0] (the NCCL path reduces into the first part)
    parts2 = parts.clone()
    _lib.call("jd_grad_reduce_local", p(parts2), D, n, p(G), 0.7, fH, fW, p(shift), stride, r0, r1, p(parts2), s)
    assert torch.equal(parts2[0], out)


def test_fp32_probe_reports_flops():
    out = torch.zeros(1, device=DEV)
    flops = _lib.load().jd_probe_fp32_fma(64, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert flops > 0 and float(out) == 0.0
