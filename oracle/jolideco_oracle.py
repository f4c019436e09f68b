"""TEST INFRASTRUCTURE — CPU oracle for the Jolideco MAP-deconvolution hot path.

A plain numpy restatement of the reference algorithm (jolideco/jolideco, file:line citations are
relative to the reference root).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module; the product (`jolideco_b200/`) must
never route through it.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against
  * the reference's own known-answer tests (patch order `utils/tests/test_torch.py:8-21`, FFT conv
    `:24-60`, NPred values `models/tests/test_core.py:63-75`, GMM log-prob vs sklearn
    `priors/patches/tests/test_gmm.py:10-35`),
  * the reference's e2e golden values (`jolideco/tests/test_core.py:71-79`, `:99-124`), and
  * outputs of the *imported* reference (`oracle/make_golden.py` -> `tests/golden/*.npz`),
    including losses, autograd gradients and N-step Adam trajectories with a GMM patch prior.

All functions take/return 2-D numpy arrays (the reference's (1,1,H,W) tensors with the two
singleton axes dropped) and compute in the dtype they are given (float64 for tight checks,
float32 to mimic the reference's working precision).
"""
import math

import numpy as np

EPS_POISSON = 1e-25  # loss.py:35-37


# --------------------------------------------------------------------------------------
# a1  flux parameterisation                                   models/core.py:583-594
# --------------------------------------------------------------------------------------
def flux_from_theta(theta, mask=None, use_log_flux=True):
    flux = np.exp(theta) if use_log_flux else theta
    if mask is not None:
        flux = flux * mask
    return flux


# --------------------------------------------------------------------------------------
# a2  setup of the NPred model                                 models/npred.py:66-115
# --------------------------------------------------------------------------------------
def interpolate_bilinear(a, f):
    """F.interpolate(scale_factor=f, mode='bilinear', align_corners=False) for a 2-D array."""
    a = np.asarray(a)
    if f == 1:
        return a.copy()
    H, W = a.shape

    def axis(n):
        dst = np.arange(n * f)
        src = np.maximum((dst + 0.5) / f - 0.5, 0.0)
        i0 = np.minimum(np.floor(src).astype(int), n - 1)
        i1 = np.minimum(i0 + 1, n - 1)
        l1 = (src - i0).astype(a.dtype)
        return i0, i1, 1 - l1, l1

    y0, y1, wy0, wy1 = axis(H)
    x0, x1, wx0, wx1 = axis(W)
    rows = a[y0] * wy0[:, None] + a[y1] * wy1[:, None]
    return rows[:, x0] * wx0[None, :] + rows[:, x1] * wx1[None, :]


def npred_setup(exposure, psf, f, correct_exposure_edges=True):
    """exposure / psf bilinear-upsampled by f, psf /= f^2, exposure /= (psf (*) 1).

    npred.py:96-113.  Returns (exposure_up, psf_up).
    """
    exposure_up = interpolate_bilinear(exposure, f)
    psf_up = interpolate_bilinear(psf, f)
    if f:
        psf_up = psf_up / f**2
    if correct_exposure_edges:
        weights = convolve_fft(np.ones_like(exposure_up), psf_up)
        exposure_up = exposure_up / weights
    return exposure_up, psf_up


# --------------------------------------------------------------------------------------
# a3  convolution + forward model      utils/torch.py:337-370, models/npred.py:160-191
# --------------------------------------------------------------------------------------
def convolve_fft(image, kernel):
    """rfft2(image, s) * rfft2(kernel, s) -> irfft2 -> centred crop (start (k-1)//2)."""
    s = [image.shape[i] + kernel.shape[i] - 1 for i in range(2)]
    res = np.fft.irfft2(np.fft.rfft2(image, s=s) * np.fft.rfft2(kernel, s=s), s=s)
    start = [(s[i] - image.shape[i]) // 2 for i in range(2)]  # _centered: trunc((curr-new)/2)
    out = res[start[0] : start[0] + image.shape[0], start[1] : start[1] + image.shape[1]]
    return out.astype(image.dtype, copy=False)


def convolve_direct(image, kernel):
    """Direct form of `convolve_fft`: c[i,j] = sum_ab k[a,b] g[i+s_y-a, j+s_x-b], s=(k-1)//2."""
    kh, kw = kernel.shape
    sy, sx = (kh - 1) // 2, (kw - 1) // 2
    H, W = image.shape
    pad = np.zeros((H + kh - 1, W + kw - 1), dtype=image.dtype)
    pad[kh - 1 - sy : kh - 1 - sy + H, kw - 1 - sx : kw - 1 - sx + W] = image
    out = np.zeros_like(image)
    for a in range(kh):
        for b in range(kw):
            out += kernel[a, b] * pad[kh - 1 - a : kh - 1 - a + H, kw - 1 - b : kw - 1 - b + W]
    return out


def correlate_adjoint(dc, kernel):
    """Adjoint of `convolve_fft` w.r.t. the image: dg[m,n] = sum_ab k[a,b] dc[m-s_y+a, n-s_x+b]."""
    kh, kw = kernel.shape
    sy, sx = (kh - 1) // 2, (kw - 1) // 2
    H, W = dc.shape
    from scipy.signal import fftconvolve

    pad = np.zeros((H + kh - 1, W + kw - 1), dtype=dc.dtype)
    pad[sy : sy + H, sx : sx + W] = dc  # pad-left s, pad-right k-1-s
    # dg[m,n] = sum_ab k[a,b] pad[m+a, n+b]: 'valid' convolution with the flipped kernel
    out = fftconvolve(pad, kernel[::-1, ::-1], mode="valid")
    return out.astype(dc.dtype, copy=False)


def sum_pool(c, f):
    """F.avg_pool2d(kernel_size=f, divisor_override=1): npred.py:181-184."""
    if f == 1:
        return c
    H, W = c.shape[0] // f, c.shape[1] // f
    return c[: H * f, : W * f].reshape(H, f, W, f).sum(axis=(1, 3))


def npred_forward(flux, exposure_up, psf_up, background, f=1, background_norm=None, return_pool=False):
    """npred = clip(sumpool_f(psf (*) (flux E)), 0, inf) + B [* exp(log b)]   npred.py:160-191, 210-261."""
    g = flux * exposure_up
    c = convolve_fft(g, psf_up) if psf_up is not None else g
    pool = sum_pool(c, f)
    bkg = background if background_norm is None else background * background_norm
    npred = np.clip(pool, 0, np.inf) + bkg
    return (npred, pool) if return_pool else npred


# --------------------------------------------------------------------------------------
# a6  Poisson cash statistic                                         loss.py:35-37
# --------------------------------------------------------------------------------------
def poisson_nll(npred, counts, eps=EPS_POISSON):
    """nn.PoissonNLLLoss(log_input=False, reduction='mean', eps=1e-25, full=True)."""
    dt = npred.dtype
    loss = npred - counts * np.log(npred + dt.type(eps))
    big = counts > 1
    c = np.where(big, counts, dt.type(2.0))
    stirling = c * np.log(c) - c + dt.type(0.5) * np.log(dt.type(2 * math.pi) * c)
    loss = loss + np.where(big, stirling, dt.type(0))
    return loss.mean(dtype=np.float64).astype(dt) if dt == np.float32 else loss.mean()


def poisson_nll_grad(npred, counts, eps=EPS_POISSON):
    """d mean-loss / d npred = (1 - c/(n+eps)) / (H W)."""
    dt = npred.dtype
    return (1 - counts / (npred + dt.type(eps))) / dt.type(npred.size)


def npred_backward(dn, pool, flux, exposure_up, psf_up, f=1):
    """Adjoint of `npred_forward` w.r.t. flux: clip mask, replicate, correlate, x E."""
    dpool = dn * (pool >= 0)
    dc = np.repeat(np.repeat(dpool, f, axis=0), f, axis=1) if f > 1 else dpool
    fH, fW = flux.shape
    if dc.shape != (fH, fW):  # trailing rows/cols dropped by the pooling get zero gradient
        full = np.zeros((fH, fW), dtype=dc.dtype)
        full[: dc.shape[0], : dc.shape[1]] = dc
        dc = full
    dg = correlate_adjoint(dc, psf_up) if psf_up is not None else dc
    return dg * exposure_up


# --------------------------------------------------------------------------------------
# a8  cycle spin + patch extraction            utils/torch.py:91-119, 226-275
# --------------------------------------------------------------------------------------
def cycle_spin_roll(image, shift_y, shift_x):
    """torch.roll(image, (shift_y, shift_x), dims=(H, W)); first randint draw is the row shift."""
    return np.roll(image, (shift_y, shift_x), axis=(0, 1))


def view_as_overlapping_patches(image, size, stride=None):
    """unfold rows then cols, flattened row-major to (P, size*size); p = iy*nx + ix."""
    if stride is None:
        stride = size // 2
    H, W = image.shape
    ny, nx = (H - size) // stride + 1, (W - size) // stride + 1
    sh, sw = image.strides
    win = np.lib.stride_tricks.as_strided(
        image, shape=(ny, nx, size, size), strides=(sh * stride, sw * stride, sh, sw), writeable=False
    )
    return win.reshape(ny * nx, size * size)


# --------------------------------------------------------------------------------------
# a9  GMM constants and log-prob      utils/numpy.py:16-79, priors/patches/gmm.py:217-299
# --------------------------------------------------------------------------------------
def compute_precision_cholesky(covariances):
    """(chol(Sigma_k)^-1)^T in float64 (scipy in the reference; numpy here)."""
    out = np.empty(covariances.shape, dtype=np.float64)
    eye = np.eye(covariances.shape[1])
    for k, cov in enumerate(np.asarray(covariances, dtype=np.float64)):
        chol = np.linalg.cholesky(cov)
        out[k] = np.linalg.solve(chol, eye).T  # lower-triangular solve; result upper-triangular
    return out


def _evaluate_trapez(x, width, slope):
    x2 = min(-width / 2.0, 0)
    x3 = max(width / 2.0, 0)
    x1 = x2 - 1.0 / slope
    x4 = x3 + 1.0 / slope
    ra = np.logical_and(x >= x1, x < x2)
    rb = np.logical_and(x >= x2, x < x3)
    rc = np.logical_and(x >= x3, x < x4)
    return np.select([ra, rb, rc], [slope * (x - x1), 1, slope * (x4 - x)])


def get_pixel_weights(patch_size, stride):
    """Trapezoid pixel weights, normalised to sum stride^2 (utils/numpy.py:54-79)."""
    width = patch_size
    overlap = width - stride
    value = (width - 1.0) / 2
    x = np.linspace(-value, value, width)
    values = _evaluate_trapez(x=x, width=(stride - overlap), slope=1.0 / overlap)
    weights = values * values[:, np.newaxis]
    return weights / weights.sum() * stride**2


class GMM:
    """Constants of `GaussianMixtureModel` (gmm.py:64-299), built as `from_numpy` does:
    precision Cholesky in float64, then everything cast to float32 buffers; derived constants
    (mu L, log-det, log-weights, pixel weights) are computed from those float32 buffers."""

    def __init__(self, means, covariances, weights, meta_stride=4, dtype=np.float32):
        prec = compute_precision_cholesky(covariances)
        self.means = np.asarray(means).astype(np.float32).astype(dtype)
        self.weights = np.asarray(weights).astype(np.float32).astype(dtype)
        self.precisions_cholesky = prec.astype(np.float32).astype(dtype)
        self.K, self.D = self.means.shape
        self.patch_size = int(round(math.sqrt(self.D)))
        self.means_precisions_cholesky = np.einsum("ki,kij->kj", self.means, self.precisions_cholesky)
        diag = self.precisions_cholesky.reshape(self.K, -1)[:, :: self.D + 1]
        self.log_det_cholesky = np.log(diag).sum(axis=1)
        self.log_weights = np.log(self.weights)
        if meta_stride is None:
            w = np.ones((self.patch_size, self.patch_size))
        else:
            w = get_pixel_weights(self.patch_size, meta_stride)
        self.pixel_weights = w.reshape(-1).astype(np.float32).astype(dtype)
        self.dtype = dtype

    def estimate_log_prob(self, x, return_y=False):
        """gmm.py:262-281: loop over components."""
        dt = self.dtype
        P = x.shape[0]
        # `log_prob = torch.empty(...)` is float32 whatever dtype x has (gmm.py:266): q is stored
        # rounded to float32 and -0.5 * (D log 2pi + q) is evaluated in float32 (gmm.py:276-281);
        # a no-op for the float32 working precision, restated so that float64 checks are exact.
        q = np.empty((P, self.K), dtype=np.float32)
        ys = [] if return_y else None
        for k in range(self.K):
            y = x @ self.precisions_cholesky[k] - self.means_precisions_cholesky[k]
            q[:, k] = np.sum(np.square(y) * self.pixel_weights, axis=1)
            if return_y:
                ys.append(y)
        log_two_pi = np.float32(1.8378770351409912)  # torch.log(torch.tensor(2 * np.pi)), float32
        half = np.float32(-0.5) * (np.float32(self.D) * log_two_pi + q)
        logp = half.astype(dt) + self.log_det_cholesky + self.log_weights
        return (logp, ys) if return_y else logp


def _logsumexp(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def gmm_patch_prior(flux, gmm, shift_y, shift_x, stride=4, marginalize=False, return_grad=False,
                    row_begin=None, row_end=None):
    """GMMPatchPrior.__call__ with identity image norm + subtract-mean patch norm.

    priors/patches/core.py:189-246.  `row_begin:row_end` restrict to a block of patch rows iy
    (used to check the row-sharded multi-GPU prior); the normalisation stays flux.size.
    Returns prior (scalar) [, d prior / d flux (same shape as flux), per-patch argmax].
    """
    dt = flux.dtype
    size = gmm.patch_size
    r = cycle_spin_roll(flux, shift_y, shift_x)
    X = view_as_overlapping_patches(r, size, stride)
    fH, fW = flux.shape
    ny, nx = (fH - size) // stride + 1, (fW - size) // stride + 1
    sel = np.arange(ny * nx)
    if row_begin is not None:
        sel = sel[row_begin * nx : row_end * nx]
        X = X[sel]
    keep = np.all(X > -1e5, axis=1)  # core.py:215-216
    X = X[keep]
    sel = sel[keep]
    Xc = X - X.mean(axis=1, keepdims=True)  # utils/norms.py:97-103 (nanmean == mean once filtered)
    logp, ys = gmm.estimate_log_prob(Xc, return_y=True)
    if marginalize:
        v = _logsumexp(logp, axis=1)
    else:
        v = logp.max(axis=1)
    c = dt.type(stride**2 / (size * size)) / dt.type(flux.size)  # core.py:222-246
    prior = v.sum() * c
    if not return_grad:
        return prior
    if marginalize:
        R = np.exp(logp - v[:, None])
    else:
        R = np.zeros_like(logp)
        R[np.arange(len(v)), logp.argmax(axis=1)] = 1
    # the gradient flows back through the float32 `log_prob` buffer (gmm.py:266): c*R is rounded
    # to float32 there (exact no-op at float32 working precision)
    coef = (R * c).astype(np.float32).astype(dt)
    G = np.zeros_like(Xc)
    for k in range(gmm.K):
        G -= (coef[:, k : k + 1] * (ys[k] * gmm.pixel_weights)) @ gmm.precisions_cholesky[k].T
    G -= G.mean(axis=1, keepdims=True)
    dr = np.zeros_like(flux)
    for n, p in enumerate(sel):
        iy, ix = divmod(int(p), nx)
        dr[stride * iy : stride * iy + size, stride * ix : stride * ix + size] += G[n].reshape(size, size)
    dflux = np.roll(dr, (-shift_y, -shift_x), axis=(0, 1))
    return prior, dflux, logp.argmax(axis=1)


def gmm_patch_prior_lean(flux, gmm, shift_y, shift_x, stride=4, marginalize=False, row_begin=None, row_end=None):
    """`gmm_patch_prior(..., return_grad=True)` for BASELINE-size inputs (P = 65 025, K = 256): the same arithmetic in
    two passes over the components instead of keeping every whitened residual y_k (P x K x 64 values).  Pass 1 = the
    log-probabilities (as `GMM.estimate_log_prob`, incl. its float32 `log_prob` buffer), pass 2 recomputes y_k only
    for components that carry gradient.  Checked against `gmm_patch_prior` in tests/test_oracle_golden.py.

    Returns dict(prior, dflux, value (P,), argmax (P,), gap (P,) = best minus second-best log-probability: patches
    with a tiny gap are the ones whose argmax may legitimately differ between float32 implementations)."""
    dt = flux.dtype
    size = gmm.patch_size
    r = cycle_spin_roll(flux, shift_y, shift_x)
    X = view_as_overlapping_patches(r, size, stride)
    fH, fW = flux.shape
    ny, nx = (fH - size) // stride + 1, (fW - size) // stride + 1
    sel = np.arange(ny * nx)
    if row_begin is not None:
        sel = sel[row_begin * nx : row_end * nx]
        X = X[sel]
    keep = np.all(X > -1e5, axis=1)
    X, sel = X[keep], sel[keep]
    Xc = X - X.mean(axis=1, keepdims=True)
    logp = gmm.estimate_log_prob(Xc)
    v = _logsumexp(logp, axis=1) if marginalize else logp.max(axis=1)
    k_star = logp.argmax(axis=1)
    top2 = np.partition(logp, -2, axis=1)[:, -2:] if gmm.K > 1 else np.stack([logp[:, 0] - np.inf, logp[:, 0]], axis=1)
    gap = top2[:, 1] - top2[:, 0]
    c = dt.type(stride**2 / (size * size)) / dt.type(flux.size)
    G = np.zeros_like(Xc)
    for k in range(gmm.K):
        if marginalize:
            rows = slice(None)
            coef = (np.exp(logp[:, k] - v) * c).astype(np.float32).astype(dt)
        else:
            rows = np.nonzero(k_star == k)[0]
            if rows.size == 0:
                continue
            coef = np.full(rows.size, np.float32(c), dtype=np.float32).astype(dt)
        y = Xc[rows] @ gmm.precisions_cholesky[k] - gmm.means_precisions_cholesky[k]
        G[rows] -= (coef[:, None] * (y * gmm.pixel_weights)) @ gmm.precisions_cholesky[k].T
    G -= G.mean(axis=1, keepdims=True)
    dr = np.zeros_like(flux)
    Gim = G.reshape(-1, size, size)
    iy, ix = np.divmod(sel, nx)
    for u in range(size):  # scatter-add, one patch element at a time (each (u, v) hits distinct pixels)
        for w_ in range(size):
            np.add.at(dr, (stride * iy + u, stride * ix + w_), Gim[:, u, w_])
    dflux = np.roll(dr, (-shift_y, -shift_x), axis=(0, 1))
    return dict(prior=v.sum() * c, dflux=dflux, value=v, argmax=k_star, gap=gap, patch_index=sel)


# --------------------------------------------------------------------------------------
# a12 Adam                                          torch.optim.Adam defaults, core.py:39-42
# --------------------------------------------------------------------------------------
class Adam:
    def __init__(self, shape, lr=0.1, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32):
        self.m = np.zeros(shape, dtype=dtype)
        self.v = np.zeros(shape, dtype=dtype)
        self.t = 0
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.dt = dtype

    def step(self, theta, grad):
        dt = self.dt
        self.t += 1
        self.m += (grad - self.m) * dt(1 - self.b1)  # exp_avg.lerp_(grad, 1-beta1)
        self.v *= dt(self.b2)
        self.v += dt(1 - self.b2) * grad * grad
        bc1 = 1 - self.b1**self.t
        bc2 = 1 - self.b2**self.t
        step_size = self.lr / bc1
        denom = np.sqrt(self.v) / dt(math.sqrt(bc2)) + dt(self.eps)
        return theta - dt(step_size) * (self.m / denom)


# --------------------------------------------------------------------------------------
# calibration: sub-pixel shift of the flux                        utils/torch.py:196-223, npred.py:226-230
# --------------------------------------------------------------------------------------
def _shift_taps(shift_y, shift_x, scale, dtype, shape):
    """`shift_image_torch` = affine_grid + grid_sample (bilinear, zeros padding, align_corners=False) with a pure
    translation: output pixel (i, j) samples the input at (i + scale * shift_y, j + scale * shift_x), i.e. a
    4-tap stencil with constant weights.  Returns (fy, fx, wy, wx, gy, gx): integer offsets, fractional weights and
    d(sample position)/d(shift).  The reference forms the normalised translation with `2 * scale / torch.tensor(
    [[W], [H]])`, a FLOAT32 tensor whatever the dtype of the image (utils/torch.py:216): that rounding of 2 scale / W
    is reproduced here (it moves the sample position by ~1e-7 pixel)."""
    H, W = shape
    # torch evaluates scalar / tensor as scalar * reciprocal(tensor), in float32
    gy = dtype.type(np.float32(2 * scale) * (np.float32(1) / np.float32(H))) * dtype.type(H) / dtype.type(2)
    gx = dtype.type(np.float32(2 * scale) * (np.float32(1) / np.float32(W))) * dtype.type(W) / dtype.type(2)
    dy, dx = gy * dtype.type(shift_y), gx * dtype.type(shift_x)
    fy, fx = int(np.floor(dy)), int(np.floor(dx))
    return fy, fx, dtype.type(dy - fy), dtype.type(dx - fx), gy, gx


def _shifted(image, oy, ox):
    """out[i, j] = image[i + oy, j + ox], zero outside."""
    H, W = image.shape
    out = np.zeros_like(image)
    i0, i1 = max(0, -oy), min(H, H - oy)
    j0, j1 = max(0, -ox), min(W, W - ox)
    if i1 > i0 and j1 > j0:
        out[i0:i1, j0:j1] = image[i0 + oy:i1 + oy, j0 + ox:j1 + ox]
    return out


def shift_image(image, shift_y, shift_x, scale=1, return_grads=False):
    """Shifted image (utils/torch.py:196-223); with return_grads also d out / d shift_y and d out / d shift_x
    (per pixel; the derivative of the bilinear interpolant, as grid_sample's backward gives it).
    The reference returns the input unchanged when both shifts are ~0 (utils/torch.py:211) - callers handle that."""
    fy, fx, wy, wx, gy, gx = _shift_taps(shift_y, shift_x, scale, image.dtype, image.shape)
    a00, a01 = _shifted(image, fy, fx), _shifted(image, fy, fx + 1)
    a10, a11 = _shifted(image, fy + 1, fx), _shifted(image, fy + 1, fx + 1)
    out = (1 - wy) * ((1 - wx) * a00 + wx * a01) + wy * ((1 - wx) * a10 + wx * a11)
    if not return_grads:
        return out
    d_dy = gy * ((1 - wx) * (a10 - a00) + wx * (a11 - a01))
    d_dx = gx * ((1 - wy) * (a01 - a00) + wy * (a11 - a10))
    return out, d_dy, d_dx


def shift_image_adjoint(d, shift_y, shift_x, scale=1):
    """Transpose of `shift_image` w.r.t. the image: dimage[m, n] = sum_ab w_ab d[m - fy - a, n - fx - b]."""
    fy, fx, wy, wx, _, _ = _shift_taps(shift_y, shift_x, scale, d.dtype, d.shape)
    return ((1 - wy) * ((1 - wx) * _shifted(d, -fy, -fx) + wx * _shifted(d, -fy, -fx - 1))
            + wy * ((1 - wx) * _shifted(d, -fy - 1, -fx) + wx * _shifted(d, -fy - 1, -fx - 1)))


def shift_is_identity(shift_y, shift_x):
    """The reference's early return (utils/torch.py:211): torch.isclose(shift, 0) with default tolerances."""
    return abs(shift_y) <= 1e-8 and abs(shift_x) <= 1e-8


# --------------------------------------------------------------------------------------
# the MAP step / run                                              core.py:209-230
# --------------------------------------------------------------------------------------
def dataset_loss_and_grad(theta, ds, mask=None, logb=None, return_dlogb=False, shift_xy=None):
    """Poisson loss of one dataset and its gradient w.r.t. theta (log flux) [and w.r.t. the log background
    norm `logb` and the sub-pixel shift `shift_xy` = (shift_x, shift_y) of an NPredCalibration,
    models/npred.py:226-237, 329-337].

    ds: dict(counts, exposure_up, psf_up, background, f).  With `shift_xy` the return value gains a last element
    dshift_xy (None when the shift is ~0: the reference then returns the unshifted image without a graph)."""
    flux = flux_from_theta(theta, mask)
    bnorm = None if logb is None else np.exp(theta.dtype.type(logb))
    shifted = shift_xy is not None and not shift_is_identity(shift_xy[1], shift_xy[0])
    flux_in = flux
    if shifted:
        flux_in, d_dy, d_dx = shift_image(flux, shift_xy[1], shift_xy[0], ds["f"], return_grads=True)
    npred, pool = npred_forward(flux_in, ds["exposure_up"], ds["psf_up"], ds["background"], ds["f"], bnorm, return_pool=True)
    loss = poisson_nll(npred, ds["counts"])
    dn = poisson_nll_grad(npred, ds["counts"])
    dflux = npred_backward(dn, pool, flux_in, ds["exposure_up"], ds["psf_up"], ds["f"])
    dshift = None
    if shifted:
        dshift = np.array([(dflux * d_dx).sum(dtype=np.float64), (dflux * d_dy).sum(dtype=np.float64)], dtype=theta.dtype)
        dflux = shift_image_adjoint(dflux, shift_xy[1], shift_xy[0], ds["f"])
    out = (loss, dflux * flux, npred)
    if return_dlogb:
        out += (float((dn * ds["background"] * (1 if bnorm is None else bnorm)).sum(dtype=np.float64)),)
    if shift_xy is not None:
        out += (dshift,)
    return out


def map_step(theta, adam, ds, n_datasets, beta, gmm=None, shifts=None, stride=4, marginalize=False, mask=None,
             cal=None):
    """One reference step: total = L_d - beta * prior / D, backward, Adam (core.py:214-229).
    cal: None or dict(logb=array(1), adam=Adam[, shift_xy=array(2), adam_shift=Adam]) - the dataset's trainable log
    background norm and (shift_x, shift_y), each stepped by its own Adam state (torch.optim.Adam keeps a step counter
    per parameter; a parameter without gradient - a shift at 0 - is skipped)."""
    if cal is None:
        loss, dtheta, _ = dataset_loss_and_grad(theta, ds, mask)
    elif "shift_xy" in cal:
        loss, dtheta, _, dlogb, dshift = dataset_loss_and_grad(theta, ds, mask, cal["logb"][0], return_dlogb=True,
                                                               shift_xy=cal["shift_xy"])
        cal["logb"] = cal["adam"].step(cal["logb"], np.array([dlogb], dtype=theta.dtype))
        if dshift is not None:
            cal["shift_xy"] = cal["adam_shift"].step(cal["shift_xy"], dshift)
    else:
        loss, dtheta, _, dlogb = dataset_loss_and_grad(theta, ds, mask, cal["logb"][0], return_dlogb=True)
        cal["logb"] = cal["adam"].step(cal["logb"], np.array([dlogb], dtype=theta.dtype))
    prior = theta.dtype.type(0)
    if gmm is not None:
        flux = flux_from_theta(theta, mask)
        prior, dflux_p, _ = gmm_patch_prior(flux, gmm, shifts[0], shifts[1], stride, marginalize, return_grad=True)
        dtheta = dtheta - theta.dtype.type(beta / n_datasets) * dflux_p * flux
    total = loss - beta * prior / n_datasets
    return adam.step(theta, dtheta), float(total), float(loss), float(prior)


def prepare_dataset(dataset, f=1, dtype=np.float32):
    """numpy dataset dict (counts, psf, exposure, background) -> oracle dataset (npred.py:263-295)."""
    exposure_up, psf_up = npred_setup(np.asarray(dataset["exposure"], dtype=np.float32),
                                      np.asarray(dataset["psf"], dtype=np.float32), f)
    return dict(
        counts=np.asarray(dataset["counts"]).astype(dtype),
        background=np.asarray(dataset["background"]).astype(dtype),
        exposure_up=exposure_up.astype(dtype),
        psf_up=psf_up.astype(dtype),
        f=f,
    )


def map_run(flux_init_up, datasets, n_epochs, lr=0.1, beta=1.0, gmm=None, shifts=None, stride=4,
            marginalize=False, dtype=np.float32, trace_shifts=None, background_norms=None, shifts_xy=None):
    """MAPDeconvolver.run restated: sequential per-dataset Adam steps (core.py:209-230) and the
    per-epoch trace (loss.py:212-250).  `shifts[step]` are the injected cycle-spin draws of the
    training steps; `trace_shifts[epoch]` those consumed by `append_trace`'s extra prior call.
    `background_norms` / `shifts_xy` (one value / (shift_x, shift_y) pair per dataset): trainable NPredCalibration
    parameters.  Returns (flux_upsampled, trace rows[, background norms[, shifts_xy]])."""
    theta = np.log(np.asarray(flux_init_up, dtype=dtype))
    adam = Adam(theta.shape, lr=lr, dtype=dtype)
    D = len(datasets)
    trace = []
    step = 0
    cals = None
    if background_norms is not None:
        cals = [dict(logb=np.log(np.array([b], dtype=dtype)), adam=Adam((1,), lr=lr, dtype=dtype)) for b in background_norms]
        if shifts_xy is not None:
            for c, sxy in zip(cals, shifts_xy):
                c["shift_xy"] = np.array(sxy, dtype=dtype)
                c["adam_shift"] = Adam((2,), lr=lr, dtype=dtype)
    for epoch in range(n_epochs):
        for i_ds, ds in enumerate(datasets):
            sh = shifts[step] if gmm is not None else None
            # `fluxes` is evaluated before the step (core.py:217) and the same (by then stale)
            # tuple is handed to append_trace after the loop (core.py:245): the trace of an
            # epoch is the loss at the parameters *before* the epoch's last Adam step.
            flux = flux_from_theta(theta)
            theta, *_ = map_step(theta, adam, ds, D, beta, gmm, sh, stride, marginalize,
                                 cal=None if cals is None else cals[i_ds])
            step += 1
        # the calibration parameters are read live by append_trace (only the flux tuple is stale)
        bn = [None] * D if cals is None else [np.exp(c["logb"][0]) for c in cals]
        fl = [flux] * D
        if cals is not None and shifts_xy is not None:
            fl = [flux if shift_is_identity(c["shift_xy"][1], c["shift_xy"][0]) else
                  shift_image(flux, c["shift_xy"][1], c["shift_xy"][0], d["f"]) for c, d in zip(cals, datasets)]
        ld = [float(poisson_nll(npred_forward(f_, d["exposure_up"], d["psf_up"], d["background"], d["f"], b),
                                d["counts"])) for f_, d, b in zip(fl, datasets, bn)]
        lp = 0.0
        if gmm is not None:
            sh = trace_shifts[epoch]
            lp = float(gmm_patch_prior(flux, gmm, sh[0], sh[1], stride, marginalize))
        trace.append({"total": sum(ld) - beta * lp, "datasets-total": sum(ld), "priors-total": -beta * lp,
                      "datasets": ld})
    if cals is not None and shifts_xy is not None:
        return (flux_from_theta(theta), trace, [float(np.exp(c["logb"][0])) for c in cals],
                [c["shift_xy"].copy() for c in cals])
    if cals is not None:
        return flux_from_theta(theta), trace, [float(np.exp(c["logb"][0])) for c in cals]
    return flux_from_theta(theta), trace


def joint_loss_and_grad(theta, datasets, beta, gmm=None, shifts=None, stride=4, marginalize=False,
                        dataset_index=None, rows=None, lean=False):
    """Value and gradient (w.r.t. theta) of the joint objective sum_d L_d - beta * prior
    (TotalLoss.__call__, loss.py:257-261).  `dataset_index` / `rows` restrict to a shard (subset of
    datasets, block of prior patch rows): shard values / gradients sum to the whole."""
    idx = range(len(datasets)) if dataset_index is None else dataset_index
    flux = flux_from_theta(theta)
    total = theta.dtype.type(0)
    dtheta = np.zeros_like(theta)
    for i in idx:
        loss, dth, _ = dataset_loss_and_grad(theta, datasets[i])
        total += loss
        dtheta += dth
    if gmm is not None:
        r0, r1 = (None, None) if rows is None else rows
        if lean and (rows is None or r1 > r0):  # BASELINE-size inputs
            res = gmm_patch_prior_lean(flux, gmm, shifts[0], shifts[1], stride, marginalize, r0, r1)
            total -= beta * res["prior"]
            dtheta -= theta.dtype.type(beta) * res["dflux"] * flux
        elif rows is None or r1 > r0:
            prior, dflux_p, _ = gmm_patch_prior(flux, gmm, shifts[0], shifts[1], stride, marginalize, True, r0, r1)
            total -= beta * prior
            dtheta -= theta.dtype.type(beta) * dflux_p * flux
    return total, dtheta


def map_run_joint(flux_init_up, datasets, n_epochs, lr=0.1, beta=1.0, gmm=None, shifts=None, stride=4,
                  marginalize=False, dtype=np.float32):
    """Joint-step MAP: one Adam step per epoch on the joint objective; shifts[2*e] is consumed by the
    step of epoch e and shifts[2*e+1] by its trace evaluation (at the pre-step flux, like core.py:245)."""
    theta = np.log(np.asarray(flux_init_up, dtype=dtype))
    adam = Adam(theta.shape, lr=lr, dtype=dtype)
    trace = []
    for epoch in range(n_epochs):
        flux = flux_from_theta(theta)
        sh = shifts[2 * epoch] if gmm is not None else None
        _, dtheta = joint_loss_and_grad(theta, datasets, beta, gmm, sh, stride, marginalize)
        theta = adam.step(theta, dtheta)
        ld = [float(poisson_nll(npred_forward(flux, d["exposure_up"], d["psf_up"], d["background"], d["f"]),
                                d["counts"])) for d in datasets]
        lp = 0.0
        if gmm is not None:
            sh = shifts[2 * epoch + 1]
            lp = float(gmm_patch_prior(flux, gmm, sh[0], sh[1], stride, marginalize))
        trace.append({"total": sum(ld) - beta * lp, "datasets": ld, "priors-total": -beta * lp})
    return flux_from_theta(theta), trace
