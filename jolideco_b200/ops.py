"""Tensor-level wrappers over the C ABI: validate, allocate outputs with torch, pass raw device
pointers and the current CUDA stream.  No arithmetic happens in Python/PyTorch here."""
import math
import os

import numpy as np
import torch

from . import _lib

PATCH = 8
PD = 64


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, name, dtype=torch.float32):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.JolidecoB200Error(f"{name}: expected a CUDA tensor (jolideco_b200 has no CPU path), got "
                                     f"{type(t).__name__} on {getattr(t, 'device', '?')}")
    if t.dtype != dtype:
        raise _lib.JolidecoB200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.JolidecoB200Error(f"{name}: tensor must be contiguous")
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def _hw(t):
    return int(t.shape[-2]), int(t.shape[-1])


def require_device(device=None):
    """Raise unless `device` (default: current) is a Blackwell (sm_100) GPU."""
    if not torch.cuda.is_available():
        raise _lib.JolidecoB200Error("no CUDA device: jolideco_b200 runs on B200 (sm_100a) only, there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    ok = _lib.load().jd_device_supported(idx)
    if ok != 1:
        raise _lib.JolidecoB200Error(f"cuda:{idx} is not an sm_100 device ({torch.cuda.get_device_name(idx)})")


# ------------------------------------------------------------------------------------------------
def flux_forward(theta, mask=None, use_log_flux=True, out=None):
    _check(theta, "theta")
    _check(mask, "mask", torch.uint8)
    out = torch.empty_like(theta) if out is None else _check(out, "out")
    _lib.call("jd_flux_forward", _ptr(theta), _ptr(mask), _ptr(out), theta.numel(), int(use_log_flux), _stream())
    return out


def conv_forward(flux, exposure, psf, out=None):
    _check(flux, "flux"), _check(exposure, "exposure"), _check(psf, "psf")
    fH, fW = _hw(flux)
    kh, kw = _hw(psf)
    out = torch.empty_like(flux) if out is None else _check(out, "out")
    _lib.call("jd_conv_forward_direct", _ptr(flux), _ptr(exposure), _ptr(psf), _ptr(out), fH, fW, kh, kw, _stream())
    return out


# PSFs at least this large take the shared-memory FFT path.  tools/conv_exp.py (back-to-back launches) puts the
# cross-over of the cp.async direct kernel near 36 x 36 on a 512^2 grid, but inside the cfg2 step (cold L2) the FFT
# path is still 7 us ahead at 34 x 34: the threshold sits between the 23 x 23 and 34 x 34 measurements.
FFT_MIN_PSF_AREA = 30 * 30


class FFTConvPlan:
    """Cached PSF spectrum + scratch workspace of one dataset for the FFT convolution path."""

    def __init__(self, psf, fH, fW):
        import ctypes

        _check(psf, "psf")
        self.kh, self.kw = _hw(psf)
        self.fH, self.fW = int(fH), int(fW)
        n_hat, n_ws = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.call("jd_fftconv_sizes", self.fH, self.fW, self.kh, self.kw, ctypes.addressof(n_hat), ctypes.addressof(n_ws))
        self.psf_hat = torch.empty(n_hat.value, dtype=torch.float32, device=psf.device)
        self.workspace = torch.empty(n_ws.value, dtype=torch.float32, device=psf.device)
        with torch.cuda.device(psf.device):
            _lib.call("jd_fftconv_prepare_psf", _ptr(psf), self.kh, self.kw, self.fH, self.fW, _ptr(self.psf_hat),
                      _ptr(self.workspace), _stream())


def conv_forward_fft(flux, exposure, plan, out=None):
    _check(flux, "flux"), _check(exposure, "exposure")
    out = torch.empty_like(flux) if out is None else _check(out, "out")
    _lib.call("jd_conv_forward_fft", _ptr(flux), _ptr(exposure), _ptr(plan.psf_hat), _ptr(plan.workspace), _ptr(out),
              plan.fH, plan.fW, plan.kh, plan.kw, _stream())
    return out


def conv_backward_fft(dpool, exposure, plan, f, out=None, accumulate=False):
    _check(dpool, "dpool"), _check(exposure, "exposure")
    H, W = _hw(dpool)
    if out is None:
        out = torch.empty_like(exposure)
        accumulate = False
    _check(out, "out")
    _lib.call("jd_conv_backward_fft", _ptr(dpool), _ptr(exposure), _ptr(plan.psf_hat), _ptr(plan.workspace), _ptr(out),
              int(accumulate), plan.fH, plan.fW, plan.kh, plan.kw, int(f), H, W, _stream())
    return out


def conv_backward(dpool, exposure, psf, f, out=None, accumulate=False):
    _check(dpool, "dpool"), _check(exposure, "exposure"), _check(psf, "psf")
    H, W = _hw(dpool)
    fH, fW = _hw(exposure)
    kh, kw = _hw(psf)
    if out is None:
        out = torch.empty_like(exposure)
        accumulate = False
    _check(out, "out")
    _lib.call("jd_conv_backward_direct", _ptr(dpool), _ptr(exposure), _ptr(psf), _ptr(out), int(accumulate), fH, fW,
              kh, kw, int(f), H, W, _stream())
    return out


def pool_clip(conv, H, W, f):
    """Pre-clip sum-pool of the convolution (the caller clamps)."""
    _check(conv, "conv")
    out = torch.empty((H, W), dtype=torch.float32, device=conv.device)
    _lib.call("jd_pool_sum", _ptr(conv), _ptr(out), H, W, int(f), int(conv.shape[-1]), _stream())
    return out


def poisson_forward_backward(conv, background, counts, f=1, bkg_log_norm=None, loss_sum=None, dlogb=None,
                             want_npred=False, want_grad=True, grad_scale=None, eps=1e-25):
    """Returns dict(loss_sum=double[1] tensor (sum over pixels), npred, dpool, dlogb)."""
    _check(conv, "conv"), _check(background, "background"), _check(counts, "counts")
    _check(bkg_log_norm, "bkg_log_norm")
    H, W = _hw(counts)
    fW = int(conv.shape[-1])
    if loss_sum is None:
        loss_sum = torch.zeros(1, dtype=torch.float64, device=conv.device)
    if dlogb is None and bkg_log_norm is not None and want_grad:
        dlogb = torch.zeros(1, dtype=torch.float64, device=conv.device)
    npred = torch.empty_like(counts) if want_npred else None
    dpool = torch.empty_like(counts) if want_grad else None
    if grad_scale is None:
        grad_scale = 1.0 / (H * W)
    _lib.call("jd_poisson_forward_backward", _ptr(conv), _ptr(background), _ptr(bkg_log_norm), _ptr(counts),
              _ptr(npred), _ptr(dpool), _ptr(loss_sum), _ptr(dlogb), H, W, int(f), fW, float(eps), float(grad_scale),
              _stream())
    return dict(loss_sum=loss_sum, npred=npred, dpool=dpool, dlogb=dlogb)


# jd_lik_dataset of include/jolideco_b200.h (96 bytes): one record per dataset of a batched likelihood launch
LIK_DTYPE = np.dtype([("flux", "u8"), ("exposure", "u8"), ("psf", "u8"), ("background", "u8"), ("counts", "u8"),
                      ("bkg_log_norm", "u8"), ("dpool", "u8"), ("loss_sum", "u8"), ("dlogb", "u8"), ("dflux", "u8"),
                      ("loss_const", "f8"), ("accumulate", "i4"), ("reserved", "i4")])


# jd_fftlik_dataset: the same record + the dataset's spectrum workspace and cached PSF spectrum (FFT path)
FFTLIK_DTYPE = np.dtype(LIK_DTYPE.descr + [("workspace", "u8"), ("psf_hat", "u8")])


def stirling_constant(counts):
    """sum_pix 1{c > 1} (c log c - c + 1/2 log(2 pi c)): the counts-only term of nn.PoissonNLLLoss(full=True)
    (loss.py:35-37), a constant of the dataset that the fused likelihood kernel adds once (setup-time, float64)."""
    c = counts.double()
    cc = c.clamp(min=1)
    return float(torch.where(c > 1, c * torch.log(cc) - c + 0.5 * torch.log(2 * math.pi * cc), torch.zeros_like(c)).sum())


def likelihood_batched(flux, datasets, f=1, want_grad=True, eps=1e-25, fft=False):
    """All datasets of a joint iteration through jd_likelihood_forward (+ _backward), or with fft=True through the
    shared-memory FFT path jd_likelihood_forward_fft (+ _backward_fft): list of dicts with CUDA tensors exposure, psf,
    background, counts [, bkg_log_norm, flux (own NPred input)].  Returns dict(loss_sum (D,) double, dlogb (D,) double,
    dpool [D x (H,W)], dflux (D,fH,fW))."""
    _check(flux, "flux")
    fH, fW = _hw(flux)
    D = len(datasets)
    kh, kw = _hw(datasets[0]["psf"])
    H, W = _hw(datasets[0]["counts"])
    dev = flux.device
    loss = torch.zeros(D, dtype=torch.float64, device=dev)
    dlogb = torch.zeros(D, dtype=torch.float64, device=dev)
    dpool = torch.zeros((D, H, W), dtype=torch.float32, device=dev) if want_grad else None
    dflux = torch.zeros((D, fH, fW), dtype=torch.float32, device=dev) if want_grad else None
    rec = np.zeros(D, dtype=FFTLIK_DTYPE if fft else LIK_DTYPE)
    plans = [FFTConvPlan(d["psf"], fH, fW) for d in datasets] if fft else []
    for i, (r, d) in enumerate(zip(rec, datasets)):
        for k in ("exposure", "psf", "background", "counts"):
            r[k] = _ptr(_check(d[k], k))
        if fft:
            r["workspace"], r["psf_hat"] = _ptr(plans[i].workspace), _ptr(plans[i].psf_hat)
        r["flux"] = _ptr(_check(d.get("flux", flux), "flux"))
        r["bkg_log_norm"] = _ptr(_check(d.get("bkg_log_norm"), "bkg_log_norm")) or 0
        r["loss_sum"] = loss.data_ptr() + 8 * i
        r["loss_const"] = stirling_constant(d["counts"])
        if want_grad:
            r["dpool"] = dpool.data_ptr() + 4 * H * W * i
            r["dlogb"] = dlogb.data_ptr() + 8 * i
            r["dflux"] = dflux.data_ptr() + 4 * fH * fW * i
    table = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(dev)
    sfx = "_fft" if fft else ""
    _lib.call("jd_likelihood_forward" + sfx, _ptr(table), D, fH, fW, kh, kw, int(f), H, W, float(eps), 1.0 / (H * W),
              _stream())
    if want_grad:
        _lib.call("jd_likelihood_backward" + sfx, _ptr(table), D, fH, fW, kh, kw, int(f), H, W, _stream())
    return dict(loss_sum=loss, dlogb=dlogb, dpool=dpool, dflux=dflux)


# ------------------------------------------------------------------------------------------------
class GMMPacked:
    """Device constants of a Gaussian mixture in the layout the kernels consume.

    Built once at setup (float64 on the host, then rounded to float32) from the buffers the
    reference keeps (priors/patches/gmm.py:83-86, 217-240, 283-299):
        Lw[k]  = L_k diag(sqrt(w))            (K, D, D)   L_k = precisions_cholesky[k]
        mw[k]  = (mu_k L_k) sqrt(w)           (K, D)
        ck[k]  = -D/2 log 2pi + sum log diag L_k + log pi_k
        Lam[k] = Lw_k Lw_k^T,  bk[k] = mw_k Lw_k^T        (backward: xc Lam - bk = (xc Lw - mw) Lw^T)
    """

    def __init__(self, means, precisions_cholesky, weights, pixel_weights, device):
        # setup-time plumbing: float64 torch ops on the device (a few ms for K=256; numpy took ~30 ms)
        device = torch.device(device)
        with torch.cuda.device(device):
            def up(a):
                return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(device)

            L32, mu32 = up(precisions_cholesky), up(means)
            L, pi = L32.double(), up(weights).double()
            w = up(np.asarray(pixel_weights).reshape(-1)).double()
            K, D, _ = L.shape
            self.K, self.D = int(K), int(D)
            sw = torch.sqrt(w)
            Lw = L * sw[None, None, :]
            # mu L evaluated in float32 like the reference's lazily cached means_precisions_cholesky
            muL = torch.einsum("ki,kij->kj", mu32, L32).double()
            mw = muL * sw[None, :]
            log_det = torch.log(torch.diagonal(L, dim1=1, dim2=2)).sum(dim=1)
            ck = -0.5 * D * math.log(2 * math.pi) + log_det + torch.log(pi)
            Lam = Lw @ Lw.transpose(1, 2)
            bk = torch.einsum("kj,kij->ki", mw, Lw)
            self.Lw, self.mw, self.ck = Lw.float().contiguous(), mw.float().contiguous(), ck.float().contiguous()
            self.Lam, self.bk = Lam.float().contiguous(), bk.float().contiguous()
            flags = torch.stack([(torch.tril(L, -1) == 0).all(), (mw == 0).all()]).cpu()
        self.device = device
        self._Bt = None
        self._Bt16 = None
        self._Btm = None
        self._Bt_lam = None
        self.upper_tri = bool(flags[0])
        self.zero_mean = bool(flags[1])

    @property
    def Bt(self):
        """Tensor-core operand image of Lw (split-TF32, swizzled K-major), packed once on the device."""
        if self._Bt is None:
            if self.D != PD:
                raise _lib.JolidecoB200Error("the tcgen05 prior kernel supports 8x8 patches (D=64) only")
            nbytes = _lib.load().jd_gmm_tc_packed_bytes(self.K)
            with torch.cuda.device(self.device):
                bt = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                _lib.call("jd_gmm_tc_pack", _ptr(self.Lw), self.K, _ptr(bt), _stream())
            self._Bt = bt
        return self._Bt


def _bt_lam(packed):
    """Tensor-core operand image of Lam_k (for the logsumexp backward)."""
    if packed._Bt_lam is None:
        nbytes = _lib.load().jd_gmm_tc_packed_bytes(packed.K)
        with torch.cuda.device(packed.device):
            bt = torch.empty(nbytes, dtype=torch.uint8, device=packed.device)
            _lib.call("jd_gmm_tc_pack", _ptr(packed.Lam), packed.K, _ptr(bt), _stream())
        packed._Bt_lam = bt
    return packed._Bt_lam


def _bt16(packed):
    """Split-FP16 operand image + inverse component scales, packed once on the device."""
    if packed._Bt16 is None:
        if packed.D != PD:
            raise _lib.JolidecoB200Error("the tcgen05 prior kernels support 8x8 patches (D=64) only")
        nbytes = _lib.load().jd_gmm_tc16_packed_bytes(packed.K)
        with torch.cuda.device(packed.device):
            bt = torch.empty(nbytes, dtype=torch.uint8, device=packed.device)
            binv = torch.empty(packed.K, dtype=torch.float32, device=packed.device)
            _lib.call("jd_gmm_tc16_pack", _ptr(packed.Lw), packed.K, _ptr(bt), _ptr(binv), _stream())
        packed._Bt16 = (bt, binv)
    return packed._Bt16


def _btm(packed):
    """Operand image of the mixed TF32 / FP16 kernel (backend 3) + inverse component scales, packed once."""
    if packed._Btm is None:
        if packed.D != PD:
            raise _lib.JolidecoB200Error("the tcgen05 prior kernels support 8x8 patches (D=64) only")
        nbytes = _lib.load().jd_gmm_tcm_packed_bytes(packed.K)
        with torch.cuda.device(packed.device):
            bt = torch.empty(nbytes, dtype=torch.uint8, device=packed.device)
            binv = torch.empty(packed.K, dtype=torch.float32, device=packed.device)
            _lib.call("jd_gmm_tcm_pack", _ptr(packed.Lw), packed.K, _ptr(bt), _ptr(binv), _stream())
        packed._Btm = (bt, binv)
    return packed._Btm


# persistent stream-K forwards: backend -> C entry point
TCM_ENTRY = {3: "jd_gmm_prior_forward_tcm", 4: "jd_gmm_prior_forward_tcm2", 5: "jd_gmm_prior_forward_tc16x2"}


def tcm_workspace(P, K, device, backend=3):
    """Zero-initialised workspace of the backend-3 / 4 forwards (arrival counters + per-segment partials)."""
    fn = _lib.load().jd_gmm_tcm2_workspace_bytes if int(backend) in (4, 5) else _lib.load().jd_gmm_tcm_workspace_bytes
    n = int(fn(int(P), int(K)))
    return torch.zeros(max(n, 256), dtype=torch.uint8, device=device)


# tcgen05 prior forward: stream-K work decomposition (jd_gmm_prior_forward_tc_sk).  JD_TC_STREAMK = 0 | 1 forces it
# off / on; default "auto": only when the one-tile-per-CTA kernel would leave more than a quarter of the SMs idle
# (row-block shards of a multi-GPU run, small images) - with every SM busy the per-segment overhead of stream-K
# (pipeline drain + gather at each tile boundary) outweighs the balance (profiles/r01_summary.md).
TC_STREAMK = {"0": False, "1": True}.get(os.environ.get("JD_TC_STREAMK", "auto"), None)


def use_stream_k(P, device):
    if TC_STREAMK is not None:
        return TC_STREAMK
    n_pairs = ((int(P) + 127) // 128 + 1) // 2
    clusters = torch.cuda.get_device_properties(device).multi_processor_count // 2
    return 0 < n_pairs <= 0.75 * clusters


def tc_sk_workspace(P, K, device):
    """Zero-initialised workspace of the stream-K forward (arrival counters + per-segment partials)."""
    n = int(_lib.load().jd_gmm_tc_sk_workspace_bytes(int(P), int(K)))
    return torch.zeros(max(n, 256), dtype=torch.uint8, device=device)


def gmm_log_prob(x, packed):
    _check(x, "x")
    P, D = x.shape
    if D != packed.D:
        raise _lib.JolidecoB200Error(f"gmm_log_prob: x has {D} features, GMM has {packed.D}")
    out = torch.empty((P, packed.K), dtype=torch.float32, device=x.device)
    _lib.call("jd_gmm_log_prob", _ptr(x), P, D, packed.K, _ptr(packed.Lw), _ptr(packed.mw), _ptr(packed.ck),
              _ptr(out), _stream())
    return out


def patch_grid(fH, fW, stride):
    return (fH - PATCH) // stride + 1, (fW - PATCH) // stride + 1


def as_shift_tensor(shift_yx, device):
    if isinstance(shift_yx, torch.Tensor):
        return _check(shift_yx, "shift_yx", torch.int32)
    return torch.tensor([int(shift_yx[0]), int(shift_yx[1])], dtype=torch.int32, device=device)


def extract_patches(flux, shift_yx=(0, 0), stride=4, rows=None):
    _check(flux, "flux")
    fH, fW = _hw(flux)
    ny, nx = patch_grid(fH, fW, stride)
    r0, r1 = (0, ny) if rows is None else rows
    shift = as_shift_tensor(shift_yx, flux.device)
    X = torch.empty(((r1 - r0) * nx, PD), dtype=torch.float32, device=flux.device)
    _lib.call("jd_extract_patches", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(X), _stream())
    return X


def gmm_prior_forward(flux, shift_yx, packed, stride=4, marginalize=False, rows=None, want_logp=None, sum_out=None,
                      backend=0, clusters=0):
    """Returns (value[P'], argmax[P'], logp[P',K] or None, sum double[1]).  clusters > 0 (backends 4 / 5): on at most
    that many CTA pairs (jd_gmm_prior_forward_tcx2_on)."""
    _check(flux, "flux")
    if packed.D != PD:
        raise _lib.JolidecoB200Error("gmm_prior_forward: only 8x8 patches (D=64) are supported")
    fH, fW = _hw(flux)
    ny, nx = patch_grid(fH, fW, stride)
    r0, r1 = (0, ny) if rows is None else rows
    P = (r1 - r0) * nx
    shift = as_shift_tensor(shift_yx, flux.device)
    value = torch.empty(P, dtype=torch.float32, device=flux.device)
    argmax = torch.empty(P, dtype=torch.int32, device=flux.device)
    if want_logp is None:
        want_logp = bool(marginalize)
    # the tensor-core forwards write logp component-major (K x P'): returned as the transposed (P', K) view
    tc = int(backend) in (1, 2, 3, 4, 5)
    logp = None
    if want_logp:
        logp = torch.empty((packed.K, P) if tc else (P, packed.K), dtype=torch.float32, device=flux.device)
    if sum_out is None:
        sum_out = torch.zeros(1, dtype=torch.float64, device=flux.device)
    if int(backend) in (4, 5) and clusters:
        bt, binv = _bt16(packed) if int(backend) == 5 else _btm(packed)
        ws = tcm_workspace(P, packed.K, flux.device, backend)
        _lib.call("jd_gmm_prior_forward_tcx2_on", int(backend) - 4, int(clusters), _ptr(flux), fH, fW, _ptr(shift),
                  int(stride), r0, r1, _ptr(bt), _ptr(binv), _ptr(packed.mw), _ptr(packed.ck), packed.K,
                  int(packed.upper_tri), int(packed.zero_mean), int(bool(marginalize)), _ptr(ws), _ptr(value),
                  _ptr(argmax), _ptr(logp), _ptr(sum_out), _stream())
    elif int(backend) in (3, 4, 5):
        bt, binv = _bt16(packed) if int(backend) == 5 else _btm(packed)
        ws = tcm_workspace(P, packed.K, flux.device, backend)
        _lib.call(TCM_ENTRY[int(backend)], _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(bt), _ptr(binv),
                  _ptr(packed.mw), _ptr(packed.ck), packed.K, int(packed.upper_tri), int(packed.zero_mean),
                  int(bool(marginalize)), _ptr(ws), _ptr(value), _ptr(argmax), _ptr(logp), _ptr(sum_out), _stream())
    elif int(backend) == 2:
        bt, binv = _bt16(packed)
        _lib.call("jd_gmm_prior_forward_tc16", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(bt), _ptr(binv),
                  _ptr(packed.mw), _ptr(packed.ck), packed.K, int(packed.upper_tri), int(packed.zero_mean),
                  int(bool(marginalize)), _ptr(value), _ptr(argmax), _ptr(logp), _ptr(sum_out), _stream())
    elif int(backend) == 1 and use_stream_k(P, flux.device):
        ws = tc_sk_workspace(P, packed.K, flux.device)
        _lib.call("jd_gmm_prior_forward_tc_sk", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(packed.Bt),
                  _ptr(packed.mw), _ptr(packed.ck), packed.K, int(packed.upper_tri), int(packed.zero_mean),
                  int(bool(marginalize)), _ptr(ws), _ptr(value), _ptr(argmax), _ptr(logp), _ptr(sum_out), _stream())
    elif int(backend) == 1:
        _lib.call("jd_gmm_prior_forward_tc", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(packed.Bt),
                  _ptr(packed.mw), _ptr(packed.ck), packed.K, int(packed.upper_tri), int(packed.zero_mean), int(bool(marginalize)), _ptr(value),
                  _ptr(argmax), _ptr(logp), _ptr(sum_out), _stream())
    else:
        _lib.call("jd_gmm_prior_forward", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(packed.Lw),
                  _ptr(packed.mw), _ptr(packed.ck), packed.K, int(bool(marginalize)), _ptr(value), _ptr(argmax),
                  _ptr(logp), _ptr(sum_out), int(backend), _stream())
    if tc and logp is not None:
        logp = logp.t()
    return value, argmax, logp, sum_out


def gmm_backward_workspace(P, K, device):
    """Zero-initialised workspace of the bucketed max-mode backward (the kernels leave it ready for the next launch)."""
    n = _lib.load().jd_gmm_backward_workspace_elems(int(P), int(K))
    return torch.zeros(n, dtype=torch.int32, device=device)


# max-mode backward bucketed by winning component (Lam_k staged once per 64 patches): from this many patches on
# (profiles/r02_summary.md); JD_BWD_BUCKETED = 0 | 1 forces it off / on
BWD_BUCKETED = {"0": False, "1": True}.get(os.environ.get("JD_BWD_BUCKETED", "auto"), None)
BWD_BUCKETED_MIN_PATCHES = 32768


def use_bwd_bucketed(P, marginalize=False):
    if marginalize:
        return False
    return (P >= BWD_BUCKETED_MIN_PATCHES) if BWD_BUCKETED is None else BWD_BUCKETED


# max-mode backward from the triangular factor Lw (12 KB of L2 reads per patch) instead of Lam = Lw Lw^T (16 KB);
# JD_BWD_TRI=0 keeps the Lam kernel
BWD_TRI = os.environ.get("JD_BWD_TRI", "1") != "0"
# ... for up to this many patches: inside the steps the triangular kernel saves 15 us of 41 at 16 129 patches (cfg2)
# but loses 8 us of 78 at 65 025 (joint1024), where the leaner Lam kernel (64 resident warps per SM) hides latency better
BWD_TRI_MAX_PATCHES = 32768


def use_bwd_tri(packed, P, marginalize=False):
    return bool(BWD_TRI and not marginalize and packed.upper_tri and packed.D == PD and P <= BWD_TRI_MAX_PATCHES)


def gmm_prior_backward(flux, shift_yx, packed, scale, stride=4, marginalize=False, rows=None, argmax=None, logp=None,
                       value=None, out=None, workspace=None, bucketed=False, tri=None):
    _check(flux, "flux")
    fH, fW = _hw(flux)
    ny, nx = patch_grid(fH, fW, stride)
    r0, r1 = (0, ny) if rows is None else rows
    P = (r1 - r0) * nx
    shift = as_shift_tensor(shift_yx, flux.device)
    G = torch.empty((P, PD), dtype=torch.float32, device=flux.device) if out is None else _check(out, "out")
    if marginalize and logp is not None and not logp.is_contiguous() and logp.t().is_contiguous():
        # component-major logp from a tensor-core forward: tensor-core backward
        _lib.call("jd_gmm_prior_backward_lse_tc", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1,
                  _ptr(_bt_lam(packed)), _ptr(packed.bk), packed.K, _ptr(logp.t()), _ptr(_check(value, "value")),
                  float(scale), _ptr(G), _stream())
        return G
    if tri is None:
        tri = use_bwd_tri(packed, P, marginalize) and not bucketed
    if not marginalize and tri and packed.upper_tri and workspace is None and packed.D == PD:
        _lib.call("jd_gmm_prior_backward_max_tri", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(packed.Lw),
                  _ptr(packed.mw), packed.K, _ptr(_check(argmax, "argmax", torch.int32)), float(scale), _ptr(G), _stream())
        return G
    if workspace is None and bucketed and not marginalize:
        workspace = gmm_backward_workspace(P, packed.K, flux.device)
    _lib.call("jd_gmm_prior_backward", _ptr(flux), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(packed.Lam),
              _ptr(packed.bk), packed.K, int(bool(marginalize)), _ptr(_check(argmax, "argmax", torch.int32)),
              _ptr(_check(logp, "logp")), _ptr(_check(value, "value")), float(scale), _ptr(G),
              _ptr(_check(workspace, "workspace", torch.int32)), _stream())
    return G


def patch_fold(G, fH, fW, shift_yx, stride=4, rows=None, out=None, accumulate=False):
    _check(G, "G")
    ny, nx = patch_grid(fH, fW, stride)
    r0, r1 = (0, ny) if rows is None else rows
    shift = as_shift_tensor(shift_yx, G.device)
    if out is None:
        out = torch.empty((fH, fW), dtype=torch.float32, device=G.device)
        accumulate = False
    _check(out, "out")
    _lib.call("jd_patch_fold", _ptr(G), fH, fW, _ptr(shift), int(stride), r0, r1, _ptr(out), int(accumulate),
              _stream())
    return out


def adam_step(theta, m, v, flux, dflux_a, dflux_b=None, scale_b=0.0, mask=None, use_log_flux=True, step=1, lr=0.1,
              beta1=0.9, beta2=0.999, eps=1e-8):
    for t, n in [(theta, "theta"), (m, "m"), (v, "v"), (flux, "flux"), (dflux_a, "dflux_a"), (dflux_b, "dflux_b")]:
        _check(t, n)
    _check(mask, "mask", torch.uint8)
    _lib.call("jd_adam_step", _ptr(theta), _ptr(m), _ptr(v), _ptr(flux), _ptr(mask), _ptr(dflux_a), _ptr(dflux_b),
              float(scale_b), int(use_log_flux), theta.numel(), int(step), float(lr), float(beta1), float(beta2),
              float(eps), _stream())
    return theta
