"""Parity of every C-ABI kernel against the oracle on seeded inputs (B200 only)."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from conftest import load_golden
from oracle import jolideco_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from jolideco_b200 import ops

DEV = "cuda"


def t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(DEV)


def rel_max(a, ref):
    return np.abs(a - ref).max() / np.abs(ref).max()


def synthetic_gmm(K, seed=0, mean_scale=0.01):
    rng = np.random.default_rng(seed)
    A = rng.normal(0, 0.05, size=(K, 64, 64))
    cov = A @ A.transpose(0, 2, 1) + 0.01 * np.eye(64)
    means = rng.normal(0, mean_scale, size=(K, 64)) if mean_scale else np.zeros((K, 64))
    w = rng.uniform(0.5, 1.5, size=K)
    return means, cov, w / w.sum()


def pack(gmm):
    return ops.GMMPacked(gmm.means, gmm.precisions_cholesky, gmm.weights, gmm.pixel_weights, DEV)


def test_device_is_blackwell():
    ops.require_device()


def test_flux_forward():
    rng = np.random.default_rng(0)
    theta = rng.normal(size=(33, 47)).astype(np.float32)
    mask = (rng.uniform(size=theta.shape) > 0.3).astype(np.uint8)
    out = ops.flux_forward(t(theta), t(mask, torch.uint8)).cpu().numpy()
    assert_allclose(out, O.flux_from_theta(theta, mask), rtol=5e-7)  # expf: <= 2 ulp
    out = ops.flux_forward(t(theta), None, use_log_flux=False).cpu().numpy()
    assert np.array_equal(out, theta)


@pytest.mark.parametrize("shape,kshape", [((37, 45), (5, 4)), ((64, 64), (17, 17)), ((130, 70), (34, 34)),
                                          ((96, 96), (64, 64)), ((20, 23), (1, 1)), ((50, 40), (41, 67))])
def test_conv_forward_and_adjoint(shape, kshape):
    rng = np.random.default_rng(1)
    flux = rng.gamma(2.0, size=shape)
    E = rng.uniform(0.5, 1.5, size=shape)
    psf = rng.uniform(size=kshape)
    psf /= psf.sum()
    ref = O.convolve_fft(flux * E, psf)
    out = ops.conv_forward(t(flux), t(E), t(psf)).cpu().numpy()
    assert rel_max(out, ref) < 5e-6
    d = rng.normal(size=shape)
    ref_b = O.correlate_adjoint(d, psf) * E
    out_b = ops.conv_backward(t(d), t(E), t(psf), 1).cpu().numpy()
    assert rel_max(out_b, ref_b) < 5e-6


@pytest.mark.parametrize("shape,kshape", [((37, 45), (5, 4)), ((64, 64), (17, 17)), ((130, 70), (34, 34)),
                                          ((96, 96), (64, 64)), ((50, 40), (41, 67)), ((256, 256), (201, 201)),
                                          ((512, 512), (34, 34))])
def test_fft_conv_forward_and_adjoint(shape, kshape):
    """Shared-memory FFT path against the oracle (and hence against the direct path)."""
    rng = np.random.default_rng(11)
    flux = rng.gamma(2.0, size=shape) * np.exp(rng.normal(0, 1, size=shape))
    E = rng.uniform(0.5, 1.5, size=shape)
    psf = rng.uniform(size=kshape) * np.outer(np.hanning(kshape[0] + 2)[1:-1], np.hanning(kshape[1] + 2)[1:-1])
    psf /= psf.sum()
    plan = ops.FFTConvPlan(t(psf), *shape)
    ref = O.convolve_fft(flux * E, psf)
    out = ops.conv_forward_fft(t(flux), t(E), plan).cpu().numpy()
    assert rel_max(out, ref) < 5e-6
    d = rng.normal(size=shape)
    ref_b = O.correlate_adjoint(d, psf) * E
    out_b = ops.conv_backward_fft(t(d), t(E), plan, 1).cpu().numpy()
    assert rel_max(out_b, ref_b) < 5e-6
    acc = t(np.ones(shape))
    ops.conv_backward_fft(t(d), t(E), plan, 1, out=acc, accumulate=True)
    assert rel_max(acc.cpu().numpy(), ref_b + 1) < 5e-6


def test_fft_conv_backward_upsampled():
    rng = np.random.default_rng(12)
    H, W, f = 21, 17, 2
    psf = rng.uniform(size=(6, 6))
    E = rng.uniform(0.5, 1.5, size=(H * f, W * f))
    dn = rng.normal(size=(H, W))
    ref = O.npred_backward(dn, np.ones((H, W)), np.ones((H * f, W * f)), E, psf, f)
    plan = ops.FFTConvPlan(t(psf), H * f, W * f)
    out = ops.conv_backward_fft(t(dn), t(E), plan, f).cpu().numpy()
    assert rel_max(out, ref) < 5e-6


def test_conv_golden_even_kernel():
    g = load_golden("kat.npz")
    out = ops.conv_forward(t(g["conv_img"]), t(np.ones_like(g["conv_img"])), t(g["conv_ker"])).cpu().numpy()
    assert rel_max(out, g["conv_out"]) < 5e-6


@pytest.mark.parametrize("f", [1, 2, 3])
def test_conv_backward_upsampled(f):
    rng = np.random.default_rng(2)
    H, W = 21, 17
    fH, fW = H * f, W * f
    psf = rng.uniform(size=(6, 6))
    E = rng.uniform(0.5, 1.5, size=(fH, fW))
    dn = rng.normal(size=(H, W))
    pool = np.ones((H, W))
    ref = O.npred_backward(dn, pool, np.ones((fH, fW)), E, psf, f)
    out = ops.conv_backward(t(dn), t(E), t(psf), f).cpu().numpy()
    assert rel_max(out, ref) < 5e-6
    acc = t(np.ones((fH, fW)))
    ops.conv_backward(t(dn), t(E), t(psf), f, out=acc, accumulate=True)
    assert rel_max(acc.cpu().numpy(), ref + 1) < 5e-6


@pytest.mark.parametrize("tx,split", [(0, 0), (8, 1), (8, 2), (8, 4), (16, 1), (16, 2), (16, 4), (28, 1), (28, 2), (28, 4)])
@pytest.mark.parametrize("shape,kshape", [((96, 128), (17, 17)), ((64, 192), (18, 19)), ((80, 64), (34, 34)),
                                          ((72, 100), (5, 20)), ((64, 64), (3, 2)), ((50, 70), (9, 9)),
                                          ((40, 48), (23, 1))])
def test_conv_v3_variants(shape, kshape, tx, split):
    """cp.async kernel: every tile width / PSF-row split, vector (16 B) and scalar (4 B) staging, all tap tails
    (kw + dx mod 4), against the oracle; and identical to the previous kernel to rounding."""
    from jolideco_b200 import _lib

    rng = np.random.default_rng(21)
    flux = rng.gamma(2.0, size=shape)
    E = rng.uniform(0.5, 1.5, size=shape)
    psf = rng.uniform(size=kshape)
    psf /= psf.sum()
    d = rng.normal(size=shape)
    ref = O.convolve_fft(flux * E, psf)
    ref_b = O.correlate_adjoint(d, psf) * E
    _lib.call("jd_conv_tuning", 1, tx, split)
    try:
        out = ops.conv_forward(t(flux), t(E), t(psf)).cpu().numpy()
        out_b = ops.conv_backward(t(d), t(E), t(psf), 1).cpu().numpy()
        acc = t(np.ones(shape))
        ops.conv_backward(t(d), t(E), t(psf), 1, out=acc, accumulate=True)
    finally:
        _lib.call("jd_conv_tuning", 1, 0, 0)
    assert rel_max(out, ref) < 5e-6
    assert rel_max(out_b, ref_b) < 5e-6
    assert rel_max(acc.cpu().numpy(), ref_b + 1) < 5e-6


@pytest.mark.parametrize("f", [2, 3])
@pytest.mark.parametrize("split", [1, 4])
def test_conv_v3_backward_upsampled_multi_tile(f, split):
    from jolideco_b200 import _lib

    rng = np.random.default_rng(22)
    H, W = 40, 36
    fH, fW = H * f, W * f
    psf = rng.uniform(size=(7 * f, 7 * f))
    E = rng.uniform(0.5, 1.5, size=(fH, fW))
    dn = rng.normal(size=(H, W))
    ref = O.npred_backward(dn, np.ones((H, W)), np.ones((fH, fW)), E, psf, f)
    _lib.call("jd_conv_tuning", 1, 8, split)
    try:
        out = ops.conv_backward(t(dn), t(E), t(psf), f).cpu().numpy()
    finally:
        _lib.call("jd_conv_tuning", 1, 0, 0)
    assert rel_max(out, ref) < 5e-6


def test_conv_previous_kernel_still_matches():
    from jolideco_b200 import _lib

    rng = np.random.default_rng(23)
    shape, kshape = (96, 128), (17, 17)
    flux, E = rng.gamma(2.0, size=shape), rng.uniform(0.5, 1.5, size=shape)
    psf = rng.uniform(size=kshape)
    ref = O.convolve_fft(flux * E, psf)
    _lib.call("jd_conv_tuning", 0, 0, 0)
    try:
        out = ops.conv_forward(t(flux), t(E), t(psf)).cpu().numpy()
    finally:
        _lib.call("jd_conv_tuning", 1, 0, 0)
    assert rel_max(out, ref) < 5e-6


@pytest.mark.parametrize("f", [1, 2])
def test_poisson_forward_backward(f):
    rng = np.random.default_rng(3)
    H, W = 40, 28
    conv = rng.gamma(1.0, size=(H * f, W * f)).astype(np.float32)
    conv[0:f, 0 : 4 * f] = 0.0
    conv[3 * f : 4 * f, :f] = -0.3  # negative pool -> clipped, zero gradient
    bkg = np.full((H, W), 0.2, dtype=np.float32)
    bkg[0, :2] = 0.0  # npred exactly 0 -> log(eps)
    counts = rng.poisson(1.5, size=(H, W)).astype(np.float32)
    pool = O.sum_pool(conv.astype(np.float64), f)
    npred = np.clip(pool, 0, np.inf) + bkg
    loss_ref = O.poisson_nll(npred, counts.astype(np.float64))
    with np.errstate(over="ignore", divide="ignore"):
        grad_ref = (O.poisson_nll_grad(npred.astype(np.float32), counts) * (pool >= 0)).astype(np.float32)
    res = ops.poisson_forward_backward(t(conv), t(bkg), t(counts), f=f, want_npred=True)
    loss = res["loss_sum"].item() / (H * W)
    assert_allclose(loss, loss_ref, rtol=1e-6)
    assert_allclose(res["npred"].cpu().numpy(), npred, rtol=1e-6)
    # 1 - c/n cancels where c ~ n: absolute tolerance 1e-6 x grad_scale
    assert_allclose(res["dpool"].cpu().numpy(), grad_ref, rtol=2e-6, atol=1e-9)
    # background-norm gradient
    logb = np.float32(0.3)
    res = ops.poisson_forward_backward(t(conv), t(bkg), t(counts), f=f, bkg_log_norm=t(np.array([logb])))
    npred_b = np.clip(pool, 0, np.inf) + bkg.astype(np.float64) * np.exp(np.float64(logb))
    assert_allclose(res["loss_sum"].item() / (H * W), O.poisson_nll(npred_b, counts.astype(np.float64)), rtol=1e-6)
    m = npred_b > 0
    dlogb_ref = ((1 - counts[m] / npred_b[m]) / (H * W) * (npred_b[m] - np.clip(pool, 0, np.inf)[m])).sum()
    assert_allclose(res["dlogb"].item(), dlogb_ref, rtol=1e-5)


def test_gmm_log_prob_generic_d_sklearn_kat():
    # reference priors/patches/tests/test_gmm.py:10-35 (D=9, identity covariance, no pixel weights)
    means = np.linspace(-1, 1, 9).reshape((1, 9))
    gmm = O.GMM(means, np.array([np.eye(9)]), np.array([1.0]), meta_stride=None)
    packed = ops.GMMPacked(gmm.means, gmm.precisions_cholesky, gmm.weights, gmm.pixel_weights, DEV)
    x = np.ones((2, 9), dtype=np.float32)
    out = ops.gmm_log_prob(t(x), packed).cpu().numpy()
    assert_allclose(out, gmm.estimate_log_prob(x), rtol=1e-6)


def test_gmm_log_prob_golden():
    g = load_golden("kat.npz")
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"])
    out = ops.gmm_log_prob(t(g["gmm_x"]), pack(gmm)).cpu().numpy()
    assert_allclose(out, g["gmm_logp"], rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("shape", [(21, 19), (64, 64), (38, 46)])
def test_patch_extraction_bit_exact(shape):
    img = np.arange(shape[0] * shape[1], dtype=np.float32).reshape(shape)
    ny = (shape[0] - 8) // 4 + 1
    for sy in range(-2, 3):
        for sx in range(-2, 3):
            ref = O.view_as_overlapping_patches(O.cycle_spin_roll(img, sy, sx), 8, 4)
            out = ops.extract_patches(t(img), (sy, sx)).cpu().numpy()
            assert np.array_equal(out, ref)
            # row-block shards (multi-GPU prior): concatenation of shards == whole
            cut = max(1, ny // 3)
            parts = [ops.extract_patches(t(img), (sy, sx), rows=r).cpu().numpy() for r in [(0, cut), (cut, ny)]]
            assert np.array_equal(np.concatenate(parts), ref)


def test_patch_extraction_golden_stride2():
    g = load_golden("kat.npz")
    out = ops.extract_patches(t(g["patches_img"]), (0, 0), stride=2).cpu().numpy()
    assert np.array_equal(out, g["patches_8_2"])
    out = ops.extract_patches(t(g["patches_img"]), (0, 0), stride=4).cpu().numpy()
    assert np.array_equal(out, g["patches_8_4"])


def prior_cuda(flux, gmm_packed, sy, sx, marginalize, rows=None, backend=0):
    fH, fW = flux.shape
    c = 16.0 / 64.0 / (fH * fW)
    fl = t(flux)
    value, argmax, logp, s = ops.gmm_prior_forward(fl, (sy, sx), gmm_packed, 4, marginalize, rows, backend=backend)
    # backend 2 also exercises the bucketed max-mode backward, the others the warp-per-patch one
    G = ops.gmm_prior_backward(fl, (sy, sx), gmm_packed, -c, 4, marginalize, rows, argmax, logp, value,
                               bucketed=backend == 2)
    dflux = ops.patch_fold(G, fH, fW, (sy, sx), 4, rows)
    return s.item() * c, dflux.cpu().numpy(), argmax.cpu().numpy()


@pytest.mark.parametrize("backend", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("marginalize", [False, True])
@pytest.mark.parametrize("shape,shift", [((38, 46), (0, 0)), ((38, 46), (-2, 1)), ((64, 80), (2, -2)), ((24, 24), (1, 2))])
def test_gmm_prior_value_and_grad(marginalize, shape, shift, backend):
    rng = np.random.default_rng(5)
    flux = rng.gamma(2.0, size=shape).astype(np.float32)
    gmm64 = O.GMM(*synthetic_gmm(11, seed=2), dtype=np.float64)
    ref_v, ref_g, ref_k = O.gmm_patch_prior(flux.astype(np.float64), gmm64, shift[0], shift[1], 4, marginalize, True)
    v, gr, k = prior_cuda(flux, pack(O.GMM(*synthetic_gmm(11, seed=2))), shift[0], shift[1], marginalize,
                          backend=backend)
    assert_allclose(v, ref_v, rtol=1e-5)
    flipped = (k != ref_k).sum()
    assert flipped <= 0.01 * len(k)
    tol = 1e-5 if flipped == 0 else 1e-3
    assert np.linalg.norm(gr - ref_g) / np.linalg.norm(ref_g) < tol
    if flipped == 0:
        assert rel_max(gr, ref_g) < 2e-5


@pytest.mark.parametrize("backend", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("case", range(8))
def test_gmm_prior_golden(case, backend):
    g = load_golden("prior_step.npz")
    sy, sx = (int(v) for v in g[f"c{case}_shift"])
    marg = bool(g[f"c{case}_marginalize"])
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"])
    v, gr, _ = prior_cuda(g["flux"], pack(gmm), sy, sx, marg, backend=backend)
    assert_allclose(v, g[f"c{case}_f64_value"], rtol=1e-5)
    ref = g[f"c{case}_f64_grad"]
    assert rel_max(gr, ref) < 2e-5


def test_gmm_prior_row_blocks_sum_to_whole():
    rng = np.random.default_rng(6)
    flux = rng.gamma(2.0, size=(72, 40)).astype(np.float32)
    packed = pack(O.GMM(*synthetic_gmm(7, seed=3)))
    v, gr, _ = prior_cuda(flux, packed, -1, 2, False)
    ny = (72 - 8) // 4 + 1
    parts = [prior_cuda(flux, packed, -1, 2, False, rows=r) for r in [(0, 5), (5, 6), (6, ny)]]
    assert_allclose(sum(p[0] for p in parts), v, rtol=1e-6)
    assert rel_max(sum(p[1] for p in parts), gr) < 1e-6


@pytest.mark.parametrize("backend", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("mean_scale", [0.05, 0.0])
@pytest.mark.parametrize("marginalize", [False, True])
def test_gmm_prior_tensor_core_vs_cuda_core_full_size(marginalize, mean_scale, backend):
    """tcgen05 split-TF32 forward against the FP32 CUDA-core forward at the BASELINE config-2 size
    (512x512 flux, 16129 patches), K=32 with non-zero means: per-patch values to FP32 accuracy,
    identical argmax except near-ties."""
    rng = np.random.default_rng(9)
    flux = t(rng.gamma(2.0, size=(512, 512)) * np.exp(rng.normal(0, 1.0, size=(512, 512))))
    packed = pack(O.GMM(*synthetic_gmm(33, seed=5, mean_scale=mean_scale)))
    assert packed.zero_mean == (mean_scale == 0.0) and packed.upper_tri
    v0, k0, lp0, s0 = ops.gmm_prior_forward(flux, (1, -2), packed, 4, marginalize, want_logp=True, backend=0)
    v1, k1, lp1, s1 = ops.gmm_prior_forward(flux, (1, -2), packed, 4, marginalize, want_logp=True, backend=backend)
    lp0, lp1 = lp0.cpu().numpy().astype(np.float64), lp1.cpu().numpy().astype(np.float64)
    scale = np.abs(lp0).max(axis=1, keepdims=True)
    assert (np.abs(lp1 - lp0) / scale).max() < 2e-6
    assert_allclose(s1.item(), s0.item(), rtol=1e-6)
    assert_allclose(v1.cpu().numpy(), v0.cpu().numpy(), rtol=1e-5, atol=1e-3)
    assert (k0 != k1).sum().item() <= 3


@pytest.mark.parametrize("marginalize", [False, True])
@pytest.mark.parametrize("shape,rows,K", [((512, 512), None, 256), ((1024, 1024), (31, 63), 256), ((200, 328), None, 40),
                                          ((64, 80), None, 3), ((1024, 1024), None, 64)])
def test_gmm_prior_stream_k_equals_tile_per_cta(monkeypatch, shape, rows, K, marginalize):
    """Stream-K decomposition (segments merged through the workspace) against the one-tile-per-CTA kernel:
    identical per-component log-probabilities, hence bit-identical max / argmax; logsumexp to rounding.
    Run twice on the same workspace-free path to check the arrival counters are left at zero."""
    rng = np.random.default_rng(15)
    flux = t(rng.gamma(2.0, size=shape) * np.exp(rng.normal(0, 0.7, size=shape)))
    packed = pack(O.GMM(*synthetic_gmm(K, seed=9, mean_scale=0.02)))
    monkeypatch.setattr(ops, "TC_STREAMK", False)
    v0, k0, lp0, s0 = ops.gmm_prior_forward(flux, (2, -1), packed, 4, marginalize, rows=rows, want_logp=True, backend=1)
    monkeypatch.setattr(ops, "TC_STREAMK", True)
    P = v0.numel()
    ws = ops.tc_sk_workspace(P, packed.K, flux.device)
    monkeypatch.setattr(ops, "tc_sk_workspace", lambda *a: ws)
    for _ in range(2):
        v1, k1, lp1, s1 = ops.gmm_prior_forward(flux, (2, -1), packed, 4, marginalize, rows=rows, want_logp=True,
                                                backend=1)
        diff = lp0 != lp1
        if bool(diff.any()):
            idx = diff.nonzero().cpu().numpy()
            a, b = lp0[diff].cpu().numpy(), lp1[diff].cpu().numpy()
            raise AssertionError(f"logp differs in {idx.shape[0]} entries: tiles {np.unique(idx[:, 0] // 128)[:16]}, rows "
                                 f"{np.unique(idx[:, 0] % 128)[:16]}, components {np.unique(idx[:, 1])[:32]}, "
                                 f"max |diff| {np.abs(a - b).max()}, samples {a[:4]} vs {b[:4]}")
        assert torch.equal(k0, k1)
        if marginalize:
            assert_allclose(v1.cpu().numpy(), v0.cpu().numpy(), rtol=2e-6)
        else:
            assert torch.equal(v0, v1)
        assert_allclose(s1.item(), s0.item(), rtol=1e-9 if not marginalize else 1e-6)
    n_tiles2 = ((P + 127) // 128 + 1) // 2 * 2  # arrival counters: one per tile, whole CTA pairs
    assert int(ws[:4 * n_tiles2].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("marginalize", [False, True])
@pytest.mark.parametrize("shape,rows,K,mean_scale", [((512, 512), None, 256, 0.0), ((1024, 1024), (31, 63), 256, 0.02),
                                                     ((200, 328), None, 40, 0.02), ((64, 80), None, 3, 0.0),
                                                     ((1024, 1024), None, 64, 0.02), ((96, 2048), None, 16, 0.02)])
@pytest.mark.parametrize("new,old", [(4, 3), (5, 2)])
def test_gmm_prior_two_tiles_per_cta_equals_one_tile(shape, rows, K, mean_scale, marginalize, new, old):
    """Backends 4 / 5 (two patch tiles per CTA and staged operand image, jd_gmm_tcm2.cu; mixed TF32 / FP16 and split-FP16
    recipes) issue the same MMAs in the same order as the one-tile kernels of their recipe (backends 3 / 2) - the
    accumulators are identical, the epilogue sums the 64 squares in packed pairs (another order): per-component
    log-probabilities agree to FP32 rounding of that sum, the argmax except on ties within that rounding.  The
    workspace's arrival counters are back at zero after every launch (two launches on the same workspace) and the two
    launches are bit-identical."""
    rng = np.random.default_rng(16)
    flux = t(rng.gamma(2.0, size=shape) * np.exp(rng.normal(0, 0.7, size=shape)))
    packed = pack(O.GMM(*synthetic_gmm(K, seed=9, mean_scale=mean_scale)))
    v0, k0, lp0, s0 = ops.gmm_prior_forward(flux, (2, -1), packed, 4, marginalize, rows=rows, want_logp=True,
                                            backend=old)
    P = v0.numel()
    ws = ops.tcm_workspace(P, packed.K, flux.device, new)
    orig = ops.tcm_workspace
    ops.tcm_workspace = lambda *a: ws
    try:
        runs = []
        for _ in range(2):
            v1, k1, lp1, s1 = ops.gmm_prior_forward(flux, (2, -1), packed, 4, marginalize, rows=rows, want_logp=True,
                                                    backend=new)
            runs.append((v1.clone(), k1.clone(), lp1.clone()))
            a, b = lp0.cpu().numpy().astype(np.float64), lp1.cpu().numpy().astype(np.float64)
            # -0.5 * sum of squares + c_k: rounding of the sum relative to the larger of the two terms
            err = np.abs(a - b)
            tol = 2e-6 * (np.abs(a) + np.abs(packed.ck.cpu().numpy().astype(np.float64))[None, :])
            if (err > tol).any():
                idx = np.argwhere(err > tol)
                raise AssertionError(f"logp differs in {idx.shape[0]} entries: tiles {np.unique(idx[:, 0] // 128)[:16]}, "
                                     f"rows {np.unique(idx[:, 0] % 128)[:16]}, components {np.unique(idx[:, 1])[:32]}, "
                                     f"max |diff| {err.max()}, samples {a[err > tol][:4]} vs {b[err > tol][:4]}")
            flipped = (k0 != k1).cpu().numpy()
            if flipped.any():  # only where the two best components tie within the rounding above
                top2 = np.sort(a[flipped], axis=1)[:, -2:]
                assert (top2[:, 1] - top2[:, 0] <= 1e-6 * np.abs(top2[:, 1])).all()
                assert flipped.sum() <= max(2, P // 2000)
            assert_allclose(v1.cpu().numpy(), v0.cpu().numpy(), rtol=2e-6)
            assert_allclose(s1.item(), s0.item(), rtol=1e-6)
        for x, y in zip(*runs):
            assert torch.equal(x, y)  # run-to-run bit-identical
    finally:
        ops.tcm_workspace = orig
    n_tiles4 = ((P + 127) // 128 + 3) // 4 * 4  # arrival counters: one per tile, whole tile groups
    assert int(ws[:4 * n_tiles4].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("backend", [4, 5])
@pytest.mark.parametrize("marginalize", [False, True])
@pytest.mark.parametrize("shape,rows,K,clusters", [((1024, 1024), None, 256, 37), ((1024, 1024), (0, 32), 256, 18),
                                                   ((512, 512), None, 256, 48), ((200, 328), None, 40, 5),
                                                   ((64, 80), None, 3, 1), ((512, 512), None, 256, 1000)])
def test_gmm_prior_forward_on_part_of_the_sm_pairs(shape, rows, K, clusters, marginalize, backend):
    """jd_gmm_prior_forward_tcx2_on: the two-tile forward on at most `clusters` CTA pairs (the split of one-dataset
    steps).  Only the chunking of the (tile group, component) space changes: per-component log-probabilities, max and
    argmax are bit-identical to the launch on every SM pair, logsumexp values agree to FP32 rounding of the merged
    partial sums; the workspace of the full launch is large enough and is left clean."""
    rng = np.random.default_rng(21)
    flux = t(rng.gamma(2.0, size=shape) * np.exp(rng.normal(0, 0.7, size=shape)))
    packed = pack(O.GMM(*synthetic_gmm(K, seed=9, mean_scale=0.0)))
    v0, k0, lp0, s0 = ops.gmm_prior_forward(flux, (1, -2), packed, 4, marginalize, rows=rows, want_logp=True,
                                            backend=backend)
    ws = ops.tcm_workspace(v0.numel(), packed.K, flux.device, backend)
    orig = ops.tcm_workspace
    ops.tcm_workspace = lambda *a: ws
    try:
        v1, k1, lp1, s1 = ops.gmm_prior_forward(flux, (1, -2), packed, 4, marginalize, rows=rows, want_logp=True,
                                                backend=backend, clusters=clusters)
    finally:
        ops.tcm_workspace = orig
    assert torch.equal(lp0, lp1) and torch.equal(k0, k1)
    if marginalize:
        assert_allclose(v1.cpu().numpy(), v0.cpu().numpy(), rtol=2e-6)
    else:
        assert torch.equal(v0, v1)
    assert_allclose(s1.item(), s0.item(), rtol=1e-9 if not marginalize else 1e-6)
    P = v0.numel()
    n_tiles4 = ((P + 127) // 128 + 3) // 4 * 4
    assert int(ws[:4 * n_tiles4].view(torch.int32).abs().sum()) == 0


def test_gmm_prior_tensor_core_dense_precision_factors():
    """Non-triangular component matrices take the untrimmed MMA schedule (upper_tri = 0)."""
    rng = np.random.default_rng(10)
    K = 9
    L = rng.normal(0, 1.0, size=(K, 64, 64)).astype(np.float32)
    L[:, np.arange(64), np.arange(64)] = np.abs(L[:, np.arange(64), np.arange(64)]) + 1
    packed = ops.GMMPacked(rng.normal(0, 0.05, size=(K, 64)), L, np.full(K, 1.0 / K), O.get_pixel_weights(8, 4), DEV)
    assert not packed.upper_tri
    flux = t(rng.gamma(2.0, size=(100, 84)))
    v0, k0, lp0, s0 = ops.gmm_prior_forward(flux, (0, 1), packed, 4, False, want_logp=True, backend=0)
    for backend in (1, 2, 3, 4, 5):
        v1, k1, lp1, s1 = ops.gmm_prior_forward(flux, (0, 1), packed, 4, False, want_logp=True, backend=backend)
        lp1 = lp1.cpu().numpy().astype(np.float64)
        ref = lp0.cpu().numpy().astype(np.float64)
        assert (np.abs(lp1 - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() < 2e-6
        assert (k0 != k1).sum().item() == 0


def test_gmm_backward_bucketed_equals_per_patch_full_size():
    rng = np.random.default_rng(13)
    flux = t(rng.gamma(2.0, size=(512, 512)))
    packed = pack(O.GMM(*synthetic_gmm(40, seed=6)))
    value, argmax, _, _ = ops.gmm_prior_forward(flux, (2, -1), packed, 4, False, backend=1)
    Ga = ops.gmm_prior_backward(flux, (2, -1), packed, -1e-3, 4, False, None, argmax, None, value, bucketed=False)
    Gb = ops.gmm_prior_backward(flux, (2, -1), packed, -1e-3, 4, False, None, argmax, None, value, bucketed=True)
    a, b = Ga.cpu().numpy(), Gb.cpu().numpy()
    assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()


@pytest.mark.parametrize("mean_scale", [0.05, 0.0])
def test_gmm_backward_max_triangular_equals_lam_kernel_full_size(mean_scale):
    """(xc Lw - mw) Lw^T from the triangular factor against xc Lam - bk, 16 129 patches, K = 40."""
    rng = np.random.default_rng(16)
    flux = t(rng.gamma(2.0, size=(512, 512)))
    packed = pack(O.GMM(*synthetic_gmm(40, seed=6, mean_scale=mean_scale)))
    value, argmax, _, _ = ops.gmm_prior_forward(flux, (2, -1), packed, 4, False, backend=1)
    argmax[5] = -1  # a filtered patch: zero row
    Ga = ops.gmm_prior_backward(flux, (2, -1), packed, -1e-3, 4, False, None, argmax, None, value, tri=False)
    Gb = ops.gmm_prior_backward(flux, (2, -1), packed, -1e-3, 4, False, None, argmax, None, value, tri=True)
    a, b = Ga.cpu().numpy(), Gb.cpu().numpy()
    assert np.abs(b[5]).max() == 0
    assert np.abs(a - b).max() <= 3e-6 * np.abs(a).max()
    assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(a)


def test_gmm_lse_backward_tensor_core_vs_cuda_core_full_size():
    """Tensor-core logsumexp backward against the FP32 CUDA-core tile kernel at config-2 size."""
    rng = np.random.default_rng(14)
    flux = t(rng.gamma(2.0, size=(512, 512)) * np.exp(rng.normal(0, 0.5, size=(512, 512))))
    packed = pack(O.GMM(*synthetic_gmm(24, seed=8, mean_scale=0.05)))
    v0, k0, lp0, _ = ops.gmm_prior_forward(flux, (-1, 2), packed, 4, True, backend=0)
    v1, k1, lp1, _ = ops.gmm_prior_forward(flux, (-1, 2), packed, 4, True, backend=1)
    assert lp0.is_contiguous() and not lp1.is_contiguous()
    G0 = ops.gmm_prior_backward(flux, (-1, 2), packed, -1e-3, 4, True, None, k0, lp0, v0).cpu().numpy()
    G1 = ops.gmm_prior_backward(flux, (-1, 2), packed, -1e-3, 4, True, None, k1, lp1, v1).cpu().numpy()
    # responsibilities exp(logp - lse) amplify the fp32 round-off of |logp| ~ 1e2..1e3 in BOTH kernels:
    # compare in rel-L2 (the float64-oracle tests above bound each kernel separately at small sizes)
    assert np.linalg.norm(G1 - G0) / np.linalg.norm(G0) < 2e-5
    assert np.abs(G1 - G0).max() <= 1e-4 * np.abs(G0).max()


def test_gmm_prior_nan_patch_is_skipped():
    rng = np.random.default_rng(7)
    flux = rng.gamma(2.0, size=(32, 32)).astype(np.float32)
    flux[5, 5] = np.nan
    gmm = O.GMM(*synthetic_gmm(5, seed=4))
    ref = O.gmm_patch_prior(flux, gmm, 0, 0)
    v, _, k = prior_cuda(flux, pack(gmm), 0, 0, False)
    assert_allclose(v, ref, rtol=1e-5)
    assert (k == -1).sum() == 4


def test_adam_matches_torch_and_oracle():
    rng = np.random.default_rng(8)
    n = (37, 29)
    theta0 = rng.normal(size=n).astype(np.float32)
    ga = rng.normal(size=(6,) + n).astype(np.float32)
    gb = rng.normal(size=(6,) + n).astype(np.float32)
    th = t(theta0.copy())
    m, v = torch.zeros_like(th), torch.zeros_like(th)
    p = torch.nn.Parameter(torch.from_numpy(theta0.copy()))
    opt = torch.optim.Adam([p], lr=0.1)
    adam = O.Adam(n, lr=0.1)
    th_o = theta0.copy()
    for s in range(6):
        flux = ops.flux_forward(th)
        ops.adam_step(th, m, v, flux, t(ga[s]), t(gb[s]), scale_b=-0.5, step=s + 1, lr=0.1)
        gtheta = (ga[s] - np.float32(0.5) * gb[s]) * np.exp(p.detach().numpy())
        p.grad = torch.from_numpy(gtheta)
        opt.step()
        th_o = adam.step(th_o, (ga[s] - np.float32(0.5) * gb[s]) * np.exp(th_o))
    assert_allclose(th.cpu().numpy(), p.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert_allclose(th.cpu().numpy(), th_o, rtol=1e-5, atol=1e-6)


def test_bad_arguments_raise():
    from jolideco_b200._lib import JolidecoB200Error

    with pytest.raises(JolidecoB200Error):
        ops.extract_patches(t(np.zeros((4, 4))))  # smaller than a patch
    with pytest.raises(JolidecoB200Error):
        ops.flux_forward(torch.zeros(3, 3, device=DEV, dtype=torch.float64))


@pytest.mark.parametrize("n,k", [(512, 64), (1024, 201)])
def test_fft_convolution_at_baseline_shapes(n, k):
    """BASELINE configs[2] and [3]: 512^2 flux with a 64 x 64 EVEN PSF (asymmetric crop, utils/torch.py:337-344) and
    1024^2 with a 201 x 201 PSF on the radix-5 1280^2 plan; forward and adjoint, two datasets' worth of PSFs."""
    rng = np.random.default_rng(100 + k)
    flux = (rng.gamma(2.0, size=(n, n)) * np.exp(rng.normal(0, 1, size=(n, n)))).astype(np.float32)
    E = rng.uniform(0.5, 1.5, size=(n, n)).astype(np.float32)
    d = rng.normal(size=(n, n)).astype(np.float32)
    for rep in range(2):
        yy, xx = np.mgrid[:k, :k]
        sig = k / (6.0 + rep)
        psf = np.exp(-0.5 * (((yy - (k - 1) / 2 - 0.7 * rep) / sig) ** 2 + ((xx - (k - 1) / 2 + 1.3 * rep) / (1.2 * sig)) ** 2))
        psf = (psf / psf.sum()).astype(np.float32)
        plan = ops.FFTConvPlan(t(psf), n, n)
        ref = O.convolve_fft(flux.astype(np.float64) * E, psf.astype(np.float64))
        out = ops.conv_forward_fft(t(flux), t(E), plan).cpu().numpy()
        assert rel_max(out, ref) < 5e-6
        ref_b = O.correlate_adjoint(d.astype(np.float64), psf.astype(np.float64)) * E
        out_b = ops.conv_backward_fft(t(d), t(E), plan, 1).cpu().numpy()
        assert rel_max(out_b, ref_b) < 5e-6
        # <conv(x), d> == <x, conv^T(d)> on the kernels themselves (adjoint consistency, float32 accumulation)
        lhs = float((out.astype(np.float64) * d).sum())
        rhs = float((flux.astype(np.float64) * out_b).sum())
        assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), abs(rhs), np.abs(out).sum() * 1e-3)
