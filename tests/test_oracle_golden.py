"""The oracle (numpy restatement) pinned against the reference: its known-answer tests, its e2e
golden values and outputs of the imported reference stored under tests/golden/ (CPU only)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from conftest import load_golden, unpack_datasets
from oracle import jolideco_oracle as O


def test_patch_order_kat():
    # reference utils/tests/test_torch.py:8-21
    x = np.arange(16).reshape(4, 4)
    p = O.view_as_overlapping_patches(x, 2)
    assert_allclose(p[0], [0, 1, 4, 5])
    assert_allclose(p[1], [1, 2, 5, 6])
    p = O.view_as_overlapping_patches(x, 2, stride=2)
    assert_allclose(p[0], [0, 1, 4, 5])
    assert_allclose(p[1], [2, 3, 6, 7])


def test_patches_golden():
    g = load_golden("kat.npz")
    assert np.array_equal(O.view_as_overlapping_patches(g["patches_img"], 8, 4), g["patches_8_4"])
    assert np.array_equal(O.view_as_overlapping_patches(g["patches_img"], 8, 2), g["patches_8_2"])


def test_convolve_kat():
    # reference utils/tests/test_torch.py:24-38: 9x9 box (*) 3x3 uniform == 'same' convolution
    image = np.zeros((9, 9))
    image[3:6, 3:6] = 1
    kernel = np.ones((3, 3)) / 9
    from scipy.signal import convolve2d

    ref = convolve2d(image, kernel, mode="same")
    assert_allclose(O.convolve_fft(image, kernel), ref, atol=1e-12)
    assert_allclose(O.convolve_direct(image, kernel), ref, atol=1e-12)


def test_convolve_golden_even_kernel():
    g = load_golden("kat.npz")
    assert_allclose(O.convolve_fft(g["conv_img"], g["conv_ker"]), g["conv_out"], atol=1e-12)
    assert_allclose(O.convolve_direct(g["conv_img"], g["conv_ker"]), g["conv_out"], atol=1e-12)


def test_adjoint_is_transpose():
    rng = np.random.default_rng(0)
    g = rng.normal(size=(13, 11))
    for kshape in [(4, 6), (5, 3), (6, 6)]:
        k = rng.uniform(size=kshape)
        d = rng.normal(size=g.shape)
        lhs = (O.convolve_fft(g, k) * d).sum()
        rhs = (g * O.correlate_adjoint(d, k)).sum()
        assert_allclose(lhs, rhs, rtol=1e-12)


def test_npred_kat():
    # reference models/tests/test_core.py:63-75: delta at (10,10), Gaussian sigma=3 PSF
    y, x = np.mgrid[-12:13, -12:13]
    psf = np.exp(-0.5 * (x**2 + y**2) / 9.0)
    psf /= psf.sum()
    flux = np.zeros((25, 25))
    flux[10, 10] = 1
    npred = O.npred_forward(flux, np.ones((25, 25)), psf, np.zeros((25, 25)), f=1)
    assert_allclose(npred[10, 10], 0.017684, atol=1e-5)
    assert_allclose(npred.sum(), 1.0, atol=1e-3)


def test_npred_setup_forward_grad_golden():
    g = load_golden("kat.npz")
    ds = {k: g[f"npred_ds_{k}"] for k in ["counts", "psf", "exposure", "background"]}
    flux_up = O.interpolate_bilinear(g["npred_flux_init"].astype(np.float32), 2)
    assert_allclose(flux_up, g["npred_flux_up"], rtol=2e-6)
    d = O.prepare_dataset(ds, f=2)
    assert_allclose(d["exposure_up"], g["npred_exposure_up"], rtol=1e-5)
    assert_allclose(d["psf_up"], g["npred_psf_up"], rtol=1e-5, atol=1e-9)
    # forward / loss / gradient in float64 from the reference's own float32 buffers
    d64 = dict(counts=ds["counts"].astype(float), background=ds["background"].astype(float),
               exposure_up=g["npred_exposure_up"].astype(float), psf_up=g["npred_psf_up"].astype(float), f=2)
    theta = np.log(g["npred_flux_up"].astype(float))
    loss, dtheta, npred = O.dataset_loss_and_grad(theta, d64)
    assert_allclose(npred, g["npred_out"], rtol=2e-5)
    assert_allclose(loss, g["npred_loss"], rtol=1e-6)
    assert_allclose(dtheta, g["npred_theta_grad"], rtol=2e-4, atol=1e-8)


def test_poisson_golden():
    g = load_golden("kat.npz")
    n, c = g["poisson_npred"], g["poisson_counts"]
    assert_allclose(O.poisson_nll(n, c), g["poisson_loss"], rtol=1e-6)
    with np.errstate(over="ignore"):
        grad = O.poisson_nll_grad(n, c)
    assert_allclose(grad, g["poisson_grad"], rtol=1e-6)


def test_gmm_vs_sklearn_kat():
    # reference priors/patches/tests/test_gmm.py:10-35
    from sklearn.mixture import GaussianMixture
    from sklearn.mixture._gaussian_mixture import _compute_precision_cholesky

    means = np.linspace(-1, 1, 9).reshape((1, 9))
    cov = np.array([np.eye(9)])
    w = np.array([1.0])
    gmm = O.GMM(means, cov, w, meta_stride=None)
    sk = GaussianMixture()
    sk.weights_, sk.covariances_, sk.means_ = w, cov, means
    sk.precisions_cholesky_ = _compute_precision_cholesky(cov, "full")
    x = np.ones((2, 9), dtype=np.float32)
    assert_allclose(gmm.estimate_log_prob(x), sk._estimate_weighted_log_prob(X=x), rtol=1e-6)


def test_gmm_constants_and_logp_golden():
    g = load_golden("kat.npz")
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], meta_stride=4)
    assert_allclose(gmm.precisions_cholesky, g["gmm_prec_chol"], rtol=1e-5, atol=1e-6)
    assert_allclose(gmm.means_precisions_cholesky, g["gmm_mu_prec"], rtol=1e-4, atol=1e-6)
    assert_allclose(gmm.log_det_cholesky, g["gmm_log_det"], rtol=1e-6)
    assert_allclose(gmm.pixel_weights, g["gmm_pixel_weights"].ravel(), rtol=1e-6)
    t = np.array([0.125, 0.375, 0.625, 0.875, 0.875, 0.625, 0.375, 0.125])
    assert_allclose(gmm.pixel_weights.reshape(8, 8), np.outer(t, t), rtol=1e-6)
    # log-probs cancel (-q/2 + log-det): fp32 round-off is absolute, ~1e-7 x |q|
    assert_allclose(gmm.estimate_log_prob(g["gmm_x"]), g["gmm_logp"], rtol=2e-5, atol=1e-4)


@pytest.mark.parametrize("case", range(8))
def test_prior_value_and_grad_golden(case):
    g = load_golden("prior_step.npz")
    sy, sx = (int(v) for v in g[f"c{case}_shift"])
    marg = bool(g[f"c{case}_marginalize"])
    # float64 oracle vs float64 reference: formulas exact
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], meta_stride=4, dtype=np.float64)
    val, grad, _ = O.gmm_patch_prior(g["flux"].astype(np.float64), gmm, sy, sx, 4, marg, return_grad=True)
    assert_allclose(val, g[f"c{case}_f64_value"], rtol=1e-12)
    assert_allclose(grad, g[f"c{case}_f64_grad"], rtol=1e-9, atol=1e-14)
    # float32 oracle vs float32 reference: within fp32 round-off
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], meta_stride=4, dtype=np.float32)
    val, grad, _ = O.gmm_patch_prior(g["flux"], gmm, sy, sx, 4, marg, return_grad=True)
    assert_allclose(val, g[f"c{case}_f32_value"], rtol=1e-5)
    ref = g[f"c{case}_f32_grad"]
    assert np.abs(grad - ref).max() <= 2e-5 * np.abs(ref).max()


def test_prior_row_blocks_sum_to_whole():
    g = load_golden("prior_step.npz")
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], dtype=np.float64)
    flux = g["flux"].astype(np.float64)
    val, grad, _ = O.gmm_patch_prior(flux, gmm, 1, -2, return_grad=True)
    ny = (flux.shape[0] - 8) // 4 + 1
    parts = [O.gmm_patch_prior(flux, gmm, 1, -2, return_grad=True, row_begin=a, row_end=b)
             for a, b in [(0, 3), (3, 4), (4, ny)]]
    assert_allclose(sum(p[0] for p in parts), val, rtol=1e-12)
    assert_allclose(sum(p[1] for p in parts), grad, rtol=1e-10, atol=1e-15)


def _run(name, f, n_epochs, rtol_flux, gmm=False):
    g = load_golden(name)
    datasets = [O.prepare_dataset(d, f=f) for d in unpack_datasets(g)]
    kw = {}
    if gmm:
        kw = dict(gmm=O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"]), shifts=g["shifts"],
                  trace_shifts=g["trace_shifts"], marginalize=bool(g["marginalize"]))
    flux_up, trace = O.map_run(g["flux_init_up"], datasets, n_epochs, **kw)
    rel = np.linalg.norm(flux_up - g["flux_up"]) / np.linalg.norm(g["flux_up"])
    assert rel < rtol_flux, rel
    assert_allclose([t["total"] for t in trace], g["trace_total"][:n_epochs], rtol=1e-5)
    assert_allclose([t["datasets"] for t in trace], g["trace_datasets"][:n_epochs], rtol=1e-5)
    return g, flux_up, trace


def test_run_uniform_reference_e2e_golden():
    g, flux_up, trace = _run("run_uniform.npz", 1, 100, 1e-3)
    # the reference's own golden numbers, jolideco/tests/test_core.py:71-79
    assert_allclose(flux_up[12, 12], 1.542659, rtol=1e-3)
    assert_allclose(flux_up[0, 0], 3.927929, rtol=1e-3)
    assert_allclose(trace[-1]["total"], 5.842237, rtol=1e-3)
    assert_allclose(trace[-1]["datasets"], [1.956523, 1.945902, 1.939812], rtol=1e-3)


def test_run_upsampling2_reference_e2e_golden():
    g, flux_up, trace = _run("run_upsampling2.npz", 2, 100, 1e-3)
    flux = O.sum_pool(flux_up, 2)
    # jolideco/tests/test_core.py:99-124
    assert flux_up.shape == (64, 64)
    assert_allclose(flux[12, 12], 3.565998, rtol=1e-3)
    assert_allclose(flux[0, 0], 1.605782, rtol=1e-3)
    assert_allclose(trace[-1]["total"], 5.844786, rtol=1e-3)
    assert_allclose(trace[-1]["datasets"], [1.946759, 1.958015, 1.940012], rtol=1e-3)


@pytest.mark.parametrize("name,f,n", [("run_gmm_max.npz", 1, 8), ("run_gmm_lse.npz", 1, 8), ("run_gmm_up2.npz", 2, 6)])
def test_run_gmm_prior_matches_imported_reference(name, f, n):
    g, flux_up, trace = _run(name, f, n, 1e-3, gmm=True)
    assert_allclose([t["priors-total"] for t in trace], g["trace_prior"], rtol=1e-4)


@pytest.mark.parametrize("name,f,n", [("run_gmm_max.npz", 1, 8), ("run_gmm_up2.npz", 2, 6), ("run_uniform.npz", 1, 20)])
def test_torch_port_matches_imported_reference(name, f, n):
    """The torch-CPU port timed by bench.py as the CPU baseline reproduces the reference's runs."""
    from oracle import torch_port as T

    g = load_golden(name)
    datasets = [T.Dataset(d, f) for d in unpack_datasets(g)]
    gmm = None
    if "gmm_means" in g:
        og = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"])
        gmm = T.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], og.pixel_weights)
    loop = T.MapLoop(g["flux_init_up"], datasets, gmm, marginalize=bool(g.get("marginalize", False)))
    step = 0
    for epoch in range(n):
        for i in range(len(datasets)):
            loop.step(i, g["shifts"][step] if gmm is not None else None)
            step += 1
        tr = loop.trace(g["trace_shifts"][epoch] if gmm is not None else None)
        assert_allclose(tr["total"], g["trace_total"][epoch], rtol=1e-5)
    if n == len(g["trace_total"]):
        assert_allclose(loop.flux_numpy(), g["flux_up"], rtol=1e-4)


def test_run_with_background_norm_calibration_matches_imported_reference():
    """NPredCalibrations with trainable background norms (shifts 0): flux, trace and fitted norms."""
    g = load_golden("run_gmm_calib.npz")
    datasets = [O.prepare_dataset(d, f=1) for d in unpack_datasets(g)]
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"])
    flux_up, trace, norms = O.map_run(g["flux_init_up"], datasets, 8, gmm=gmm, shifts=g["shifts"],
                                      trace_shifts=g["trace_shifts"], background_norms=g["background_norm_init"])
    assert np.linalg.norm(flux_up - g["flux_up"]) / np.linalg.norm(g["flux_up"]) < 1e-3
    assert_allclose([t["total"] for t in trace], g["trace_total"], rtol=1e-5)
    assert_allclose(norms, g["background_norm"], rtol=1e-5)


def test_shift_image_kat_against_imported_reference():
    """Sub-pixel shift of the flux (utils/torch.py:196-223): values, image gradient and shift gradient of the
    4-tap restatement against `shift_image_torch` + autograd, fp64 to round-off and fp32 to its own noise."""
    g = load_golden("shift_kat.npz")
    image, cot = g["image"], g["cot"]
    for i, (sx, sy, scale) in enumerate(g["cases"]):
        out, d_dy, d_dx = O.shift_image(image, sy, sx, scale, return_grads=True)
        assert_allclose(out, g[f"c{i}_f64_out"], rtol=1e-9, atol=1e-12)
        assert_allclose(O.shift_image_adjoint(cot, sy, sx, scale), g[f"c{i}_f64_dimage"], rtol=1e-9, atol=1e-12)
        assert_allclose([(cot * d_dx).sum(), (cot * d_dy).sum()], g[f"c{i}_f64_dshift_xy"], rtol=1e-9)
        out32 = O.shift_image(image.astype(np.float32), sy, sx, scale)
        assert np.abs(out32 - g[f"c{i}_f32_out"]).max() <= 2e-5 * np.abs(out32).max()
    assert_allclose(O.shift_image(image, -1.0, 2.0, 1), g["whole_out"], rtol=1e-12)
    assert np.array_equal(g["zero_is_identity"], image) and O.shift_is_identity(0.0, 0.0)
    # the shifted image leaves zeros where it samples outside (zeros padding)
    assert np.all(O.shift_image(image, 0.0, 3.0, 1)[:, -3:] == 0)


@pytest.mark.parametrize("name,f", [("run_gmm_shift.npz", 1), ("run_gmm_shift_up2.npz", 2)])
def test_run_with_shift_calibration_matches_imported_reference(name, f):
    """NPredCalibrations with trainable non-zero sub-pixel shifts AND background norms: flux, trace, fitted
    norms and fitted shifts of the restated loop against the imported reference."""
    g = load_golden(name)
    datasets = [O.prepare_dataset(d, f=f) for d in unpack_datasets(g)]
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"])
    flux_up, trace, norms, shifts_xy = O.map_run(g["flux_init_up"], datasets, 6, gmm=gmm, shifts=g["shifts"],
                                                 trace_shifts=g["trace_shifts"],
                                                 background_norms=g["background_norm_init"],
                                                 shifts_xy=g["shift_xy_init"])
    assert np.linalg.norm(flux_up - g["flux_up"]) / np.linalg.norm(g["flux_up"]) < 1e-3
    assert_allclose([t["total"] for t in trace], g["trace_total"], rtol=2e-5)
    assert_allclose(norms, g["background_norm"], rtol=1e-4)
    assert_allclose(np.stack(shifts_xy), g["shift_xy"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("tag,marginalize", [("max", False), ("lse", True)])
def test_joint_objective_matches_imported_reference(tag, marginalize):
    """The objective of `mode="joint"` (one Adam step on sum_d L_d - beta * prior, sharded over GPUs): value against
    the reference's `TotalLoss.__call__` (loss.py:257-261), gradient against autograd through the reference's own
    components (its `__call__` detaches the dataset terms, loss.py:71)."""
    g = load_golden("joint_objective.npz")
    datasets = [O.prepare_dataset(d, f=1) for d in unpack_datasets(g)]
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"])
    flux = g["flux"].astype(np.float32)
    theta = np.log(flux)
    total, dtheta = O.joint_loss_and_grad(theta, datasets, float(g["beta"]), gmm, g[f"{tag}_shift"], 4, marginalize)
    assert_allclose(total, g[f"{tag}_total"], rtol=2e-6)
    assert_allclose(g[f"{tag}_total"], g[f"{tag}_total_components"], rtol=1e-6)
    dflux = dtheta / flux  # the oracle differentiates w.r.t. theta = log flux
    ref = g[f"{tag}_dflux"]
    assert np.abs(dflux - ref).max() <= 2e-5 * np.abs(ref).max()
    # shards (datasets dealt to 2 ranks, prior rows split) add up to the whole: what the all-reduce relies on
    ny = (flux.shape[0] - 8) // 4 + 1
    parts = [O.joint_loss_and_grad(theta, datasets, float(g["beta"]), gmm, g[f"{tag}_shift"], 4, marginalize,
                                   dataset_index=idx, rows=rows)
             for idx, rows in (([0, 2], (0, ny // 2)), ([1], (ny // 2, ny)))]
    assert_allclose(sum(p[0] for p in parts), total, rtol=1e-6)
    assert np.abs(sum(p[1] for p in parts) - dtheta).max() <= 1e-6 * np.abs(dtheta).max()


@pytest.mark.parametrize("marginalize", [False, True])
@pytest.mark.parametrize("rows", [None, (2, 7)])
def test_lean_prior_oracle_equals_the_pinned_one(marginalize, rows):
    """`gmm_patch_prior_lean` (two passes over the components; what the BASELINE-size GPU tests compare with) is the
    same function as `gmm_patch_prior`, which the reference goldens pin."""
    rng = np.random.default_rng(0)
    K = 7
    A = rng.normal(0, 0.05, size=(K, 64, 64))
    gmm = O.GMM(rng.normal(0, 0.01, size=(K, 64)), A @ A.transpose(0, 2, 1) + 0.01 * np.eye(64), np.full(K, 1 / K),
                dtype=np.float64)
    flux = rng.gamma(2.0, size=(45, 52))
    r0, r1 = (None, None) if rows is None else rows
    prior, dflux, k = O.gmm_patch_prior(flux, gmm, 2, -1, 4, marginalize, True, r0, r1)
    res = O.gmm_patch_prior_lean(flux, gmm, 2, -1, 4, marginalize, r0, r1)
    assert abs(prior - res["prior"]) <= 1e-15 * abs(prior)
    assert np.abs(dflux - res["dflux"]).max() <= 1e-14 * np.abs(dflux).max()
    assert np.array_equal(k, res["argmax"]) and (res["gap"] >= 0).all()
