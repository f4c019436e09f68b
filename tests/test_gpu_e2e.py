"""End-to-end parity of the drop-in API (MAPDeconvolver.run, loss classes, priors) on B200 against
the reference's golden values and the runs of the imported reference stored in tests/golden/."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from conftest import load_golden, unpack_datasets
from oracle import jolideco_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import jolideco_b200 as J

DEV = "cuda"


def as_datasets(g):
    return {str(i): d for i, d in enumerate(unpack_datasets(g))}


def make_prior(g, seed, backend=None):
    gmm = J.GaussianMixtureModel.from_numpy(g["gmm_means"], g["gmm_cov"], g["gmm_w"],
                                            meta=J.GaussianMixtureModelMeta(stride=4))
    gen = torch.Generator().manual_seed(seed)
    return J.GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=bool(g["marginalize"]), backend=backend)


def run(g, f, n_epochs, prior, fused=True, graph=True):
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=f, prior=prior)
    deco = J.MAPDeconvolver(n_epochs=n_epochs, learning_rate=0.1, display_progress=False, device=DEV, fused=fused,
                            use_cuda_graph=graph)
    return deco.run(datasets=as_datasets(g), components=comps)


def check(res, g, n_epochs, rtol_flux=1e-3, rtol_trace=2e-5):
    flux_up = res.flux_upsampled_total
    rel = np.linalg.norm(flux_up - g["flux_up"]) / np.linalg.norm(g["flux_up"])
    assert rel < rtol_flux, rel
    tr = res.trace_loss
    assert len(tr) == n_epochs
    assert_allclose(tr["total"], g["trace_total"], rtol=rtol_trace)
    for i in range(g["trace_datasets"].shape[1]):
        assert_allclose(tr[f"dataset-{i}"], g["trace_datasets"][:, i], rtol=rtol_trace)
    assert_allclose(tr["priors-total"], g["trace_prior"], rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("fused,graph", [(True, True), (True, False), (False, False)])
def test_reference_e2e_golden_uniform(fused, graph):
    g = load_golden("run_uniform.npz")
    res = run(g, 1, 100, J.UniformPrior(), fused, graph)
    check(res, g, 100)
    # the reference's own golden numbers (jolideco/tests/test_core.py:71-79)
    assert_allclose(res.flux_total[12, 12], 1.542659, rtol=1e-3)
    assert_allclose(res.flux_total[0, 0], 3.927929, rtol=1e-3)
    t = res.trace_loss[-1]
    assert_allclose(t["total"], 5.842237, rtol=1e-3)
    assert_allclose([t["dataset-0"], t["dataset-1"], t["dataset-2"]], [1.956523, 1.945902, 1.939812], rtol=1e-3)


def test_reference_e2e_golden_upsampling2():
    g = load_golden("run_upsampling2.npz")
    res = run(g, 2, 100, J.UniformPrior())
    check(res, g, 100)
    # jolideco/tests/test_core.py:99-124
    assert res.flux_upsampled_total.shape == (64, 64)
    assert res.components["flux-1"].upsampling_factor == 2
    assert_allclose(res.flux_total[12, 12], 3.565998, rtol=1e-3)
    assert_allclose(res.flux_total[0, 0], 1.605782, rtol=1e-3)
    assert_allclose(res.trace_loss[-1]["total"], 5.844786, rtol=1e-3)


@pytest.mark.parametrize("name,f,n,seed", [("run_gmm_max.npz", 1, 8, 4), ("run_gmm_lse.npz", 1, 8, 4),
                                           ("run_gmm_up2.npz", 2, 6, 5)])
@pytest.mark.parametrize("fused", [True, False])
def test_gmm_prior_run_matches_imported_reference(name, f, n, seed, fused):
    g = load_golden(name)
    prior = make_prior(g, seed)
    res = run(g, f, n, prior, fused=fused, graph=fused)
    check(res, g, n)


def test_seeded_generator_reproduces_reference_shifts():
    g = load_golden("run_gmm_max.npz")
    prior = make_prior(g, 4)
    draws = [prior.draw_shifts() for _ in range(6)]
    # consumption order: D training draws then 1 trace draw per epoch (D = 2)
    expect = [tuple(g["shifts"][0]), tuple(g["shifts"][1]), tuple(g["trace_shifts"][0]),
              tuple(g["shifts"][2]), tuple(g["shifts"][3]), tuple(g["trace_shifts"][1])]
    assert [tuple(int(v) for v in d) for d in draws] == [tuple(int(v) for v in e) for e in expect]


def test_per_iteration_loss_and_gradient_vs_oracle_autograd_api():
    """One reference iteration through the class API: loss and theta.grad within 1e-5 of the oracle."""
    g = load_golden("run_gmm_max.npz")
    prior = make_prior(g, 4)
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=prior)
    comps = comps.to(DEV)
    datasets = as_datasets(g)
    total = J.TotalLoss.from_datasets_and_components(datasets=datasets, components=comps, beta=1.0, device=DEV)
    counts, npred_model = next(iter(total.poisson_loss.iter_by_dataset))
    fluxes = comps.to_flux_tuple()
    npred = npred_model.evaluate(fluxes=fluxes)
    loss = total.poisson_loss.loss_function(npred, counts)
    sh = tuple(int(v) for v in g["shifts"][0])
    loss_prior = prior(flux=fluxes[0], shift_yx=sh)
    loss_total = loss - 1.0 * loss_prior / total.prior_weight
    loss_total.backward()
    grad = comps["flux-1"]._flux_upsampled.grad.cpu().numpy()[0, 0]
    # oracle in float64
    ds = O.prepare_dataset(unpack_datasets(g)[0], f=1, dtype=np.float64)
    theta = np.log(g["flux_init_up"].astype(np.float64))
    l_ref, dth_ref, _ = O.dataset_loss_and_grad(theta, ds)
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], dtype=np.float64)
    p_ref, dfl_ref, _ = O.gmm_patch_prior(np.exp(theta), gmm, sh[0], sh[1], return_grad=True)
    g_ref = dth_ref - 0.5 * dfl_ref * np.exp(theta)
    assert_allclose(loss.item(), l_ref, rtol=1e-5)
    assert_allclose(loss_prior.item(), p_ref, rtol=1e-5)
    assert np.abs(grad - g_ref).max() <= 1e-5 * np.abs(g_ref).max()


def test_cpu_device_is_refused():
    with pytest.raises(J.JolidecoB200Error):
        J.MAPDeconvolver(device="cpu")


def test_bad_arguments_match_reference_errors():
    with pytest.raises(ValueError):
        J.MAPDeconvolver(optimizer_type="lbfgs", device=DEV)
    with pytest.raises(ValueError):
        J.MAPDeconvolver(stop_early=True, device=DEV).run(datasets={}, components=None)
    with pytest.raises(ValueError):
        J.SpatialFluxComponent(flux_upsampled=torch.ones(4, 4))


def test_batched_independent_runs_equal_single_runs():
    """run_many (runs interleaved on CUDA streams) gives exactly the result of running each job alone (with the same
    prior kernel: batched runs take the one-tile-per-CTA tcgen05 kernel, backend 1)."""
    g = load_golden("run_gmm_max.npz")
    jobs, singles = [], []
    for seed in (4, 5, 6):
        for store in (jobs, singles):
            comps = J.FluxComponents()
            comps["flux-1"] = J.SpatialFluxComponent.from_numpy(
                flux=g["flux_init"] * (1 + 0.1 * seed), upsampling_factor=1,
                prior=make_prior(g, seed, backend=1 if store is singles else None))
            store.append(dict(datasets=as_datasets(g), components=comps))
    res = J.run_many(jobs, n_epochs=5, n_streams=2, device=DEV)
    assert sorted(res) == [0, 1, 2]
    for j, job in enumerate(singles):
        ref = J.MAPDeconvolver(n_epochs=5, display_progress=False, device=DEV).run(**job)
        assert np.array_equal(res[j].flux_upsampled_total, ref.flux_upsampled_total)
        assert_allclose(res[j].trace_loss["total"], ref.trace_loss["total"], rtol=1e-12)
    # dealing jobs to ranks: rank 1 of 2 gets job 1 only
    part = J.run_many(jobs, n_epochs=1, rank=1, world=2, device=DEV)
    assert sorted(part) == [1]


@pytest.mark.parametrize("fused", [True, False])
def test_background_norm_calibration_matches_imported_reference(fused):
    """NPredCalibrations with trainable background norms: fused engine (scalar Adam kernel) and autograd path."""
    g = load_golden("run_gmm_calib.npz")
    prior = make_prior(g, 6)
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=prior)
    cals = J.NPredCalibrations()
    for name, b in zip(as_datasets(g), g["background_norm_init"]):
        cals[name] = J.NPredCalibration(background_norm=float(b))
    deco = J.MAPDeconvolver(n_epochs=8, learning_rate=0.1, display_progress=False, device=DEV, fused=fused)
    res = deco.run(datasets=as_datasets(g), components=comps, calibrations=cals)
    assert hasattr(deco, "engine") == fused
    check(res, g, 8)
    norms = [float(c.background_norm) for c in res.calibrations.values()]
    assert_allclose(norms, g["background_norm"], rtol=1e-4)


@pytest.mark.parametrize("name,f", [("run_gmm_shift.npz", 1), ("run_gmm_shift_up2.npz", 2)])
def test_shift_calibration_matches_imported_reference(name, f):
    """trainable non-zero sub-pixel shifts + background norms.  Not covered by the fused engine:
    `MAPDeconvolver.run` must route it to the autograd path (grid_sample + the CUDA kernels) and still reproduce
    the imported reference's run (flux, trace, fitted norms and shifts)."""
    g = load_golden(name)
    prior = make_prior(g, 8)
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=f, prior=prior)
    cals = J.NPredCalibrations()
    for ds_name, b, (sx, sy) in zip(as_datasets(g), g["background_norm_init"], g["shift_xy_init"]):
        cals[ds_name] = J.NPredCalibration(shift_x=float(sx), shift_y=float(sy), background_norm=float(b))
    deco = J.MAPDeconvolver(n_epochs=6, learning_rate=0.1, display_progress=False, device=DEV, fused=False)
    res = deco.run(datasets=as_datasets(g), components=comps, calibrations=cals)
    assert not hasattr(deco, "engine")  # autograd path: grid_sample + the CUDA kernels
    check(res, g, 6)
    assert_allclose([float(c.background_norm) for c in res.calibrations.values()], g["background_norm"], rtol=1e-4)
    shifts = np.stack([c.shift_xy.detach().cpu().numpy()[0] for c in res.calibrations.values()])
    assert_allclose(shifts, g["shift_xy"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("name,f", [("run_gmm_shift.npz", 1), ("run_gmm_shift_up2.npz", 2)])
def test_shift_calibration_fused_engine_matches_imported_reference(monkeypatch, name, f):
    """The same runs through the fused engine (jd_shift_forward / jd_shift_backward + scalar Adam on the shift pair),
    the default route."""
    monkeypatch.delenv("JD_FUSED_SHIFT", raising=False)
    g = load_golden(name)
    prior = make_prior(g, 8)
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=f, prior=prior)
    cals = J.NPredCalibrations()
    for ds_name, b, (sx, sy) in zip(as_datasets(g), g["background_norm_init"], g["shift_xy_init"]):
        cals[ds_name] = J.NPredCalibration(shift_x=float(sx), shift_y=float(sy), background_norm=float(b))
    deco = J.MAPDeconvolver(n_epochs=6, learning_rate=0.1, display_progress=False, device=DEV)
    res = deco.run(datasets=as_datasets(g), components=comps, calibrations=cals)
    assert hasattr(deco, "engine")
    # The fused shift is the exact 4-tap stencil; the reference goes through affine_grid + grid_sample, whose float32
    # normalise / unnormalise of the pixel coordinates perturbs the bilinear weights by ~1e-7 * grid size.  Adam on the
    # two scalar shift parameters amplifies that into a 2.4e-5 trace difference on the 48 x 48 (f = 2) grid after the
    # first epoch (the first epoch agrees to 2e-7); fitted norms and shifts below agree to 1e-4 / 1e-3.
    check(res, g, 6, rtol_trace=2e-5 if f == 1 else 1e-4)
    assert_allclose([float(c.background_norm) for c in res.calibrations.values()], g["background_norm"], rtol=1e-4)
    shifts = np.stack([c.shift_xy.detach().cpu().numpy()[0] for c in res.calibrations.values()])
    assert_allclose(shifts, g["shift_xy"], rtol=1e-3, atol=1e-4)


def test_shift_kernels_match_oracle():
    """jd_shift_forward / jd_shift_backward against the oracle's 4-tap restatement (pinned to shift_image_torch)."""
    from jolideco_b200 import _lib

    g = load_golden("shift_kat.npz")
    image, cot = g["image"].astype(np.float32), g["cot"].astype(np.float32)
    H, W = image.shape
    dev = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(DEV)  # noqa: E731
    s = torch.cuda.current_stream().cuda_stream
    for sx, sy, scale in g["cases"]:
        shift = dev(np.array([sx, sy]))
        img, d = dev(image), dev(cot)
        out, dflux = torch.empty_like(img), torch.ones_like(img)
        dshift = torch.zeros(2, dtype=torch.float64, device=DEV)
        _lib.call("jd_shift_forward", img.data_ptr(), shift.data_ptr(), int(scale), H, W, out.data_ptr(), s)
        _lib.call("jd_shift_backward", d.data_ptr(), img.data_ptr(), shift.data_ptr(), int(scale), H, W, dflux.data_ptr(), 1,
                  dshift.data_ptr(), s)
        ref, d_dy, d_dx = O.shift_image(image.astype(np.float64), sy, sx, int(scale), return_grads=True)
        assert np.abs(out.cpu().numpy() - ref).max() <= 3e-6 * np.abs(ref).max()
        adj = O.shift_image_adjoint(cot.astype(np.float64), sy, sx, int(scale)) + 1
        assert np.abs(dflux.cpu().numpy() - adj).max() <= 3e-6 * np.abs(adj).max()
        assert_allclose(dshift.cpu().numpy(), [(cot * d_dx).sum(), (cot * d_dy).sum()], rtol=2e-5)


# ---------------------------------------------------------------------------------------------------------------
# validation datasets + early stopping (core.py:251-261, loss.py:244-248); goldens from the imported
# reference (oracle/make_golden.py::golden_validation_early_stop)
# ---------------------------------------------------------------------------------------------------------------
def _validation_run(g, fused, **kwargs):
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=J.UniformPrior())
    datasets = {str(i): d for i, d in enumerate(unpack_datasets(g))}
    validation = {"2": unpack_datasets(g, "dv")[0]}
    deco = J.MAPDeconvolver(display_progress=False, device=DEV, fused=fused, **kwargs)
    res = deco.run(datasets=datasets, components=comps, datasets_validation=validation)
    assert hasattr(deco, "engine") == fused
    return res


@pytest.mark.parametrize("fused", [True, False])
def test_validation_datasets_trace_matches_imported_reference(fused):
    g = load_golden("run_validation.npz")
    res = _validation_run(g, fused, n_epochs=30, learning_rate=0.1)
    tr = res.trace_loss
    assert len(tr) == 30 == int(g["a_n_epochs_run"])
    assert_allclose(tr["datasets-validation-total"], g["a_trace_validation"], rtol=2e-5)
    assert_allclose(tr["total"], g["a_trace_total"], rtol=2e-5)
    for i in range(2):
        assert_allclose(tr[f"dataset-{i}"], g["a_trace_datasets"][:, i], rtol=2e-5)
    rel = np.linalg.norm(res.flux_upsampled_total - g["a_flux_up"]) / np.linalg.norm(g["a_flux_up"])
    assert rel < 1e-3


@pytest.mark.parametrize("fused", [True, False])
def test_early_stopping_stops_where_the_reference_stops(fused):
    """stop_early with a 5-epoch running mean: the reference ends after 16 of 100 epochs on this input."""
    g = load_golden("run_validation.npz")
    res = _validation_run(g, fused, n_epochs=100, learning_rate=0.3, stop_early=True, stop_early_n_average=5)
    tr = res.trace_loss
    assert len(tr) == int(g["b_n_epochs_run"]) == 16
    assert_allclose(tr["datasets-validation-total"], g["b_trace_validation"], rtol=5e-5)
    assert_allclose(tr["total"], g["b_trace_total"], rtol=5e-5)
    rel = np.linalg.norm(res.flux_upsampled_total - g["b_flux_up"]) / np.linalg.norm(g["b_flux_up"])
    assert rel < 1e-3


def test_early_stop_leaves_the_prior_generator_where_the_reference_does():
    """the engine pre-draws n_epochs x (D + 1) cycle-spin shifts; after an early stop the generator must
    sit exactly behind the draws that were consumed (2 randint calls per evaluation of the prior)."""
    g = load_golden("run_validation.npz")
    gmm_g = load_golden("run_gmm_max.npz")
    gmm = J.GaussianMixtureModel.from_numpy(gmm_g["gmm_means"], gmm_g["gmm_cov"], gmm_g["gmm_w"],
                                            meta=J.GaussianMixtureModelMeta(stride=4))
    gen = torch.Generator().manual_seed(21)
    expect = torch.Generator().manual_seed(21)
    prior = J.GMMPatchPrior(gmm=gmm, stride=4, generator=gen)
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=prior)
    datasets = {str(i): d for i, d in enumerate(unpack_datasets(g))}
    # a 2-epoch average stops at the first increase of the validation loss; learning rate 0.5 makes it oscillate early
    deco = J.MAPDeconvolver(n_epochs=200, learning_rate=0.5, stop_early=True, stop_early_n_average=2,
                            display_progress=False, device=DEV)
    res = deco.run(datasets=datasets, components=comps, datasets_validation={"2": unpack_datasets(g, "dv")[0]})
    n = len(res.trace_loss)
    assert 2 < n < 200
    for _ in range(n * (len(datasets) + 1)):  # D training draws + 1 trace draw per completed epoch
        torch.randint(-2, 3, (1,), generator=expect)
        torch.randint(-2, 3, (1,), generator=expect)
    assert torch.equal(gen.get_state(), expect.get_state())


# ---------------------------------------------------------------------------------------------------------------
# SURVEY 8f row 4 / reference tests/test_core.py:191-220: GMM patch prior behind a non-identity, TRAINABLE image norm
# (ASinhImageNorm) with upsampling 2 - the CUDA prior op inside torch autograd, norm parameters in the optimiser
# ---------------------------------------------------------------------------------------------------------------
def test_gmm_prior_with_asinh_norm_matches_imported_reference(tmp_path):
    g = load_golden("run_gmm_asinh.npz")
    gmm = J.GaussianMixtureModel.from_numpy(g["gmm_means"], g["gmm_cov"], g["gmm_w"],
                                            meta=J.GaussianMixtureModelMeta(stride=4))
    norm = J.ASinhImageNorm()
    prior = J.GMMPatchPrior(gmm=gmm, stride=4, generator=torch.Generator().manual_seed(13), norm=norm)
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=2, prior=prior)
    deco = J.MAPDeconvolver(n_epochs=6, learning_rate=0.1, display_progress=False, device=DEV, checkpoint_path=tmp_path)
    res = deco.run(datasets=as_datasets(g), components=comps)
    assert not hasattr(deco, "engine")  # non-identity norm: the reference loop on the autograd bindings
    assert res.flux_upsampled_total.shape == (64, 64)
    check(res, g, 6, rtol_trace=5e-5)
    final = [float(res.components["flux-1"].prior.norm.alpha), float(res.components["flux-1"].prior.norm.beta)]
    assert_allclose(final, g["norm_final"], rtol=1e-3)
    # per-epoch checkpoints (core.py:234-243) are results that read back
    back = J.MAPDeconvolverResult.read(tmp_path / res.trace_loss["filename"][-1])
    assert back.flux_upsampled_total.shape == (64, 64)
    assert_allclose(back.flux_upsampled_total, res.flux_upsampled_total, rtol=1e-6)


def test_fused_engine_checkpoints_hold_the_current_flux(tmp_path):
    """checkpoints written by the fused engine carry the flux of THAT epoch (not the initial one) together
    with the trace so far."""
    g = load_golden("run_gmm_max.npz")
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=make_prior(g, 4))
    deco = J.MAPDeconvolver(n_epochs=3, learning_rate=0.1, display_progress=False, device=DEV, checkpoint_path=tmp_path)
    res = deco.run(datasets=as_datasets(g), components=comps)
    assert hasattr(deco, "engine")
    first = res.read_checkpoint(0).flux_upsampled_total
    last = res.read_checkpoint(2).flux_upsampled_total
    assert not np.allclose(first, g["flux_init_up"]) and not np.allclose(first, last)
    assert_allclose(last, res.flux_upsampled_total, rtol=1e-6)


@pytest.mark.parametrize("f", [1, 2])
def test_npred_setup_on_the_gpu_equals_the_host_setup(f):
    """NPredModel.from_numpy(device='cuda') (upload, bilinear upsampling, PSF / f^2, edge correction through the
    library's convolution) against the host path, which tests/test_host_api.py pins to the reference (kat.npz)."""
    g = load_golden("kat.npz")
    ds = {k[len("npred_ds_"):]: g[k] for k in g if k.startswith("npred_ds_")}
    host = J.NPredModel.from_numpy(ds["exposure"], ds["psf"], upsampling_factor=f)
    dev = J.NPredModel.from_numpy(ds["exposure"], ds["psf"], upsampling_factor=f, device=DEV)
    assert dev.exposure.is_cuda and dev.psf.is_cuda and dev.exposure.shape == host.exposure.shape
    assert_allclose(dev.psf.cpu().numpy(), host.psf.numpy(), rtol=1e-6, atol=1e-12)
    assert_allclose(dev.exposure.cpu().numpy(), host.exposure.numpy(), rtol=2e-6)
    if f == 2:
        assert_allclose(dev.exposure.cpu().numpy()[0, 0], g["npred_exposure_up"], rtol=2e-6)


@pytest.mark.parametrize("name,f,n,seed", [("run_gmm_max.npz", 1, 8, 4), ("run_gmm_lse.npz", 1, 8, 4),
                                           ("run_gmm_up2.npz", 2, 6, 5)])
def test_goldens_through_the_batched_likelihood_kernels(monkeypatch, name, f, n, seed):
    """The golden images are a single 64 x 64 tile, which the engine would route to the separate conv / Poisson kernels
    (engine.LIK_MIN_CTAS); forced through jd_likelihood_forward / _backward they reproduce the reference run too."""
    from jolideco_b200 import engine as E

    monkeypatch.setattr(E, "LIK_MIN_CTAS", 0)
    seen = []
    real = E._lib.call
    monkeypatch.setattr(E._lib, "call", lambda name_, *a: (seen.append(name_), real(name_, *a))[1])
    g = load_golden(name)
    res = run(g, f, n, make_prior(g, seed), fused=True, graph=False)
    check(res, g, n)
    # (jd_conv_forward_direct still appears once per dataset: the exposure edge correction of the setup)
    assert "jd_likelihood_forward" in seen and "jd_likelihood_backward" in seen
    assert "jd_poisson_forward_backward" not in seen and "jd_conv_backward_direct" not in seen


@pytest.mark.parametrize("name,f,n,seed", [("run_gmm_max.npz", 1, 8, 4), ("run_gmm_lse.npz", 1, 8, 4),
                                           ("run_gmm_up2.npz", 2, 6, 5)])
@pytest.mark.parametrize("graph", [True, False])
def test_goldens_with_the_prior_forward_beside_the_likelihood_chain(monkeypatch, name, f, n, seed, graph):
    """One-dataset steps with the prior forward on part of the SM pairs (here forced to one pair; normally tuned at
    warm-up) and the likelihood chain on the side stream reproduce the reference run: the two chains only share the
    flux."""
    from jolideco_b200 import engine as E

    monkeypatch.setenv("JD_SPLIT_CLUSTERS", "1")
    seen = []
    real = E._lib.call
    monkeypatch.setattr(E._lib, "call", lambda name_, *a: (seen.append(name_), real(name_, *a))[1])
    g = load_golden(name)
    res = run(g, f, n, make_prior(g, seed), fused=True, graph=graph)
    check(res, g, n)
    assert "jd_gmm_prior_forward_tcx2_on" in seen


def test_calibration_goldens_through_the_batched_likelihood_kernels(monkeypatch):
    from jolideco_b200 import engine as E

    monkeypatch.setattr(E, "LIK_MIN_CTAS", 0)
    g = load_golden("run_gmm_calib.npz")
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=make_prior(g, 6))
    cals = J.NPredCalibrations()
    for name, b in zip(as_datasets(g), g["background_norm_init"]):
        cals[name] = J.NPredCalibration(background_norm=float(b))
    res = J.MAPDeconvolver(n_epochs=8, learning_rate=0.1, display_progress=False, device=DEV).run(
        datasets=as_datasets(g), components=comps, calibrations=cals)
    check(res, g, 8)
    assert_allclose([float(c.background_norm) for c in res.calibrations.values()], g["background_norm"], rtol=1e-4)
    g = load_golden("run_gmm_shift.npz")
    comps = J.FluxComponents()
    comps["flux-1"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=make_prior(g, 8))
    cals = J.NPredCalibrations()
    for ds_name, b, (sx, sy) in zip(as_datasets(g), g["background_norm_init"], g["shift_xy_init"]):
        cals[ds_name] = J.NPredCalibration(shift_x=float(sx), shift_y=float(sy), background_norm=float(b))
    res = J.MAPDeconvolver(n_epochs=6, learning_rate=0.1, display_progress=False, device=DEV).run(
        datasets=as_datasets(g), components=comps, calibrations=cals)
    check(res, g, 6)
    shifts = np.stack([c.shift_xy.detach().cpu().numpy()[0] for c in res.calibrations.values()])
    assert_allclose(shifts, g["shift_xy"], rtol=1e-3, atol=1e-4)
