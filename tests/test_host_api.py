"""CPU checks of the host-side mirror of the reference interface (no kernel is launched): constructor arguments and
error behaviour (jolideco/core.py:79-112, 173-174; models/core.py:394-409), setup-time numpy helpers against the
oracle, the cycle-spin draw order (utils/torch.py:108-116), the trace table, and the refusal to run without CUDA."""
import pickle

import numpy as np
import pytest
import torch

import jolideco_b200 as J
from jolideco_b200 import priors
from jolideco_b200.table import TraceTable
from oracle import jolideco_oracle as O


def small_gmm(K=5, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.normal(0, 0.05, size=(K, 64, 64))
    cov = A @ A.transpose(0, 2, 1) + 0.01 * np.eye(64)
    w = rng.uniform(0.5, 1.5, size=K)
    return rng.normal(0, 0.01, size=(K, 64)), cov, w / w.sum()


def test_deconvolver_arguments_and_errors():
    deco = J.MAPDeconvolver(n_epochs=7, beta=0.5, learning_rate=0.2, display_progress=False)
    assert deco.n_epochs == 7 and deco.beta == 0.5 and deco.optimizer_kwargs["lr"] == 0.2
    assert "MAPDeconvolver" in str(deco) and deco.to_dict()["device"].startswith("cuda")
    with pytest.raises(ValueError, match="Unknown optimizer"):
        J.MAPDeconvolver(optimizer_type="lbfgs")
    with pytest.raises(ValueError, match="Unknown mode"):
        J.MAPDeconvolver(mode="sideways")
    with pytest.raises(J.JolidecoB200Error, match="no CPU fallback"):
        J.MAPDeconvolver(device="cpu")  # the reference degrades to CPU with a warning; this backend refuses
    with pytest.raises(ValueError, match="Early stopping requires"):
        J.MAPDeconvolver(stop_early=True).run(datasets={}, components=None)


def test_run_refuses_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    comp = J.SpatialFluxComponent.from_numpy(flux=np.ones((16, 16)), prior=J.UniformPrior())
    ds = dict(counts=np.ones((16, 16)), psf=np.ones((3, 3)) / 9, exposure=np.ones((16, 16)), background=np.ones((16, 16)))
    with pytest.raises(J.JolidecoB200Error):
        J.MAPDeconvolver(n_epochs=1, display_progress=False).run(datasets={"a": ds}, components=comp)


def test_flux_component_shapes_and_log_parameterisation():
    flux = np.random.default_rng(1).gamma(2.0, size=(12, 10))
    comp = J.SpatialFluxComponent.from_numpy(flux=flux, upsampling_factor=2, prior=J.UniformPrior())
    assert comp.upsampling_factor == 2 and tuple(comp._flux_upsampled.shape) == (1, 1, 24, 20)
    assert comp.use_log_flux and not comp.frozen
    # theta stores log(flux_init) (models/core.py:399-402)
    up = torch.nn.functional.interpolate(torch.from_numpy(flux[None, None].astype(np.float32)), scale_factor=2,
                                         mode="bilinear")
    np.testing.assert_allclose(comp._flux_upsampled.detach().numpy(), np.log(up.numpy()), rtol=1e-6)
    with pytest.raises(ValueError):
        J.SpatialFluxComponent(flux_upsampled=torch.ones(4, 4))  # 4-D enforced (models/core.py:394-397)
    comps = J.FluxComponents()
    comps["a"] = comp
    assert list(comps) == ["a"] and len(list(comps.parameters())) == 1


def test_gmm_setup_helpers_match_oracle():
    means, cov, w = small_gmm()
    gmm = J.GaussianMixtureModel.from_numpy(means, cov, w, meta=J.GaussianMixtureModelMeta(stride=4))
    assert gmm.n_components == 5 and gmm.n_features == 64 and gmm.patch_shape == (8, 8)
    np.testing.assert_allclose(priors.compute_precision_cholesky(cov), O.compute_precision_cholesky(cov), rtol=1e-12)
    L = gmm.precisions_cholesky.numpy()
    assert np.all(np.tril(L, -1) == 0)  # upper triangular: what the trimmed MMA schedule and the tri backward rely on
    pw = priors.get_pixel_weights((8, 8), 4)
    np.testing.assert_allclose(pw, O.get_pixel_weights(8, 4), rtol=1e-12)
    np.testing.assert_allclose(pw.sum(), 16.0, rtol=1e-12)
    t = np.array([.125, .375, .625, .875, .875, .625, .375, .125])
    np.testing.assert_allclose(pw, np.outer(t, t), rtol=1e-12)
    ones = J.GaussianMixtureModel.from_numpy(means, cov, w)  # meta.stride None -> all-ones weights (gmm.py:291-299)
    assert np.all(ones.pixel_weights_numpy == 1)
    ref = O.GMM(means, cov, w)
    np.testing.assert_allclose(gmm.log_det_cholesky.numpy(), np.log(np.diagonal(ref.precisions_cholesky, axis1=1,
                                                                             axis2=2)).sum(1), rtol=2e-6)
    with pytest.raises(ValueError, match="Cholesky"):
        priors.compute_precision_cholesky(-np.eye(4)[None])


def test_patch_prior_arguments_and_shift_draws():
    gmm = J.GaussianMixtureModel.from_numpy(*small_gmm(), meta=J.GaussianMixtureModelMeta(stride=4))
    prior = J.GMMPatchPrior(gmm=gmm, generator=torch.Generator().manual_seed(7))
    assert prior.stride == 4 and prior.log_like_weight == 16 / 64 and not prior.marginalize
    # two randint(-2, 3) draws per call, row shift first (utils/torch.py:108-116)
    g = torch.Generator().manual_seed(7)
    want = [(int(torch.randint(-2, 3, (1,), generator=g)), int(torch.randint(-2, 3, (1,), generator=g)))
            for _ in range(6)]
    assert [prior.draw_shifts() for _ in range(6)] == want
    assert J.GMMPatchPrior(gmm=gmm, cycle_spin=False).draw_shifts() == (0, 0)
    with pytest.raises(ValueError, match="stride"):
        J.GMMPatchPrior(gmm=J.GaussianMixtureModel.from_numpy(*small_gmm()))  # meta.stride None, no stride given
    with pytest.raises(NotImplementedError):
        J.GMMPatchPrior(gmm=gmm, jitter=True)
    with pytest.raises(J.JolidecoB200Error):
        J.GMMPatchPrior()  # packaged GMM library is not available offline
    with pytest.raises(J.JolidecoB200Error, match="CUDA"):
        prior(torch.ones(1, 1, 16, 16))  # CPU tensor: refused, never evaluated on the host
    # the generator is pickled by state (priors/core.py:28-47)
    clone = pickle.loads(pickle.dumps(prior))
    assert clone.draw_shifts() == prior.draw_shifts()


def test_trace_table_indexing_like_the_reference():
    t = TraceTable(names=["total", "dataset-0"], dtype=[float, float])
    for i in range(4):
        t.add_row({"total": float(i), "dataset-0": 2.0 * i})
    assert len(t) == 4 and t[-1]["total"] == 3.0 and t.colnames == ["total", "dataset-0"]
    np.testing.assert_array_equal(t["dataset-0"], [0.0, 2.0, 4.0, 6.0])
    assert len(t[-2:]) == 2 and t[-2:][0]["total"] == 2.0
    assert t.to_dict()["total"] == [0.0, 1.0, 2.0, 3.0]


def test_gmm_numpy_accessors_and_host_log_prob():
    means, cov, w = small_gmm(K=6, seed=3)
    gmm = J.GaussianMixtureModel.from_numpy(means, cov, w, meta=J.GaussianMixtureModelMeta(stride=4))
    assert gmm.covariances_numpy.shape == (6, 64, 64) and gmm.weights_numpy.shape == (6,)
    np.testing.assert_allclose(gmm.log_weights_numpy, np.log(w), rtol=1e-6)
    np.testing.assert_allclose(gmm.log_det_cholesky_numpy, gmm.log_det_cholesky.numpy(), rtol=1e-5)
    assert gmm.is_equal(gmm) and not gmm.is_equal(gmm.reduce_to_topk(3))
    top = gmm.reduce_to_topk(2)
    assert top.n_components == 2
    np.testing.assert_allclose(np.sort(top.weights_numpy), np.sort(w)[-2:], rtol=1e-6)
    x = np.random.default_rng(0).normal(0, 0.1, size=(7, 64))
    lp = gmm.estimate_log_prob_numpy(x)
    ref = O.GMM(means, cov, w, dtype=np.float64).estimate_log_prob(x.astype(np.float64))
    assert lp.shape == (7, 6)
    np.testing.assert_allclose(lp, ref, rtol=1e-4)
    prior = J.GMMPatchPrior(gmm=gmm)
    assert prior.overlap == 4
    comps = J.FluxComponents()
    comps["a"] = J.SpatialFluxComponent.from_numpy(flux=np.full((8, 8), 2.0), prior=J.UniformPrior())
    comps["b"] = J.SpatialFluxComponent.from_numpy(flux=np.full((8, 8), 3.0), prior=J.UniformPrior())
    np.testing.assert_allclose(comps.flux_upsampled_total.detach().numpy()[0, 0], 5.0, rtol=1e-6)


def test_result_write_read_roundtrip(tmp_path):
    """MAPDeconvolverResult.write / read / read_checkpoint (core.py:329-343, 435-471): config, trace, the stored
    parameter (lossless under a mask), errors and calibration parameters survive; an existing file is not overwritten."""
    from jolideco_b200.core import MAPDeconvolverResult
    from jolideco_b200.table import TraceTable

    rng = np.random.default_rng(0)
    mask = np.ones((8, 8), dtype=bool)
    mask[0, 0] = False
    comps = J.FluxComponents()
    comps["flux"] = J.SpatialFluxComponent.from_numpy(flux=rng.gamma(2.0, size=(8, 8)), upsampling_factor=2, mask=mask)
    cals = J.NPredCalibrations()
    cals["obs"] = J.NPredCalibration(shift_x=0.3, shift_y=-0.1, background_norm=1.2)
    trace = TraceTable(names=["total", "dataset-obs", "filename"])
    trace.add_row({"total": 1.5, "dataset-obs": 1.25, "filename": "checkpoint-epoch-0.npz"})
    res = MAPDeconvolverResult(config={"n_epochs": 3, "device": "cuda", "checkpoint_path": str(tmp_path)},
                               components=comps, trace_loss=trace, calibrations=cals)
    fn = tmp_path / "checkpoint-epoch-0.npz"
    res.write(fn)
    with pytest.raises(IOError):
        res.write(fn)
    res.write(fn, overwrite=True)
    back = MAPDeconvolverResult.read(fn)
    assert np.array_equal(back.flux_upsampled_total, res.flux_upsampled_total)
    assert back.components["flux"].upsampling_factor == 2 and back.components["flux"].use_log_flux
    assert back.trace_loss[-1]["total"] == 1.5 and back.trace_loss[-1]["filename"] == "checkpoint-epoch-0.npz"
    assert back.calibrations["obs"].to_dict()["shift_x"] == pytest.approx(0.3)
    assert back.config["n_epochs"] == 3
    again = res.read_checkpoint(0)
    assert np.array_equal(again.flux_upsampled_total, res.flux_upsampled_total)


def test_datasets_are_cast_to_float32():
    """Integer counts / float64 exposure, background and PSF (which the reference promotes on the fly) are cast at the
    API boundary instead of failing inside the kernels' dtype checks."""
    rng = np.random.default_rng(1)
    ds = {"a": dict(counts=rng.poisson(2.0, size=(16, 16)), psf=np.full((3, 3), 1 / 9.0), exposure=np.ones((16, 16)),
                    background=np.full((16, 16), 0.5))}
    comps = J.FluxComponents()
    comps["flux"] = J.SpatialFluxComponent.from_numpy(flux=np.ones((16, 16)))
    loss = J.PoissonLoss.from_datasets(ds, comps, device="cpu")
    assert loss.counts_all[0].dtype == torch.float32
    model = loss.npred_models_all[0]
    assert model.background.dtype == torch.float32 and model["flux"].psf.dtype == torch.float32
    assert model["flux"].exposure.dtype == torch.float32
