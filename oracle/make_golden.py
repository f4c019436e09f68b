"""TEST INFRASTRUCTURE — generate `tests/golden/*.npz` by running the UNMODIFIED reference
(imported from /root/reference through `oracle/ref_shim.py`).  Run in the build container:

    python oracle/make_golden.py

The GPU box has no copy of the reference, so the committed fixtures are what pins the oracle
(`tests/test_oracle_golden.py`) and the CUDA path (`tests/test_gpu_*.py`) to the reference.
Cycle-spin shifts are not monkey-patched: the reference draws them from its seeded
`torch.Generator` (utils/torch.py:108-119); we peek at a clone of the generator state before
each prior call to record what it is about to draw.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

ref_shim.install()

from jolideco.core import MAPDeconvolver  # noqa: E402
from jolideco.data import disk_source_gauss_psf, gauss_and_point_sources_gauss_psf  # noqa: E402
from jolideco.loss import PoissonLoss, TotalLoss  # noqa: E402
from jolideco.models import FluxComponents, NPredModels, SpatialFluxComponent  # noqa: E402
from jolideco.priors import GMMPatchPrior, UniformPrior  # noqa: E402
from jolideco.priors.patches.gmm import GaussianMixtureModel, GaussianMixtureModelMeta  # noqa: E402
from jolideco.utils.torch import convolve_fft_torch, shift_image_torch, view_as_overlapping_patches_torch  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def synthetic_gmm_arrays(K, D=64, seed=0, mean_scale=0.01):
    rng = np.random.default_rng(seed)
    A = rng.normal(0, 0.05, size=(K, D, D))
    cov = A @ A.transpose(0, 2, 1) + 0.01 * np.eye(D)
    means = rng.normal(0, mean_scale, size=(K, D))
    w = rng.uniform(0.5, 1.5, size=K)
    return means, cov, w / w.sum()


def make_gmm(K, seed=0, meta_stride=4):
    means, cov, w = synthetic_gmm_arrays(K, seed=seed)
    gmm = GaussianMixtureModel.from_numpy(means, cov, w, meta=GaussianMixtureModelMeta(stride=meta_stride))
    return gmm, (means, cov, w)


def peek_shifts(generator):
    g = torch.Generator()
    g.set_state(generator.get_state())
    sy = int(torch.randint(-2, 3, (1,), generator=g))
    sx = int(torch.randint(-2, 3, (1,), generator=g))
    return sy, sx


def synthetic_dataset(rng, H, W, kh, kw, f=1, bkg=0.5, level=20.0):
    y, x = np.mgrid[:H, :W]
    flux = 0.2 + level * np.exp(-0.5 * ((y - H / 2) ** 2 + (x - W / 3) ** 2) / (H / 8) ** 2)
    flux[H // 4, W // 2] += 50
    ky, kx = np.mgrid[:kh, :kw]
    psf = np.exp(-0.5 * (((ky - (kh - 1) / 2) / (kh / 5)) ** 2 + ((kx - (kw - 1) / 2) / (kw / 5)) ** 2))
    psf /= psf.sum()
    exposure = 1 + 0.5 * np.linspace(-1, 1, H).reshape(-1, 1) * np.ones((1, W))
    background = bkg * np.ones((H, W))
    from scipy.signal import fftconvolve

    npred = background + fftconvolve(flux * exposure, psf, mode="same")
    counts = rng.poisson(np.clip(npred, 0, None))
    return {
        "counts": counts.astype(np.float32),
        "psf": psf.astype(np.float32),
        "exposure": exposure.astype(np.float32),
        "background": background.astype(np.float32),
    }


def pack_datasets(datasets, prefix, out):
    for i, (name, d) in enumerate(datasets.items()):
        for key in ["counts", "psf", "exposure", "background"]:
            out[f"{prefix}{i}_{key}"] = d[key]
    out[f"{prefix}n"] = len(datasets)


# ----------------------------------------------------------------------------------------
def golden_kat():
    out = {}
    rng = np.random.default_rng(1)
    # patch order (utils/tests/test_torch.py:8-21) + a ragged image with dropped trailing rows
    img = rng.normal(size=(21, 19)).astype(np.float32)
    out["patches_img"] = img
    out["patches_8_4"] = view_as_overlapping_patches_torch(torch.from_numpy(img[None, None]), (8, 8), 4).numpy()
    out["patches_8_2"] = view_as_overlapping_patches_torch(torch.from_numpy(img[None, None]), (8, 8), 2).numpy()
    # convolution: odd x even kernel, float64
    im = rng.normal(size=(20, 23))
    ke = rng.uniform(size=(5, 4))
    out["conv_img"], out["conv_ker"] = im, ke
    out["conv_out"] = convolve_fft_torch(torch.from_numpy(im[None, None]), torch.from_numpy(ke[None, None])).numpy()[0, 0]
    # NPred setup + forward with upsampling 2 and an even PSF (6x6 -> 12x12)
    ds = synthetic_dataset(rng, 24, 20, 6, 6)
    comps = FluxComponents()
    flux0 = rng.gamma(5.0, size=(24, 20))
    comps["flux"] = SpatialFluxComponent.from_numpy(flux=flux0, upsampling_factor=2, prior=UniformPrior())
    npm = NPredModels.from_dataset_numpy(ds, comps)
    for k in ds:
        out[f"npred_ds_{k}"] = ds[k]
    out["npred_flux_init"] = flux0
    out["npred_flux_up"] = comps["flux"].flux_upsampled.detach().numpy()[0, 0]
    out["npred_exposure_up"] = npm["flux"].exposure.numpy()[0, 0]
    out["npred_psf_up"] = npm["flux"].psf.numpy()[0, 0]
    fluxes = comps.to_flux_tuple()
    npred = npm.evaluate(fluxes=fluxes)
    out["npred_out"] = npred.detach().numpy()[0, 0]
    pl = PoissonLoss([torch.from_numpy(ds["counts"][None, None])], [npm], ["0"])
    loss = pl.loss_function(npred, pl.counts_all[0])
    loss.backward()
    out["npred_loss"] = loss.item()
    out["npred_theta_grad"] = comps["flux"]._flux_upsampled.grad.numpy()[0, 0]
    # Poisson loss with exact zeros in npred and counts in {0,1,2,...}
    n = rng.gamma(1.0, size=(16, 16)).astype(np.float32)
    n[0, :4] = 0
    c = rng.poisson(1.0, size=(16, 16)).astype(np.float32)
    nt = torch.from_numpy(n[None, None]).requires_grad_()
    l2 = pl.loss_function(nt, torch.from_numpy(c[None, None]))
    l2.backward()
    out["poisson_npred"], out["poisson_counts"] = n, c
    out["poisson_loss"] = l2.item()
    out["poisson_grad"] = nt.grad.numpy()[0, 0]
    # GMM constants + log-prob (float32, the reference's working precision)
    gmm, (means, cov, w) = make_gmm(5, seed=3)
    x = rng.normal(0, 0.3, size=(37, 64)).astype(np.float32)
    x -= x.mean(axis=1, keepdims=True)
    out["gmm_means"], out["gmm_cov"], out["gmm_w"] = means, cov, w
    out["gmm_x"] = x
    out["gmm_logp"] = gmm.estimate_log_prob(torch.from_numpy(x)).numpy()
    out["gmm_prec_chol"] = gmm.precisions_cholesky.numpy()
    out["gmm_mu_prec"] = gmm.means_precisions_cholesky.numpy()
    out["gmm_log_det"] = gmm.log_det_cholesky.numpy()
    out["gmm_pixel_weights"] = gmm.pixel_weights.numpy()
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **out)
    print("kat.npz", len(out))


# ----------------------------------------------------------------------------------------
def golden_prior_step():
    """Value and autograd gradient of GMMPatchPrior + one dataset's Poisson loss, fp32 and fp64."""
    out = {}
    rng = np.random.default_rng(2)
    K = 6
    _, (means, cov, w) = make_gmm(K, seed=5)
    out["gmm_means"], out["gmm_cov"], out["gmm_w"] = means, cov, w
    H, W = 38, 46  # (38-8)%4 != 0: trailing rows dropped
    flux = rng.gamma(2.0, size=(H, W)).astype(np.float32)
    out["flux"] = flux
    case = 0
    for marginalize in [False, True]:
        for seed in [0, 1, 2, 7]:
            for dtype in [torch.float32, torch.float64]:
                gmm = GaussianMixtureModel.from_numpy(means, cov, w, meta=GaussianMixtureModelMeta(stride=4))
                gen = torch.Generator().manual_seed(seed)
                prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=marginalize)
                if dtype == torch.float64:
                    prior = prior.double()
                sy, sx = peek_shifts(gen)
                f = torch.from_numpy(flux[None, None]).to(dtype).requires_grad_()
                val = prior(flux=f)
                val.backward()
                tag = f"c{case}_{'f32' if dtype == torch.float32 else 'f64'}"
                out[f"{tag}_value"] = val.item()
                out[f"{tag}_grad"] = f.grad.numpy()[0, 0]
            out[f"c{case}_shift"] = np.array([sy, sx])
            out[f"c{case}_marginalize"] = marginalize
            case += 1
    out["n_cases"] = case
    np.savez_compressed(os.path.join(OUT, "prior_step.npz"), **out)
    print("prior_step.npz", case, "cases")


# ----------------------------------------------------------------------------------------
def run_reference(datasets, flux_init, f, n_epochs, gmm_arrays=None, marginalize=False, seed=0, beta=1.0):
    if gmm_arrays is not None:
        gmm = GaussianMixtureModel.from_numpy(*gmm_arrays, meta=GaussianMixtureModelMeta(stride=4))
        gen = torch.Generator().manual_seed(seed)
        prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=marginalize)
        # record the draws: D per epoch for training + 1 for the trace (loss.py:224)
        g = torch.Generator()
        g.set_state(gen.get_state())
        D = len(datasets)
        shifts, trace_shifts = [], []
        for _ in range(n_epochs):
            for _ in range(D):
                shifts.append(peek_and_advance(g))
            trace_shifts.append(peek_and_advance(g))
    else:
        prior, shifts, trace_shifts = UniformPrior(), [], []
    comps = FluxComponents()
    comps["flux-1"] = SpatialFluxComponent.from_numpy(flux=flux_init, upsampling_factor=f, prior=prior)
    flux_init_up = comps["flux-1"].flux_upsampled.detach().numpy()[0, 0].copy()
    deco = MAPDeconvolver(n_epochs=n_epochs, learning_rate=0.1, beta=beta, display_progress=False)
    res = deco.run(datasets=datasets, components=comps)
    tr = res.trace_loss
    D = len(datasets)
    return dict(
        flux_init_up=flux_init_up,
        flux_up=res.flux_upsampled_total,
        flux=res.flux_total,
        trace_total=np.asarray(tr["total"]),
        trace_datasets=np.stack([np.asarray(tr[f"dataset-{n}"]) for n in datasets], axis=1),
        trace_prior=np.asarray(tr["priors-total"]),
        shifts=np.array(shifts).reshape(-1, 2),
        trace_shifts=np.array(trace_shifts).reshape(-1, 2),
    )


def peek_and_advance(g):
    sy = int(torch.randint(-2, 3, (1,), generator=g))
    sx = int(torch.randint(-2, 3, (1,), generator=g))
    return sy, sx


def golden_runs():
    # (1) the reference's own e2e golden (tests/test_core.py:71-79): 3 toy datasets, uniform prior
    out = {}
    rs = np.random.RandomState(642020)
    datasets = {str(i): gauss_and_point_sources_gauss_psf(random_state=rs) for i in range(3)}
    rs = np.random.RandomState(642020)
    flux_init = rs.gamma(20, size=(32, 32))
    pack_datasets(datasets, "ds", out)
    out["flux_init"] = flux_init
    for k, v in run_reference(datasets, flux_init, 1, 100).items():
        out[k] = v
    np.savez_compressed(os.path.join(OUT, "run_uniform.npz"), **out)
    print("run_uniform.npz flux[12,12]=%.6f (ref test: 1.542659) total=%.6f (5.842237)" % (
        out["flux"][12, 12], out["trace_total"][-1]))

    # (2) upsampling 2 (tests/test_core.py:99-124): disk datasets
    out = {}
    rs = np.random.RandomState(642020)
    datasets = {str(i): disk_source_gauss_psf(random_state=rs) for i in range(3)}
    rs = np.random.RandomState(642020)
    flux_init = rs.gamma(20, size=(32, 32))
    pack_datasets(datasets, "ds", out)
    out["flux_init"] = flux_init
    for k, v in run_reference(datasets, flux_init, 2, 100).items():
        out[k] = v
    np.savez_compressed(os.path.join(OUT, "run_upsampling2.npz"), **out)
    print("run_upsampling2.npz flux[12,12]=%.6f (ref test: 3.565998) total=%.6f (5.844786)" % (
        out["flux"][12, 12], out["trace_total"][-1]))

    # (3) GMM patch prior runs (no stored golden exists in the reference for a synthetic GMM:
    # pinned by this live run of the imported reference), both max and logsumexp modes
    rng = np.random.default_rng(11)
    datasets = {str(i): synthetic_dataset(rng, 40, 36, 7, 7) for i in range(2)}
    flux_init = rng.gamma(20, size=(40, 36)) / 10
    gmm_arrays = synthetic_gmm_arrays(8, seed=9)
    for marginalize in [False, True]:
        out = {}
        pack_datasets(datasets, "ds", out)
        out["flux_init"] = flux_init
        out["gmm_means"], out["gmm_cov"], out["gmm_w"] = gmm_arrays
        out["marginalize"] = marginalize
        for k, v in run_reference(datasets, flux_init, 1, 8, gmm_arrays, marginalize, seed=4).items():
            out[k] = v
        name = "run_gmm_lse.npz" if marginalize else "run_gmm_max.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print(name, "total", out["trace_total"][-1], "shifts", out["shifts"][:3].tolist())

    # (4) GMM prior + upsampling 2 + even PSF
    rng = np.random.default_rng(12)
    datasets = {str(i): synthetic_dataset(rng, 24, 24, 6, 6) for i in range(2)}
    flux_init = rng.gamma(20, size=(24, 24)) / 10
    out = {}
    pack_datasets(datasets, "ds", out)
    out["flux_init"] = flux_init
    out["gmm_means"], out["gmm_cov"], out["gmm_w"] = gmm_arrays
    out["marginalize"] = False
    for k, v in run_reference(datasets, flux_init, 2, 6, gmm_arrays, False, seed=5).items():
        out[k] = v
    np.savez_compressed(os.path.join(OUT, "run_gmm_up2.npz"), **out)
    print("run_gmm_up2.npz total", out["trace_total"][-1])


def golden_calibration():
    """GMM prior + NPredCalibrations with trainable background norms (shifts 0, as in the Chandra example)."""
    from jolideco.models import NPredCalibration, NPredCalibrations

    rng = np.random.default_rng(21)
    datasets = {str(i): synthetic_dataset(rng, 40, 36, 7, 7) for i in range(2)}
    flux_init = rng.gamma(20, size=(40, 36)) / 10
    gmm_arrays = synthetic_gmm_arrays(8, seed=9)
    n_epochs, norms = 8, [1.3, 0.7]
    gmm = GaussianMixtureModel.from_numpy(*gmm_arrays, meta=GaussianMixtureModelMeta(stride=4))
    gen = torch.Generator().manual_seed(6)
    prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen)
    g = torch.Generator()
    g.set_state(gen.get_state())
    shifts, trace_shifts = [], []
    for _ in range(n_epochs):
        for _ in range(2):
            shifts.append(peek_and_advance(g))
        trace_shifts.append(peek_and_advance(g))
    comps = FluxComponents()
    comps["flux-1"] = SpatialFluxComponent.from_numpy(flux=flux_init, upsampling_factor=1, prior=prior)
    flux_init_up = comps["flux-1"].flux_upsampled.detach().numpy()[0, 0].copy()
    cals = NPredCalibrations()
    for name, b in zip(datasets, norms):
        cals[name] = NPredCalibration(background_norm=b)
    res = MAPDeconvolver(n_epochs=n_epochs, learning_rate=0.1, display_progress=False).run(
        datasets=datasets, components=comps, calibrations=cals)
    out = {}
    pack_datasets(datasets, "ds", out)
    out["flux_init"], out["flux_init_up"] = flux_init, flux_init_up
    out["gmm_means"], out["gmm_cov"], out["gmm_w"] = gmm_arrays
    out["marginalize"] = False
    out["background_norm_init"] = np.array(norms)
    out["background_norm"] = np.array([float(c.background_norm) for c in res.calibrations.values()])
    out["flux_up"] = res.flux_upsampled_total
    tr = res.trace_loss
    out["trace_total"] = np.asarray(tr["total"])
    out["trace_datasets"] = np.stack([np.asarray(tr[f"dataset-{n}"]) for n in datasets], axis=1)
    out["trace_prior"] = np.asarray(tr["priors-total"])
    out["shifts"] = np.array(shifts).reshape(-1, 2)
    out["trace_shifts"] = np.array(trace_shifts).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "run_gmm_calib.npz"), **out)
    print("run_gmm_calib.npz norms", out["background_norm"], "total", out["trace_total"][-1])


def golden_calibration_shift():
    """GMM prior + NPredCalibrations with trainable NON-ZERO sub-pixel shifts and background norms
    (the Chandra example fits both, examples/chandra-e0102-filament.py:197-206), f = 1 and f = 2."""
    from jolideco.models import NPredCalibration, NPredCalibrations

    for f, tag in ((1, "run_gmm_shift.npz"), (2, "run_gmm_shift_up2.npz")):
        rng = np.random.default_rng(41 + f)
        datasets = {str(i): synthetic_dataset(rng, 32, 28, 6 if f == 2 else 7, 6 if f == 2 else 7) for i in range(2)}
        flux_init = rng.gamma(20, size=(32, 28)) / 10
        gmm_arrays = synthetic_gmm_arrays(8, seed=9)
        n_epochs, norms, shifts_xy = 6, [1.2, 0.8], [(0.4, -0.7), (-0.25, 0.3)]
        gmm = GaussianMixtureModel.from_numpy(*gmm_arrays, meta=GaussianMixtureModelMeta(stride=4))
        gen = torch.Generator().manual_seed(8)
        prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen)
        g = torch.Generator()
        g.set_state(gen.get_state())
        shifts, trace_shifts = [], []
        for _ in range(n_epochs):
            for _ in range(2):
                shifts.append(peek_and_advance(g))
            trace_shifts.append(peek_and_advance(g))
        comps = FluxComponents()
        comps["flux-1"] = SpatialFluxComponent.from_numpy(flux=flux_init, upsampling_factor=f, prior=prior)
        flux_init_up = comps["flux-1"].flux_upsampled.detach().numpy()[0, 0].copy()
        cals = NPredCalibrations()
        for name, b, (sx, sy) in zip(datasets, norms, shifts_xy):
            cals[name] = NPredCalibration(shift_x=sx, shift_y=sy, background_norm=b)
        res = MAPDeconvolver(n_epochs=n_epochs, learning_rate=0.1, display_progress=False).run(
            datasets=datasets, components=comps, calibrations=cals)
        out = {}
        pack_datasets(datasets, "ds", out)
        out["flux_init"], out["flux_init_up"], out["upsampling"] = flux_init, flux_init_up, f
        out["gmm_means"], out["gmm_cov"], out["gmm_w"] = gmm_arrays
        out["marginalize"] = False
        out["background_norm_init"] = np.array(norms)
        out["shift_xy_init"] = np.array(shifts_xy)
        out["background_norm"] = np.array([float(c.background_norm) for c in res.calibrations.values()])
        out["shift_xy"] = np.stack([c.shift_xy.detach().numpy()[0] for c in res.calibrations.values()])
        out["flux_up"] = res.flux_upsampled_total
        tr = res.trace_loss
        out["trace_total"] = np.asarray(tr["total"])
        out["trace_datasets"] = np.stack([np.asarray(tr[f"dataset-{n}"]) for n in datasets], axis=1)
        out["trace_prior"] = np.asarray(tr["priors-total"])
        out["shifts"] = np.array(shifts).reshape(-1, 2)
        out["trace_shifts"] = np.array(trace_shifts).reshape(-1, 2)
        np.savez_compressed(os.path.join(OUT, tag), **out)
        print(tag, "norms", out["background_norm"], "shift_xy", out["shift_xy"].tolist(), "total", out["trace_total"][-1])


def golden_joint_objective():
    """The joint objective of `mode="joint"`: TotalLoss.__call__ = sum_d L_d - beta * prior (loss.py:257-261).
    Value from the reference's own call; gradient w.r.t. the flux assembled from the reference's components with
    autograd (per-dataset `loss_function(npred_model.evaluate(fluxes), counts)` and the prior) - `TotalLoss.__call__`
    itself detaches the dataset terms (loss.py:71), so its own backward would miss them."""
    rng = np.random.default_rng(51)
    datasets = {str(i): synthetic_dataset(rng, 32, 36, 7, 7) for i in range(3)}
    flux = rng.gamma(20, size=(32, 36)) / 10
    gmm_arrays = synthetic_gmm_arrays(8, seed=9)
    beta = 0.7
    out = {}
    pack_datasets(datasets, "ds", out)
    out["flux"], out["beta"] = flux, beta
    out["gmm_means"], out["gmm_cov"], out["gmm_w"] = gmm_arrays
    for marginalize, tag in ((False, "max"), (True, "lse")):
        gmm = GaussianMixtureModel.from_numpy(*gmm_arrays, meta=GaussianMixtureModelMeta(stride=4))
        gen = torch.Generator().manual_seed(12)
        prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=marginalize)
        comps = FluxComponents()
        comps["flux-1"] = SpatialFluxComponent.from_numpy(flux=flux, upsampling_factor=1, prior=prior)
        total_loss = TotalLoss.from_datasets_and_components(datasets=datasets, components=comps, beta=beta)
        g = torch.Generator()
        g.set_state(gen.get_state())
        out[f"{tag}_shift"] = np.array(peek_and_advance(g))
        fluxes = comps.to_flux_tuple()
        out[f"{tag}_total"] = float(total_loss(fluxes))
        # intended gradient, from the reference's components (same shift: rewind the generator)
        gen.manual_seed(12)
        fl = tuple(f.detach().clone().requires_grad_(True) for f in fluxes)
        value = -beta * total_loss.prior_loss(fluxes=fl)
        for counts, npred_model in total_loss.poisson_loss.iter_by_dataset:
            value = value + total_loss.poisson_loss.loss_function(npred_model.evaluate(fluxes=fl), counts)
        value.backward()
        out[f"{tag}_total_components"] = float(value)
        out[f"{tag}_dflux"] = fl[0].grad.numpy()[0, 0]
    np.savez_compressed(os.path.join(OUT, "joint_objective.npz"), **out)
    print("joint_objective.npz", out["max_total"], out["max_total_components"], out["lse_total"])


def golden_shift():
    """`shift_image_torch` (utils/torch.py:196-223) values and autograd gradients (w.r.t. the image and shift_xy)
    for non-trivial sub-pixel shifts, fp32 and fp64; plus one whole-pixel shift (values only: the interpolant has a
    kink there)."""
    rng = np.random.default_rng(31)
    out = {}
    cases = [(0.3, -1.7, 1), (2.25, 0.6, 2), (-0.6, 0.05, 1), (-3.4, 2.2, 3)]
    image = rng.gamma(2.0, size=(13, 17))
    cot = rng.normal(size=(13, 17))
    out["image"], out["cot"] = image, cot
    out["cases"] = np.array(cases, dtype=np.float64)
    for i, (sx, sy, scale) in enumerate(cases):
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            img = torch.tensor(image[None, None], dtype=dt, requires_grad=True)
            shift_xy = torch.tensor([[sx, sy]], dtype=dt, requires_grad=True)
            res = shift_image_torch(img, shift_xy, scale=scale)
            (res * torch.tensor(cot[None, None], dtype=dt)).sum().backward()
            out[f"c{i}_{tag}_out"] = res.detach().numpy()[0, 0]
            out[f"c{i}_{tag}_dimage"] = img.grad.numpy()[0, 0]
            out[f"c{i}_{tag}_dshift_xy"] = shift_xy.grad.numpy()[0]
    img = torch.tensor(image[None, None], dtype=torch.float64)
    out["whole_out"] = shift_image_torch(img, torch.tensor([[2.0, -1.0]], dtype=torch.float64), scale=1).numpy()[0, 0]
    out["zero_is_identity"] = shift_image_torch(img, torch.zeros((1, 2), dtype=torch.float64)).numpy()[0, 0]
    np.savez_compressed(os.path.join(OUT, "shift_kat.npz"), **out)
    print("shift_kat.npz", {k: v.shape for k, v in out.items() if k.startswith("c0")})


def golden_validation_early_stop():
    """`datasets_validation` + `stop_early` (core.py:251-261, loss.py:244-248), uniform prior.
    Two runs on the reference's point-source test datasets: (a) 30 epochs, learning rate 0.1, with a validation
    dataset, no early stop; (b) learning rate 0.3 and stop_early with a 5-epoch average: the validation loss starts to
    oscillate (amplitude 1e-2) and the run ends after 16 of the 100 epochs, when it exceeds its running mean."""
    rs = np.random.RandomState(642020)
    all_ds = {str(i): gauss_and_point_sources_gauss_psf(random_state=rs) for i in range(3)}
    rs = np.random.RandomState(642020)
    flux_init = rs.gamma(20, size=(32, 32))
    datasets = {n: all_ds[n] for n in ["0", "1"]}
    validation = {n: all_ds[n] for n in ["2"]}
    out = {}
    pack_datasets(datasets, "ds", out)
    pack_datasets(validation, "dv", out)
    out["flux_init"] = flux_init
    for tag, kwargs in [("a", dict(n_epochs=30, learning_rate=0.1)),
                        ("b", dict(n_epochs=100, learning_rate=0.3, stop_early=True, stop_early_n_average=5))]:
        comps = FluxComponents()
        comps["flux-1"] = SpatialFluxComponent.from_numpy(flux=flux_init, upsampling_factor=1, prior=UniformPrior())
        res = MAPDeconvolver(display_progress=False, **kwargs).run(
            datasets=datasets, components=comps, datasets_validation=validation)
        tr = res.trace_loss
        out[f"{tag}_flux_up"] = res.flux_upsampled_total
        out[f"{tag}_trace_total"] = np.asarray(tr["total"])
        out[f"{tag}_trace_datasets"] = np.stack([np.asarray(tr[f"dataset-{n}"]) for n in datasets], axis=1)
        out[f"{tag}_trace_validation"] = np.asarray(tr["datasets-validation-total"])
        out[f"{tag}_n_epochs_run"] = len(tr)
    np.savez_compressed(os.path.join(OUT, "run_validation.npz"), **out)
    print("run_validation.npz epochs run:", out["a_n_epochs_run"], out["b_n_epochs_run"], "validation[-1]",
          out["a_trace_validation"][-1], out["b_trace_validation"][-3:])


def golden_gmm_asinh_norm():
    """Analogue of the reference's only GMM e2e test (tests/test_core.py:191-220: GMMPatchPrior(norm=ASinhImageNorm()),
    upsampling 2; its .mat GMM is not available offline): synthetic GMM, ASinhImageNorm with its two trainable
    parameters in the optimiser (utils/norms.py:235-257), upsampling 2, 6 epochs."""
    from jolideco.utils.norms import ASinhImageNorm

    rs = np.random.RandomState(642020)
    datasets = {str(i): disk_source_gauss_psf(random_state=rs) for i in range(3)}
    rs = np.random.RandomState(642020)
    flux_init = rs.gamma(20, size=(32, 32))
    gmm_arrays = synthetic_gmm_arrays(8, seed=9)
    gmm = GaussianMixtureModel.from_numpy(*gmm_arrays, meta=GaussianMixtureModelMeta(stride=4))
    n_epochs, D = 6, len(datasets)
    gen = torch.Generator().manual_seed(13)
    g = torch.Generator()
    g.set_state(gen.get_state())
    shifts, trace_shifts = [], []
    for _ in range(n_epochs):
        for _ in range(D):
            shifts.append(peek_and_advance(g))
        trace_shifts.append(peek_and_advance(g))
    norm = ASinhImageNorm()
    prior = GMMPatchPrior(gmm=gmm, stride=4, generator=gen, norm=norm)
    comps = FluxComponents()
    comps["flux-1"] = SpatialFluxComponent.from_numpy(flux=flux_init, upsampling_factor=2, prior=prior)
    out = {}
    pack_datasets(datasets, "ds", out)
    out["flux_init"] = flux_init
    out["flux_init_up"] = comps["flux-1"].flux_upsampled.detach().numpy()[0, 0].copy()
    out["gmm_means"], out["gmm_cov"], out["gmm_w"] = gmm_arrays
    out["marginalize"] = False
    out["norm_init"] = np.array([float(norm.alpha), float(norm.beta)])
    res = MAPDeconvolver(n_epochs=n_epochs, learning_rate=0.1, display_progress=False).run(datasets=datasets, components=comps)
    tr = res.trace_loss
    out["flux_up"] = res.flux_upsampled_total
    out["trace_total"] = np.asarray(tr["total"])
    out["trace_datasets"] = np.stack([np.asarray(tr[f"dataset-{n}"]) for n in datasets], axis=1)
    out["trace_prior"] = np.asarray(tr["priors-total"])
    out["norm_final"] = np.array([float(norm.alpha), float(norm.beta)])
    out["shifts"] = np.array(shifts).reshape(-1, 2)
    out["trace_shifts"] = np.array(trace_shifts).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "run_gmm_asinh.npz"), **out)
    print("run_gmm_asinh.npz total", out["trace_total"][-1], "norm", out["norm_init"], "->", out["norm_final"])


if __name__ == "__main__":
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        golden_validation_early_stop()
        golden_gmm_asinh_norm()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "shift":
        golden_shift()
        golden_calibration_shift()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "joint":
        golden_joint_objective()
        sys.exit(0)
    golden_kat()
    golden_prior_step()
    golden_runs()
    golden_calibration()
    golden_shift()
    golden_calibration_shift()
    golden_joint_objective()
    golden_validation_early_stop()
    golden_gmm_asinh_norm()
