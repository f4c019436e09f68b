"""Seeded synthetic workloads of the BASELINE.json shapes (no network, no packaged GMM library):
Poisson counts of a blobs + point-sources sky, Gaussian PSF, linear exposure gradient, constant
background, and a synthetic zero-mean K-component GMM over 8x8 patches."""
import numpy as np

WORKLOADS = {
    # name: (H, f, psf, K, D, description)
    "cfg1": dict(H=128, f=1, psf=17, K=0, D=1, desc="first-steps toy: 128x128, Gaussian PSF 17x17, uniform prior"),
    "cfg2": dict(H=256, f=2, psf=17, K=256, D=1,
                 desc="256x256 single dataset, oversample=2 (512x512 flux, 34x34 PSF), GMM patch prior K=256"),
    "cfg3": dict(H=512, f=1, psf=64, K=256, D=8, desc="Chandra-like joint: 8 datasets of 512x512, 64x64 PSFs, GMM K=256"),
    "cfg4": dict(H=1024, f=1, psf=201, K=256, D=20, desc="Fermi-LAT-like: 20 datasets of 1024x1024, 201x201 PSFs (FFT path)"),
    "cfg5": dict(H=256, f=1, psf=17, K=256, D=1, desc="one of 64 independent 256x256 GMM-prior runs"),
    "joint1024": dict(H=1024, f=1, psf=17, K=256, D=8,
                      desc="north-star: 1024x1024 8-dataset GMM-prior joint deconvolution, 17x17 PSFs, K=256"),
    "joint_tiny": dict(H=64, f=1, psf=7, K=8, D=2, desc="joint-step smoke size: 2 datasets of 64x64, 7x7 PSFs, K=8"),
    "tiny": dict(H=48, f=1, psf=7, K=8, D=2, desc="smoke-test size"),
}


def gaussian_psf(size, sigma=None):
    sigma = size / 8.0 if sigma is None else sigma
    x = np.arange(size) - (size - 1) / 2.0
    g = np.exp(-0.5 * (x / sigma) ** 2)
    psf = np.outer(g, g)
    return (psf / psf.sum()).astype(np.float32)


def synthetic_gmm(K, D=64, seed=0, mean_scale=0.0):
    """cov_k = A_k A_k^T + 0.01 I, A_k ~ N(0, 0.05^2); zero (or small) means; normalised weights."""
    rng = np.random.default_rng(seed)
    A = rng.normal(0, 0.05, size=(K, D, D))
    cov = A @ A.transpose(0, 2, 1) + 0.01 * np.eye(D)
    means = rng.normal(0, mean_scale, size=(K, D)) if mean_scale else np.zeros((K, D))
    w = rng.uniform(0.5, 1.5, size=K)
    return means, cov, w / w.sum()


def synthetic_sky(H, W, rng):
    y, x = np.mgrid[:H, :W]
    flux = np.full((H, W), 0.2)
    for _ in range(6):
        cy, cx = rng.uniform(0.15, 0.85) * H, rng.uniform(0.15, 0.85) * W
        s = rng.uniform(0.02, 0.08) * H
        flux += rng.uniform(5, 30) * np.exp(-0.5 * ((y - cy) ** 2 + (x - cx) ** 2) / s**2)
    for _ in range(12):
        flux[rng.integers(4, H - 4), rng.integers(4, W - 4)] += rng.uniform(20, 200)
    return flux


def synthetic_datasets(H, W, psf_size, D, seed=0, background=0.5):
    """D datasets sharing one sky: dict name -> dict(counts, psf, exposure, background) of float32 arrays."""
    from scipy.signal import fftconvolve

    rng = np.random.default_rng(seed)
    sky = synthetic_sky(H, W, rng)
    datasets = {}
    for i in range(D):
        psf = gaussian_psf(psf_size, psf_size / 8.0 * rng.uniform(0.8, 1.2))
        grad = np.linspace(-1, 1, H).reshape(-1, 1) if i % 2 == 0 else np.linspace(-1, 1, W).reshape(1, -1)
        exposure = np.ones((H, W)) + 0.5 * rng.uniform(0.5, 1.0) * grad
        bkg = np.full((H, W), background * rng.uniform(0.5, 1.5))
        npred = bkg + fftconvolve(sky * exposure, psf, mode="same")
        counts = rng.poisson(np.clip(npred, 0, None))
        datasets[f"obs-{i}"] = {
            "counts": counts.astype(np.float32),
            "psf": psf,
            "exposure": exposure.astype(np.float32),
            "background": bkg.astype(np.float32),
        }
    return datasets, sky


def make_workload(name, seed=0, n_datasets=None, K=None, gmm_mean_scale=0.0):
    """Returns dict(datasets, flux_init, f, gmm_arrays or None, cfg).  gmm_mean_scale > 0: mixture components with
    non-zero means (the kernels' general path; SURVEY 8d's synthetic mixture is zero-mean)."""
    cfg = dict(WORKLOADS[name])
    if n_datasets is not None:
        cfg["D"] = n_datasets
    if K is not None:
        cfg["K"] = K
    H = cfg["H"]
    datasets, _ = synthetic_datasets(H, H, cfg["psf"], cfg["D"], seed=seed)
    rng = np.random.default_rng(seed + 1000)
    flux_init = rng.gamma(20.0, size=(H, H)) / 20.0
    gmm_arrays = synthetic_gmm(cfg["K"], seed=seed + 7, mean_scale=gmm_mean_scale) if cfg["K"] else None
    return dict(datasets=datasets, flux_init=flux_init, f=cfg["f"], gmm_arrays=gmm_arrays, cfg=cfg, name=name)
