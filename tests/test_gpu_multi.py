"""Joint-step mode: single GPU against the oracle, and (when >= 2 GPUs are visible) the dataset-sharded
2-rank run over NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, unpack_datasets
from oracle import jolideco_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import jolideco_b200 as J


@pytest.mark.parametrize("marginalize", [False, True])
def test_joint_mode_single_gpu_matches_oracle(marginalize):
    g = load_golden("run_gmm_max.npz")
    raw = unpack_datasets(g)
    n_epochs = 6
    gmm = J.GaussianMixtureModel.from_numpy(g["gmm_means"], g["gmm_cov"], g["gmm_w"],
                                            meta=J.GaussianMixtureModelMeta(stride=4))
    gen = torch.Generator().manual_seed(3)
    probe = torch.Generator()
    probe.set_state(gen.get_state())
    shifts = [(int(torch.randint(-2, 3, (1,), generator=probe)), int(torch.randint(-2, 3, (1,), generator=probe)))
              for _ in range(2 * n_epochs)]
    prior = J.GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=marginalize)
    comps = J.FluxComponents()
    comps["flux"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=prior)
    res = J.MAPDeconvolver(n_epochs=n_epochs, display_progress=False, device="cuda", mode="joint").run(
        datasets={str(i): d for i, d in enumerate(raw)}, components=comps)
    ods = [O.prepare_dataset(d, f=1) for d in raw]
    flux_ref, trace_ref = O.map_run_joint(g["flux_init_up"], ods, n_epochs,
                                          gmm=O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"]), shifts=shifts,
                                          marginalize=marginalize)
    flux = res.flux_upsampled_total
    assert np.linalg.norm(flux - flux_ref) / np.linalg.norm(flux_ref) < 1e-3
    np.testing.assert_allclose(res.trace_loss["total"], [t["total"] for t in trace_ref], rtol=2e-5)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("collective", ["nccl", "peer"])
def test_dataset_sharded_joint_run_two_ranks(collective):
    """NCCL all-reduce + replicated Adam, and the fused peer-memory reduce + Adam + theta broadcast kernel."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "dist_worker.py"),
           collective]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
    assert res.returncode == 0 and "DIST_WORKER_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
