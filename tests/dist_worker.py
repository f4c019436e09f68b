"""Worker of tests/test_gpu_multi.py: launched with torch.distributed.run, one rank per GPU.
Runs the dataset-sharded joint deconvolution (NCCL all-reduce of the flux gradient) and checks the
result on every rank against the oracle's joint run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import jolideco_b200 as J  # noqa: E402
from conftest import load_golden, unpack_datasets  # noqa: E402
from oracle import jolideco_oracle as O  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    g = load_golden("run_gmm_max.npz")
    raw = unpack_datasets(g)
    raw = raw + [dict(d, counts=np.roll(d["counts"], 3, axis=1)) for d in raw] + raw[:1]  # 5 datasets: uneven shards
    datasets = {f"d{i}": d for i, d in enumerate(raw)}
    n_epochs = 6
    collectives = sys.argv[1:] or ["nccl"]
    for marginalize, collective in [(m, c) for m in (False, True) for c in collectives]:
        gmm = J.GaussianMixtureModel.from_numpy(g["gmm_means"], g["gmm_cov"], g["gmm_w"],
                                                meta=J.GaussianMixtureModelMeta(stride=4))
        gen = torch.Generator().manual_seed(11)
        probe = torch.Generator()
        probe.set_state(gen.get_state())
        shifts = [(int(torch.randint(-2, 3, (1,), generator=probe)), int(torch.randint(-2, 3, (1,), generator=probe)))
                  for _ in range(2 * n_epochs)]
        prior = J.GMMPatchPrior(gmm=gmm, stride=4, generator=gen, marginalize=marginalize)
        comps = J.FluxComponents()
        comps["flux"] = J.SpatialFluxComponent.from_numpy(flux=g["flux_init"], upsampling_factor=1, prior=prior)
        deco = J.MAPDeconvolver(n_epochs=n_epochs, display_progress=False, device=f"cuda:{local}", mode="joint",
                                collective=collective)
        res = deco.run(datasets=datasets, components=comps)
        assert deco.engine.world == dist.get_world_size() and len(deco.engine.datasets) <= 3
        assert deco.engine.collective == collective
        ods = [O.prepare_dataset(d, f=1) for d in raw]
        flux_ref, trace_ref = O.map_run_joint(g["flux_init_up"], ods, n_epochs, gmm=O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"]),
                                              shifts=shifts, marginalize=marginalize)
        flux = res.flux_upsampled_total
        rel = np.linalg.norm(flux - flux_ref) / np.linalg.norm(flux_ref)
        assert rel < 1e-3, (rank, rel)
        np.testing.assert_allclose(res.trace_loss["total"], [t["total"] for t in trace_ref], rtol=2e-5)
        for i in range(len(raw)):
            np.testing.assert_allclose(res.trace_loss[f"dataset-d{i}"], [t["datasets"][i] for t in trace_ref], rtol=2e-5)
        # replicas stay bit-identical: same all-reduced gradient, same Adam
        t = torch.from_numpy(flux.copy()).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "replicas diverged"
    if rank == 0:
        print("DIST_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
