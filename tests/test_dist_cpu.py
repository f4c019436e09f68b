"""Host-side logic of the multi-GPU path on CPU: partition helpers, and (world_size 2, gloo) that
dataset-shard + prior-row-block partial gradients all-reduce to the whole gradient."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist_t
import torch.multiprocessing as mp

from conftest import load_golden, unpack_datasets
from jolideco_b200 import dist
from oracle import jolideco_oracle as O


def test_shard_indices_partition():
    for n in [0, 1, 5, 8, 20]:
        for world in [1, 2, 3, 8]:
            parts = [dist.shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_row_blocks_tile_exactly_with_halos():
    for ny in [1, 7, 127, 255]:
        for world in [1, 2, 4, 8]:
            blocks = [dist.row_block(ny, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == ny
            for a, b in zip(blocks[:-1], blocks[1:]):
                assert a[1] == b[0]
            for lo, hi in blocks:
                if hi > lo:
                    y0, y1 = dist.halo_rows(lo, hi, 4)
                    assert y0 == 4 * lo and y1 - y0 == 4 * (hi - lo) + 4  # own rows + a 4-row halo


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist_t.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("run_gmm_max.npz")
    datasets = [O.prepare_dataset(d, f=1, dtype=np.float64) for d in unpack_datasets(g)]
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], dtype=np.float64)
    theta = np.log(g["flux_init_up"].astype(np.float64))
    ny = (theta.shape[0] - 8) // 4 + 1
    idx = dist.shard_indices(len(datasets), rank, world)
    rows = dist.row_block(ny, rank, world)
    val, grad = O.joint_loss_and_grad(theta, datasets, 1.0, gmm, (1, -2), dataset_index=idx, rows=rows)
    t = torch.from_numpy(np.concatenate([[val], grad.ravel()]))
    dist_t.all_reduce(t)
    if rank == 0:
        np.save(out, t.numpy())
    dist_t.destroy_process_group()


def test_sharded_gradient_allreduces_to_whole_gloo(tmp_path):
    out = str(tmp_path / "reduced.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    red = np.load(out)
    g = load_golden("run_gmm_max.npz")
    datasets = [O.prepare_dataset(d, f=1, dtype=np.float64) for d in unpack_datasets(g)]
    gmm = O.GMM(g["gmm_means"], g["gmm_cov"], g["gmm_w"], dtype=np.float64)
    theta = np.log(g["flux_init_up"].astype(np.float64))
    val, grad = O.joint_loss_and_grad(theta, datasets, 1.0, gmm, (1, -2))
    np.testing.assert_allclose(red[0], val, rtol=1e-12)
    np.testing.assert_allclose(red[1:], grad.ravel(), rtol=1e-9, atol=1e-14)
