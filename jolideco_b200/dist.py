"""Multi-GPU partitioning of the MAP path (one process per GPU, torch.distributed / NCCL).

The reference is single-process; the path shards in three independent ways:
  1. datasets of a joint deconvolution  -> rank d mod G          (`shard_indices`)
  2. patch rows of the GMM prior        -> contiguous row blocks  (`row_block`), exact 4-row halos
     are implicit because the flux is replicated: a block reads image rows
     [s*iy0, s*(iy1-1)+8) at rolled coordinates; the overlapping gradient rows of neighbouring
     blocks are summed by the same all-reduce as (1)
  3. independent runs (bootstrap/restarts) -> `shard_indices` over runs, no collective.
Only host-side index arithmetic lives here so that it is testable on CPU (gloo).
"""
import os


def shard_indices(n, rank, world):
    """Indices of the items (datasets / runs) owned by `rank`: item i -> rank i mod world."""
    return [i for i in range(n) if i % world == rank]


def row_block(ny, rank, world):
    """Contiguous block [lo, hi) of patch rows for `rank`; blocks tile [0, ny) exactly."""
    return (ny * rank) // world, (ny * (rank + 1)) // world


def halo_rows(lo, hi, stride, patch=8):
    """Rolled-image rows read by patch rows [lo, hi): [stride*lo, stride*(hi-1)+patch)."""
    if hi <= lo:
        return (0, 0)
    return stride * lo, stride * (hi - 1) + patch


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
