"""Multi-GPU plumbing: one process per GPU, rows sharded contiguously.

Scoring and preprocessing need NO collective (independent rows).  The PCA fit exchanges
one small payload: the partial Gram (p x p fp64), column sums and row count are summed
with an NCCL all-reduce over NVLink (`engine.covariance(..., group=...)`); every rank then
finalises covariance + eigensolve redundantly.  CV folds are dealt round-robin to ranks and
the per-(fold, A) confusion counts are all-gathered.  The same code runs on `gloo` for the
CPU tests of the host logic (tests/test_host_cpu.py::test_gloo_world2_moment_allreduce) and for the GPU test
that runs two ranks on one device (tests/test_gpu_parity.py::test_cv_grid_folds_over_ranks).
"""
from __future__ import annotations

import os


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend="nccl"):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif backend == "nccl":
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_rows(n, rank, world):
    """Contiguous block [lo, hi) of n rows owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def folds_of_rank(n_folds, rank, world):
    return list(range(rank, n_folds, world))


def allreduce_moments(G, s, n, group=None):
    """Sum the partial Gram / column sums / count of all ranks in place; returns the total count."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(n)
    cnt = torch.tensor([float(n)], dtype=torch.float64, device=G.device)
    dist.all_reduce(G, group=group)
    dist.all_reduce(s, group=group)
    dist.all_reduce(cnt, group=group)
    return int(round(cnt.item()))


def allgather_counts(counts, group=None):
    """Gather per-rank int64 confusion-count tables [m,4] -> list over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [counts]
    out = [torch.empty_like(counts) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, counts, group=group)
    return out


def max_over_ranks(value, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
