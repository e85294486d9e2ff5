"""Sharding of the (profile point x repetition) chain grid over the GPUs of one box and the one collective
the path needs: an NCCL all-gather of per-chain bestfits at iteration boundaries.

Replaces the MPI / process-pool task farm of prospect/communication.py:50-157 and prospect/run.py:27-44:
chains never interact, so every rank owns a fixed block of chains for the whole run and there is no
data-path collective; the all-gather only makes every rank's records complete (status files, snapshots,
min over repetitions).  The gathered unit is the contiguous per-iteration record block the annealing kernel
writes (engine.Records), so no packing kernel sits between compute and the collective.  One process per GPU (torchrun), backend nccl on GPUs / gloo in the CPU tests.
"""
import os

import numpy as np


def dist_info():
    """(rank, world_size, initialised)"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), True
    except ImportError:
        pass
    return 0, 1, False


def init_from_env(backend='nccl'):
    """Initialise torch.distributed from torchrun's environment (no-op for a single process)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1 or dist.is_initialized():
        return dist_info()
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    local = int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0')))
    # a mismatched collective must fail within minutes, not hold the GPUs for the default 10: PB200_DIST_TIMEOUT_S
    from datetime import timedelta
    limit = timedelta(seconds=int(os.environ.get('PB200_DIST_TIMEOUT_S', '180')))
    if backend == 'nccl':
        torch.cuda.set_device(local)
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local), timeout=limit)
    else:
        dist.init_process_group(backend=backend, timeout=limit)
    return dist_info()


def shard_grid(n_points, reps, rank, world):
    """Chains of `rank` as (point [B_r], rep [B_r]) in rep-major order.

    P >= world: contiguous blocks of profile points (all repetitions of a point share one proposal factor, so
    they stay on one GPU); otherwise the repetitions of every point are split."""
    if world == 1:
        pts, rps = np.arange(n_points), np.arange(reps)
    elif n_points >= world:
        pts, rps = np.array_split(np.arange(n_points), world)[rank], np.arange(reps)
    else:
        pts, rps = np.arange(n_points), np.array_split(np.arange(reps), world)[rank]
    point = np.tile(pts, len(rps)).astype(np.int64)
    rep = np.repeat(rps, len(pts)).astype(np.int64)
    return point, rep


def global_chain_index(point, rep, n_points):
    """c = rep * P + point: the reference's task enumeration (tasks/initialise_profile_task.py:16-26)."""
    return rep * n_points + point
