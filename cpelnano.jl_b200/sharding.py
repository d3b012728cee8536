"""Sharding of estimation regions over ranks (one process per GPU) and the final result gather.

Regions are independent (the reference farms them out with Distributed.pmap, src/nanopore.jl:406), so each rank owns a
contiguous block of the region list — contiguous so that outputs stay in input order — balanced by the cost proxy
n_init·M_r·L_r.  No collective is on the data path; only the per-region result records are gathered.
"""
import numpy as np


def shard_bounds(costs, world_size):
    """Contiguous partition of len(costs) items into world_size blocks with near-equal total cost.
    Returns int64 array b[world_size+1]; rank k owns items b[k]:b[k+1]."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if world_size <= 1 or n == 0:
        return np.array([0, n] + [n] * (world_size - 1), dtype=np.int64)[:world_size + 1]
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    targets = cum[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(cum, targets, side="left")
    # choose the nearer of the two neighbouring boundaries
    cuts = np.where((cuts > 0) & (np.abs(cum[np.maximum(cuts - 1, 0)] - targets) <= np.abs(cum[np.minimum(cuts, n)] - targets)), cuts - 1, cuts)
    b = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def region_costs(grp_off, read_off, n_init):
    return n_init * np.diff(np.asarray(grp_off)).astype(np.float64) * np.diff(np.asarray(read_off)).astype(np.float64)


def gather_records(local, group=None, dst=0):
    """Gather per-rank result arrays (dict of numpy arrays, row = region) on rank `dst`, concatenated in rank order.
    Uses torch.distributed (NCCL on GPUs, gloo in the CPU tests); with a single process it is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return {k: np.concatenate([p[k] for p in parts]) for k in local}
