"""N>1 host logic on CPU: two gloo ranks shard a region list by cost, each 'processes' its block, rank 0 gathers.
(The compute itself needs a B200; what is covered here is the partition + gather plumbing used by multi-GPU runs.)"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import cpelnano_b200 as cpn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = cpn.simulations.make_nano_batch(37, M=7, seed=3)
    costs = cpn.sharding.region_costs(nb.grp_off, nb.read_off, 10)
    b = cpn.sharding.shard_bounds(costs, world)
    mine = np.arange(b[rank], b[rank + 1])
    sub = nb.subset(mine)
    # stand-in for the per-region result record: something only the owner of the region can know
    local = {"region": mine.astype(np.int64), "n_obs": np.diff(sub.obs_off), "first_lu": np.array([sub.log_pyx_u[o] for o in sub.obs_off[:-1]])}
    out = cpn.sharding.gather_records(local, dst=0)
    dist.barrier()
    if rank == 0:
        q.put((b.tolist(), {k: v.tolist() for k, v in out.items()}, costs.tolist()))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    import torch.multiprocessing as mp
    import cpelnano_b200 as cpn
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    b, out, costs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nb = cpn.simulations.make_nano_batch(37, M=7, seed=3)
    assert out["region"] == list(range(37))                                  # every region exactly once, input order kept
    assert out["n_obs"] == np.diff(nb.obs_off).tolist()
    assert np.array_equal(np.array(out["first_lu"]), nb.log_pyx_u[nb.obs_off[:-1]], equal_nan=True)
    c = np.array(costs)
    assert abs(c[:b[1]].sum() - c[b[1]:].sum()) <= c.max()                   # balanced to within one region


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_bounds_properties(world):
    import cpelnano_b200 as cpn
    rng = np.random.default_rng(world)
    for n in (0, 1, 5, 100, 1000):
        costs = rng.integers(1, 100, n).astype(float)
        b = cpn.sharding.shard_bounds(costs, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        if n >= 50 * world:
            loads = np.array([costs[b[k]:b[k + 1]].sum() for k in range(world)])
            assert loads.max() - loads.min() <= 2 * costs.max()
