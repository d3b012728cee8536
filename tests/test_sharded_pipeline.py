"""File-level multi-GPU path on CPU: two gloo ranks run analyze_nano_csr(rank, world) on their slice of every chromosome,
barrier, rank 0 merges the part files — the result must be the bytes of a single-process run.

The device call (cpn_analyze_batch) needs a B200, so it is replaced here by a deterministic stand-in that depends on every
input array of the batch (so that a wrong slice, offset or start would change the files); everything around it — native FASTA
/ CpG index / partition / features / calls / coverage filters / sub-region tables / writers, the sharding and the merge — is
the product code.  The same comparison with the real fit runs on the GPU in tests/test_pipeline_gpu.py.
"""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")


def _stand_in(grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, *a, **k):
    R = len(grp_off) - 1
    phi = np.asarray(phi0).reshape(R, n_init, 3)[:, 0, :].copy()
    llr = np.nan_to_num(np.asarray(lm) - np.asarray(lu))
    per_reg = np.add.reduceat(llr, np.concatenate([[0], np.cumsum(np.diff(read_off) * np.diff(grp_off))])[:-1])
    phi[:, 1] += 1e-3 * per_reg + np.add.reduceat(np.asarray(Nl) * np.asarray(rho_l), grp_off[:-1])
    pick = np.zeros(R, dtype=np.int32)
    pick[phi[:, 0] > 0.8] = -1
    nk = np.diff(sub_off)
    return dict(phi_hat=phi, pick=pick, ex=np.tanh(np.asarray(rho_n) * 10) * np.repeat(phi[:, 0], np.diff(cpg_off)), exx=np.tanh(np.asarray(d_n) / 100),
                mml=np.repeat(phi[:, 1], nk) * 0.01 + np.asarray(sub_p) * 1e-3, nme=np.repeat(phi[:, 2], nk) * 0.1 + np.asarray(sub_q) * 1e-3)


def _starts(name, st):
    return np.stack([np.random.default_rng(int(s)).uniform(-1, 1, (3, 3)) for s in st])


class _Eng:
    def __init__(self, lib):
        self.lib, self.device = lib, 0


def _run(out_dir, rank, world, block):
    import cpelnano_b200 as cpn
    lib = cpn.Library()
    lib.analyze_batch = _stand_in
    cfg = cpn.CpelNanoConfig(max_em_init=3)
    return cpn.analyze_nano_csr(os.path.join(INP, "full_example_calls.tsv.gz"), os.path.join(INP, "full_example.fa.gz"), cfg, out_dir, "x",
                                engine=_Eng(lib), phi0_by_region=_starts, rank=rank, world=world, block_regions=block)


def _worker(rank, world, port, out_dir, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import cpelnano_b200 as cpn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = _run(out_dir, rank, world, 7)
    dist.barrier()                                   # every rank's part files are on disk
    if rank == 0:
        cpn.merge_shards(out_dir, "x", world)
    dist.barrier()
    q.put((rank, st))
    dist.destroy_process_group()


def test_two_rank_file_pipeline_equals_single_process(tmp_path):
    one = str(tmp_path / "one")
    s1 = _run(one, 0, 1, 50_000)
    assert s1["with_data"] == 27 and 0 < s1["fitted"] <= 27
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    two = str(tmp_path / "two")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, two, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tot = {k: sum(st[k] for _, st in got) for k in s1}
    assert tot == s1                                             # every region analysed exactly once
    for suffix in ("theta", "ex", "exx", "mml", "nme"):
        a, b = open(os.path.join(one, f"x_{suffix}.txt"), "rb").read(), open(os.path.join(two, f"x_{suffix}.txt"), "rb").read()
        assert a == b and len(a) > 0, suffix
    assert not os.path.exists(os.path.join(two, "_shards"))
    # a second run refuses to overwrite (check_output_exists, src/input_output.jl:36-58)
    assert _run(one, 0, 1, 50_000) is None
