"""diff_grp_comp_csr (arrays end to end) against diff_grp_comp (per-region structs) on CPU: the device call
cpn_grp_test_batch is replaced by a stand-in that depends on every array it is handed (ϕ̂ of the samples in order, region
features, sub-region tables, labelings), so equal output files mean the two host paths hand the device identical batches and
order, round and print its results identically — including samples that lack some regions (batches per (n1, n2)), matched and
unmatched.  With the real kernels both paths are compared with the reference's golden files in tests/test_pipeline_gpu.py.
"""
import os
import tarfile

import numpy as np
import pytest

import cpelnano_b200 as cpn
from cpelnano_b200 import pipeline

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")


def _stand_in(n1, n2, matched, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, lab, exact, device=0):
    phi = np.asarray(phi); R = len(cpg_off) - 1; K = np.diff(sub_off); ns = 2 if matched else 3
    sK = int(sub_off[-1])
    base = phi.reshape(R, -1) @ np.linspace(0.1, 1.0, phi.shape[1] * 3) + np.add.reduceat(np.asarray(rho_n), cpg_off[:-1]) + \
        np.add.reduceat(np.asarray(d_n), (cpg_off[:-1] - np.arange(R))) * 1e-4 + len(lab) * 1e-3 + n1 * 0.01 + n2 * 0.001 + (0.5 if exact else 0.0)
    T = np.empty(ns * sK); p = np.empty(ns * sK); cnt = np.zeros(ns * sK, dtype=np.int64)
    for r in range(R):
        for j in range(ns):
            sl = slice(ns * sub_off[r] + j * K[r], ns * sub_off[r] + (j + 1) * K[r])
            T[sl] = np.sin(base[r] * (j + 1) + np.asarray(sub_p)[sub_off[r]:sub_off[r + 1]] * 0.1)
            p[sl] = (np.asarray(sub_q)[sub_off[r]:sub_off[r + 1]] % 7 + j) / 10.0
    T[np.abs(T) > 0.97] = np.nan
    tcmd = np.abs(np.cos(np.repeat(base, K)))
    return (T, cnt, p, tcmd) if matched else (T, cnt, p)


@pytest.fixture(scope="module")
def model_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("mods")
    with tarfile.open(os.path.join(INP, "golden_outputs.tar.gz")) as tar:
        tar.extractall(d, filter="data")
    g1 = [f"{d}/grp_comp_example/mod_files/g1_rep{i}.txt" for i in range(1, 7)]
    g2 = [f"{d}/grp_comp_example/mod_files/g2_rep{i}.txt" for i in range(1, 7)]

    def drop(src, name, which):
        lines = open(src).read().splitlines(True)
        dst = f"{d}/{name}"
        open(dst, "w").write("".join(l for i, l in enumerate(lines) if i not in which))
        return dst

    rag1 = [g1[0], drop(g1[1], "g1b.txt", {0, 5, 6}), g1[2], g1[3], drop(g1[4], "g1e.txt", {26}), g1[5]]
    rag2 = [g2[0], g2[1], g2[2], drop(g2[3], "g2d.txt", {5, 20}), g2[4], g2[5]]
    return {"full": (g1, g2), "ragged": (rag1, rag2)}


@pytest.mark.parametrize("which", ["full", "ragged"])
@pytest.mark.parametrize("matched", [False, True])
def test_csr_path_hands_the_device_the_same_batches(model_files, tmp_path, which, matched):
    lib = cpn.Library()
    lib.grp_test_batch = _stand_in
    eng = object.__new__(cpn.Engine)
    eng.lib, eng.device = lib, 0
    g1, g2 = model_files[which]
    fa = os.path.join(INP, "full_example.fa.gz")
    cfg = cpn.CpelNanoConfig(matched=matched)
    f1 = pipeline.diff_grp_comp(g1, g2, fa, cfg, str(tmp_path / "a"), "x", engine=eng)
    f2 = pipeline.diff_grp_comp_csr(g1, g2, fa, cfg, str(tmp_path / "b"), "x", engine=eng, block_regions=10)
    for x, y in zip(f1.all(), f2.all()):
        assert open(x).read() == open(y).read() and os.path.getsize(x) > 0, os.path.basename(x)
    assert pipeline.diff_grp_comp_csr(g1, g2, fa, cfg, str(tmp_path / "b"), "x", engine=eng) is None      # existing outputs stop the run
