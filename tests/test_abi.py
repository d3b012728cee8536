"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/cpelnano_b200.h declares, refuses to
compute without a GPU (no fallback), and its host-only integer routine is bit-exact against the oracle and goldens."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, region_slices


def test_header_symbols_exported(cpn):
    hdr = open(os.path.join(ROOT, "include", "cpelnano_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cpn_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(cpn.capi.SYMBOLS)
    lib = cpn.Library()
    for s in declared:
        assert hasattr(lib.dll, s), s
    assert "sm_100a" in lib.version()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cpelnano.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in txt and "cpn_oracle" not in txt and "libcpn_oracle" not in txt, f


def test_no_gpu_means_loud_failure(cpn):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(cpn.CpnError):
        cpn.Engine(0)


@pytest.mark.parametrize("name", ["full_example", "targeted_example", "grp_comp_example"])
def test_nls_reg_info_bit_exact(cpn, orc, name):
    """cpn_nls_reg_info (host, integer) == oracle restatement of genome_analysis.jl:472-538 on every golden region,
    and the intervals are exactly those written in the golden MML files."""
    fx = load_golden(name)
    lib = cpn.Library()
    mss = int(fx["max_size_subreg"])
    gold = {(int(r[0]), int(r[1])) for r in fx["mml"]} if "mml" in fx else None
    for r in range(len(fx["chrst"])):
        s = region_slices(fx, r)
        a = lib.nls_reg_info(int(fx["chrst"][r]), int(fx["chrend"][r]), mss, s["pos"])
        b = orc.get_nls_reg_info(int(fx["chrst"][r]), int(fx["chrend"][r]), mss, s["pos"])
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        if gold is not None:
            for k in range(len(a[0])):
                if a[2][k] != 0:
                    assert (int(a[0][k]), int(a[1][k])) in gold


def test_nls_reg_info_edge_cases(cpn, orc):
    lib = cpn.Library()
    rng = np.random.default_rng(0)
    for _ in range(300):
        st = int(rng.integers(1, 10_000))
        size = int(rng.integers(1, 5000))
        mss = int(rng.integers(1, 700))
        n = int(rng.integers(0, 40))
        pos = np.sort(rng.choice(np.arange(st, st + size), size=min(n, size), replace=False)).astype(np.int64)
        a = lib.nls_reg_info(st, st + size - 1, mss, pos)
        b = orc.get_nls_reg_info(st, st + size - 1, mss, pos)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # CpG exactly on the right end belongs to the last sub-region; empty sub-regions are 0:0; last one 0:N when empty
    a = lib.nls_reg_info(1, 3000, 350, np.array([25, 50, 400, 2999, 3000]))
    assert list(a[2]) == [1, 3, 0, 0, 0, 0, 0, 0, 4] and list(a[3]) == [2, 3, 0, 0, 0, 0, 0, 0, 5]
    a = lib.nls_reg_info(1, 3000, 350, np.array([25, 50]))
    assert a[2][-1] == 0 and a[3][-1] == 2
