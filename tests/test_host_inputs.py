"""Host-side inputs of the path: FASTA → regions → CpG features, nanopolish TSV → emissions.

Pins: the committed fixtures tests/golden/{targeted,full}_example.npz were produced by a regex-per-window restatement of the
reference (tests/golden/make_fixtures.py) and reproduce the golden theta files' α, β bit-exactly (test_oracle_golden.py); the
vectorised implementation in cpelnano.jl_b200/genome.py must give the same integers and the same doubles.  Calls ingestion is
checked against the line-by-line restatement in oracle/calls_ref.py of get_dic_nanopolish + get_calls_nanopolish! (src/input_output.jl:536-570,
src/nanopore.jl:172-230) on the bundled full_example calls.
"""
import os

import numpy as np
import pytest

import calls_ref            # oracle/calls_ref.py (conftest puts oracle/ on sys.path)
import cpelnano_b200 as cpn
from cpelnano_b200 import genome, nanopolish

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INP = os.path.join(GOLD, "inputs")


@pytest.fixture(scope="module")
def chr_targeted():
    return genome.load_genome(os.path.join(INP, "targeted_example.fa.gz"))["hg38_dna"]


@pytest.fixture(scope="module")
def chr_full():
    return genome.load_genome(os.path.join(INP, "full_example.fa.gz"))["hg38_dna"]


@pytest.fixture(scope="module")
def calls_full():
    return nanopolish.NanopolishCalls(os.path.join(INP, "full_example_calls.tsv.gz"))


def _check_features(chrom, fx):
    for r in range(len(fx["chrst"])):
        gi = chrom.grp_info(int(fx["chrst"][r]), int(fx["chrend"][r]))
        n0, n1 = fx["cpg_off"][r], fx["cpg_off"][r + 1]
        g0, g1 = fx["grp_off"][r], fx["grp_off"][r + 1]
        assert np.array_equal(gi["cpg_pos"], fx["cpg_pos"][n0:n1])
        assert np.array_equal(gi["rho_n"], fx["rho_n"][n0:n1])                      # bit-exact doubles
        assert np.array_equal(gi["d_n"], fx["d_n"][n0 - r:n1 - r - 1])
        assert np.array_equal(gi["Nl"], fx["Nl"][g0:g1])
        assert np.array_equal(gi["rho_l"], fx["rho_l"][g0:g1])
        assert np.array_equal(gi["d_l"], fx["d_l"][g0 - r:g1 - r - 1])


def test_region_features_bit_exact_targeted(chr_targeted):
    _check_features(chr_targeted, np.load(os.path.join(GOLD, "targeted_example.npz")))


def test_region_features_bit_exact_full(chr_full):
    _check_features(chr_full, np.load(os.path.join(GOLD, "full_example.npz")))


def test_partition_bed_contains_golden_regions(chr_targeted):
    """get_chr_part with brca1.bed: every region of the golden theta file is a window of the partition, windows tile the
    BED interval without gaps, and no boundary cuts a CpG group (genome_analysis.jl:391-458)."""
    targets = genome.read_bed(os.path.join(INP, "brca1.bed"))["hg38_dna"]
    part = chr_targeted.chr_part(3000, 10, targets)
    fx = np.load(os.path.join(GOLD, "targeted_example.npz"))
    assert {(int(a), int(b)) for a, b in zip(fx["chrst"], fx["chrend"])} <= set(part)
    assert part[0][0] == 155510 and part[-1][1] == 255510
    for (a, b), (c, d) in zip(part[:-1], part[1:]):
        assert c == b + 1
    cg = chr_targeted.cg
    first, last = genome.cpg_groups(cg, 10)
    gst, gend = cg[first] - 5, cg[last] + 5
    for (_, b) in part[:-1]:
        assert not np.any((gst <= b) & (b <= gend)), b


def test_partition_whole_chromosome_matches_layout_fixture(chr_targeted):
    """Whole-chromosome partition (genome_analysis.jl:327-390) + the N>9, L>1 filters of pmap_analyze_reg reproduce the
    137 layouts the synthetic workload cycles through."""
    lay = np.load(os.path.join(ROOT, "cpelnano.jl_b200", "data", "layouts_targeted.npz"))
    part = chr_targeted.chr_part(3000, 10)
    keep = []
    for a, b in part:
        gi = chr_targeted.grp_info(a, b)
        if gi["N"] > 9 and len(gi["Nl"]) > 1:
            keep.append((a, b))
    assert keep == list(zip(lay["chrst"].tolist(), lay["chrend"].tolist()))
    _check_features(chr_targeted, lay)


def test_cpg_scan_equals_regex(chr_full):
    import re
    s = chr_full.seq.tobytes().decode()
    assert np.array_equal(chr_full.cg, np.array([m.start() + 1 for m in re.finditer(r"[Cc][Gg]", s)]))
    # window queries = regex on the substring (a CG cut by the window end is not a match)
    rng = np.random.default_rng(0)
    for _ in range(200):
        st = int(rng.integers(1, chr_full.size - 10)); en = int(min(chr_full.size, st + rng.integers(0, 1200)))
        lo, hi = chr_full.cpgs_in(st, en)
        assert hi - lo == len(re.findall(r"[Cc][Gg]", s[st - 1:en]))


@pytest.mark.parametrize("mod_nse", [True, False])
def test_calls_ingestion_matches_line_by_line_restatement(chr_full, calls_full, mod_nse):
    fx = np.load(os.path.join(GOLD, "full_example.npz"))
    path = os.path.join(INP, "full_example_calls.tsv.gz")
    for r in (0, 5, 13, 26):
        rs = chr_full.region_struct(int(fx["chrst"][r]), int(fx["chrend"][r]))
        lu, lm, names = calls_full.region_calls(rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        ref = calls_ref.region_calls_reference(path, rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        assert set(names) == set(ref) and len(names) == len(ref)
        for i, nm in enumerate(names):
            assert np.array_equal(lu[i], ref[nm][0], equal_nan=True)
            assert np.array_equal(lm[i], ref[nm][1], equal_nan=True)
        assert lu.shape[0] >= 10 and np.isfinite(lu).sum() > 0


def test_calls_edge_cases(chr_full, calls_full):
    rs = chr_full.region_struct(1, 3000)
    lu, lm, names = calls_full.region_calls("no_such_chr", rs.chrst, rs.chrend, rs.grp_st, rs.grp_end)
    assert lu.shape == (0, rs.L) and len(names) == 0
    # region with no overlapping line
    c = calls_full.by_chr["hg38_dna"]
    far = int(c["end"].max()) + 10_000
    lu, lm, names = calls_full.region_calls("hg38_dna", far, far + 3000, rs.grp_st, rs.grp_end)
    assert lu.shape == (0, rs.L)
    # attach + the coverage filters of pmap_analyze_reg (structures.jl:200-203)
    calls_full.attach_calls(rs)
    assert rs.m == rs.log_pyx_u.shape[0] and 0.0 <= rs.perc_gprs_obs() <= 1.0
    assert rs.get_ave_depth() == np.isfinite(rs.log_pyx_u).sum() / rs.L
