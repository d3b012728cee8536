"""Native calls loader (csrc/cpn_calls.cpp behind cpn_calls_* of the C ABI) — host-only, runs without a GPU.

Bit-exact against (i) the numpy implementation (nanopolish.NanopolishCalls) and (ii) the line-by-line restatement of the
reference's get_dic_nanopolish + get_calls_nanopolish! (src/input_output.jl:536-570, src/nanopore.jl:172-230) on the bundled
full_example calls, plus the edge cases the reference's semantics define: unsorted files, several chromosomes, repeated
(read, group) lines (the last one wins), calls that match no group, reads with no matching call (still a row), hard
emissions (mod_nse = false), header-only and malformed files.
"""
import gzip
import os

import numpy as np
import pytest

import calls_ref            # oracle/calls_ref.py (conftest puts oracle/ on sys.path)
import cpelnano_b200 as cpn
from cpelnano_b200 import genome, nanopolish

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")
TSV = os.path.join(INP, "full_example_calls.tsv.gz")


@pytest.fixture(scope="module")
def lib():
    return cpn.Library()


@pytest.fixture(scope="module")
def regions():
    ch = genome.load_genome(os.path.join(INP, "full_example.fa.gz"))["hg38_dna"]
    return [ch.region_struct(a, b) for (a, b) in ch.chr_part(3000, 10)]


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("mod_nse", [True, False])
def test_native_equals_numpy_and_restatement_on_bundled_calls(lib, regions, mod_nse):
    nat = nanopolish.NativeCalls(TSV, lib)
    ref = nanopolish.NanopolishCalls(TSV)
    assert nat.n_lines == ref.n_lines == 13176 and nat.handle.n_reads == 13 - 1     # 12 reads + header column name excluded
    for i, rs in enumerate(regions):
        a = nat.region_calls(rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        b = ref.region_calls(rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        assert _same(a[0], b[0]) and _same(a[1], b[1]) and list(a[2]) == list(b[2]), i
    for i in (7, 20):
        rs = regions[i]
        a = nat.region_calls(rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        line = calls_ref.region_calls_reference(TSV, rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        assert set(a[2]) == set(line)
        for k, nm in enumerate(a[2]):
            assert np.array_equal(a[0][k], line[nm][0], equal_nan=True) and np.array_equal(a[1][k], line[nm][1], equal_nan=True)


def test_batch_fill_equals_per_region_and_is_thread_independent(lib, regions, tmp_path):
    plain = tmp_path / "calls.tsv"
    with gzip.open(TSV, "rb") as f:
        plain.write_bytes(f.read())
    outs = []
    for path, threads in ((TSV, 1), (str(plain), 1), (str(plain), 8), (str(plain), 3)):
        nat = nanopolish.NativeCalls(path, lib, n_threads=threads)
        regs = [genome.RegStruct(chr=r.chr, chrst=r.chrst, chrend=r.chrend) for r in regions]
        for a, r in zip(regs, regions):
            a.grp_st, a.grp_end, a.L = r.grp_st, r.grp_end, r.L
        nat.attach_calls_batch(regs)
        outs.append(regs)
        one = nanopolish.NativeCalls(path, lib, n_threads=1)
        for a in regs[::5]:
            lu, lm, _ = one.region_calls(a.chr, a.chrst, a.chrend, a.grp_st, a.grp_end)
            assert _same(lu, a.log_pyx_u) and _same(lm, a.log_pyx_m) and a.m == lu.shape[0]
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert _same(a.log_pyx_u, b.log_pyx_u) and _same(a.log_pyx_m, b.log_pyx_m)
    # the CSR arrays go straight into cpn_fit_batch: offsets are consistent with the per-region shapes
    h = nanopolish.NativeCalls(TSV, lib).handle
    grp_off = np.concatenate([[0], np.cumsum([r.L for r in regions])])
    read_off, lu, lm, ids = h.fill("hg38_dna", [r.chrst for r in regions], [r.chrend for r in regions], grp_off,
                                   np.concatenate([r.grp_st for r in regions]), np.concatenate([r.grp_end for r in regions]), want_ids=True)
    assert len(lu) == len(lm) == int(np.sum(np.diff(read_off) * np.diff(grp_off))) and len(ids) == read_off[-1]
    assert np.array_equal(np.isnan(lu), np.isnan(lm))
    assert np.array_equal(np.diff(read_off), h.count_reads("hg38_dna", [r.chrst for r in regions], [r.chrend for r in regions]))


HDR = "chromosome\tstrand\tstart\tend\tread_name\tlog_lik_ratio\tlog_lik_methylated\tlog_lik_unmethylated\tnum_calling_strands\tnum_motifs\tsequence\n"


def _line(chrom, start, end, read, lm, lu, extra=True):
    s = f"{chrom}\t+\t{start}\t{end}\t{read}\t{lm - lu:.2f}\t{lm:.2f}\t{lu:.2f}"
    return s + ("\t1\t1\tACGT\n" if extra else "\n")


def test_semantics_on_a_hand_made_file(lib, tmp_path):
    # groups of the region [100, 400] on chrA: CpG groups with intervals [105,125] and [195,205] (grp_st = first CpG − 5 …)
    grp_st, grp_end = np.array([105, 195]), np.array([125, 205])
    # a call matches group g when start + 1 − 5 == grp_st and end + 1 + 5 == grp_end  →  (start, end) = (109, 119) and (199, 199)
    lines = [
        _line("chrB", 109, 119, "rB", -1.0, -2.0),                 # other chromosome
        _line("chrA", 199, 199, "r2", -3.0, -4.0),                 # file not sorted by start: sorted stably on load
        _line("chrA", 109, 119, "r1", -5.25, -6.5),
        _line("chrA", 109, 119, "r2", -7.0, -8.0),
        _line("chrA", 109, 119, "r1", -9.0, -1.5, extra=False),    # same (read, group) again: the later line wins; 8 columns only
        _line("chrA", 150, 160, "r3", -1.0, -1.0),                 # overlaps the region, matches no group: r3 is an all-NaN row
        _line("chrA", 109, 121, "r1", -2.0, -2.0),                 # start matches group 0 but end does not: skipped
        _line("chrA", 20, 99, "r4", -1.0, -1.0),                   # ends before the region starts: not in the region
        _line("chrA", 401, 410, "r5", -1.0, -1.0),                 # starts after the region ends
        _line("chrA", 90, 100, "r6", -1.0, -3.0),                  # end == chrst: overlaps (end ≥ chrst)
    ]
    path = tmp_path / "hand.tsv"
    path.write_text(HDR + "".join(lines))
    for cls in (lambda p: nanopolish.NativeCalls(p, lib), nanopolish.NanopolishCalls):
        c = cls(str(path))
        lu, lm, names = c.region_calls("chrA", 100, 400, grp_st, grp_end)
        assert list(names) == ["r6", "r1", "r2", "r3"]                                     # first appearance after the stable sort by start
        want_lm = np.array([[np.nan, np.nan], [-9.0, np.nan], [-7.0, -3.0], [np.nan, np.nan]])
        want_lu = np.array([[np.nan, np.nan], [-1.5, np.nan], [-8.0, -4.0], [np.nan, np.nan]])
        assert _same(lm, want_lm) and _same(lu, want_lu)
        lu, lm, _ = c.region_calls("chrA", 100, 400, grp_st, grp_end, mod_nse=False)       # hard emissions (src/CpelNano.jl:32-33)
        assert _same(lm, np.array([[np.nan, np.nan], [-50.0, np.nan], [0.0, 0.0], [np.nan, np.nan]]))
        assert _same(lu, np.array([[np.nan, np.nan], [0.0, np.nan], [-50.0, -50.0], [np.nan, np.nan]]))
        lu, lm, names = c.region_calls("chrC", 100, 400, grp_st, grp_end)
        assert lu.shape == (0, 2) and len(names) == 0
        lu, lm, names = c.region_calls("chrB", 100, 400, grp_st, grp_end)
        assert list(names) == ["rB"] and lm[0, 0] == -1.0
    h = nanopolish.NativeCalls(str(path), lib).handle
    assert h.chromosomes == {"chrB": 1, "chrA": 9} and h.n_lines == 10 and h.n_reads == 7
    # CRLF line endings and a header-only file
    crlf = tmp_path / "crlf.tsv"
    crlf.write_bytes((HDR + "".join(lines)).replace("\n", "\r\n").encode())
    lu2, lm2, _ = nanopolish.NativeCalls(str(crlf), lib).region_calls("chrA", 100, 400, grp_st, grp_end)
    assert _same(lm2, want_lm)
    empty = tmp_path / "empty.tsv"
    empty.write_text(HDR)
    e = nanopolish.NativeCalls(str(empty), lib)
    assert e.n_lines == 0 and e.region_calls("chrA", 1, 10, grp_st, grp_end)[0].shape == (0, 2)


def test_errors_are_reported_not_swallowed(lib, tmp_path):
    with pytest.raises(cpn.CpnError, match="cannot open"):
        nanopolish.NativeCalls(str(tmp_path / "missing.tsv"), lib)
    bad = tmp_path / "bad.tsv"
    bad.write_text(HDR + _line("chrA", 1, 2, "r", -1.0, -2.0) + "chrA\t+\tnot_a_number\t5\tr\t0\t-1\t-2\n")
    with pytest.raises(cpn.CpnError, match="malformed line 3"):
        nanopolish.NativeCalls(str(bad), lib)
    short = tmp_path / "short.tsv"
    short.write_text(HDR + "chrA\t+\t1\t2\tr\n")
    with pytest.raises(cpn.CpnError, match="malformed line 2"):
        nanopolish.NativeCalls(str(short), lib)
    # a read_off that does not belong to the regions is refused, not written out of bounds
    h = nanopolish.NativeCalls(TSV, lib).handle
    with pytest.raises(cpn.CpnError, match="read_off"):
        h.fill("hg38_dna", [20988], [23987], [0, 3], [1, 2, 3], [4, 5, 6], read_off=np.array([0, 1]))


def test_decimal_text_is_correctly_rounded(lib, tmp_path):
    """from_chars must agree with Python's float() (both correctly rounded, like Julia's parse) on awkward decimals."""
    vals = ["-134.40", "0.1", "-0.30000000000000004", "1e-320", "123456789.123456789", "-2.2250738585072014e-308", "17.005", "-99.995"]
    path = tmp_path / "dec.tsv"
    path.write_text(HDR + "".join(f"chrA\t+\t{109}\t{119}\tr{i}\t0\t{v}\t{v}\n" for i, v in enumerate(vals)))
    lu, lm, names = nanopolish.NativeCalls(str(path), lib).region_calls("chrA", 100, 400, np.array([105]), np.array([125]))
    assert [float(v) for v in vals] == list(lm[:, 0]) == list(lu[:, 0])
