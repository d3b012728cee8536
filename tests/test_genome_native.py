"""Native genome features (csrc/cpn_genome.cpp behind cpn_cpg_scan / cpn_chr_part / cpn_grp_info_*) — host-only, no GPU.

Three implementations must agree bit for bit (integers AND doubles): the native batch code, the numpy code in genome.py, and
the regex-per-window restatement of the reference used to build the golden-pinned fixtures (tests/golden/make_fixtures.py:
get_grp_info! src/genome_analysis.jl:226-277, get_chr_part :327-458).  Inputs: the two bundled hg38 windows and random
sequences built to hit the edge cases: CpG-poor sequence (windows with zero or exactly one CpG — whose group interval the
reference leaves unpadded), CpG islands, lower-case runs, N runs, regions at both chromosome ends, regions with N ≤ 1.
"""
import os
import sys

import numpy as np
import pytest

import cpelnano_b200 as cpn
from cpelnano_b200 import genome, capi

HERE = os.path.dirname(os.path.abspath(__file__))
INP = os.path.join(HERE, "golden", "inputs")
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_fixtures as mf  # noqa: E402   (regex restatement; only its pure functions are used)


@pytest.fixture(scope="module")
def lib():
    return cpn.Library()


def _random_seq(rng, n, p_cg, lower=0.2):
    s = rng.choice(list("ACGTN"), size=n, p=[0.3, 0.2, 0.2, 0.29, 0.01])
    # break accidental CGs, then plant them at rate p_cg, partly in dense islands
    for i in np.nonzero((s[:-1] == "C") & (s[1:] == "G"))[0]:
        s[i + 1] = "A"
    k = int(n * p_cg)
    pos = rng.integers(0, n - 1, k)
    isl = rng.integers(0, n - 400)
    pos = np.concatenate([pos, isl + rng.integers(0, 300, 40)])
    for i in pos:
        s[i], s[i + 1] = "C", "G"
    low = rng.random(n) < lower
    s = np.where(low, np.char.lower(s), s)
    return "".join(s)


def _check_all(lib, seq_str, size_est_reg, targets=None):
    seq = np.frombuffer(seq_str.encode(), dtype=np.uint8)
    ch_np = genome.Chromosome("c", seq)
    ch_nat = genome.Chromosome("c", seq, lib)
    assert np.array_equal(ch_np.cg, ch_nat.cg)
    part_np = ch_np.chr_part(size_est_reg, 10, targets)
    part_nat = ch_nat.chr_part_native(lib, size_est_reg, 10, targets)
    if targets is None:
        part_re = mf.chr_part_full(seq_str, size_est_reg, 10)
    else:
        part_re = [w for (a, b) in targets for w in mf.chr_part_targeted(seq_str, a, b, size_est_reg, 10)]
    assert part_np == part_nat == [(int(a), int(b)) for a, b in part_re]
    regs = ch_nat.region_structs(lib, part_nat, 10)
    for (a, b), rs in zip(part_nat, regs):
        ref = mf.grp_info(seq_str, a, b, 10)
        alt = ch_np.grp_info(a, b, 10)
        assert np.array_equal(rs.cpg_pos, ref["cpg_pos"]) and rs.N == ref["N"] == alt["N"]
        if ref["N"] <= 1:
            assert rs.L == 0
            continue
        for mine, key in ((rs.rho_n, "rho_n"), (rs.dn, "d_n"), (rs.Nl, "Nl"), (rs.rho_l, "rho_l"), (rs.dl, "d_l")):
            assert np.array_equal(mine, alt[key]), key                     # bit-identical doubles
            if key != "rho_l":
                assert np.array_equal(mine, ref[key]), key
        # ρ_l = sum(ρ_n[group]) / N_l: Julia's sum adds strictly left to right below 16 elements (Base._mapreduce); the fixture
        # script uses numpy's pairwise sum, which only agrees for short groups — compare with an explicit left-to-right sum
        for g, (a0, b0) in enumerate(zip(ref["grp_first"], ref["grp_last"])):
            acc = ref["rho_n"][a0]
            for i in range(a0 + 1, b0 + 1):
                acc = acc + ref["rho_n"][i]
            assert rs.rho_l[g] == acc / (b0 - a0 + 1)
            if b0 - a0 + 1 < 8:
                assert rs.rho_l[g] == ref["rho_l"][g]
        assert np.array_equal(rs.grp_st, alt["grp_st"]) and np.array_equal(rs.grp_end, alt["grp_end"])
        assert np.array_equal(rs.grp_st, ref["cpg_pos"][ref["grp_first"]] - 5) and np.array_equal(rs.grp_end, ref["cpg_pos"][ref["grp_last"]] + 5)
    return len(part_nat)


def test_bundled_windows(lib):
    for fa in ("full_example.fa.gz", "targeted_example.fa.gz"):
        seq = genome.read_fasta(os.path.join(INP, fa))["hg38_dna"].tobytes().decode()
        assert _check_all(lib, seq, 3000) > 40
    seq = genome.read_fasta(os.path.join(INP, "targeted_example.fa.gz"))["hg38_dna"].tobytes().decode()
    assert _check_all(lib, seq, 3000, genome.read_bed(os.path.join(INP, "brca1.bed"))["hg38_dna"]) == 34


@pytest.mark.parametrize("seed,p_cg,size", [(1, 0.02, 3000), (2, 0.0004, 3000), (3, 0.0002, 1500), (4, 0.05, 3000), (5, 0.001, 2000), (6, 0.0, 3000)])
def test_random_sequences_incl_lone_cpg_windows(lib, seed, p_cg, size):
    rng = np.random.default_rng(seed)
    seq = _random_seq(rng, 60_000 + int(rng.integers(0, 2999)), p_cg)
    _check_all(lib, seq, size)
    _check_all(lib, seq, size, [(1200, 9100), (20_000, 20_001), (30_000, len(seq))])


def test_lone_cpg_interval_is_unpadded(lib):
    """A window holding exactly ONE CpG, 3 bp left of the candidate boundary: with the ±5 padding every other group gets, the
    boundary would be moved; the reference returns a lone CpG's interval unpadded (genome_analysis.jl:67) and leaves it."""
    s = ["A"] * 8000
    s[2995], s[2996] = "C", "G"                  # 1-based position 2996, boundary candidate at 3000
    seq = "".join(s)
    ch = genome.Chromosome("c", np.frombuffer(seq.encode(), dtype=np.uint8), lib)
    assert ch.chr_part_native(lib, 3000, 10)[0] == ch.chr_part(3000, 10)[0] == tuple(mf.chr_part_full(seq, 3000, 10)[0]) == (1, 3000)
    s[2990], s[2991] = "C", "G"                  # a second CpG joins the group: now padded, the boundary moves past it
    seq = "".join(s)
    ch = genome.Chromosome("c", np.frombuffer(seq.encode(), dtype=np.uint8), lib)
    assert ch.chr_part_native(lib, 3000, 10)[0] == ch.chr_part(3000, 10)[0] == tuple(mf.chr_part_full(seq, 3000, 10)[0]) == (1, 3003)


def test_scan_edge_cases(lib):
    for s, want in (("", []), ("C", []), ("CG", [1]), ("cg", [1]), ("cGCg", [1, 3]), ("CCGG", [2]), ("GC", []), ("ACGCGT" * 3, [2, 4, 8, 10, 14, 16])):
        assert capi.cpg_scan(lib, np.frombuffer(s.encode(), dtype=np.uint8)).tolist() == want
    big = np.frombuffer(("ACGT" * 1_000_003).encode(), dtype=np.uint8)           # several threads, cut points inside CGs
    got = capi.cpg_scan(lib, big)
    assert np.array_equal(got, genome.cpg_starts(big)) and len(got) == 1_000_003
    o = capi.grp_info_batch(lib, got[:0], 100, [1], [50])                         # chromosome without CpGs
    assert o["cpg_off"].tolist() == [0, 0] and o["grp_off"].tolist() == [0, 0]


def test_native_fasta_loader_equals_python_reader(lib, tmp_path):
    for fa in ("full_example.fa.gz", "targeted_example.fa.gz"):
        p = os.path.join(INP, fa)
        a, b = genome.read_fasta(p), genome.load_genome(p, lib)
        assert list(a) == list(b) and all(np.array_equal(a[k], b[k].seq) for k in a)
    p = tmp_path / "multi.fa"
    p.write_bytes(b">chr1 some description\r\nACGT\r\nacgn\r\n>chr2\tx\nCG\n\nCGCG\n>empty\n>last\nTTTT")      # CRLF, blank line, no final newline
    a, b = genome.read_fasta(str(p)), genome.load_genome(str(p), lib)
    assert {k: v.tobytes() for k, v in a.items()} == {k: v.seq.tobytes() for k, v in b.items()} == \
        {"chr1": b"ACGTacgn", "chr2": b"CGCGCG", "empty": b"", "last": b"TTTT"}
    assert b["chr2"].cg.tolist() == [1, 3, 5] and b["empty"].cg.tolist() == []
    with pytest.raises(cpn.CpnError, match="cannot open"):
        capi.FastaHandle(lib, str(tmp_path / "missing.fa"))
