"""WGBS front-end, BAM side: the native reader csrc/cpn_bam.cpp (get_calls_bam for all regions of a
chromosome) against oracle/wgbs_py.py (line-by-line restatement of src/wgbs.jl:23-323) on the reference's example BAM files
and on synthetic single- and paired-end files; analyze_bam on the reference's example.  CPU tests except the last one."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import bam_writer as bw       # noqa: E402
import wgbs_py as orc         # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "wgbs")


@pytest.fixture(scope="module")
def cpn():
    import cpelnano_b200
    return cpelnano_b200


@pytest.fixture(scope="module")
def lib(cpn):
    return cpn.Library()


def native_calls(cpn, lib, path, chrom, regions, cpgs, pe=False, trim=(0, 0, 0, 0)):
    """regions [(st, end)], cpgs [array per region] → list of int8 [m, N] tables"""
    h = cpn.capi.BamHandle(lib, path)
    cpg_off = np.concatenate([[0], np.cumsum([len(c) for c in cpgs])]).astype(np.int64)
    flat_cpg = np.concatenate(cpgs) if cpgs else np.zeros(0, dtype=np.int64)
    read_off, flat = h.calls(chrom, [r[0] for r in regions], [r[1] for r in regions], cpg_off, flat_cpg, pe, trim)
    n = h.count_reads(chrom, [r[0] for r in regions], [r[1] for r in regions], cpg_off, flat_cpg, pe, trim, n_threads=1)
    assert np.array_equal(np.diff(read_off), n)
    out, o = [], 0
    for i, c in enumerate(cpgs):
        m = int(read_off[i + 1] - read_off[i])
        out.append(flat[o:o + m * len(c)].reshape(m, len(c)))
        o += m * len(c)
    assert o == flat.size
    return out


def oracle_calls(bam, chrom, regions, cpgs, pe=False, trim=(0, 0, 0, 0)):
    out = []
    for (st, en), c in zip(regions, cpgs):
        m, calls = orc.get_calls_bam(bam, chrom, st, en, [int(x) for x in c], pe, trim)
        assert m == len(calls)
        out.append(np.array(calls, dtype=np.int8).reshape(m, len(c)))
    return out


def ref_cpgs(st, en):
    seq = "".join(l.strip() for l in open(os.path.join(GOLD, "ref.fa")) if not l.startswith(">"))
    return np.array([m.start() + st for m in re.finditer("[Cc][Gg]", seq[st - 1:en])], dtype=np.int64)


def test_docstring_example():
    assert orc.get_align_strand(True, 99, 147) == "OT"          # src/wgbs.jl:18-20


@pytest.mark.parametrize("name", ["meth_sample", "unmeth_sample"])
def test_reader_matches_the_specification_restatement(cpn, lib, name):
    path = os.path.join(GOLD, name + ".bam")
    refs, recs = orc.read_bam(path)
    h = cpn.capi.BamHandle(lib, path)
    assert list(h.refs.items()) == refs and h.n_records == len(recs) == 1600
    for i, r in enumerate(recs):
        g = h.record(i)
        assert (g["ref"], g["pos"], g["rightpos"], g["flag"], g["mapq"], g["qname"]) == (r["ref"], r["pos"], r["rightpos"], r["flag"], r["mapq"], r["qname"])
        assert g["has_xm"] == ("XM" in r["tags"]) and g["has_xs"] == ("XS" in r["tags"]) and g["xm"] == r["tags"].get("XM", "")
    h.close()
    with pytest.raises(cpn.CpnError):
        cpn.capi.BamHandle(lib, os.path.join(GOLD, "ref.fa"))          # not BGZF


@pytest.mark.parametrize("name", ["meth_sample", "unmeth_sample"])
@pytest.mark.parametrize("trim", [(0, 0, 0, 0), (5, 3, 0, 0), (0, 149, 7, 7), (100, 60, 0, 0)])
def test_example_single_end_calls_equal_the_oracle(cpn, lib, name, trim):
    path = os.path.join(GOLD, name + ".bam")
    bam = orc.read_bam(path)
    regions = [(1, 849), (1, 1000), (1, 300), (200, 600), (457, 457), (700, 1000), (850, 1000), (900, 850)]
    cpgs = [ref_cpgs(st, max(st, en)) for st, en in regions]
    cpgs[4] = np.array([457], dtype=np.int64)
    got = native_calls(cpn, lib, path, "chrTest", regions, cpgs, False, trim)
    want = oracle_calls(bam, "chrTest", regions, cpgs, False, trim)
    n_rows = 0
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
        n_rows += g.shape[0]
    if trim == (0, 0, 0, 0):
        assert n_rows > 2000                                            # the comparison is not vacuous
        meth = got[0][got[0] != 0].mean()                               # the two samples are what their names say
        assert (meth > 0.5) == (name == "meth_sample")
    # a chromosome the BAM does not know: no reads (the reference fails its header lookup)
    none = native_calls(cpn, lib, path, "chrOther", [(1, 500)], [np.array([10, 20], dtype=np.int64)])
    assert none[0].shape == (0, 2)


def xm(length, calls):
    """XM string of `length` dots with calls {1-based index: char}."""
    s = ["."] * length
    for k, ch in calls.items():
        s[k - 1] = ch
    return "".join(s)


CPG = np.array([100, 110, 150, 200, 260, 300, 305, 400, 520, 640, 700], dtype=np.int64)


def paired_records():
    R = []

    def rec(pos, flag, name, xms, mapq=42, cigar=None, extra=()):
        n = len(xms)
        R.append((pos, bw.record(0, pos, flag, mapq, name, cigar or [(n, "M")], n, [("NM", "C", 0), ("XM", "Z", xms)] + list(extra))))

    # T1 OT, mates apart: R1 95..144 (Z at 100, z at 110), R2 180..229 (Z at 200, h elsewhere)
    rec(95, 99, "t1", xm(50, {6: "Z", 16: "z", 30: "h"})); rec(180, 147, "t1", xm(50, {21: "Z", 3: "x"}))
    # T2 OB, mates overlapping: R1 250..289 (z ↔ 260), R2 280..319 (Z ↔ 300, z ↔ 305; its first 10 bases repeat R1's end)
    rec(250, 83, "t2", xm(40, {12: "z"})); rec(280, 163, "t2", xm(40, {22: "Z", 27: "z", 5: "Z"}))
    # T3 CTOT (147 first, 99 second), adjacent mates
    rec(380, 147, "t3", xm(30, {21: "Z"})); rec(410, 99, "t3", xm(30, {2: "z"}))
    # T4 CTOB (163 first, 83 second)
    rec(500, 163, "t4", xm(40, {22: "z"})); rec(600, 83, "t4", xm(60, {42: "Z"}))
    # T5 mate with low MAPQ → singleton → dropped; T6 mate with XS → dropped
    rec(90, 99, "t5", xm(50, {11: "Z"})); rec(160, 147, "t5", xm(50, {41: "Z"}), mapq=39)
    rec(92, 99, "t6", xm(50, {9: "z"})); rec(170, 147, "t6", xm(50, {31: "z"}), extra=[("XS", "i", -3)])
    # T7 three kept records with one name → not a template
    rec(96, 99, "t7", xm(20, {5: "Z"})); rec(100, 147, "t7", xm(20, {11: "Z"})); rec(104, 147, "t7", xm(20, {7: "Z"}))
    # T8 both mates start at the same position: the first in the file is R1
    rec(690, 99, "t8", xm(20, {11: "Z"})); rec(690, 147, "t8", xm(25, {11: "z", 12: "H"}))
    # T9 unmapped mate (flag 4 + XM) and a flag outside FLAGS_ALLOWED (1123 = 99 + duplicate)
    rec(640, 99, "t9", xm(20, {1: "Z"})); rec(650, 4, "t9", xm(20, {2: "Z"}))
    rec(100, 1123, "t10", xm(20, {1: "Z"})); rec(150, 147, "t10", xm(20, {1: "Z"}))
    # T11 OT pair whose R1 has a deletion: 20M5D30M covers 900..954, so it overlaps a window that starts at 952
    rec(900, 99, "t11", xm(50, {3: "Z"}), cigar=[(20, "M"), (5, "D"), (30, "M")]); rec(960, 147, "t11", xm(30, {4: "z"}))
    # record without XM, record with a B-array tag before XM
    R.append((300, bw.record(0, 300, 99, 42, "t12", [(30, "M")], 30, [("NM", "C", 1)])))
    R.append((305, bw.record(0, 305, 147, 42, "t12", [(30, "M")], 30, [("ZB", "B", ("s", [1, -2, 3])), ("XM", "Z", xm(30, {1: "Z"}))])))
    R.sort(key=lambda t: t[0])
    return [r for _, r in R]


@pytest.fixture(scope="module")
def pe_bam(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("bam") / "pe.bam")
    bw.write_bam(path, [("chrA", 1200), ("chrB", 800)], paired_records(), block=700)
    return path


def test_paired_end_known_answers(cpn, lib, pe_bam):
    """Hand-computed tables for the templates of paired_records() over CPG = [100 110 150 200 260 300 305 400 520 640 700]."""
    got = native_calls(cpn, lib, pe_bam, "chrA", [(1, 1200)], [CPG], True)[0]
    rows = {
        "t1": [1, -1, 0, 1, 0, 0, 0, 0, 0, 0, 0],      # OT: index k ↔ 95 + k − 1; R2's part starts at index 86 ↔ 180
        "t2": [0, 0, 0, 0, -1, 1, -1, 0, 0, 0, 0],     # OB: index k ↔ 250 + k − 2; R2's first 10 characters are dropped
        "t3": [0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0],       # CTOT: 380 + 21 − 1 = 400; R2's z at 411 is no CpG of the region
        "t4": [0, 0, 0, 0, 0, 0, 0, 0, -1, 1, 0],      # CTOB: 500 + 22 − 2 = 520, 600 + 42 − 2 = 640
        "t8": [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1],       # equal starts: R1 = first in file (Z at 700); R2 adds only its tail beyond R1
    }
    # templates in order of first appearance among the position-sorted records: t1 (95), t2 (250), t3 (380), t4 (500), t8 (690), t11 (900: no CpG)
    assert got.shape == (5, len(CPG))
    assert np.array_equal(got, np.array([rows[k] for k in ("t1", "t2", "t3", "t4", "t8")], dtype=np.int8))


@pytest.mark.parametrize("trim", [(0, 0, 0, 0), (2, 3, 1, 4), (6, 0, 0, 0), (0, 0, 4, 0)])
def test_paired_end_calls_equal_the_oracle(cpn, lib, pe_bam, trim):
    bam = orc.read_bam(pe_bam)
    regions = [(1, 1200), (1, 170), (150, 400), (952, 1049), (955, 1049), (600, 700), (1, 1)]
    cpgs = [CPG[(CPG >= st) & (CPG <= en)] for st, en in regions]
    cpgs[3] = np.array([902, 963], dtype=np.int64); cpgs[4] = np.array([902, 963], dtype=np.int64)
    got = native_calls(cpn, lib, pe_bam, "chrA", regions, cpgs, True, trim)
    want = oracle_calls(bam, "chrA", regions, cpgs, True, trim)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    if trim == (0, 0, 0, 0):
        assert got[3].shape[0] == 1 and got[4].shape[0] == 0      # the deletion makes R1 of t11 reach 954: in at 952, out at 955
    # the other reference of the file has no records
    assert native_calls(cpn, lib, pe_bam, "chrB", [(1, 600)], [np.array([5, 50], dtype=np.int64)], True)[0].shape == (0, 2)


def test_unexpected_flags_are_errors(cpn, lib, pe_bam, tmp_path):
    bam = orc.read_bam(pe_bam)
    # single-end mode on paired flags: the reference prints and exits (src/wgbs.jl:31-33)
    with pytest.raises(orc.ReferenceExit):
        orc.get_calls_bam(bam, "chrA", 1, 1200, [int(c) for c in CPG], False, (0, 0, 0, 0))
    with pytest.raises(cpn.CpnError, match="FLAG 0 or 16"):
        native_calls(cpn, lib, pe_bam, "chrA", [(1, 1200)], [CPG], False)
    # paired-end mode with mates 99 / 163 (:45-47)
    path = str(tmp_path / "bad.bam")
    bw.write_bam(path, [("chrA", 1200)], [bw.record(0, 95, 99, 42, "b", [(50, "M")], 50, [("XM", "Z", xm(50, {6: "Z"}))]),
                                          bw.record(0, 180, 163, 42, "b", [(50, "M")], 50, [("XM", "Z", xm(50, {21: "Z"}))])])
    with pytest.raises(orc.ReferenceExit):
        orc.get_calls_bam(orc.read_bam(path), "chrA", 1, 1200, [int(c) for c in CPG], True, (0, 0, 0, 0))
    with pytest.raises(cpn.CpnError, match="unexpected FLAG combination 99/163"):
        native_calls(cpn, lib, path, "chrA", [(1, 1200)], [CPG], True)


def test_mate_inside_the_other(cpn, lib, tmp_path):
    """R2 ends before R1's string does: the reference indexes R2's string past its end (BoundsError); here R2 adds nothing."""
    path = str(tmp_path / "in.bam")
    bw.write_bam(path, [("chrA", 1200)], [bw.record(0, 95, 99, 42, "c", [(100, "M")], 100, [("XM", "Z", xm(100, {6: "Z", 56: "z"}))]),
                                          bw.record(0, 105, 147, 42, "c", [(50, "M")], 50, [("XM", "Z", xm(50, {6: "Z"}))])])
    with pytest.raises(IndexError):
        orc.get_calls_bam(orc.read_bam(path), "chrA", 1, 1200, [int(c) for c in CPG], True, (0, 0, 0, 0))
    got = native_calls(cpn, lib, path, "chrA", [(1, 1200)], [CPG], True)[0]
    assert got.tolist() == [[1, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0]]


def synthetic_wgbs(tmp_path, seed=5, G=9000, read_len=100, depth=40, p_high=0.7):
    """A random chromosome with ≈ 2 % CpG dinucleotides and a coordinate-sorted single-end BAM of OT / OB reads whose XM strings carry
    calls drawn per read: a read is mostly methylated (probability p_high) or mostly unmethylated, so neighbouring calls are
    correlated; p_high = 1: independent calls, methylated with probability 0.9."""
    rng = np.random.default_rng(seed)
    seq = rng.choice(list("ATG"), G)                      # no C: CpGs only where planted
    cpos = np.sort(rng.choice(np.arange(50, G - 400, 9), 150, replace=False))      # C of each CpG, 0-based
    seq[cpos] = "C"; seq[cpos + 1] = "G"
    fasta = str(tmp_path / "syn.fa")
    with open(fasta, "w") as f:
        f.write(">chrS\n")
        s = "".join(seq)
        for o in range(0, G, 60):
            f.write(s[o:o + 60] + "\n")
    cpg1 = cpos + 1                                      # 1-based C positions
    n_reads = depth * G // read_len
    starts = np.sort(rng.integers(1, G - read_len, n_reads))
    recs = []
    for k, p in enumerate(starts):
        ob = bool(rng.integers(0, 2))
        high = rng.random() < p_high
        s = ["."] * read_len
        lo = p - (1 if ob else 0)                         # OB: the call sits on the G, one base after the C
        for c in cpg1[(cpg1 >= lo) & (cpg1 <= lo + read_len - 1)]:
            idx = c - p + (2 if ob else 1)
            if 1 <= idx <= read_len:
                s[idx - 1] = "Z" if rng.random() < (0.9 if high else 0.15) else "z"
        recs.append(bw.record(0, int(p), 16 if ob else 0, 42, f"r{k}", [(read_len, "M")], read_len, [("XM", "Z", "".join(s))]))
    bam = str(tmp_path / "syn.bam")
    bw.write_bam(bam, [("chrS", G)], recs, block=60000)
    return fasta, bam


def test_synthetic_single_end_calls_equal_the_oracle(cpn, lib, tmp_path):
    fasta, bam = synthetic_wgbs(tmp_path, G=4000, depth=12)
    chrom = cpn.genome.load_genome(fasta, lib)["chrS"]
    part = chrom.chr_part_native(lib, 1500, 10, None)
    regs = chrom.region_structs(lib, part, 1)
    assert len(regs) >= 2
    regions = [(rs.chrst, rs.chrend) for rs in regs]
    cpgs = [np.asarray(rs.cpg_pos, dtype=np.int64) for rs in regs]
    got = native_calls(cpn, lib, bam, "chrS", regions, cpgs)
    want = oracle_calls(orc.read_bam(bam), "chrS", regions, cpgs)
    for g, w in zip(got, want):
        assert g.shape == w.shape and g.shape[0] > 50 and np.array_equal(g, w)


def test_many_regions_thread_invariance(cpn, lib):
    """3 000 random windows of the example, 1 / 8 / all threads: identical tables (a row is written only once it is known to be
    one; regions are independent)."""
    h = cpn.capi.BamHandle(lib, os.path.join(GOLD, "meth_sample.bam"))
    rng = np.random.default_rng(0)
    st = rng.integers(1, 900, 3000); en = st + rng.integers(0, 300, 3000)
    cp = ref_cpgs(1, 1000)
    cpgs = [cp[(cp >= a) & (cp <= b)] for a, b in zip(st, en)]
    cpg_off = np.concatenate([[0], np.cumsum([len(c) for c in cpgs])]).astype(np.int64)
    outs = [h.calls("chrTest", st, en, cpg_off, np.concatenate(cpgs), False, (3, 2, 0, 0), n_threads=t) for t in (1, 8, 0)]
    assert outs[0][1].size > 100000
    for ro, calls in outs[1:]:
        assert np.array_equal(ro, outs[0][0]) and np.array_equal(calls, outs[0][1])


def test_analyze_bam_reference_example_is_empty(cpn, tmp_path):
    """examples/wgbs: the analysed window of the 1 kb chromosome holds 7 CpG sites; pmap_analyze_reg_bam skips regions with ≤ 9
    (src/wgbs.jl:353), so the reference's expected theta file is empty — and no GPU is needed to reproduce it."""
    out = str(tmp_path / "out")
    done = cpn.analyze_bam(os.path.join(GOLD, "meth_sample.bam"), os.path.join(GOLD, "ref.fa"), cpn.CpelNanoConfig(), out, "test")
    assert done == []
    assert open(os.path.join(out, "test_theta.txt")).read() == open(os.path.join(GOLD, "test_theta.cpelont")).read() == ""
    assert len(ref_cpgs(1, 849)) == 7


@pytest.mark.gpu
def test_analyze_bam_equals_the_batch_call_on_its_calls(cpn, lib, tmp_path):
    """File level: analyze_bam on a synthetic chromosome = analyze_regions_wgbs on the calls the reader extracts, with the same
    starts; the fitted regions are methylated where the reads are."""
    fasta, bam = synthetic_wgbs(tmp_path, depth=20, p_high=1.0)
    # hard calls make the EM slow: with the default cap of 20 iterations most starts end unconverged (Q = −Inf, em_algorithm.jl:344);
    # the reference's own WGBS script sets 100 (calls/GeneralTesting/wgbs_tests.jl:58)
    cfg = cpn.CpelNanoConfig(max_em_iters=100)
    n_init = cfg.max_em_init
    starts = {}

    def phi0_fn(n):
        starts[n] = cpn.simulations.gen_phi0s(np.random.default_rng(11), n, n_init)
        return starts[n]

    chrom = cpn.genome.load_genome(fasta, lib)["chrS"]
    part = chrom.chr_part_native(lib, cfg.size_est_reg, cfg.min_grp_dist, None)
    regs = [rs for rs in chrom.region_structs(lib, part, 1) if len(rs.cpg_pos) > 9]
    calls = native_calls(cpn, lib, bam, "chrS", [(rs.chrst, rs.chrend) for rs in regs], [np.asarray(rs.cpg_pos, dtype=np.int64) for rs in regs])
    cpn.analyze_regions_wgbs(regs, calls, cfg, phi0s=phi0_fn(len(regs)))
    ref = [rs for rs in regs if rs.proc]
    assert len(ref) >= 2, [(rs.chrst, rs.m, len(rs.cpg_pos), rs.proc) for rs in regs]
    out = str(tmp_path / "out")
    done = cpn.analyze_bam(bam, fasta, cfg, out, "syn", phi0_fn=phi0_fn)
    assert [(a.chrst, a.chrend) for a in done] == [(b.chrst, b.chrend) for b in ref]
    frac = np.mean(np.concatenate([c[c != 0] for c in calls if c.size]) > 0)
    assert abs(frac - 0.9) < 0.02
    for a, b in zip(done, ref):
        assert np.array_equal(a.phihat, b.phihat) and np.array_equal(a.mml, b.mml, equal_nan=True) and np.array_equal(a.nme, b.nme, equal_nan=True)
        assert abs(np.nanmean(a.mml) - frac) < 0.05                   # MML ≈ fraction of methylated calls
    lines = open(os.path.join(out, "syn_theta.txt")).read().splitlines()
    assert len(lines) == len(done)
