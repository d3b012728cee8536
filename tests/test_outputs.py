"""Output text and BH correction against the reference's own golden files.

The golden files were written by Julia: every floating-point token in them is Julia's `string(::Float64)`.  Parsing a token
and printing it with outputs.julia_str must give the token back; BH q-values recomputed from the golden p columns must
print to the golden q columns.  GPU-side end-to-end file comparisons are in test_pipeline_gpu.py.
"""
import io
import os
import tarfile

import numpy as np
import pytest

import cpelnano_b200 as cpn
from cpelnano_b200 import outputs

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")


@pytest.fixture(scope="module")
def golden():
    out = {}
    with tarfile.open(os.path.join(INP, "golden_outputs.tar.gz")) as tar:
        for m in tar.getmembers():
            out[m.name] = tar.extractfile(m).read().decode()
    return out


def _float_tokens(text, first_col):
    for line in text.splitlines():
        for field in line.split("\t")[first_col:]:
            for tok in field.split(","):
                yield tok


def test_julia_float_text_round_trips_every_golden_number(golden):
    n = 0
    for name, text in golden.items():
        for tok in _float_tokens(text, 3):
            assert outputs.julia_str(float(tok)) == tok, (name, tok)
            n += 1
    assert n > 20_000


def test_julia_float_text_special_cases():
    cases = {1.0: "1.0", 100000.0: "100000.0", 1e6: "1.0e6", 1234567.0: "1.234567e6", 1e-5: "1.0e-5", 0.0001: "0.0001",
             -3.25e-7: "-3.25e-7", 1e22: "1.0e22", 5e-324: "5.0e-324", 0.1 + 0.2: "0.30000000000000004", -0.0: "-0.0",
             float("inf"): "Inf", float("-inf"): "-Inf", 2.0 ** 70: "1.1805916207174113e21", 999999.9: "999999.9"}
    for x, s in cases.items():
        assert outputs.julia_str(x) == s
    assert outputs.julia_str(float("nan")) == "NaN"


def test_julia_round_ties_to_even_on_the_scaled_value():
    assert outputs.julia_round(0.00005) == np.rint(0.00005 * 1e4) / 1e4
    assert outputs.julia_round(2.5, 0) == 2.0 and outputs.julia_round(3.5, 0) == 4.0
    assert outputs.julia_round(-0.93145) == -0.9314 or outputs.julia_round(-0.93145) == -0.9315   # whichever x·1e4 rounds to
    assert outputs.julia_round(0.12344999) == 0.1234


@pytest.mark.parametrize("name", ["unmat_tmml", "unmat_tnme", "unmat_tcmd", "mat_tmml", "mat_tnme"])
def test_bh_qvalues_reprint_the_golden_q_column(golden, name):
    rows = [l.split("\t") for l in golden[f"grp_comp_example/tests/cpelnano_grp_comp_{name}.txt"].splitlines()]
    p = np.array([float(r[4]) for r in rows])
    q = outputs.bh_adjust(p)
    assert [outputs.julia_str(v) for v in q] == [r[5] for r in rows]


def test_bh_properties():
    rng = np.random.default_rng(1)
    p = rng.random(1000)
    q = outputs.bh_adjust(p)
    o = np.argsort(p)
    assert np.all(np.diff(q[o]) >= 0) and np.all(q >= p) and q.max() <= 1.0
    assert np.array_equal(outputs.bh_adjust([0.2]), [0.2]) and len(outputs.bh_adjust([])) == 0


def test_add_qvalues_rewrites_file_like_the_reference(golden, tmp_path):
    """Strip the q column of a golden file, run add_qvalues_to_path: the file must come back byte-identical (NaN p-values of
    the matched Tcmd file are never corrected: mult_hyp_corr skips it, src/multiple_hypothesis.jl:60)."""
    text = golden["grp_comp_example/tests/cpelnano_grp_comp_unmat_tnme.txt"]
    path = tmp_path / "t_tnme.txt"
    path.write_text("".join("\t".join(l.split("\t")[:5]) + "\n" for l in text.splitlines()))
    outputs.add_qvalues_to_path(str(path))
    assert path.read_text() == text
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    outputs.add_qvalues_to_path(str(empty))
    assert empty.read_text() == ""


def test_theta_reader_and_alpha_beta_print_identically(golden, tmp_path):
    """read_theta_lines + genome features + get_αβ_from_ϕ + julia_str reproduce the golden theta file byte for byte
    (the α, β columns are per-CpG: write_output_ϕ runs after rscle_grp_mod!)."""
    from cpelnano_b200 import genome
    text = golden["full_example/cpelnano/full_example_mod_nse_true_theta.txt"]
    src = tmp_path / "theta.txt"
    src.write_text(text)
    chrom = genome.load_genome(os.path.join(INP, "full_example.fa.gz"))["hg38_dna"]
    regs = []
    for (c, st, end, phi) in outputs.read_theta_lines(str(src), "hg38_dna"):
        rs = chrom.region_struct(st, end)
        rs.phihat = phi
        cpn.rscle_grp_mod(rs)
        rs.proc = True
        regs.append(rs)
    assert len(regs) == 27
    dst = tmp_path / "out_theta.txt"
    outputs.write_output_phi(str(dst), regs)
    assert dst.read_text() == text


# ---- native writers (csrc/cpn_text.cpp) ------------------------------------------------------------------------------------------
def test_native_float_text_equals_julia_text_on_every_golden_number(golden):
    from cpelnano_b200 import capi
    lib = cpn.Library()
    for name, text in golden.items():
        toks = list(_float_tokens(text, 3))
        vals = np.array([float(t) for t in toks])
        assert capi.format_f64(lib, vals, sep="\n").split("\n") == toks, name
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000) * 10.0 ** rng.integers(-12, 12, 20000), [0.0, -0.0, np.nan, np.inf, -np.inf, 1e-5, 1e6, 999999.9,
                        5e-324, 1.7976931348623157e308, 0.1 + 0.2, 100000.0, 123456.7, 1e22, 2.0 ** 70]])
    assert capi.format_f64(lib, x, sep=";").split(";") == [outputs.julia_str(v) for v in x]
    r4 = capi.format_f64(lib, x[:20000], sep=";", digits=4).split(";")
    assert r4 == [outputs.julia_str(outputs.julia_round(v)) for v in x[:20000]]


def test_native_writers_reproduce_golden_theta_and_python_writers(golden, tmp_path):
    from cpelnano_b200 import genome
    lib = cpn.Library()
    text = golden["full_example/cpelnano/full_example_mod_nse_true_theta.txt"]
    src = tmp_path / "theta.txt"
    src.write_text(text)
    chrom = genome.load_genome(os.path.join(INP, "full_example.fa.gz"), lib)["hg38_dna"]
    lines = outputs.read_theta_lines(str(src), "hg38_dna")
    regs = chrom.region_structs(lib, [(l[1], l[2]) for l in lines])
    rng = np.random.default_rng(3)
    for rs, l in zip(regs, lines):
        rs.phihat = l[3]
        cpn.rscle_grp_mod(rs)
        rs.proc = True
        # summaries are the GPU's job; here any doubles will do to compare the two writers (NaN rows must be skipped by both)
        rs.ex, rs.exx = rng.uniform(-1, 1, rs.N), rng.uniform(-1, 1, rs.N - 1)
        K = 9
        rs.nls_rgs = cpn.host.AnalysisRegions(K, np.arange(K) * 300 + rs.chrst, np.arange(K) * 300 + rs.chrst + 299, np.zeros(K, dtype=np.int64),
                                              np.zeros(K, dtype=np.int64))
        rs.mml, rs.nme = rng.uniform(0, 1, K), rng.uniform(0, 1, K)
        rs.mml[rng.integers(0, K)] = np.nan; rs.nme[np.isnan(rs.mml)] = np.nan
        rs.ex[0] = np.nan
    regs[3].proc = False
    fa, fb = outputs.OutputFiles(str(tmp_path), "py"), outputs.OutputFiles(str(tmp_path), "native")
    outputs.write_output(regs, fa)
    outputs.write_output(regs, fb, lib)
    for a, b in zip(fa.all(), fb.all()):
        assert open(a).read() == open(b).read() and os.path.getsize(a) > 0, a
    regs[3].proc = True
    fc = outputs.OutputFiles(str(tmp_path), "native_all")
    outputs.write_output(regs, fc, lib)
    assert open(fc.theta_file).read() == text                                  # the reference's own theta file, byte for byte


def test_native_diff_writer_equals_python_writer(tmp_path):
    lib = cpn.Library()
    rng = np.random.default_rng(2)
    tests, chrom = [], []
    for r in range(40):
        K = int(rng.integers(1, 10))
        co = cpn.host.AnalysisRegions(K, np.arange(K) * 350 + 1000 * r, np.arange(K) * 350 + 1000 * r + 349, np.zeros(K, dtype=np.int64), np.zeros(K, dtype=np.int64))
        T = rng.standard_normal((3, K)); T[2] = np.abs(T[2]) - 0.05                  # some Tcmd slightly negative: clamped to 0.0
        T[:, rng.integers(0, K)] = np.nan
        p = rng.integers(0, 925, (3, K)) / 924.0
        p[rng.random((3, K)) < 0.1] = np.nan
        tests.append(dict(T=T, pval=p, tcmd=np.abs(T[2]), coords=co)); chrom.append("chrA" if r < 25 else "chrB")
    for matched in (False, True):
        fa, fb = outputs.OutputDiffFiles(str(tmp_path), f"py{matched}"), outputs.OutputDiffFiles(str(tmp_path), f"nat{matched}")
        outputs.write_diff_output(chrom, tests, fa, matched)
        outputs.write_diff_output(chrom, tests, fb, matched, lib)
        for a, b in zip(fa.all(), fb.all()):
            assert open(a).read() == open(b).read() and os.path.getsize(a) > 0


def test_writers_create_empty_files_when_nothing_was_processed(tmp_path):
    """open(path, "a") in the reference creates all five files even if every region was skipped; both writers do the same."""
    lib = cpn.Library()
    for tag, l in (("py", None), ("nat", lib)):
        f = outputs.OutputFiles(str(tmp_path), tag)
        outputs.write_output([cpn.RegStruct(chr="c", chrst=1, chrend=10)], f, l)
        assert all(os.path.exists(p) and os.path.getsize(p) == 0 for p in f.all())
        d = outputs.OutputDiffFiles(str(tmp_path), tag)
        outputs.write_diff_output([], [], d, False, l)
        assert all(os.path.exists(p) and os.path.getsize(p) == 0 for p in d.all())
        outputs.mult_hyp_corr(d)


def test_native_qvalues_equal_python_and_golden(golden, tmp_path):
    from cpelnano_b200 import capi
    lib = cpn.Library()
    for name in ("unmat_tmml", "unmat_tnme", "unmat_tcmd", "mat_tmml", "mat_tnme"):
        text = golden[f"grp_comp_example/tests/cpelnano_grp_comp_{name}.txt"]
        stripped = "".join("\t".join(l.split("\t")[:5]) + "\n" for l in text.splitlines())
        a, b = tmp_path / f"{name}_py.txt", tmp_path / f"{name}_nat.txt"
        a.write_text(stripped); b.write_text(stripped)
        outputs.add_qvalues_to_path(str(a))
        capi.add_qvalues(lib, str(b))
        assert a.read_text() == b.read_text() == text, name           # the reference's file, byte for byte
    # NaN p-values, ties, a single row, empty and missing files
    rng = np.random.default_rng(4)
    rows = []
    for i in range(500):
        p = rng.integers(0, 30) / 29.0 if rng.random() > 0.1 else float("nan")
        rows.append(f"chr1\t{100 * i}\t{100 * i + 99}\t{outputs.julia_str(outputs.julia_round(rng.standard_normal()))}\t{outputs.julia_str(p)}\n")
    a, b = tmp_path / "r_py.txt", tmp_path / "r_nat.txt"
    a.write_text("".join(rows)); b.write_text("".join(rows))
    outputs.add_qvalues_to_path(str(a)); capi.add_qvalues(lib, str(b))
    assert a.read_text() == b.read_text() and "NaN\tNaN" in a.read_text()
    one = tmp_path / "one.txt"
    one.write_text("chr1\t1\t2\t0.5\t0.25\n")
    capi.add_qvalues(lib, str(one))
    assert one.read_text() == "chr1\t1\t2\t0.5\t0.25\t0.25\n"
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    capi.add_qvalues(lib, str(empty)); capi.add_qvalues(lib, str(tmp_path / "missing.txt"))
    assert empty.read_text() == "" and not (tmp_path / "missing.txt").exists()


def test_native_theta_reader_equals_python_reader(golden, tmp_path):
    from cpelnano_b200 import capi, pipeline
    lib = cpn.Library()
    for name in ("full_example/cpelnano/full_example_mod_nse_true_theta.txt", "grp_comp_example/mod_files/g1_rep3.txt",
                 "targeted_example/cpelnano/targeted_example_theta.txt"):
        p = tmp_path / "t.txt"
        p.write_text(golden[name])
        a, b = pipeline._read_theta_columns(str(p)), capi.read_theta(lib, str(p))
        assert list(a) == list(b)
        for k in a:
            assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k]))
    # two chromosomes interleaved, 4-column lines, NaN, CRLF, blank line
    p = tmp_path / "m.txt"
    p.write_bytes(b"chr2\t10\t20\t0.5,-1.25,3.0\r\nchr1\t1\t5\tNaN,NaN,NaN\talpha\tbeta\n\nchr2\t21\t30\t1e-3,2,3\n")
    b = capi.read_theta(lib, str(p))
    assert list(b) == ["chr2", "chr1"] and b["chr2"][0].tolist() == [10, 21] and b["chr2"][2].tolist() == [[0.5, -1.25, 3.0], [1e-3, 2.0, 3.0]]
    assert np.isnan(b["chr1"][2]).all()
    bad = tmp_path / "bad.txt"
    bad.write_text("chr1\t1\t5\t0.5,0.1\n")
    with pytest.raises(cpn.CpnError, match="malformed line 1"):
        capi.read_theta(lib, str(bad))
    with pytest.raises(cpn.CpnError, match="cannot open"):
        capi.read_theta(lib, str(tmp_path / "missing.txt"))
