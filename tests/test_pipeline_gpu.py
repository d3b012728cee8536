"""File-in → file-out entry points on the GPU against the reference's golden output FILES (run with -m gpu).

  * diff_grp_comp on the bundled 6 vs 6 model files (recipe: calls/GeneralTesting/permutation_tests_tests.jl:118-156) must
    write the golden tmml / tnme rows character for character (T at 4 dp, exact p from the full 924-labeling / 64-sign sweeps, BH q)
    and the golden T column of tcmd; unmatched Tcmd p-values may differ by one count of 924 where complementary labelings tie
    at the last ulp (the reference's own result there depends on its summation order).
  * summaries recomputed from the golden theta file must re-write the golden ex / exx / mml / nme files: mml / nme (rounded
    to 4 dp) byte for byte, E[X] / E[XX] (full precision) to 1e-12.
  * analyze_nano on the bundled full_example calls selects exactly the golden file's 27 regions (integer filters, bit-exact),
    converges on all of them, and lands on the golden ϕ̂ as closely as the reference's stopping rule (‖Δϕ‖ ≤ 3e-2, random
    unseeded starts) allows.
"""
import os
import tarfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")


@pytest.fixture(scope="module")
def golden_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_outputs")
    with tarfile.open(os.path.join(INP, "golden_outputs.tar.gz")) as tar:
        tar.extractall(d, filter="data")
    return str(d)


@pytest.fixture(scope="module")
def fasta_full(tmp_path_factory):
    return os.path.join(INP, "full_example.fa.gz")


def _rows(path, sort=False):
    rows = [l.rstrip("\n").split("\t") for l in open(path)]
    return sorted(rows, key=lambda r: (int(r[1]), int(r[2]))) if sort else rows


@pytest.mark.parametrize("matched", [False, True])
def test_diff_grp_comp_writes_the_golden_files(engine, cpn, golden_dir, fasta_full, tmp_path, matched):
    g1 = [f"{golden_dir}/grp_comp_example/mod_files/g1_rep{i}.txt" for i in range(1, 7)]
    g2 = [f"{golden_dir}/grp_comp_example/mod_files/g2_rep{i}.txt" for i in range(1, 7)]
    cfg = cpn.CpelNanoConfig(matched=matched)
    kind = "mat" if matched else "unmat"
    files = cpn.diff_grp_comp(g1, g2, fasta_full, cfg, str(tmp_path), f"cpelnano_grp_comp_{kind}", engine=engine)
    gold = f"{golden_dir}/grp_comp_example/tests/cpelnano_grp_comp_{kind}"
    for stat in ("tmml", "tnme"):
        # the reference writes regions in Julia Dict order (find_comm_regs); rows are compared in coordinate order, as text
        assert _rows(getattr(files, stat + "_file")) == _rows(f"{gold}_{stat}.txt", sort=True), stat
    got, ref = _rows(files.tcmd_file), _rows(f"{gold}_tcmd.txt", sort=True)
    assert len(got) == len(ref) == 225
    if matched:
        assert got == ref                                     # T only, p = NaN, no q column
    else:
        assert [r[:4] for r in got] == [r[:4] for r in ref]   # coordinates and T
        dp = np.array([float(a[4]) - float(b[4]) for a, b in zip(got, ref)]) * 924
        assert np.all(np.abs(dp) <= 1.0 + 1e-9), np.abs(dp).max()
    # an existing output stops a second run (check_diff_output_exists)
    assert cpn.diff_grp_comp(g1, g2, fasta_full, cfg, str(tmp_path), f"cpelnano_grp_comp_{kind}", engine=engine) is None
    # the array path writes the same bytes
    files2 = cpn.diff_grp_comp_csr(g1, g2, fasta_full, cfg, str(tmp_path / "csr"), f"cpelnano_grp_comp_{kind}", engine=engine)
    for a, b in zip(files.all(), files2.all()):
        assert open(a).read() == open(b).read()


@pytest.mark.parametrize("example,fa,prefix,mss", [
    ("targeted_example", "targeted_example.fa.gz", "targeted_example/cpelnano/targeted_example", 351),
])
def test_summaries_from_golden_theta_rewrite_the_golden_files(engine, cpn, golden_dir, tmp_path, example, fa, prefix, mss):
    from cpelnano_b200 import genome, outputs
    chrom = genome.load_genome(os.path.join(INP, fa))["hg38_dna"]
    cfg = cpn.CpelNanoConfig(max_size_subreg=mss)             # 5-argument ctor of the example stores 350 + 1
    regs = outputs.read_mod_file_chr(f"{golden_dir}/{prefix}_theta.txt", chrom, cfg, engine)
    regs = sorted(regs.values(), key=lambda r: r.chrst)
    files = outputs.OutputFiles(str(tmp_path), example)
    outputs.write_output(regs, files)
    assert open(files.theta_file).read() == open(f"{golden_dir}/{prefix}_theta.txt").read()
    assert open(files.mml_file).read() == open(f"{golden_dir}/{prefix}_mml.txt").read()
    assert open(files.nme_file).read() == open(f"{golden_dir}/{prefix}_nme.txt").read()
    # the golden ex / exx files hold four appended runs of the example (open(path, "a")), only the last of which belongs to the
    # golden theta file: a row written here must equal one of the golden rows with the same coordinates
    for stat in ("ex", "exx"):
        ref = {}
        for r in _rows(f"{golden_dir}/{prefix}_{stat}.txt"):
            ref.setdefault((r[0], r[1], r[2]), []).append(float(r[3]))
        got = _rows(getattr(files, stat + "_file"))
        assert len(got) == sum(r.N - (stat == "exx") for r in regs) and {tuple(r[:3]) for r in got} <= set(ref)
        d = np.array([min(abs(float(r[3]) - v) for v in ref[tuple(r[:3])]) for r in got])
        assert d.max() <= 1e-12, d.max()


def test_analyze_nano_full_example(engine, cpn, golden_dir, fasta_full, tmp_path):
    from cpelnano_b200 import outputs
    cfg = cpn.CpelNanoConfig()                                # defaults: min_cov 10, 10 starts, ≤ 20 EM iterations
    rng = np.random.default_rng(20211111)
    done = cpn.analyze_nano(os.path.join(INP, "full_example_calls.tsv.gz"), fasta_full, cfg, str(tmp_path), "full_example", rng=rng,
                            engine=engine)
    gold = outputs.read_theta_lines(f"{golden_dir}/full_example/cpelnano/full_example_mod_nse_true_theta.txt")
    mine = outputs.read_theta_lines(str(tmp_path / "full_example_theta.txt"))
    assert [(g[1], g[2]) for g in mine] == [(g[1], g[2]) for g in gold]          # the same 27 regions, in order
    assert len(done) == 27
    dphi = np.array([m[3] - g[3] for m, g in zip(mine, gold)])
    # the golden run used different (unseeded) starts: same optimum up to the EM stopping rule on most regions
    med = np.median(np.abs(dphi), axis=0)
    assert med[0] <= 0.02 and med[1] <= 1.0 and med[2] <= 0.5, med
    # outputs are internally consistent: mml file has one row per non-empty sub-region, all within [0, 1]
    mml = np.array([float(r[3]) for r in _rows(tmp_path / "full_example_mml.txt")])
    nme = np.array([float(r[3]) for r in _rows(tmp_path / "full_example_nme.txt")])
    assert len(mml) == len(nme) > 200 and mml.min() >= 0.0 and mml.max() <= 1.0 and nme.min() >= 0.0 and nme.max() <= 1.0
    # the MML of the golden model and of this fit agree closely (MML is a smooth function of the optimum)
    with open(tmp_path / "gold_theta.txt", "w") as f:
        f.write(open(f"{golden_dir}/full_example/cpelnano/full_example_mod_nse_true_theta.txt").read())
    from cpelnano_b200 import genome
    chrom = genome.load_genome(fasta_full)["hg38_dna"]
    gregs = sorted(outputs.read_mod_file_chr(str(tmp_path / "gold_theta.txt"), chrom, cfg, engine).values(), key=lambda r: r.chrst)
    gm = np.concatenate([r.mml[~np.isnan(r.mml)] for r in gregs])
    assert len(gm) == len(mml) and np.median(np.abs(gm - mml)) <= 0.01, np.median(np.abs(gm - mml))
    # mod_nse = false: hard emissions (0, −50) go through the same kernels (the WGBS-style likelihood, src/CpelNano.jl:32-33)
    done2 = cpn.analyze_nano(os.path.join(INP, "full_example_calls.tsv.gz"), fasta_full, cfg, str(tmp_path), "full_example_hard",
                             mod_nse=False, rng=np.random.default_rng(1), engine=engine)
    assert len(done2) >= 20


def test_diff_two_samp_comp_same_sample_reproduces_the_golden_files(engine, cpn, golden_dir, fasta_full, tmp_path):
    """Recipe of calls/GeneralTesting/permutation_tests_tests.jl:160-190: a sample against itself (4 regions of the full
    example, 4 starts, ≤ 25 EM iterations).  T = 0 for every statistic, every valid permutation statistic is ≥ 0 in absolute
    value, so p = (1+n)/(1+n) = 1 and q = 1: the golden files are fully determined and must be reproduced as text."""
    theta = tmp_path / "two_samp_theta.txt"
    lines = open(f"{golden_dir}/full_example/cpelnano/full_example_mod_nse_true_theta.txt").read().splitlines()[:4]
    theta.write_text("".join("\t".join(l.split("\t")[:4]) + "\n" for l in lines))      # model files may omit α, β (4 columns)
    nano = os.path.join(INP, "full_example_calls.tsv.gz")
    cfg = cpn.CpelNanoConfig(max_em_iters=25, max_em_init=4, pval_comp=True)
    files = cpn.diff_two_samp_comp(str(theta), str(theta), nano, nano, fasta_full, cfg, str(tmp_path), "cpelnano_two_samp",
                                   rng=np.random.default_rng(5), engine=engine)
    for stat in ("tmml", "tnme", "tcmd"):
        got = _rows(getattr(files, stat + "_file"))
        ref = _rows(f"{golden_dir}/two_samp_example/tests/cpelnano_two_samp_{stat}.txt", sort=True)
        assert got == ref, stat
    # without p-values: T columns only, p = NaN, no q column (two_samp_perm_tests.jl:452)
    cfg2 = cpn.CpelNanoConfig(pval_comp=False)
    files2 = cpn.diff_two_samp_comp(str(theta), str(theta), nano, nano, fasta_full, cfg2, str(tmp_path), "no_pvals", engine=engine)
    rows = _rows(files2.tmml_file)
    assert len(rows) == 34 and all(r[3] == "0.0" and r[4] == "NaN" and len(r) == 5 for r in rows)


def test_file_pipeline_equals_the_csr_api_bit_for_bit(engine, cpn, tmp_path):
    """FASTA + calls file → analyze_nano (native CpG index, partition, features, calls loader, batched GPU fit, writers) against
    cpn_analyze_batch on the arrays the file was written from.  Every group is observed (pobs = 1) so that the file's read
    order (first appearance) is the batch's read order; doubles are written with 17 significant digits.  Equality is exact."""
    from cpelnano_b200 import genome, simulations, outputs
    fa = os.path.join(INP, "targeted_example.fa.gz")
    chrom = genome.load_genome(fa, engine.lib)["hg38_dna"]
    lay, grp_st, grp_end = simulations.layouts_from_fasta(engine.lib, chrom)
    ref = simulations.load_layouts()
    assert all(np.array_equal(lay[k], ref[k]) for k in ref)                      # native features = the committed, golden-pinned layouts
    R, n_init = len(lay["chrst"]), 3
    nb = simulations.make_nano_batch(R, M=20, seed=5, pobs=1.0, layouts=lay)
    tsv = tmp_path / "calls.tsv"
    simulations.write_nanopolish_tsv(str(tsv), nb, "hg38_dna", grp_st, grp_end)
    phi0 = simulations.gen_phi0s(np.random.default_rng(9), R, n_init)
    cfg = cpn.CpelNanoConfig(max_em_init=n_init)
    subs = [engine.lib.nls_reg_info(int(nb.chrst[r]), int(nb.chrend[r]), cfg.max_size_subreg, nb.cpg_pos[nb.cpg_off[r]:nb.cpg_off[r + 1]]) for r in range(R)]
    sub_off = np.concatenate([[0], np.cumsum([len(s[0]) for s in subs])]).astype(np.int64)
    direct = engine.lib.analyze_batch(nb.grp_off, nb.Nl, nb.rho_l, nb.d_l, nb.read_off, nb.log_pyx_u, nb.log_pyx_m, phi0, n_init, nb.cpg_off,
                                      nb.rho_n, nb.d_n, sub_off, np.concatenate([s[2] for s in subs]), np.concatenate([s[3] for s in subs]))

    def starts(n):
        assert n == R
        return phi0

    done = cpn.analyze_nano(str(tsv), fa, cfg, str(tmp_path), "sim", phi0_fn=starts, engine=engine)
    ok = np.nonzero(direct["pick"] >= 0)[0]
    assert len(done) == len(ok) > 100
    for rs, r in zip(done, ok):
        assert (rs.chrst, rs.chrend) == (int(nb.chrst[r]), int(nb.chrend[r]))
        assert np.array_equal(rs.phihat, direct["phi_hat"][r])
        assert np.array_equal(rs.mml, direct["mml"][sub_off[r]:sub_off[r + 1]], equal_nan=True)
        assert np.array_equal(rs.nme, direct["nme"][sub_off[r]:sub_off[r + 1]], equal_nan=True)
        assert np.array_equal(rs.ex, direct["ex"][nb.cpg_off[r]:nb.cpg_off[r + 1]])
    # and the theta file holds exactly those fits, as text
    lines = outputs.read_theta_lines(str(tmp_path / "sim_theta.txt"))
    assert len(lines) == len(ok) and all(np.array_equal(l[3], direct["phi_hat"][r]) for l, r in zip(lines, ok))


def test_csr_fast_path_writes_the_same_files(engine, cpn, tmp_path):
    """analyze_nano_csr (block-wise CSR arrays end to end, vectorised filters, one cpn_analyze_batch per block) against analyze_nano
    (per-region structs): the five output files must be byte-identical — on the bundled calls (14 of 41 regions rejected by
    the coverage filters) and on a simulated file with unobserved groups, with small blocks to cross block boundaries."""
    from cpelnano_b200 import genome, simulations
    cases = []
    cases.append((os.path.join(INP, "full_example_calls.tsv.gz"), os.path.join(INP, "full_example.fa.gz"), cpn.CpelNanoConfig(max_em_init=4), 41, 27))
    fa = os.path.join(INP, "targeted_example.fa.gz")
    chrom = genome.load_genome(fa, engine.lib)["hg38_dna"]
    lay, grp_st, grp_end = simulations.layouts_from_fasta(engine.lib, chrom)
    nb = simulations.make_nano_batch(len(lay["chrst"]), M=15, seed=8, pobs=0.8, layouts=lay)
    tsv = tmp_path / "sim.tsv"
    simulations.write_nanopolish_tsv(str(tsv), nb, "hg38_dna", grp_st, grp_end, fmt="{:.2f}".format)
    cases.append((str(tsv), fa, cpn.CpelNanoConfig(max_em_init=3, min_cov=5.0), None, 137))
    for i, (calls, fasta, cfg, n_cand, n_fit_max) in enumerate(cases):
        store = {}

        def starts(n, store=store, cfg=cfg):
            if n not in store:
                store[n] = simulations.gen_phi0s(np.random.default_rng(100 + n), n, cfg.max_em_init)
            return store[n]

        a = cpn.analyze_nano(calls, fasta, cfg, str(tmp_path / f"a{i}"), "x", phi0_fn=starts, engine=engine)
        for block in (50_000, 16):
            if block == 16:                       # starts are handed out per block of candidates: rebuild them block-wise from the full set
                full = next(iter(store.values()))
                state = {"o": 0}

                def starts_blk(n, full=full, state=state):
                    out = full[state["o"]:state["o"] + n]; state["o"] += n
                    return out
                fn = starts_blk
            else:
                fn = starts
            tag = f"b{i}_{block}"
            st = cpn.analyze_nano_csr(calls, fasta, cfg, str(tmp_path / tag), "x", phi0_fn=fn, engine=engine, block_regions=block)
            assert st["fitted"] == len(a) <= n_fit_max and (n_cand is None or st["candidates"] == n_cand)
            if block == 16 and i == 0:
                continue                          # blocks of partition windows ≠ blocks of candidates only when windows are skipped (N ≤ 9)
            for suffix in ("theta", "ex", "exx", "mml", "nme"):
                fa_, fb_ = tmp_path / f"a{i}" / f"x_{suffix}.txt", tmp_path / tag / f"x_{suffix}.txt"
                assert fa_.read_bytes() == fb_.read_bytes() and fa_.stat().st_size > 0, (i, block, suffix)


def test_sharded_run_merges_to_the_single_process_files(engine, cpn, tmp_path):
    """analyze_nano_csr(rank, world = 3) run shard by shard on this GPU with the real fit, merged: the bytes of the unsharded
    run (regions are independent and the starts are a function of the region)."""
    nano, fasta = os.path.join(INP, "full_example_calls.tsv.gz"), os.path.join(INP, "full_example.fa.gz")
    cfg = cpn.CpelNanoConfig(max_em_init=3)

    def starts(name, st):
        from cpelnano_b200 import simulations
        return np.stack([simulations.gen_phi0s(np.random.default_rng(int(s)), 1, 3)[0] for s in st])

    s1 = cpn.analyze_nano_csr(nano, fasta, cfg, str(tmp_path / "one"), "x", engine=engine, phi0_by_region=starts)
    tot = 0
    for r in range(3):
        tot += cpn.analyze_nano_csr(nano, fasta, cfg, str(tmp_path / "three"), "x", engine=engine, phi0_by_region=starts, rank=r, world=3)["fitted"]
    cpn.merge_shards(str(tmp_path / "three"), "x", 3)
    assert tot == s1["fitted"] >= 20          # with 3 starts a region may end without a converged start (rs.ϕhat = [])
    for suffix in ("theta", "ex", "exx", "mml", "nme"):
        assert (tmp_path / "one" / f"x_{suffix}.txt").read_bytes() == (tmp_path / "three" / f"x_{suffix}.txt").read_bytes()
