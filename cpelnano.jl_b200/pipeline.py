"""User entry points of the reference, file in → file out, with the per-region `pmap` bodies replaced by one device batch
per chromosome (SURVEY.md §8b "batch points"):

    analyze_nano(nano, fasta, config)                          src/nanopore.jl:362-416
    diff_grp_comp(mod_fls_g1, mod_fls_g2, fasta, config)       src/grp_perm_tests.jl:637-687
    diff_two_samp_comp(mod1, mod2, nano1, nano2, fasta, config) src/two_samp_perm_tests.jl:408-460

Same arguments, same output files (outputs.py restates Julia's number formatting).  Two documented differences, both about
ORDER only: the reference iterates Julia `Dict`s (reads of a region in get_calls_nanopolish!, common regions in
find_comm_regs), whose order depends on the Julia build; here reads follow their first appearance in the calls file and
common regions are written in coordinate order.  One difference about RANDOMNESS, only when a comparison has too many labelings to
enumerate (≥ 1000: more than 6 vs 6 unmatched, or ≥ 10 matched pairs): the reference draws its random subset of labelings inside
every region's test, here one subset is drawn per (n1, n2) batch and shared by the regions of the batch — each region's p-value
has the same distribution, but their Monte-Carlo errors are correlated across regions (exact tests are unaffected).  Host work is native and batched per chromosome: CpG index, partition and region features (csrc/cpn_genome.cpp), calls
parsing and binning (csrc/cpn_calls.cpp); every number comes from libcpelnano_b200.so.
"""
import os

import numpy as np

from . import genome, outputs, sharding
from .host import default_engine, analyze_regions
from .nanopolish import NativeCalls


def _print_log(config, msg):
    if getattr(config, "verbose", False):
        print(f"[cpelnano_b200] {msg}", flush=True)


def _outputs_exist(paths):
    """check_output_exists / check_diff_output_exists (src/input_output.jl:36-88): any existing output file stops the run."""
    return any(os.path.exists(p) for p in paths)


def analyze_nano(nano, fasta, config, out_dir, out_prefix, bed_reg=None, mod_nse=True, rng=None, phi0_fn=None, engine=None, timings=None):
    """Estimate ϕ̂ for every estimation region of every chromosome and write theta / ex / exx / mml / nme files.
    out_dir, out_prefix, bed_reg, mod_nse are fields of the reference's config struct (src/structures.jl:36-69) passed
    explicitly here.  phi0_fn(n_regions) → [n, max_em_init, 3] replaces the unseeded gen_ϕ0s draws (tests); default: rng.
    timings: optional dict that receives the wall seconds of each stage.  Returns the list of processed RegStructs (the
    reference returns nothing)."""
    import time
    tm = timings if timings is not None else {}

    def lap(key, t0):
        tm[key] = tm.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    eng = engine or default_engine()
    t = time.perf_counter()
    chroms = genome.load_genome(fasta, eng.lib)
    t = lap("fasta_and_cpg_index_s", t)
    os.makedirs(out_dir, exist_ok=True)
    files = outputs.OutputFiles(out_dir, out_prefix)
    if _outputs_exist(files.all()):
        return None
    targets = genome.read_bed(bed_reg) if bed_reg else None
    t = time.perf_counter()
    calls = NativeCalls(nano, eng.lib)
    t = lap("calls_parse_s", t)
    done = []
    for name, chrom in chroms.items():
        if targets is not None and name not in targets:
            continue
        t = time.perf_counter()
        part = chrom.chr_part_native(eng.lib, config.size_est_reg, config.min_grp_dist, None if targets is None else targets[name])
        _print_log(config, f"chr {name}: {len(part)} estimation regions")
        regs = chrom.region_structs(eng.lib, part, config.min_grp_dist)
        t = lap("partition_and_features_s", t)
        cand = [rs for rs in regs if len(rs.cpg_pos) > 9]      # pmap_analyze_reg returns before get_calls! otherwise (:308)
        calls.attach_calls_batch(cand, mod_nse)
        t = lap("calls_fill_s", t)
        phi0s = None
        if phi0_fn is not None:
            phi0s = phi0_fn(len(cand))
        analyze_regions(cand, config, phi0s=phi0s, rng=rng, engine=eng)
        t = lap("fit_and_summaries_s", t)
        # write_output_ϕ runs after rscle_grp_mod!, so the theta file holds the per-CpG α, β (as the reference's does)
        outputs.write_output(regs, files, eng.lib)
        t = lap("write_outputs_s", t)
        done.extend(rs for rs in regs if rs.proc)
        tm["regions"] = tm.get("regions", 0) + len(part)
        tm["regions_fitted"] = tm.get("regions_fitted", 0) + sum(1 for rs in regs if rs.proc)
    return done


def analyze_bam(bam, fasta, config, out_dir, out_prefix, bed_reg=None, rng=None, phi0_fn=None, engine=None, lib=None, timings=None):
    """analyze_bam (src/wgbs.jl:401-464): ϕ̂, MML and NME of every estimation region from a Bismark / Arioc BAM file.  Per
    chromosome: partition (get_chr_part with config.min_grp_dist, :441), per-CpG features (get_grp_info!(rs, fa_rec, 1), :351),
    the calls of ALL its regions in one pass over the BAM (capi.BamHandle.calls = get_calls_bam, :357), then the body of
    pmap_analyze_reg_bam for the whole batch (host.analyze_regions_wgbs) and the writers.  config.pe / config.trim as in the
    reference.  The library is used without a device context until a region passes the data filters, so a run in which none
    does (the reference's own example: 7 CpG sites) needs no GPU.  Returns the list of processed RegStructs."""
    import time
    from . import capi
    from .host import analyze_regions_wgbs
    tm = timings if timings is not None else {}

    def lap(key, t0):
        tm[key] = tm.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    lib = lib or (engine.lib if engine is not None else capi.Library())
    holder = {"eng": engine}

    def get_engine():
        if holder["eng"] is None:
            holder["eng"] = default_engine()
        return holder["eng"]

    t = time.perf_counter()
    chroms = genome.load_genome(fasta, lib)
    t = lap("fasta_and_cpg_index_s", t)
    os.makedirs(out_dir, exist_ok=True)
    files = outputs.OutputFiles(out_dir, out_prefix)
    if _outputs_exist(files.all()):
        return None
    targets = genome.read_bed(bed_reg) if bed_reg else None
    t = time.perf_counter()
    reads = capi.BamHandle(lib, bam)
    t = lap("bam_inflate_and_index_s", t)
    done = []
    for name, chrom in chroms.items():
        if targets is not None and name not in targets:
            continue
        t = time.perf_counter()
        part = chrom.chr_part_native(lib, config.size_est_reg, config.min_grp_dist, None if targets is None else targets[name])
        _print_log(config, f"chr {name}: {len(part)} estimation regions")
        regs = chrom.region_structs(lib, part, 1)                 # every CpG its own group (:351)
        t = lap("partition_and_features_s", t)
        cand = [rs for rs in regs if len(rs.cpg_pos) > 9]          # pmap_analyze_reg_bam returns before get_calls_bam otherwise (:353)
        calls = []
        if cand:
            cpg_off = np.concatenate([[0], np.cumsum([len(rs.cpg_pos) for rs in cand])]).astype(np.int64)
            read_off, flat = reads.calls(name, [rs.chrst for rs in cand], [rs.chrend for rs in cand], cpg_off,
                                         np.concatenate([rs.cpg_pos for rs in cand]), config.pe, config.trim)
            cell = np.concatenate([[0], np.cumsum(np.diff(read_off) * np.diff(cpg_off))])
            calls = [flat[cell[i]:cell[i + 1]].reshape(int(read_off[i + 1] - read_off[i]), int(cpg_off[i + 1] - cpg_off[i])) for i in range(len(cand))]
        t = lap("calls_s", t)
        phi0s = phi0_fn(len(cand)) if (phi0_fn is not None and cand) else None
        if cand:
            analyze_regions_wgbs(cand, calls, config, phi0s=phi0s, rng=rng, engine=get_engine)
        t = lap("fit_and_summaries_s", t)
        outputs.write_output(regs, files, lib)
        t = lap("write_outputs_s", t)
        done.extend(rs for rs in regs if rs.proc)
        tm["regions"] = tm.get("regions", 0) + len(part)
        tm["regions_fitted"] = tm.get("regions_fitted", 0) + sum(1 for rs in regs if rs.proc)
    return done


def analyze_nano_csr(nano, fasta, config, out_dir, out_prefix, bed_reg=None, mod_nse=True, rng=None, phi0_fn=None, engine=None, timings=None,
                     block_regions=50_000, rank=0, world=1, phi0_by_region=None):
    """analyze_nano without per-region Python objects: every stage works on the CSR arrays of a block of estimation regions —
    native features → native calls → filters of pmap_analyze_reg (vectorised) → ONE cpn_analyze_batch → native writers.
    Writes the same bytes as analyze_nano (tests/test_pipeline_gpu.py); this is the path for whole chromosomes, where a Python
    loop over 10⁵–10⁶ regions would cost more than everything else together.  Returns a dict of counters.

    Multi-GPU (one process per GPU): with world > 1 rank k analyses the k-th contiguous slice of every
    chromosome's estimation regions on its own device and writes part files; after a barrier one rank calls merge_shards,
    which concatenates the parts in (chromosome, rank) order — the bytes of a single-process run.  No collective is needed.
    phi0_by_region(chr_name, chrst[n]) → [n, max_em_init, 3] makes the starts a function of the region (used by tests so that
    sharded and unsharded runs see the same starts)."""
    import time
    from . import capi
    from .host import AMAX, BMAX, CMAX
    from .simulations import gen_phi0s
    tm = timings if timings is not None else {}

    def lap(key, t0):
        tm[key] = tm.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    eng = engine or default_engine()
    lib = eng.lib
    t = time.perf_counter()
    chroms = genome.load_genome(fasta, lib)
    t = lap("fasta_and_cpg_index_s", t)
    os.makedirs(out_dir, exist_ok=True)
    final = outputs.OutputFiles(out_dir, out_prefix)
    if _outputs_exist(final.all()):
        return None
    if world == 1:
        files = final
        for f in files.all():
            open(f, "a").close()
    else:
        os.makedirs(os.path.join(out_dir, SHARD_DIR), exist_ok=True)
    targets = genome.read_bed(bed_reg) if bed_reg else None
    calls = NativeCalls(nano, lib)
    t = lap("calls_parse_s", t)
    stats = dict(regions=0, candidates=0, with_data=0, fitted=0)
    for ci, (name, chrom) in enumerate(chroms.items()):
        if targets is not None and name not in targets:
            continue
        t = time.perf_counter()
        part = chrom.chr_part_native(lib, config.size_est_reg, config.min_grp_dist, None if targets is None else targets[name])
        st_all = np.array([p[0] for p in part], dtype=np.int64); en_all = np.array([p[1] for p in part], dtype=np.int64)
        if world > 1:                                           # this rank's contiguous slice of the chromosome's regions
            b = (len(part) * np.arange(world + 1)) // world
            st_all, en_all = st_all[b[rank]:b[rank + 1]], en_all[b[rank]:b[rank + 1]]
            files = outputs.OutputFiles(os.path.join(out_dir, SHARD_DIR), _shard_prefix(out_prefix, ci, rank))
            for f in files.all():
                open(f, "w").close()
        stats["regions"] += len(st_all)
        for b0 in range(0, len(st_all), block_regions):
            st, en = st_all[b0:b0 + block_regions], en_all[b0:b0 + block_regions]
            N = np.empty(len(st), dtype=np.int64); L = np.empty(len(st), dtype=np.int64)
            capi.grp_info_counts(lib, chrom.cg, st, en, config.min_grp_dist, N, L)
            cand = np.nonzero(N > 9)[0]                          # length(rs.cpg_pos) > 9 || return rs (:308)
            stats["candidates"] += len(cand)
            if len(cand) == 0:
                continue
            c = capi.grp_info_batch(lib, chrom.cg, chrom.size, st[cand], en[cand], config.min_grp_dist)
            t = lap("partition_and_features_s", t)
            read_off, lu, _ = calls.handle.fill(name, st[cand], en[cand], c["grp_off"], c["grp_st"], c["grp_end"], mod_nse)
            # get_ave_depth >= min_cov, perc_gprs_obs >= 0.66, L > 1 (nanopore.jl:322-326; structures.jl:200-203), vectorised
            Lc, m = np.diff(c["grp_off"]), np.diff(read_off)
            n_obs, n_grp = capi.calls_coverage(lib, c["grp_off"], read_off, lu)
            keep = (m > 0) & (n_obs / np.maximum(Lc, 1) >= config.min_cov) & (n_grp / np.maximum(Lc, 1) >= 0.66) & (Lc > 1)
            del lu
            sel = cand[keep]
            stats["with_data"] += len(sel)
            if len(sel) == 0:
                t = lap("calls_fill_s", t)
                continue
            k = c if keep.all() else capi.grp_info_batch(lib, chrom.cg, chrom.size, st[sel], en[sel], config.min_grp_dist)
            read_off, lu, lm = calls.handle.fill(name, st[sel], en[sel], k["grp_off"], k["grp_st"], k["grp_end"], mod_nse)
            t = lap("calls_fill_s", t)
            if phi0_by_region is not None:
                phi0 = np.asarray(phi0_by_region(name, st[sel]))
            elif phi0_fn is not None:
                phi0 = np.asarray(phi0_fn(len(cand)))[keep]
            else:
                phi0 = gen_phi0s(rng or np.random.default_rng(), len(sel), config.max_em_init)
            sub_off, sub_st, sub_en, sub_p, sub_q = capi.nls_reg_info_batch(lib, st[sel], en[sel], config.max_size_subreg, k["cpg_off"], k["cpg_pos"])
            out = lib.analyze_batch(k["grp_off"], k["Nl"], k["rho_l"], k["d_l"], read_off, lu, lm, phi0, config.max_em_init, k["cpg_off"], k["rho_n"],
                                    k["d_n"], sub_off, sub_p, sub_q, config.max_em_iters, (AMAX, BMAX, CMAX), config.mstep_mode, device=eng.device)
            t = lap("fit_and_summaries_s", t)
            ok = out["pick"] >= 0
            stats["fitted"] += int(ok.sum())
            Nk, Kk = np.diff(k["cpg_off"]), np.diff(sub_off)
            site = np.repeat(ok, Nk)                                         # per-CpG rows of fitted regions
            first = np.zeros(len(site), dtype=bool); first[k["cpg_off"][:-1]] = True
            last = np.zeros(len(site), dtype=bool); last[k["cpg_off"][1:] - 1] = True
            pair = np.repeat(ok, Nk - 1)
            pos = k["cpg_pos"]
            capi.write_theta(lib, files.theta_file, name, st[sel][ok], en[sel][ok], out["phi_hat"][ok], np.concatenate([[0], np.cumsum(Nk[ok])]), None,
                             k["rho_n"][site], k["d_n"][pair])
            capi.write_rows(lib, files.exx_file, name, pos[~last][pair], pos[~first][pair], out["exx"][pair])
            capi.write_rows(lib, files.ex_file, name, pos[site], pos[site], out["ex"][site])
            sub = np.repeat(ok, Kk)
            capi.write_rows(lib, files.mml_file, name, sub_st[sub], sub_en[sub], out["mml"][sub], digits=4)
            capi.write_rows(lib, files.nme_file, name, sub_st[sub], sub_en[sub], out["nme"][sub], digits=4)
            t = lap("write_outputs_s", t)
    tm.update(stats)
    return stats


SHARD_DIR = "_shards"


def _shard_prefix(out_prefix, ci, rank):
    return f"{out_prefix}.c{ci:05d}.r{rank:04d}"


def merge_shards(out_dir, out_prefix, world, n_chr=None):
    """Concatenate the part files the ranks of a sharded analyze_nano_csr run wrote, in (chromosome, rank) order, into the five
    output files, and remove the parts.  Call on one rank after all ranks have finished."""
    import glob
    import shutil
    final = outputs.OutputFiles(out_dir, out_prefix)
    sd = os.path.join(out_dir, SHARD_DIR)
    if n_chr is None:
        found = glob.glob(os.path.join(sd, f"{out_prefix}.c*.r0000_theta.txt"))
        cis = sorted(int(os.path.basename(f).split(".c")[-1].split(".r")[0]) for f in found)
    else:
        cis = list(range(n_chr))
    for suffix, dst in zip(("theta", "ex", "exx", "mml", "nme"), final.all()):
        with open(dst, "ab") as out:
            for ci in cis:
                for r in range(world):
                    src = os.path.join(sd, f"{_shard_prefix(out_prefix, ci, r)}_{suffix}.txt")
                    if not os.path.exists(src):
                        if os.path.exists(os.path.join(sd, f"{_shard_prefix(out_prefix, ci, 0)}_{suffix}.txt")):
                            raise FileNotFoundError(f"shard of rank {r} missing: {src}")
                        continue
                    with open(src, "rb") as f:
                        shutil.copyfileobj(f, out, 1 << 24)
    shutil.rmtree(sd, ignore_errors=True)
    return final


def _find_comm_regs(ms_g1, ms_g2):
    """find_comm_regs (src/grp_perm_tests.jl:527-556): region ids present in at least one sample of each group."""
    k1 = set().union(*[set(d) for d in ms_g1]) if ms_g1 else set()
    k2 = set().union(*[set(d) for d in ms_g2]) if ms_g2 else set()
    return k1 & k2


def diff_grp_comp(mod_fls_g1, mod_fls_g2, fasta, config, out_dir, out_prefix, rng=None, engine=None):
    """Group comparison from model (theta) files: per chromosome read every sample's ϕ̂, find the common regions, run
    unmat_est_reg_test / mat_est_reg_test (config.matched) for all of them on the device, write tmml / tnme / tcmd and add BH
    q-values.  Regions are batched by their (n1, n2) sample counts (one cpn_grp_test_batch call per distinct pair)."""
    eng = engine or default_engine()
    chroms = genome.load_genome(fasta)
    os.makedirs(out_dir, exist_ok=True)
    files = outputs.OutputDiffFiles(out_dir, out_prefix)
    if _outputs_exist(files.all()):
        return None
    for name, chrom in chroms.items():
        ms_g1 = [outputs.read_mod_file_chr(f, chrom, config, eng, summaries=False) for f in mod_fls_g1]
        ms_g2 = [outputs.read_mod_file_chr(f, chrom, config, eng, summaries=False) for f in mod_fls_g2]
        comm = _find_comm_regs(ms_g1, ms_g2)
        if not comm:
            continue
        batches = {}
        for rid in comm:
            if config.matched:                           # pairs present in both groups (:576-586)
                reps = [i for i in range(min(len(ms_g1), len(ms_g2))) if rid in ms_g1[i] and rid in ms_g2[i]]
                g1, g2 = [ms_g1[i][rid] for i in reps], [ms_g2[i][rid] for i in reps]
            else:
                g1, g2 = [d[rid] for d in ms_g1 if rid in d], [d[rid] for d in ms_g2 if rid in d]
            if not g1 or not g2:
                continue
            batches.setdefault((len(g1), len(g2)), []).append((g1[0].chrst, g1[0].chrend, g1, g2))
        tests, keys = [], []
        for (n1, n2), items in batches.items():
            items.sort(key=lambda t: (t[0], t[1]))
            out = eng.grp_comp_batch([t[2] for t in items], [t[3] for t in items], matched=config.matched, rng=rng)
            tests.extend(out); keys.extend((t[0], t[1]) for t in items)
        order = sorted(range(len(tests)), key=lambda i: keys[i])
        tests = [tests[i] for i in order]
        outputs.write_diff_output([name] * len(tests), tests, files, matched=config.matched, lib=eng.lib)
    outputs.mult_hyp_corr(files, matched=config.matched)
    return files


def _read_theta_columns(path):
    """{chr: (chrst[n], chrend[n], phi[n, 3])} of a model file; only the first four tab-separated fields of a line are split
    (the α, β lists that follow are kilobytes long and are not needed: read_mod_file_chr recomputes everything from ϕ̂)."""
    cols = {}
    with open(path) as f:
        for line in f:
            v = line.split("\t", 4)
            if len(v) < 4:
                continue
            c = cols.setdefault(v[0], ([], [], []))
            c[0].append(int(v[1])); c[1].append(int(v[2])); c[2].append(v[3].rstrip("\n"))
    out = {}
    for name, (a, b, p) in cols.items():
        out[name] = (np.array(a, dtype=np.int64), np.array(b, dtype=np.int64),
                     np.array([t.split(",") for t in p], dtype=np.float64).reshape(len(p), 3))
    return out


def diff_grp_comp_csr(mod_fls_g1, mod_fls_g2, fasta, config, out_dir, out_prefix, rng=None, engine=None, block_regions=200_000, timings=None):
    """diff_grp_comp on arrays: model files → (chrst, chrend, ϕ̂) columns; the features of the common regions come from ONE native
    feature batch per block (they depend on the region, not on the sample); one cpn_grp_test_batch per distinct (n1, n2) and block;
    native writers.  Writes the same bytes as diff_grp_comp (tests/test_pipeline_gpu.py)."""
    from itertools import combinations
    import math
    from . import capi
    from .host import LMAX
    import time
    tm = timings if timings is not None else {}

    def lap(key, t0):
        tm[key] = tm.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    eng = engine or default_engine()
    lib = eng.lib
    t = time.perf_counter()
    chroms = genome.load_genome(fasta, lib)
    t = lap("fasta_and_cpg_index_s", t)
    os.makedirs(out_dir, exist_ok=True)
    files = outputs.OutputDiffFiles(out_dir, out_prefix)
    if _outputs_exist(files.all()):
        return None
    for f in files.all():
        open(f, "a").close()
    cols = [[capi.read_theta(lib, f) for f in grp] for grp in (mod_fls_g1, mod_fls_g2)]        # native; _read_theta_columns is the Python statement
    t = lap("read_model_files_s", t)
    matched = bool(config.matched)
    for name, chrom in chroms.items():
        # union of the regions of all samples, in coordinate order, and each sample's row for every region (−1 = absent)
        keys = set()
        for grp in cols:
            for c in grp:
                if name in c:
                    keys.update(zip(c[name][0].tolist(), c[name][1].tolist()))
        if not keys:
            continue
        regs = sorted(keys)
        index = {k: i for i, k in enumerate(regs)}
        rows = []
        for grp in cols:
            rg = np.full((len(grp), len(regs)), -1, dtype=np.int64)
            for s_i, c in enumerate(grp):
                if name in c:
                    rg[s_i, [index[k] for k in zip(c[name][0].tolist(), c[name][1].tolist())]] = np.arange(len(c[name][0]))
            rows.append(rg)
        if matched:                                          # replicate i of both groups must hold the region (:576-586)
            n = min(len(rows[0]), len(rows[1]))
            both = (rows[0][:n] >= 0) & (rows[1][:n] >= 0)
            pres = [both, both]
            rows = [rows[0][:n], rows[1][:n]]
        else:
            pres = [rows[0] >= 0, rows[1] >= 0]
        n1s, n2s = pres[0].sum(axis=0), pres[1].sum(axis=0)
        st_all = np.array([k[0] for k in regs], dtype=np.int64); en_all = np.array([k[1] for k in regs], dtype=np.int64)
        results = []                                         # (region indices, T, p[, tcmd], sub_off, sub_st, sub_en) per batch
        for (n1, n2) in sorted(set(zip(n1s.tolist(), n2s.tolist()))):
            if n1 == 0 or n2 == 0:
                continue
            ridx_all = np.nonzero((n1s == n1) & (n2s == n2))[0]
            S = n1 + n2
            if matched:
                if 0.5 ** n1 < 0.05:
                    exact = 2 ** n1 < LMAX
                    lab = np.arange(2 ** n1, dtype=np.int64) if exact else (rng or np.random.default_rng()).integers(0, 2 ** n1, size=LMAX)
                else:
                    exact, lab = True, np.zeros(0, dtype=np.int64)
            else:
                Lc = math.comb(S, n1)
                if Lc > 20:
                    lab, exact = sharding.labelings_for(S, n1, LMAX, rng)
                else:
                    exact, lab = True, np.zeros((0, n1), dtype=np.int32)
            for b0 in range(0, len(ridx_all), block_regions):
                ridx = ridx_all[b0:b0 + block_regions]
                phi = np.empty((len(ridx), S, 3))
                for g, (grp, off) in enumerate(((cols[0], 0), (cols[1], n1))):
                    # samples holding each region, in file order (the reference pushes them in that order, :590-597)
                    order = np.argsort(~pres[g][:, ridx], axis=0, kind="stable")[: (n1 if g == 0 else n2)]       # [n_g, R] sample ids
                    for j in range(order.shape[0]):
                        smp = order[j]
                        for s_i in np.unique(smp):
                            m = smp == s_i
                            phi[m, off + j] = grp[s_i][name][2][rows[g][s_i, ridx[m]]]
                t = lap("match_regions_and_samples_s", t)
                k = capi.grp_info_batch(lib, chrom.cg, chrom.size, st_all[ridx], en_all[ridx], config.min_grp_dist)
                sub_off, sub_st, sub_en, sub_p, sub_q = capi.nls_reg_info_batch(lib, st_all[ridx], en_all[ridx], config.max_size_subreg, k["cpg_off"],
                                                                                 k["cpg_pos"])
                t = lap("features_s", t)
                out = lib.grp_test_batch(n1, n2, matched, phi, k["cpg_off"], k["rho_n"], k["d_n"], sub_off, sub_p, sub_q, lab, exact, device=eng.device)
                t = lap("summaries_cmd_and_sweep_gpu_s", t)
                tm["regions"] = tm.get("regions", 0) + len(ridx)
                results.append((ridx, out, sub_off, sub_st, sub_en))
        if not results:
            continue
        # rows of all batches in coordinate order of their regions
        ns = 2 if matched else 3
        reg_of_row, st_r, en_r, Ts, Ps = [], [], [], [[] for _ in range(3)], [[] for _ in range(3)]
        for (ridx, out, sub_off, sub_st, sub_en) in results:
            K = np.diff(sub_off)
            within = np.arange(sub_off[-1]) - np.repeat(sub_off[:-1], K)
            reg_of_row.append(np.repeat(ridx, K)); st_r.append(sub_st); en_r.append(sub_en)
            for j in range(ns):
                idx = np.repeat(ns * sub_off[:-1] + j * K, K) + within
                Ts[j].append(out[0][idx]); Ps[j].append(out[2][idx])
            if matched:
                Ts[2].append(out[3]); Ps[2].append(np.full(len(out[3]), np.nan))
        reg_of_row = np.concatenate(reg_of_row)
        o = np.argsort(reg_of_row, kind="stable")
        st_r, en_r = np.concatenate(st_r)[o], np.concatenate(en_r)[o]
        for j, (path, clamp) in enumerate(((files.tmml_file, False), (files.tnme_file, False), (files.tcmd_file, True))):
            capi.write_rows(lib, path, name, st_r, en_r, np.concatenate(Ts[j])[o], 4, clamp, np.concatenate(Ps[j])[o])
        t = lap("write_outputs_s", t)
    t = time.perf_counter()
    outputs.mult_hyp_corr(files, matched=matched, lib=lib)
    lap("bh_qvalues_s", t)
    return files


def diff_two_samp_comp(mods_path_s1, mods_path_s2, nano_s1, nano_s2, fasta, config, out_dir, out_prefix, mod_nse=True, rng=None,
                       engine=None):
    """Two-sample comparison from two model files and the two samples' calls: observed Tmml / Tnme / Tcmd per sub-region, and
    (config.pval_comp) permutation p-values from refits of the pooled reads — every region of a chromosome in one device
    batch (the reference maps serially over regions and pmaps over permutations, :445 / :220).  BH q-values only when
    p-values were computed (:452-455)."""
    eng = engine or default_engine()
    chroms = genome.load_genome(fasta)
    os.makedirs(out_dir, exist_ok=True)
    files = outputs.OutputDiffFiles(out_dir, out_prefix)
    if _outputs_exist(files.all()):
        return None
    calls1 = NativeCalls(nano_s1, eng.lib) if config.pval_comp else None
    calls2 = (calls1 if nano_s2 == nano_s1 else NativeCalls(nano_s2, eng.lib)) if config.pval_comp else None
    for name, chrom in chroms.items():
        m1 = outputs.read_mod_file_chr(mods_path_s1, chrom, config, eng)
        m2 = outputs.read_mod_file_chr(mods_path_s2, chrom, config, eng)
        comm = sorted(set(m1) & set(m2), key=lambda k: (m1[k].chrst, m1[k].chrend))
        if not comm:
            continue
        items = []
        for rid in comm:
            a, b = m1[rid], m2[rid]
            if config.pval_comp:                         # get_calls! on the GROUP intervals of the region (:187-188)
                ga, gb = a.group_view or a, b.group_view or b
                c1 = calls1.region_calls(a.chr, a.chrst, a.chrend, ga.grp_st, ga.grp_end, mod_nse)[:2]
                c2 = calls2.region_calls(b.chr, b.chrst, b.chrend, gb.grp_st, gb.grp_end, mod_nse)[:2]
            else:
                c1 = c2 = (np.zeros((0, 0)), np.zeros((0, 0)))
            items.append((a, b, c1, c2))
        tests = eng.diff_two_samp_comp_batch(items, config, rng=rng)
        outputs.write_diff_output([name] * len(tests), tests, files, matched=False, lib=eng.lib)
    if config.pval_comp:
        outputs.mult_hyp_corr(files, matched=False)
    return files
