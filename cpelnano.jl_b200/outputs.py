"""Output files of the path, byte-compatible with the reference's writers, the model-file reader and the BH correction
(SURVEY.md §8f row 3).

    write_output_{ϕ,ex,exx,mml,nme}, write_output       src/input_output.jl:95-333
    write_output_{tmml,tnme,tcmd}, write_diff_output    src/input_output.jl:340-462
    read_mod_file_chr                                   src/input_output.jl:581-630   (summaries batched on the GPU)
    add_qvalues_to_path, mult_hyp_corr                  src/multiple_hypothesis.jl:14-63 (MultipleTesting.BenjaminiHochberg)

Julia prints a Float64 as its shortest round-trip digits, in positional notation when the decimal exponent is in
[-5, 5] and as `d.ddde±x` otherwise (`1.0e-5`, `1.234567e6`), `NaN`, `Inf`; `round(x, digits=4)` is
`round(x·10⁴)/10⁴` with ties to even.  `julia_str` / `julia_round` restate both so files written here are the text Julia
writes for the same doubles (checked by re-printing every number of the reference's golden files,
tests/test_outputs.py).
"""
import math
import os
from decimal import Decimal

import numpy as np

from .host import RegStruct, rscle_grp_mod


def julia_str(x):
    """string(x::Float64) of Julia (Grisu/Ryu shortest digits; Base/grisu/grisu.jl `_show`: exponential form when the
    decimal point position pt <= -4 or pt > 6)."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    _, dg, ex = Decimal(repr(abs(x))).as_tuple()   # repr = shortest round-trip digits, the digit string Grisu/Ryu produce
    digits = "".join(map(str, dg)).lstrip("0")
    pt = len(digits) + ex                            # |x| = 0.DIGITS × 10^pt
    digits = digits.rstrip("0")
    n = len(digits)
    neg = "-" if x < 0 else ""
    if pt <= -4 or pt > 6:
        out = digits[0] + "." + (digits[1:] if n > 1 else "0") + "e" + str(pt - 1)
    elif pt <= 0:
        out = "0." + "0" * (-pt) + digits
    elif pt >= n:
        out = digits + "0" * (pt - n) + ".0"
    else:
        out = digits[:pt] + "." + digits[pt:]
    return neg + out


def julia_round(x, digits=4):
    """round(x, digits=4) of Julia: round-half-even of x·10^digits, divided back (Base/floatfuncs.jl `_round_digits`)."""
    og = 10.0 ** digits
    r = np.rint(np.asarray(x, dtype=np.float64) * og) / og
    return float(r) if np.ndim(r) == 0 else r


def _join(v):
    return ",".join(julia_str(t) for t in v)


# ---- single-sample outputs -------------------------------------------------------------------------------------------------
class OutputFiles:
    """src/structures.jl:4-20."""

    def __init__(self, out_dir, out_prefix):
        self.theta_file = f"{out_dir}/{out_prefix}_theta.txt"
        self.ex_file = f"{out_dir}/{out_prefix}_ex.txt"
        self.exx_file = f"{out_dir}/{out_prefix}_exx.txt"
        self.mml_file = f"{out_dir}/{out_prefix}_mml.txt"
        self.nme_file = f"{out_dir}/{out_prefix}_nme.txt"

    def all(self):
        return [self.theta_file, self.ex_file, self.exx_file, self.mml_file, self.nme_file]


class OutputDiffFiles:
    """src/structures.jl:23-33."""

    def __init__(self, out_dir, out_prefix):
        self.tmml_file = f"{out_dir}/{out_prefix}_tmml.txt"
        self.tnme_file = f"{out_dir}/{out_prefix}_tnme.txt"
        self.tcmd_file = f"{out_dir}/{out_prefix}_tcmd.txt"

    def all(self):
        return [self.tmml_file, self.tnme_file, self.tcmd_file]


def alpha_beta(phi, Nl, rho_l, dl):
    """get_αβ_from_ϕ (src/link_functions.jl:16-25) — printing only; the kernels have their own."""
    a, b, c = (float(t) for t in phi)
    return (a + b * np.asarray(rho_l)) * np.asarray(Nl), c / np.asarray(dl)


def write_output_phi(path, regs):
    with open(path, "a") as io:
        for rs in regs:
            if not rs.proc:
                continue
            al, be = alpha_beta(rs.phihat, rs.Nl, rs.rho_l, rs.dl)
            io.write(f"{rs.chr}\t{rs.chrst}\t{rs.chrend}\t{_join(rs.phihat)}\t{_join(al)}\t{_join(be)}\n")


def write_output_ex(path, regs):
    with open(path, "a") as io:
        for rs in regs:
            if not rs.proc:
                continue
            for l in range(rs.L):
                if not math.isnan(rs.ex[l]):
                    io.write(f"{rs.chr}\t{rs.cpg_pos[l]}\t{rs.cpg_pos[l]}\t{julia_str(rs.ex[l])}\n")


def write_output_exx(path, regs):
    with open(path, "a") as io:
        for rs in regs:
            if not rs.proc:
                continue
            for l in range(rs.L - 1):
                if not math.isnan(rs.exx[l]):
                    io.write(f"{rs.chr}\t{rs.cpg_pos[l]}\t{rs.cpg_pos[l + 1]}\t{julia_str(rs.exx[l])}\n")


def _write_sub(path, regs, field):
    with open(path, "a") as io:
        for rs in regs:
            if not rs.proc:
                continue
            v = getattr(rs, field)
            for k in range(rs.nls_rgs.num):
                if not math.isnan(v[k]):
                    io.write(f"{rs.chr}\t{rs.nls_rgs.chr_st[k]}\t{rs.nls_rgs.chr_end[k]}\t{julia_str(julia_round(v[k]))}\n")


def write_output_mml(path, regs):
    _write_sub(path, regs, "mml")


def write_output_nme(path, regs):
    _write_sub(path, regs, "nme")


def write_output(regs, files, lib=None):
    """write_output (src/input_output.jl:314-333): theta, exx, ex, mml, nme — all opened in append mode.  With `lib` the text
    is produced by the native writers (cpn_write_theta / cpn_write_rows: same bytes, no per-number Python work)."""
    if lib is not None:
        return _write_output_native(regs, files, lib)
    write_output_phi(files.theta_file, regs)
    write_output_exx(files.exx_file, regs)
    write_output_ex(files.ex_file, regs)
    write_output_mml(files.mml_file, regs)
    write_output_nme(files.nme_file, regs)


def _by_chr(items, key):
    """Consecutive runs of items sharing a chromosome name (files are written in region order)."""
    run, name = [], None
    for it in items:
        k = key(it)
        if run and k != name:
            yield name, run
            run = []
        name = k
        run.append(it)
    if run:
        yield name, run


def _write_output_native(regs, files, lib):
    from . import capi
    done = [rs for rs in regs if rs.proc]
    for f in files.all():                      # open(path, "a") of the reference creates every file, even when no region was processed
        open(f, "a").close()
    for name, rr in _by_chr(done, lambda rs: rs.chr):
        cat = lambda f: np.concatenate([np.asarray(f(rs)) for rs in rr])      # noqa: E731
        grp_off = np.concatenate([[0], np.cumsum([rs.L for rs in rr])]).astype(np.int64)
        capi.write_theta(lib, files.theta_file, name, [rs.chrst for rs in rr], [rs.chrend for rs in rr], np.array([rs.phihat for rs in rr]), grp_off,
                         cat(lambda rs: rs.Nl), cat(lambda rs: rs.rho_l), cat(lambda rs: rs.dl))
        capi.write_rows(lib, files.exx_file, name, cat(lambda rs: rs.cpg_pos[:rs.L - 1]), cat(lambda rs: rs.cpg_pos[1:rs.L]), cat(lambda rs: rs.exx[:rs.L - 1]))
        capi.write_rows(lib, files.ex_file, name, cat(lambda rs: rs.cpg_pos[:rs.L]), cat(lambda rs: rs.cpg_pos[:rs.L]), cat(lambda rs: rs.ex[:rs.L]))
        st, en = cat(lambda rs: rs.nls_rgs.chr_st), cat(lambda rs: rs.nls_rgs.chr_end)
        capi.write_rows(lib, files.mml_file, name, st, en, cat(lambda rs: rs.mml), digits=4)
        capi.write_rows(lib, files.nme_file, name, st, en, cat(lambda rs: rs.nme), digits=4)


# ---- differential outputs --------------------------------------------------------------------------------------------------
def _write_test(path, chrom, tests, row, clamp=False):
    """tests: list of dicts from Engine.grp_comp_batch / *_est_reg_test_batch; row selects the statistic in T / pval."""
    with open(path, "a") as io:
        for chr_name, ts in zip(chrom, tests):
            co = ts["coords"]
            if row == "tcmd_matched":
                T, p = ts["tcmd"], np.full(co.num, np.nan)
            else:
                T, p = ts["T"][row], ts["pval"][row]
            for k in range(co.num):
                if math.isnan(T[k]):
                    continue
                t = max(0.0, T[k]) if clamp else T[k]
                io.write(f"{chr_name}\t{co.chr_st[k]}\t{co.chr_end[k]}\t{julia_str(julia_round(t))}\t{julia_str(p[k])}\n")


def write_diff_output(chrom, tests, files, matched=False, lib=None):
    """write_diff_output (src/input_output.jl:448-462).  `chrom[i]` is the chromosome name of tests[i]."""
    if lib is not None:
        from . import capi
        for f in files.all():
            open(f, "a").close()
        for name, run in _by_chr(list(zip(chrom, tests)), lambda ct: ct[0]):
            ts = [t for _, t in run]
            st = np.concatenate([t["coords"].chr_st for t in ts]); en = np.concatenate([t["coords"].chr_end for t in ts])
            capi.write_rows(lib, files.tmml_file, name, st, en, np.concatenate([t["T"][0] for t in ts]), 4, False, np.concatenate([t["pval"][0] for t in ts]))
            capi.write_rows(lib, files.tnme_file, name, st, en, np.concatenate([t["T"][1] for t in ts]), 4, False, np.concatenate([t["pval"][1] for t in ts]))
            if matched:
                tc = np.concatenate([t["tcmd"] for t in ts]); pc = np.full(len(tc), np.nan)
            else:
                tc = np.concatenate([t["T"][2] for t in ts]); pc = np.concatenate([t["pval"][2] for t in ts])
            capi.write_rows(lib, files.tcmd_file, name, st, en, tc, 4, True, pc)
        return
    _write_test(files.tmml_file, chrom, tests, 0)
    _write_test(files.tnme_file, chrom, tests, 1)
    _write_test(files.tcmd_file, chrom, tests, "tcmd_matched" if matched else 2, clamp=True)


def bh_adjust(p):
    """MultipleTesting.adjust(p, BenjaminiHochberg()): q_(i) = min_{j ≥ i} min(1, p_(j)·n/j) over the sorted p-values."""
    p = np.asarray(p, dtype=np.float64)
    n = len(p)
    if n <= 1:
        return p.copy()
    order = np.argsort(p, kind="stable")
    ps = p[order] * (n / np.arange(1, n + 1))          # p·(n/rank): MultipleTesting's multiplier form, bit for bit
    q = np.minimum(1.0, np.minimum.accumulate(ps[::-1])[::-1])
    out = np.empty(n)
    out[order] = q
    return out


def add_qvalues_to_path(path):
    """add_qvalues_to_path (src/multiple_hypothesis.jl:14-49): append a BH q-value column (NaN where p is NaN); the file is
    re-written through readdlm/writedlm, i.e. every float is re-printed in Julia's shortest form."""
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        return
    rows = [line.rstrip("\n").split("\t") for line in open(path)]
    p = np.array([float(r[4]) for r in rows])
    q = np.full(len(rows), np.nan)
    ok = ~np.isnan(p)
    if ok.any():
        q[ok] = bh_adjust(p[ok])
    tmp = path + ".tmp"
    with open(tmp, "w") as io:
        for r, qv in zip(rows, q):
            io.write("\t".join([r[0], r[1], r[2], julia_str(float(r[3])), julia_str(float(r[4])), julia_str(qv)]) + "\n")
    os.replace(tmp, path)


def mult_hyp_corr(files, matched=False, lib=None):
    """mult_hyp_corr (src/multiple_hypothesis.jl:55-63); with `lib` through the native cpn_add_qvalues (same bytes)."""
    if lib is not None:
        from . import capi
        add = lambda p: capi.add_qvalues(lib, p)      # noqa: E731
    else:
        add = add_qvalues_to_path
    add(files.tmml_file)
    add(files.tnme_file)
    if not matched:
        add(files.tcmd_file)


# ---- model files ---------------------------------------------------------------------------------------------------------------
def read_theta_lines(path, chrom=None):
    """[(chr, chrst, chrend, phi[3])] of a model (theta) file, optionally for one chromosome."""
    out = []
    with open(path) as f:
        for line in f:
            v = line.rstrip("\n").split("\t")
            if len(v) < 4 or (chrom is not None and v[0] != chrom):
                continue
            out.append((v[0], int(v[1]), int(v[2]), np.array([float(t) for t in v[3].split(",")])))
    return out


def read_mod_file_chr(path, chromosome, config, engine=None, summaries=True):
    """read_mod_file_chr (src/input_output.jl:581-630): {"chr_st_end": RegStruct} of one sample on one chromosome, every
    model rescaled to the per-CpG chain.  With summaries=True get_stat_sums! runs for all regions of the file in ONE device
    call; summaries=False leaves ϕ̂ + features only (enough for Engine.grp_comp_batch, which recomputes them on the device)."""
    from .host import default_engine
    regs = {}
    for (c, st, end, phi) in read_theta_lines(path, chromosome.name):
        rs = chromosome.region_struct(st, end, config.min_grp_dist)
        rs.phihat = phi
        rscle_grp_mod(rs)
        rs.proc = True
        regs[f"{c}_{st}_{end}"] = rs
    eng = engine or default_engine()
    if summaries:
        eng.get_stat_sums_batch(list(regs.values()), config)
    else:
        for rs in regs.values():
            eng.get_nls_reg_info(rs, config)
    return regs
