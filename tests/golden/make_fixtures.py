#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from /root/reference.

TEST INFRASTRUCTURE.  Run once in the build container (the reference checkout does not exist on the
GPU box):  python tests/golden/make_fixtures.py

What it does
  * restates the host-side genomic feature extraction of the reference (which stays Julia in the
    product) just far enough to turn the bundled FASTA windows into per-region inputs:
      get_grp_info!        src/genome_analysis.jl:226-277  (CpG positions, ρ_n, d_n, groups)
      get_grps_from_cgs    src/genome_analysis.jl:63-111
      get_ρn / get_ρl      src/genome_analysis.jl:122-168
      get_chr_part (BED)   src/genome_analysis.jl:391-458
  * copies the reference's golden OUTPUT numbers (theta/ex/exx/mml/nme files, group-comparison
    test tables) into compact .npz files next to the inputs they belong to.
No reference source code is copied; only data files are read.
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CG = re.compile(r"[Cc][Gg]")


def read_fasta(path):
    seq = []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                continue
            seq.append(line.strip())
    return "".join(seq)


def grps_from_cgs(cpg_pos, min_grp_dist):
    """genome_analysis.jl:63-111 → list of (first_idx, last_idx) 0-based inclusive, and padded intervals."""
    if len(cpg_pos) == 1:
        return [(0, 0)], [(cpg_pos[0], cpg_pos[0])]
    grps, ints = [], []
    ind0, ind1 = 0, 0
    g0, g1 = cpg_pos[0], cpg_pos[0]
    i = 1
    n = len(cpg_pos)
    while i < n:
        if cpg_pos[i] - g1 > min_grp_dist:
            grps.append((ind0, ind1)); ints.append((g0, g1))
            ind0 = ind1 = i
            g0 = g1 = cpg_pos[i]
        else:
            ind1 = i
            g1 = cpg_pos[i]
        if i == n - 1:
            grps.append((ind0, ind1)); ints.append((g0, g1))
        i += 1
    ints = [(a - 5, b + 5) for a, b in ints]
    return grps, ints


def grp_info(seq, chrst, chrend, min_grp_dist=10):
    """genome_analysis.jl:226-277 (1-based closed coordinates)."""
    chr_size = len(seq)
    dna = seq[chrst - 1:chrend]
    cpg_pos = np.array([m.start() + 1 + chrst - 1 for m in CG.finditer(dna)], dtype=np.int64)
    out = {"cpg_pos": cpg_pos, "N": len(cpg_pos)}
    if len(cpg_pos) <= 1:
        return out
    grps, _ = grps_from_cgs(list(cpg_pos), min_grp_dist)
    cont_st = max(1, chrst - 500)
    cont_end = min(chr_size, chrend + 500)
    cont = seq[cont_st - 1:cont_end]
    off = cpg_pos - cont_st                       # get_ρn is handed cpg_pos .- dna_cont_st
    rho_n = np.empty(len(cpg_pos))
    for n, o in enumerate(off):                   # genome_analysis.jl:122-140, width=1000
        win_st = max(1, int(o) - 500)
        win_end = min(len(cont), int(o) + 500)
        rho_n[n] = len(CG.findall(cont[win_st - 1:win_end]))
    rho_n = rho_n / 1000
    d_n = np.diff(cpg_pos).astype(np.float64)
    Nl = np.array([b - a + 1 for a, b in grps], dtype=np.float64)
    rho_l = np.array([rho_n[a:b + 1].sum() / (b - a + 1) for a, b in grps])     # sum then divide (:159-160)
    d_l = np.array([cpg_pos[grps[l][0]] - cpg_pos[grps[l - 1][1]] for l in range(1, len(grps))], dtype=np.float64)
    out.update(rho_n=rho_n, d_n=d_n, Nl=Nl, rho_l=rho_l, d_l=d_l, grp_first=np.array([a for a, _ in grps]),
               grp_last=np.array([b for _, b in grps]))
    return out


def chr_part_targeted(seq, reg_st, reg_end, size_est_reg=3000, min_grp_dist=10):
    """genome_analysis.jl:391-458 for one BED interval."""
    chr_size = len(seq)
    part = []
    i = reg_st
    while i < reg_end:
        bndy = i + size_est_reg - 1
        if bndy >= reg_end:
            part.append((i, reg_end))
            break
        lo = bndy - 1000
        hi = min(chr_size, bndy + 1000)
        dna = seq[lo - 1:hi]
        cpg = [m.start() + 1 + bndy - 1001 for m in CG.finditer(dna)]
        if cpg:
            _, ints = grps_from_cgs(cpg, min_grp_dist)
            for a, b in ints:
                if a <= bndy <= b:
                    bndy = b + 2 if 2 * bndy - a > b else a - 1
                    break
        part.append((i, bndy))
        i = bndy + 1
    return part


def chr_part_full(seq, size_est_reg=3000, min_grp_dist=10):
    """genome_analysis.jl:327-390 (whole chromosome)."""
    chr_size = len(seq)
    part = []
    i = 1
    while i < chr_size:
        bndy = i + size_est_reg - 1
        if bndy >= chr_size:
            part.append((i, chr_size))
            break
        lo = bndy - 1000
        hi = min(chr_size, bndy + 1000)
        dna = seq[lo - 1:hi]
        cpg = [m.start() + 1 + bndy - 1001 for m in CG.finditer(dna)]
        if cpg:
            _, ints = grps_from_cgs(cpg, min_grp_dist)
            for a, b in ints:
                if a <= bndy <= b:
                    bndy = b + 2 if 2 * bndy - a > b else a - 1
                    break
        part.append((i, bndy))
        i = bndy + 1
    return part


def read_theta(path):
    rows = []
    with open(path) as f:
        for line in f:
            v = line.rstrip("\n").split("\t")
            row = {"chrst": int(v[1]), "chrend": int(v[2]), "phi": np.array([float(x) for x in v[3].split(",")])}
            if len(v) > 4:
                row["alpha"] = np.array([float(x) for x in v[4].split(",")])
                row["beta"] = np.array([float(x) for x in v[5].split(",")])
            rows.append(row)
    return rows


def read_tab(path, ncol):
    rows = []
    with open(path) as f:
        for line in f:
            v = line.rstrip("\n").split("\t")
            rows.append([float(x) if x != "NaN" else np.nan for x in v[1:1 + ncol]])
    return np.array(rows, dtype=np.float64)


def pack_regions(seq, theta_rows):
    """CSR-pack per-region inputs + golden α/β of a theta file."""
    d = {k: [] for k in ("chrst", "chrend", "phi", "cpg_pos", "rho_n", "d_n", "alpha", "beta", "Nl", "rho_l", "d_l")}
    cpg_off, grp_off = [0], [0]
    for row in theta_rows:
        gi = grp_info(seq, row["chrst"], row["chrend"])
        d["chrst"].append(row["chrst"]); d["chrend"].append(row["chrend"]); d["phi"].append(row["phi"])
        d["cpg_pos"].append(gi["cpg_pos"]); d["rho_n"].append(gi["rho_n"]); d["d_n"].append(gi["d_n"])
        d["Nl"].append(gi["Nl"]); d["rho_l"].append(gi["rho_l"]); d["d_l"].append(gi["d_l"])
        if "alpha" in row:
            assert len(row["alpha"]) == gi["N"], (len(row["alpha"]), gi["N"])
            d["alpha"].append(row["alpha"]); d["beta"].append(row["beta"])
        cpg_off.append(cpg_off[-1] + gi["N"]); grp_off.append(grp_off[-1] + len(gi["Nl"]))
    out = {
        "chrst": np.array(d["chrst"], dtype=np.int64), "chrend": np.array(d["chrend"], dtype=np.int64),
        "phi": np.array(d["phi"]), "cpg_off": np.array(cpg_off, dtype=np.int64), "grp_off": np.array(grp_off, dtype=np.int64),
        "cpg_pos": np.concatenate(d["cpg_pos"]), "rho_n": np.concatenate(d["rho_n"]), "d_n": np.concatenate(d["d_n"]),
        "Nl": np.concatenate(d["Nl"]), "rho_l": np.concatenate(d["rho_l"]), "d_l": np.concatenate(d["d_l"]),
    }
    if d["alpha"]:
        out["alpha"] = np.concatenate(d["alpha"]); out["beta"] = np.concatenate(d["beta"])
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures are already committed")

    # ---- targeted example (config: max_size_subreg 350 → 351 via the 5-arg ctor, structures.jl:66) ----
    seq_t = read_fasta(f"{REF}/examples/targeted_example/reference/hg38_chr17_42879379_43290399.fa")
    base = f"{REF}/examples/targeted_example/cpelnano/targeted_example"
    fx = pack_regions(seq_t, read_theta(base + "_theta.txt"))
    fx["ex"] = read_tab(base + "_ex.txt", 3); fx["exx"] = read_tab(base + "_exx.txt", 3)
    fx["mml"] = read_tab(base + "_mml.txt", 3); fx["nme"] = read_tab(base + "_nme.txt", 3)
    fx["max_size_subreg"] = np.int64(351)
    np.savez_compressed(f"{OUT}/targeted_example.npz", **fx)
    print("targeted:", len(fx["chrst"]), "regions", len(fx["ex"]), "ex rows", len(fx["mml"]), "mml rows")

    # ---- synthetic-workload layouts: every ≥10-CpG 3 kb region of the whole targeted FASTA window ----
    part_t = chr_part_targeted(seq_t, 155510, 255510)
    theta_set = {(int(a), int(b)) for a, b in zip(fx["chrst"], fx["chrend"])}
    assert theta_set <= set(part_t), "theta-file regions must be a subset of the restated partition"
    part = chr_part_full(seq_t)
    lay = {k: [] for k in ("chrst", "chrend", "cpg_pos", "rho_n", "d_n", "Nl", "rho_l", "d_l")}
    cpg_off, grp_off = [0], [0]
    for a, b in part:
        gi = grp_info(seq_t, a, b)
        if gi["N"] <= 9 or len(gi["Nl"]) <= 1:
            continue
        lay["chrst"].append(a); lay["chrend"].append(b)
        for k in ("cpg_pos", "rho_n", "d_n", "Nl", "rho_l", "d_l"):
            lay[k].append(gi[k])
        cpg_off.append(cpg_off[-1] + gi["N"]); grp_off.append(grp_off[-1] + len(gi["Nl"]))
    np.savez_compressed(f"{OUT}/../../cpelnano.jl_b200/data/layouts_targeted.npz", chrst=np.array(lay["chrst"], dtype=np.int64), chrend=np.array(lay["chrend"], dtype=np.int64),
                        cpg_off=np.array(cpg_off, dtype=np.int64), grp_off=np.array(grp_off, dtype=np.int64),
                        **{k: np.concatenate(lay[k]) for k in ("cpg_pos", "rho_n", "d_n", "Nl", "rho_l", "d_l")})
    Ls = np.diff(grp_off); Ns = np.diff(cpg_off)
    print("layouts:", len(lay["chrst"]), "regions; L mean/max", Ls.mean(), Ls.max(), "N mean/max", Ns.mean(), Ns.max())

    # ---- full example (default config, max_size_subreg 350... written by examples/full_example.jl-like run) ----
    seq_f = read_fasta(f"{REF}/examples/full_example/reference/hg38_chr17_43023997_43145780.fa")
    base = f"{REF}/examples/full_example/cpelnano/full_example_mod_nse_true"
    fx = pack_regions(seq_f, read_theta(base + "_theta.txt"))
    fx["ex"] = read_tab(base + "_ex.txt", 3); fx["exx"] = read_tab(base + "_exx.txt", 3)
    fx["mml"] = read_tab(base + "_mml.txt", 3); fx["nme"] = read_tab(base + "_nme.txt", 3)
    fx["max_size_subreg"] = np.int64(350)
    np.savez_compressed(f"{OUT}/full_example.npz", **fx)
    print("full:", len(fx["chrst"]), "regions", len(fx["ex"]), "ex rows", len(fx["mml"]), "mml rows")

    # ---- group comparison example: 6 vs 6 model files + golden test tables (recipe:
    #      calls/GeneralTesting/permutation_tests_tests.jl:118-156, CpelNanoConfig() defaults) ----
    g = {}
    for grp in (1, 2):
        phis = []
        for rep in range(1, 7):
            rows = read_theta(f"{REF}/examples/grp_comp_example/mod_files/g{grp}_rep{rep}.txt")
            phis.append(np.array([r["phi"] for r in rows]))
            coords = [(r["chrst"], r["chrend"]) for r in rows]
            if "coords" in g:
                assert coords == g["coords"]
            g["coords"] = coords
        g[f"phi_g{grp}"] = np.array(phis)                     # [6][R][3]
    rows = [{"chrst": a, "chrend": b, "phi": np.zeros(3)} for a, b in g["coords"]]
    fx = pack_regions(seq_f, rows)
    del fx["phi"]
    fx["phi_g1"] = g["phi_g1"]; fx["phi_g2"] = g["phi_g2"]
    for kind in ("unmat", "mat"):
        for st in ("tmml", "tnme", "tcmd"):
            fx[f"{kind}_{st}"] = read_tab(f"{REF}/examples/grp_comp_example/tests/cpelnano_grp_comp_{kind}_{st}.txt", 5)
    fx["max_size_subreg"] = np.int64(350)
    np.savez_compressed(f"{OUT}/grp_comp_example.npz", **fx)
    print("grp_comp:", len(fx["chrst"]), "regions", fx["unmat_tmml"].shape, fx["mat_tcmd"].shape)


if __name__ == "__main__":
    main()
