"""Genome-side inputs of the estimation path: FASTA → estimation regions → CpG sites, CpG groups and densities.

Host-side mirror of ("next" row):

    get_genome_info / get_chr_fa_rec      src/genome_analysis.jl:19-61, :283-312   → read_fasta
    get_cpg_grps                          src/genome_analysis.jl:63-111            → cpg_groups
    get_ρn                                src/genome_analysis.jl:122-140           → inside grp_info (prefix counts, not regex)
    get_grp_info!                         src/genome_analysis.jl:226-277           → grp_info / region_features
    get_chr_part (whole chr / BED)        src/genome_analysis.jl:327-390, :391-458 → chr_part

Integer work, bit-exact by construction: CpG positions are found once per chromosome on the byte array and every
window count of the reference (one regex scan of a 1 kb window per CpG) becomes two binary searches into that array.
Outputs are pinned by the reference's golden theta files (α, β reproduce exactly; tests/test_host_inputs.py).
Coordinates are 1-based closed, as in the reference.
"""
import gzip

import numpy as np

from .host import RegStruct


def read_fasta(path):
    """{record name: upper/lower-case sequence string as bytes array (uint8)} in file order."""
    opener = gzip.open if str(path).endswith(".gz") else open
    recs, name, chunks = {}, None, []
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    recs[name] = np.frombuffer("".join(chunks).encode(), dtype=np.uint8)
                name, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.strip())
    if name is not None:
        recs[name] = np.frombuffer("".join(chunks).encode(), dtype=np.uint8)
    return recs


def cpg_starts(seq):
    """1-based positions of the C of every CG dinucleotide (the regex r"[Cc][Gg]" of the reference; matches cannot overlap)."""
    up = seq & 0xDF                                              # ASCII upper-case
    return (np.nonzero((up[:-1] == ord("C")) & (up[1:] == ord("G")))[0] + 1).astype(np.int64)


def cpg_groups(cpg_pos, min_grp_dist):
    """get_cpg_grps (genome_analysis.jl:63-111): consecutive CpGs closer than or at min_grp_dist share a group.
    Returns (first, last) 0-based inclusive index arrays; the group interval is [pos[first]-5, pos[last]+5]."""
    cpg_pos = np.asarray(cpg_pos, dtype=np.int64)
    if len(cpg_pos) == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    brk = np.nonzero(np.diff(cpg_pos) > min_grp_dist)[0]
    first = np.concatenate([[0], brk + 1]).astype(np.int64)
    last = np.concatenate([brk, [len(cpg_pos) - 1]]).astype(np.int64)
    return first, last


class Chromosome:
    """One FASTA record with its CpG index."""

    def __init__(self, name, seq, lib=None):
        self.name, self.seq, self.size = name, seq, len(seq)
        if lib is not None:
            from . import capi
            self.cg = capi.cpg_scan(lib, seq)       # native, multi-threaded (cpn_cpg_scan)
        else:
            self.cg = cpg_starts(seq)

    def cpgs_in(self, st, end):
        """CG matches lying completely inside seq[st..end] (what a regex scan of that substring finds)."""
        lo = np.searchsorted(self.cg, st, side="left")
        hi = np.searchsorted(self.cg, end - 1, side="right")
        return lo, max(lo, hi)

    # ---- get_chr_part ---------------------------------------------------------------------------------------------------
    def chr_part(self, size_est_reg=3000, min_grp_dist=10, targets=None):
        """Estimation regions of the chromosome (targets=None, :327-390) or of BED intervals [(st, end), …] (:391-458):
        windows of size_est_reg whose right boundary is moved off any CpG group it would split."""
        spans = [(1, self.size)] if targets is None else list(targets)
        part = []
        for (reg_st, reg_end) in spans:
            i = reg_st
            while i < reg_end:
                bndy = i + size_est_reg - 1
                if bndy >= reg_end:
                    part.append((i, reg_end))
                    break
                lo, hi = self.cpgs_in(bndy - 1000, min(self.size, bndy + 1000))
                if hi > lo:
                    pos = self.cg[lo:hi]
                    first, last = cpg_groups(pos, min_grp_dist)
                    pad = 0 if len(pos) == 1 else 5          # a lone CpG's interval is returned unpadded (early return, :67)
                    a, b = pos[first] - pad, pos[last] + pad
                    hit = np.nonzero((a <= bndy) & (bndy <= b))[0]
                    if len(hit):
                        a0, b0 = int(a[hit[0]]), int(b[hit[0]])
                        bndy = b0 + 2 if 2 * bndy - a0 > b0 else a0 - 1
                part.append((i, bndy))
                i = bndy + 1
        return part

    # ---- get_grp_info! ---------------------------------------------------------------------------------------------------
    def grp_info(self, chrst, chrend, min_grp_dist=10):
        """CpG sites of [chrst, chrend], their densities ρ_n (CpGs per bp in a ±500 bp window clipped to the region's ±500 bp
        context, :122-140, :249-255), distances, and the CpG groups with N_l, ρ_l (mean of ρ_n, sum then divide) and d_l."""
        lo, hi = self.cpgs_in(chrst, chrend)
        cpg_pos = self.cg[lo:hi].copy()
        out = dict(cpg_pos=cpg_pos, N=len(cpg_pos))
        if len(cpg_pos) <= 1:
            return out
        cont_st, cont_end = max(1, chrst - 500), min(self.size, chrend + 500)
        clen = cont_end - cont_st + 1
        o = cpg_pos - cont_st                                   # the offset the reference hands to get_ρn
        win_st = np.maximum(1, o - 500) + cont_st - 1           # window in chromosome coordinates
        win_end = np.minimum(clen, o + 500) + cont_st - 1
        cnt = np.searchsorted(self.cg, win_end - 1, side="right") - np.searchsorted(self.cg, win_st, side="left")
        rho_n = np.maximum(cnt, 0) / 1000
        first, last = cpg_groups(cpg_pos, min_grp_dist)
        Nl = (last - first + 1).astype(np.float64)
        # sum(ρ_n[first:last]) / N_l with the reference's left-to-right sum (cumulative-sum differences are not bit-identical);
        # groups are short, so add one column of group members at a time
        rho_l = rho_n[first].copy()
        for k in range(1, int(Nl.max())):
            more = Nl > k
            rho_l[more] += rho_n[first[more] + k]
        rho_l /= Nl
        d_l = (cpg_pos[first[1:]] - cpg_pos[last[:-1]]).astype(np.float64)
        out.update(rho_n=rho_n, d_n=np.diff(cpg_pos).astype(np.float64), Nl=Nl, rho_l=rho_l, d_l=d_l, grp_first=first, grp_last=last,
                   grp_st=cpg_pos[first] - 5, grp_end=cpg_pos[last] + 5)
        return out

    def region_struct(self, chrst, chrend, min_grp_dist=10):
        """RegStruct with the genomic fields get_grp_info! fills (group scale); calls are added by nanopolish.attach_calls."""
        gi = self.grp_info(chrst, chrend, min_grp_dist)
        rs = RegStruct(chr=self.name, chrst=int(chrst), chrend=int(chrend), N=int(gi["N"]), cpg_pos=gi["cpg_pos"])
        if gi["N"] > 1:
            rs.L = len(gi["Nl"])
            rs.Nl, rs.rho_l, rs.dl, rs.rho_n, rs.dn = gi["Nl"], gi["rho_l"], gi["d_l"], gi["rho_n"], gi["d_n"]
            rs.grp_st, rs.grp_end = gi["grp_st"], gi["grp_end"]
        return rs


    # ---- native batch path (csrc/cpn_genome.cpp through the C ABI) --------------------------------------------------------
    def chr_part_native(self, lib, size_est_reg=3000, min_grp_dist=10, targets=None):
        from . import capi
        spans = [(1, self.size)] if targets is None else list(targets)
        part = []
        for (a, b) in spans:
            st, en = capi.chr_part(lib, self.cg, self.size, a, b, size_est_reg, min_grp_dist)
            part.extend(zip(st.tolist(), en.tolist()))
        return part

    def region_structs(self, lib, part, min_grp_dist=10, n_threads=0):
        """region_struct for every (chrst, chrend) of `part` in two native calls (cpn_grp_info_count / _fill)."""
        from . import capi
        if not part:
            return []
        st = np.array([p[0] for p in part], dtype=np.int64); en = np.array([p[1] for p in part], dtype=np.int64)
        o = capi.grp_info_batch(lib, self.cg, self.size, st, en, min_grp_dist, n_threads)
        regs = []
        for r in range(len(part)):
            n0, n1, g0, g1 = o["cpg_off"][r], o["cpg_off"][r + 1], o["grp_off"][r], o["grp_off"][r + 1]
            rs = RegStruct(chr=self.name, chrst=int(st[r]), chrend=int(en[r]), N=int(n1 - n0), cpg_pos=o["cpg_pos"][n0:n1])
            if n1 - n0 > 1:
                rs.L = int(g1 - g0)
                rs.Nl, rs.rho_l, rs.rho_n = o["Nl"][g0:g1], o["rho_l"][g0:g1], o["rho_n"][n0:n1]
                rs.dn, rs.dl = o["d_n"][o["dn_off"][r]:o["dn_off"][r + 1]], o["d_l"][o["dl_off"][r]:o["dl_off"][r + 1]]
                rs.grp_st, rs.grp_end = o["grp_st"][g0:g1], o["grp_end"][g0:g1]
            regs.append(rs)
        return regs


def read_bed(path):
    """read_bed_reg (src/input_output.jl:494-525): {chr: [(st, end), …]} in file order; lines with end < st are dropped."""
    regs = {}
    with open(path) as f:
        for line in f:
            v = line.rstrip("\n").split("\t")
            if len(v) < 3:
                continue
            st, end = int(v[1]), int(v[2])
            if end >= st:
                regs.setdefault(v[0], []).append((st, end))
    return regs


def load_genome(path, lib=None):
    """{record name: Chromosome}.  With `lib` the file is read by the native FASTA loader (cpn_fasta_open) and scanned by
    cpn_cpg_scan; the Chromosome objects keep the loader's memory alive."""
    if lib is None:
        return {name: Chromosome(name, seq) for name, seq in read_fasta(path).items()}
    from . import capi
    fa = capi.FastaHandle(lib, path)
    out = {}
    for name, seq in fa.records.items():
        ch = Chromosome(name, seq, lib)
        ch._fasta = fa
        out[name] = ch
    return out
