"""Calls ingestion: nanopolish `call-methylation` TSV → per-region emission matrices ("next" row).

Host-side mirror of

    get_dic_nanopolish        src/input_output.jl:536-570   lines of the region's chromosome with start ≤ chrend and end ≥ chrst
    get_calls_nanopolish!     src/nanopore.jl:172-230       line → CpG group whose interval is [start+1-5, end+1+5] exactly
    get_calls!                src/nanopore.jl:241-257

The reference re-reads the whole TSV for every estimation region (its early `break` sits behind a `continue` and never
fires) and looks groups up with a linear `findfirst` per call: O(regions × file).  Here the file is parsed once per
chromosome into columns; a region is two binary searches on the (sorted) start column plus a vectorised filter, and a
call finds its group with a binary search on the group starts.

Semantics kept: every read with at least one overlapping line becomes a row — even if none of its lines matches a
group (all-unobserved row, it still counts in rs.m); a later line for the same (read, group) overwrites an earlier one;
mod_nse=false replaces the likelihoods by the hard pair (0, −50) (src/CpelNano.jl:32-33).  One deliberate difference:
the reference orders the reads of a region by Julia's Dict iteration order (hash order, not reproducible outside that
Julia build); rows here follow the first appearance of a read in the file.  The read order only enters through the
floating-point sum over reads in the E-step.
"""
import gzip

import numpy as np

LOG_PYX_RIGHT_X = 0.0       # src/CpelNano.jl:33
LOG_PYX_WRONG_X = -50.0     # src/CpelNano.jl:32


class NanopolishCalls:
    """Columns of a sorted nanopolish methylation-calls file, split by chromosome."""

    def __init__(self, path):
        import pandas as pd
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rt") as f:
            df = pd.read_csv(f, sep="\t", header=0, usecols=[0, 2, 3, 4, 6, 7], names=["chr", "start", "end", "read", "lm", "lu"],
                             dtype={"chr": str, "start": np.int64, "end": np.int64, "read": str, "lm": np.float64, "lu": np.float64})
        self.n_lines = len(df)
        self.by_chr = {}
        for name, g in df.groupby("chr", sort=False):
            start = g["start"].to_numpy()
            if np.any(np.diff(start) < 0):                      # the reference assumes a sorted file; sort stably if it is not
                g = g.iloc[np.argsort(start, kind="stable")]
                start = g["start"].to_numpy()
            codes, names = pd.factorize(g["read"], sort=False)
            end = g["end"].to_numpy()
            self.by_chr[name] = dict(start=start, end=end, read=codes.astype(np.int64), names=np.asarray(names),
                                     lm=g["lm"].to_numpy(), lu=g["lu"].to_numpy(), max_span=int((end - start).max()) if len(end) else 0)

    def region_calls(self, chrom, chrst, chrend, grp_st, grp_end, mod_nse=True):
        """(lu, lm, read_names): [m, L] matrices with NaN where a group is not observed in a read."""
        L = len(grp_st)
        c = self.by_chr.get(chrom)
        if c is None:
            return np.zeros((0, L)), np.zeros((0, L)), np.zeros(0, dtype=object)
        hi = np.searchsorted(c["start"], chrend, side="right")            # start ≤ chrend (file sorted by start)
        lo = np.searchsorted(c["start"], chrst - c["max_span"], side="left")   # end ≥ chrst needs start ≥ chrst − longest call
        sel = lo + np.nonzero(c["end"][lo:hi] >= chrst)[0]                 # end ≥ chrst
        if len(sel) == 0:
            return np.zeros((0, L)), np.zeros((0, L)), np.zeros(0, dtype=object)
        reads = c["read"][sel]
        uniq, first_at, inv = np.unique(reads, return_index=True, return_inverse=True)
        order = np.argsort(first_at, kind="stable")                        # rows in order of first appearance
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[order] = np.arange(len(uniq))
        row = rank[inv]
        lu = np.full((len(uniq), L), np.nan); lm = np.full((len(uniq), L), np.nan)
        st = c["start"][sel] + 1 - 5                                       # 0-based, no ±5 flanks in nanopolish (:181-184)
        en = c["end"][sel] + 1 + 5
        g = np.searchsorted(grp_st, st)
        ok = (g < L)
        g = np.minimum(g, L - 1)
        ok &= (grp_st[g] == st) & (grp_end[g] == en)                       # the same group must match both ends (:187-197)
        vm, vu = c["lm"][sel][ok], c["lu"][sel][ok]
        if not mod_nse:
            meth = vm > vu
            vm, vu = np.where(meth, LOG_PYX_RIGHT_X, LOG_PYX_WRONG_X), np.where(meth, LOG_PYX_WRONG_X, LOG_PYX_RIGHT_X)
        lm[row[ok], g[ok]] = vm                                            # file order: the last line for a (read, group) wins
        lu[row[ok], g[ok]] = vu
        return lu, lm, c["names"][uniq[order]]

    def attach_calls(self, rs, mod_nse=True):
        """get_calls!(rs, nano, config): fills rs.log_pyx_u/m and rs.m from the region's group intervals."""
        lu, lm, names = self.region_calls(rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        rs.set_calls(lu, lm)
        rs.read_names = names
        return rs


class NativeCalls:
    """Same contract as NanopolishCalls on top of the native loader (csrc/cpn_calls.cpp through the C ABI: cpn_calls_open /
    cpn_calls_count_reads / cpn_calls_fill): the file is parsed once by all host threads, and the calls of ALL regions of a
    chromosome are written in one call, directly in the CSR layout cpn_fit_batch takes."""

    def __init__(self, path, library=None, n_threads=0):
        from .capi import Library, CallsHandle
        self.handle = CallsHandle(library or Library(), path, n_threads)
        self.n_lines = self.handle.n_lines
        self.n_threads = n_threads

    def region_calls(self, chrom, chrst, chrend, grp_st, grp_end, mod_nse=True):
        L = len(grp_st)
        read_off, lu, lm, ids = self.handle.fill(chrom, [chrst], [chrend], [0, L], grp_st, grp_end, mod_nse, n_threads=1, want_ids=True)
        m = int(read_off[1])
        return lu.reshape(m, L), lm.reshape(m, L), np.array([self.handle.read_name(i) for i in ids], dtype=object)

    def attach_calls(self, rs, mod_nse=True):
        lu, lm, names = self.region_calls(rs.chr, rs.chrst, rs.chrend, rs.grp_st, rs.grp_end, mod_nse)
        rs.set_calls(lu, lm)
        rs.read_names = names
        return rs

    def attach_calls_batch(self, regs, mod_nse=True):
        """get_calls! for a list of RegStructs of ONE chromosome (group-scale features set) in two native calls."""
        if not regs:
            return regs
        chrom = regs[0].chr
        assert all(r.chr == chrom for r in regs)
        Ls = [len(r.grp_st) for r in regs]
        grp_off = np.concatenate([[0], np.cumsum(Ls)]).astype(np.int64)
        read_off, lu, lm = self.handle.fill(chrom, [r.chrst for r in regs], [r.chrend for r in regs], grp_off,
                                            np.concatenate([r.grp_st for r in regs]), np.concatenate([r.grp_end for r in regs]), mod_nse,
                                            n_threads=self.n_threads)
        o = 0
        for i, r in enumerate(regs):
            m, L = int(read_off[i + 1] - read_off[i]), Ls[i]
            r.set_calls(lu[o:o + m * L].reshape(m, L), lm[o:o + m * L].reshape(m, L))
            o += m * L
        return regs
