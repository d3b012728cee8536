#!/usr/bin/env python3
"""Throughput of the native calls loader on a synthetic nanopolish file (host only; no GPU).
Writes a TSV with the shape of real call-methylation output (11 columns, 36-character read names, reads of ~8.5 kb covering a
synthetic chromosome at the requested depth), then times: cpn_calls_open (parse), cpn_calls_count_reads + cpn_calls_fill for
every 3 kb estimation region, the numpy/pandas implementation, and — on a sample of regions — the reference's access pattern
(whole file re-read per region).   python tools/calls_throughput.py [--mbp 3] [--depth 30] > profiles/...json
"""
import argparse
import json
import os
import sys
import tempfile
import time
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import calls_ref  # noqa: E402  (test oracle)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=3.0)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--ref-sample", type=int, default=3)
    args = ap.parse_args()
    import cpelnano_b200 as cpn
    from cpelnano_b200 import nanopolish, genome
    rng = np.random.default_rng(1)
    size = int(args.mbp * 1e6)
    # CpG groups: one every ~70 bp (≈ 43 per 3 kb, the density of the bundled hg38 window), 1–3 CpGs each
    gaps = rng.integers(30, 110, size // 70)
    first = np.cumsum(gaps)
    first = first[first < size - 50]
    width = rng.integers(0, 12, len(first))
    grp_st, grp_end = first - 5, first + width + 5
    n_reads = int(args.depth * size / 8500)
    rst = rng.integers(0, size, n_reads)
    rows = []
    names = [str(uuid.UUID(int=int(x))) for x in rng.integers(0, 2 ** 62, n_reads)]
    for i in range(n_reads):
        lo, hi = np.searchsorted(first, rst[i]), np.searchsorted(first, rst[i] + 8500)
        g = np.arange(lo, hi)
        g = g[rng.random(len(g)) < 0.9]
        lm = rng.normal(-120, 30, len(g)); lu = lm - rng.normal(0, 5, len(g))
        rows.append((first[g] - 1 + 0, first[g] + width[g] - 1, np.full(len(g), i), lm, lu))
    st = np.concatenate([r[0] for r in rows]); en = np.concatenate([r[1] for r in rows]); rd = np.concatenate([r[2] for r in rows])
    lm = np.concatenate([r[3] for r in rows]); lu = np.concatenate([r[4] for r in rows])
    o = np.argsort(st, kind="stable")
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "calls.tsv")
    with open(path, "w") as f:
        f.write("chromosome\tstrand\tstart\tend\tread_name\tlog_lik_ratio\tlog_lik_methylated\tlog_lik_unmethylated\tnum_calling_strands\tnum_motifs\tsequence\n")
        for k in o:
            f.write(f"chrS\t+\t{st[k]}\t{en[k]}\t{names[rd[k]]}\t{lm[k] - lu[k]:.2f}\t{lm[k]:.2f}\t{lu[k]:.2f}\t1\t1\tACGTACGCGTAGCTAGCATCG\n")
    fsize = os.path.getsize(path)
    # estimation regions and their groups
    bounds = np.arange(0, size, 3000)
    chrst, chrend = bounds + 1, np.minimum(bounds + 3000, size)
    g_lo, g_hi = np.searchsorted(first, chrst), np.searchsorted(first, chrend, side="right")
    keep = g_hi - g_lo > 1
    chrst, chrend, g_lo, g_hi = chrst[keep], chrend[keep], g_lo[keep], g_hi[keep]
    grp_off = np.concatenate([[0], np.cumsum(g_hi - g_lo)])
    gs = np.concatenate([grp_st[a:b] for a, b in zip(g_lo, g_hi)]); ge = np.concatenate([grp_end[a:b] for a, b in zip(g_lo, g_hi)])
    lib = cpn.Library()
    out = {"file_bytes": fsize, "lines": int(len(st)), "regions": int(len(chrst)), "depth": args.depth, "cores": os.cpu_count()}
    for T in (1, os.cpu_count()):
        t = time.perf_counter(); h = cpn.capi.CallsHandle(lib, path, T); t_open = time.perf_counter() - t
        t = time.perf_counter(); read_off, lu_n, lm_n = h.fill("chrS", chrst, chrend, grp_off, gs, ge, n_threads=T); t_fill = time.perf_counter() - t
        out[f"native_{T}_threads"] = {"parse_s": t_open, "parse_MB_per_s": fsize / t_open / 1e6, "lines_per_s": len(st) / t_open,
                                      "count_and_fill_s": t_fill, "regions_per_s": len(chrst) / t_fill}
        assert h.n_lines == len(st)
    t = time.perf_counter(); npc = nanopolish.NanopolishCalls(path); t_np = time.perf_counter() - t
    t = time.perf_counter()
    for r in range(len(chrst)):
        a, b = grp_off[r], grp_off[r + 1]
        lu_r, lm_r, _ = npc.region_calls("chrS", int(chrst[r]), int(chrend[r]), gs[a:b], ge[a:b])
        if r % 97 == 0:
            m = lu_r.shape[0]
            o0 = int(np.sum(np.diff(read_off)[:r] * np.diff(grp_off)[:r]))
            assert np.array_equal(lu_r.reshape(-1), lu_n[o0:o0 + m * (b - a)], equal_nan=True)
    t_np_fill = time.perf_counter() - t
    out["numpy_pandas"] = {"parse_s": t_np, "per_region_s_total": t_np_fill, "regions_per_s": len(chrst) / t_np_fill}
    t = time.perf_counter()
    for r in range(args.ref_sample):
        a, b = grp_off[r], grp_off[r + 1]
        calls_ref.region_calls_reference(path, "chrS", int(chrst[r]), int(chrend[r]), gs[a:b], ge[a:b])
    t_ref = (time.perf_counter() - t) / args.ref_sample
    out["reference_access_pattern"] = {"s_per_region": t_ref, "regions_per_s": 1 / t_ref,
                                       "note": "whole file re-read for every region (src/input_output.jl:536-570), Python restatement; Julia would be faster per line, the O(regions x file) shape is the point"}
    out["observed_fraction"] = float(np.mean(~np.isnan(lu_n)))
    print(json.dumps(out, indent=1))
    os.remove(path)


if __name__ == "__main__":
    main()
