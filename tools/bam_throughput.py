#!/usr/bin/env python3
"""Throughput of the native BAM reader (csrc/cpn_bam.cpp) on a synthetic single-end WGBS file (host-only, no GPU):
cpn_bam_open (BGZF inflate in parallel + record index) and get_calls_bam for all 3 kb regions of the chromosome in one call pair.
Usage: python tools/bam_throughput.py [--reads 2000000] [--threads 0] > profiles/r02_bam_reader_throughput.json"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bam_writer as bw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    import cpelnano_b200 as cpn
    lib = cpn.Library()
    rng = np.random.default_rng(1)
    read_len, depth = 100, 30
    G = args.reads * read_len // depth
    cpos = np.sort(rng.choice(np.arange(50, G - 400, 9), G // 60, replace=False)) + 1         # 1-based C positions, ≈ 50 per 3 kb
    starts = np.sort(rng.integers(1, G - read_len, args.reads))
    d = tempfile.mkdtemp()
    path = os.path.join(d, "big.bam")
    t0 = time.perf_counter()
    recs = []
    lo = np.searchsorted(cpos, starts); hi = np.searchsorted(cpos, starts + read_len - 1, side="right")
    meth = rng.random(len(cpos)) < 0.7
    for k in range(args.reads):
        s = bytearray(b"." * read_len)
        p = int(starts[k])
        for c in range(lo[k], hi[k]):
            s[int(cpos[c]) - p] = 90 if meth[c] else 122
        recs.append(bw.record(0, p, 0, 42, f"r{k}", [(read_len, "M")], read_len, [("NM", "C", 0), ("XM", "Z", s.decode())]))
    bw.write_bam(path, [("chrS", G)], recs, block=65000)
    t_write = time.perf_counter() - t0
    size = os.path.getsize(path)
    t0 = time.perf_counter()
    h = cpn.capi.BamHandle(lib, path, n_threads=args.threads)
    t_open = time.perf_counter() - t0
    st = np.arange(1, G - 3200, 3000, dtype=np.int64); en = st + 2999
    a = np.searchsorted(cpos, st); b = np.searchsorted(cpos, en, side="right")
    cpg_off = np.concatenate([[0], np.cumsum(b - a)]).astype(np.int64)
    flat = np.concatenate([cpos[i:j] for i, j in zip(a, b)]).astype(np.int64)
    t0 = time.perf_counter()
    read_off, calls = h.calls("chrS", st, en, cpg_off, flat, False, (0, 0, 0, 0), n_threads=args.threads)
    t_calls = time.perf_counter() - t0
    inflated = sum(len(r) for r in recs)
    print(json.dumps({"what": "native BAM reader on a synthetic single-end WGBS file (100-base reads, 30x, one chromosome)", "reads": args.reads,
                      "bam_bytes": size, "inflated_bytes": inflated, "threads": args.threads or os.cpu_count(), "open_s": round(t_open, 3),
                      "open_MBps_compressed": round(size / t_open / 1e6, 1), "open_MBps_inflated": round(inflated / t_open / 1e6, 1),
                      "open_reads_per_s": round(args.reads / t_open), "regions": int(len(st)), "rows": int(read_off[-1]), "cells": int(calls.size),
                      "calls_s": round(t_calls, 3), "calls_regions_per_s": round(len(st) / t_calls), "calls_rows_per_s": round(int(read_off[-1]) / t_calls),
                      "python_fixture_write_s": round(t_write, 1)}))


if __name__ == "__main__":
    main()
