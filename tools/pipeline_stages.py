#!/usr/bin/env python3
"""Where the time goes in the file-in → file-out entry point analyze_nano (src/nanopore.jl:362) once the estimation runs on the
GPU.  Builds a chromosome by repeating the bundled hg38 window, simulates 30× nanopore calls for every estimation region
(cpel_samp_ont restatement, two-decimal nanopolish text), writes FASTA + TSV, then times analyze_nano stage by stage.
   python tools/pipeline_stages.py [--tiles 24] > gpurun_out/pipeline_stages.json
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=24)
    ap.add_argument("--reads", type=int, default=30)
    args = ap.parse_args()
    import cpelnano_b200 as cpn
    from cpelnano_b200 import genome, simulations
    eng = cpn.Engine(0)
    base = genome.read_fasta(os.path.join(ROOT, "tests", "golden", "inputs", "targeted_example.fa.gz"))["hg38_dna"]
    seq = np.tile(base, args.tiles)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "chr.fa")
    with open(fa, "w") as f:
        f.write(">chrSim\n")
        s = seq.tobytes().decode()
        for i in range(0, len(s), 60):
            f.write(s[i:i + 60] + "\n")
    chrom = genome.Chromosome("chrSim", seq, eng.lib)
    lay, gs, ge = simulations.layouts_from_fasta(eng.lib, chrom)
    R = len(lay["chrst"])
    t0 = time.perf_counter()
    nb = simulations.make_nano_batch(R, M=args.reads, seed=3, layouts=lay)
    tsv = os.path.join(tmp, "calls.tsv")
    simulations.write_nanopolish_tsv(tsv, nb, "chrSim", gs, ge, fmt="{:.2f}".format)
    t_gen = time.perf_counter() - t0
    cfg = cpn.CpelNanoConfig()
    out = {"what": f"analyze_nano on a simulated chromosome: bundled hg38 window x{args.tiles} = {len(seq)} bp, {R} fittable regions, {args.reads} reads each, "
                   "10 EM starts, <= 20 EM iterations; stage wall times in seconds", "fasta_bytes": os.path.getsize(fa), "calls_bytes": os.path.getsize(tsv),
           "workload_generation_s": t_gen}
    for rep in range(2):                       # second repetition = warm library (arenas allocated, file cache hot)
        tm = {}
        t0 = time.perf_counter()
        done = cpn.analyze_nano(tsv, fa, cfg, os.path.join(tmp, f"out{rep}"), "sim", rng=np.random.default_rng(1), engine=eng, timings=tm)
        tm["total_s"] = time.perf_counter() - t0
        tm["regions_per_s_whole_job"] = tm["regions_fitted"] / tm["total_s"]
        out["per_region_structs_run%d" % rep] = tm
    for rep in range(2):
        tm = {}
        t0 = time.perf_counter()
        st = cpn.analyze_nano_csr(tsv, fa, cfg, os.path.join(tmp, f"csr{rep}"), "sim", rng=np.random.default_rng(1), engine=eng, timings=tm)
        tm["total_s"] = time.perf_counter() - t0
        tm["regions_per_s_whole_job"] = st["fitted"] / tm["total_s"]
        out["csr_path_run%d" % rep] = tm
    phi = np.array([rs.phihat for rs in done]); true = nb.phi_true[: len(phi)] if len(phi) == R else None
    if true is not None:
        out["median_abs_error_vs_simulated_phi"] = np.median(np.abs(phi - true), axis=0).tolist()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
