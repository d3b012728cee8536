#!/usr/bin/env python3
"""File-level group comparison (diff_grp_comp, src/grp_perm_tests.jl:637) at scale: 6 vs 6 simulated model files on a chromosome
made of the bundled hg38 window repeated; stage wall times of diff_grp_comp_csr.
   python tools/grp_stages.py [--tiles 120] > gpurun_out/grp_stages.json
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
    ap.add_argument("--tiles", type=int, default=120)
    args = ap.parse_args()
    import cpelnano_b200 as cpn
    from cpelnano_b200 import genome, simulations, outputs
    eng = cpn.Engine(0)
    base = genome.read_fasta(os.path.join(ROOT, "tests", "golden", "inputs", "targeted_example.fa.gz"))["hg38_dna"]
    seq = np.tile(base, args.tiles)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "chr.fa")
    with open(fa, "w") as f:
        f.write(">chrSim\n")
        s = seq.tobytes().decode()
        f.write("\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    chrom = genome.Chromosome("chrSim", seq, eng.lib)
    lay, _, _ = simulations.layouts_from_fasta(eng.lib, chrom)
    R = len(lay["chrst"])
    rng = np.random.default_rng(7)
    phi1 = np.stack([rng.uniform(-1, 1, R), rng.uniform(-20, 20, R), rng.uniform(0, 5, R)], axis=1)
    phi2 = phi1.copy()
    flip = rng.random(R) < 0.1
    phi2[flip, :2] *= -1.0                                        # 10 % of the regions differ between the groups
    g1, g2 = simulations.gen_grp_phis(rng, phi1, phi2, 6, 6)
    files = [[], []]
    for g, phis in enumerate((g1, g2)):
        for s_i in range(6):
            p = os.path.join(tmp, f"g{g + 1}_rep{s_i + 1}.txt")
            with open(p, "w") as f:
                for r in range(R):
                    f.write(f"chrSim\t{lay['chrst'][r]}\t{lay['chrend'][r]}\t{outputs.julia_str(phis[s_i, r, 0])},{outputs.julia_str(phis[s_i, r, 1])},"
                            f"{outputs.julia_str(phis[s_i, r, 2])}\n")
            files[g].append(p)
    out = {"what": f"diff_grp_comp_csr, 6 vs 6 model files, {R} regions of a {len(seq)} bp chromosome, all 924 labelings (unmatched) / 64 sign patterns "
                   "(matched); stage wall times in seconds"}
    for matched in (False, True):
        for rep in range(2):
            tm = {}
            t0 = time.perf_counter()
            cpn.diff_grp_comp_csr(files[0], files[1], fa, cpn.CpelNanoConfig(matched=matched), os.path.join(tmp, f"o{matched}{rep}"), "x", engine=eng,
                                  timings=tm)
            tm["total_s"] = time.perf_counter() - t0
            tm["regions_per_s_whole_job"] = tm["regions"] / tm["total_s"]
            out[("matched" if matched else "unmatched") + f"_run{rep}"] = tm
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
