#!/usr/bin/env python3
"""Copy the reference's bundled example INPUT DATA (FASTA windows, BED, nanopolish calls, model files, golden output
tables) into tests/golden/inputs/ as compressed fixtures, so the host-side tests of the "next" rows (SURVEY.md §8f:
calls ingestion, genome features, model-file round trip, output writers, BH) run where /root/reference does not exist.

TEST INFRASTRUCTURE.  Run once in the build container:  python tests/golden/make_input_fixtures.py
Only data files are read (no reference source code).  The nanopolish TSV loses its last three columns
(num_calling_strands, num_motifs, sequence — never read by the reference: src/nanopore.jl:172-230 uses columns 1-8).
"""
import gzip
import os
import shutil
import sys
import tarfile

REF = "/root/reference/examples"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs")


def gz_copy(src, dst, transform=None):
    with open(src, "rt") as f, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as g:
        for line in f:
            g.write((transform(line) if transform else line).encode())


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures are already committed")
    os.makedirs(OUT, exist_ok=True)
    gz_copy(f"{REF}/full_example/reference/hg38_chr17_43023997_43145780.fa", f"{OUT}/full_example.fa.gz")
    gz_copy(f"{REF}/targeted_example/reference/hg38_chr17_42879379_43290399.fa", f"{OUT}/targeted_example.fa.gz")
    shutil.copy(f"{REF}/targeted_example/regions_of_interest/brca1.bed", f"{OUT}/brca1.bed")
    gz_copy(f"{REF}/full_example/nanopolish/full_example_noise2.0_methylation_calls.sorted.tsv", f"{OUT}/full_example_calls.tsv.gz",
            lambda line: "\t".join(line.rstrip("\n").split("\t")[:8]) + "\n")
    # golden OUTPUT files (text, as the reference wrote them) for the writer / reader / BH tests
    with tarfile.open(f"{OUT}/golden_outputs.tar.gz", "w:gz", compresslevel=9) as tar:
        for rel in ["targeted_example/cpelnano/targeted_example_theta.txt", "targeted_example/cpelnano/targeted_example_ex.txt",
                    "targeted_example/cpelnano/targeted_example_exx.txt", "targeted_example/cpelnano/targeted_example_mml.txt",
                    "targeted_example/cpelnano/targeted_example_nme.txt",
                    "full_example/cpelnano/full_example_mod_nse_true_theta.txt"] + \
                   [f"grp_comp_example/mod_files/g{g}_rep{r}.txt" for g in (1, 2) for r in range(1, 7)] + \
                   [f"grp_comp_example/tests/cpelnano_grp_comp_{k}_{s}.txt" for k in ("unmat", "mat") for s in ("tmml", "tnme", "tcmd")] + \
                   [f"two_samp_example/tests/cpelnano_two_samp_{s}.txt" for s in ("tmml", "tnme", "tcmd")]:
            tar.add(f"{REF}/{rel}", arcname=rel)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
