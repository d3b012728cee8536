#!/usr/bin/env python3
"""Copies the reference's WGBS example (data files, not code) into tests/golden/wgbs/: two single-end Bismark BAM files of 1 600
reads each on a 1 kb chromosome, the FASTA, and the expected theta file of the example (empty: the analysed window holds 7 CpG
sites and pmap_analyze_reg_bam skips regions with ≤ 9, src/wgbs.jl:353).  Run once where /root/reference exists."""
import os
import shutil

SRC = "/root/reference/examples/wgbs"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wgbs")
os.makedirs(DST, exist_ok=True)
for rel in ("bam/meth_sample.bam", "bam/unmeth_sample.bam", "reference/ref.fa", "cpelnano/test_theta.cpelont"):
    shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, os.path.basename(rel)))
    print(rel, os.path.getsize(os.path.join(DST, os.path.basename(rel))))
