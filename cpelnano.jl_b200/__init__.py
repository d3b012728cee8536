"""cpelnano.jl_b200 — B200-native (sm_100a) per-region Ising estimation hot path of CpelNano.jl.

The product is `libcpelnano_b200.so` (hand-written CUDA, C ABI in include/cpelnano_b200.h).  This package
is the host-side mirror of the reference's Julia interface for that path (same names, ϕ→phi, `!` dropped):

    CpelNanoConfig, RegStruct                      src/structures.jl:36-69, :150-197
    get_phihat(rs, config)                         src/em_algorithm.jl:371     (get_ϕhat!)
    rscle_grp_mod(rs)                              src/nanopore.jl:269
    get_nls_reg_info(rs, config)                   src/genome_analysis.jl:472
    get_stat_sums(rs, config)                      src/statistical_summaries.jl:470
    comp_cmd(rs1, rs2)                             src/statistical_summaries.jl:512
    unmat_est_reg_test / mat_est_reg_test          src/grp_perm_tests.jl:200 / :432
    pmap_two_samp_null_stats                       src/two_samp_perm_tests.jl:41
    analyze_regions (batched pmap_analyze_reg)     src/nanopore.jl:292,406
    analyze_regions_wgbs (pmap_analyze_reg_bam)    src/wgbs.jl:335, get_ϕhat_wgbs! src/em_algorithm.jl:556

There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA library and a
B200; a missing library raises at first use.

The directory name contains a dot, so it is loaded through `cpelnano_b200.py` at the repository root
(`import cpelnano_b200`) or `importlib`; inside the package only relative imports are used.
"""
from .capi import Library, CpnError, lib_path, MSTEP_FD, MSTEP_ANALYTIC   # noqa: F401
from .host import (CpelNanoConfig, RegStruct, Engine, get_phihat, rscle_grp_mod, get_nls_reg_info, get_stat_sums,   # noqa: F401
                   comp_cmd, unmat_est_reg_test, mat_est_reg_test, pmap_two_samp_null_stats, analyze_regions, analyze_regions_wgbs, default_engine)
from . import simulations   # noqa: F401
from . import sharding      # noqa: F401
from . import genome        # noqa: F401
from . import nanopolish    # noqa: F401
from . import outputs       # noqa: F401
from . import pipeline      # noqa: F401
from .pipeline import analyze_nano, analyze_bam, analyze_nano_csr, merge_shards, diff_grp_comp, diff_grp_comp_csr, diff_two_samp_comp   # noqa: F401

__all__ = ["Library", "CpnError", "Engine", "CpelNanoConfig", "RegStruct", "get_phihat", "rscle_grp_mod", "get_nls_reg_info",
           "get_stat_sums", "comp_cmd", "unmat_est_reg_test", "mat_est_reg_test", "pmap_two_samp_null_stats", "analyze_regions", "analyze_regions_wgbs",
           "simulations", "sharding", "genome", "nanopolish", "outputs", "pipeline", "analyze_nano", "analyze_bam", "analyze_nano_csr", "merge_shards", "diff_grp_comp_csr", "diff_grp_comp", "diff_two_samp_comp", "default_engine", "MSTEP_FD", "MSTEP_ANALYTIC"]
