# CpelNanoB200.jl — ccall shim that routes CpelNano's estimation hot path to libcpelnano_b200.so.
#
# Delivered UNTESTED BY EXECUTION: this build image has no Julia.  The executable proof of the ABI is the Python host
# (cpelnano.jl_b200/host.py, pipeline.py) and tests/, which bind exactly the same symbols with the same argument order; every
# ccall below names the header declaration it binds (include/cpelnano_b200.h).
#
# Usage inside the package (after the last `include` of src/CpelNano.jl):
#     include("CpelNanoB200.jl")
#     CpelNanoB200.enable!("/path/to/libcpelnano_b200.so"; devices=[0,1,2,3,4,5,6,7], seed=20211111)
# After enable! the package's own entry points keep their signatures, config struct and output files:
#     analyze_nano(nano, fasta, config)                                     src/nanopore.jl:362
#     diff_grp_comp(mod_fls_g1, mod_fls_g2, fasta, config)                   src/grp_perm_tests.jl:637
#     diff_two_samp_comp(mods_s1, mods_s2, nano_s1, nano_s2, fasta, config)  src/two_samp_perm_tests.jl:408
#     analyze_bam(bam, fasta, config)                                       src/wgbs.jl:401   (WGBS; get_calls_bam :222 too)
# and so do the per-region functions they are built from:
#     get_ϕhat!  get_stat_sums!  comp_cmd  unmat_est_reg_test  mat_est_reg_test  pmap_two_samp_null_stats
# What changes is the batching: each `pmap` over regions (nanopore.jl:406, grp_perm_tests.jl:674) and the serial `map` over regions
# with its inner `pmap` over permutations (two_samp_perm_tests.jl:445, :220) becomes ONE library call per chromosome; with several
# devices the library shards that call over the GPUs itself (cpn_create_multi), so the single Julia master drives all of them.
module CpelNanoB200

using Combinatorics                                  # combinations(): the labelings of the exact unmatched test (grp_perm_tests.jl:233)
using ..CpelNano: RegStruct, CpelNanoConfig, MethCallCpgGrp, AnalysisRegions, RegStatTestStruct, OutputFiles, OutputDiffFiles,
                  get_grp_info!, get_genome_info, get_chr_part, get_chr_fa_rec, read_bed_reg, check_output_exists, check_diff_output_exists,
                  read_mod_file_chr, read_in_grp_mods, find_comm_regs, write_output, write_diff_output, mult_hyp_corr, rscle_grp_mod!,
                  gen_ϕ0s, print_log, get_nls_reg_info!, log_pyx_right_x, log_pyx_wrong_x
import ..CpelNano

const LIB = Ref{String}("libcpelnano_b200")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
const SEED = Ref{Union{Nothing,UInt64}}(nothing)    # nothing: starts / permutations come from Julia's RNG, as in the reference

# status words of include/cpelnano_b200.h
const CPN_FIT_NONE, CPN_FIT_NUMERICAL, CPN_FIT_FILTERED = Int32(-1), Int32(-2), Int32(-16)

"""
    enable!(lib; devices=[0], seed=nothing)

Loads the library and creates one context over `devices` (cpn_create / cpn_create_multi).  With `seed` the EM starts and the
two-sample permutations are drawn on the GPU from a counter-based stream (cpn_set_seed): nothing random crosses PCIe and runs
are reproducible; without it they are drawn by Julia exactly where the reference draws them.
"""
function enable!(lib::String; devices::Vector{Int}=[0], seed::Union{Nothing,Integer}=nothing)
    LIB[] = lib
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    ids = Cint.(devices)
    rc = length(ids) == 1 ? ccall((:cpn_create, LIB[]), Cint, (Cint, Ref{Ptr{Cvoid}}), ids[1], ctx) :
                            ccall((:cpn_create_multi, LIB[]), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), ids, length(ids), ctx)
    rc == 0 || error("cpn_create failed ($rc): no sm_100 GPU; there is no CPU fallback")
    CTX[] = ctx[]
    SEED[] = isnothing(seed) ? nothing : UInt64(seed)
    isnothing(seed) || check(ccall((:cpn_set_seed, LIB[]), Cint, (Ptr{Cvoid}, UInt64), CTX[], UInt64(seed)))
    return nothing
end

check(rc) = rc == 0 || error(unsafe_string(ccall((:cpn_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[])))

# ---- packing ---------------------------------------------------------------------------------------------------------------
# calls::Vector{Vector{MethCallCpgGrp}} (structures.jl:121-128,168) → read-major doubles, NaN = not observed
function pack_calls(callss::Vector{Vector{Vector{MethCallCpgGrp}}}, Ls::Vector{Int})
    n = sum(length(c) * L for (c, L) in zip(callss, Ls))
    lu = Vector{Float64}(undef, n); lm = Vector{Float64}(undef, n)
    k = 0
    for calls in callss, call in calls, c in call
        k += 1
        lu[k] = c.obs ? c.log_pyx_u : NaN
        lm[k] = c.obs ? c.log_pyx_m : NaN
    end
    return Int64[0; cumsum([length(c) for c in callss])], lu, lm
end
pack_calls(rss::Vector{RegStruct}) = pack_calls([rs.calls for rs in rss], [rs.L for rs in rss])
pack_groups(rss) = (Int64[0; cumsum([rs.L for rs in rss])], Float64.(vcat([rs.Nl for rs in rss]...)),
                    Float64.(vcat([rs.ρl for rs in rss]...)), Float64.(vcat([rs.dl for rs in rss]...)))
# per-CpG tables of regions whose structs are at CpG scale (after rscle_grp_mod!: ρl = ρn, dl = dn) or still at group scale
pack_cpgs(rss) = (Int64[0; cumsum([rs.N for rs in rss])], Float64.(vcat([rs.ρn for rs in rss]...)), Float64.(vcat([rs.dn for rs in rss]...)))

# get_nls_reg_info!(rs, config)  genome_analysis.jl:472  ↔ cpn_nls_reg_info
function nls_reg_info!(rs::RegStruct, config::CpelNanoConfig)
    sig = (Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64})
    K = ccall((:cpn_nls_reg_info, LIB[]), Int64, sig, rs.chrst, rs.chrend, config.max_size_subreg, rs.cpg_pos, length(rs.cpg_pos),
              C_NULL, C_NULL, C_NULL, C_NULL)
    st = Vector{Int64}(undef, K); en = similar(st); p = similar(st); q = similar(st)
    ccall((:cpn_nls_reg_info, LIB[]), Int64, sig, rs.chrst, rs.chrend, config.max_size_subreg, rs.cpg_pos, length(rs.cpg_pos), st, en, p, q)
    rs.nls_rgs = AnalysisRegions([st[k]:en[k] for k = 1:K], [p[k]:q[k] for k = 1:K])
    return p, q
end
function pack_subregs(rss, config)
    pq = [nls_reg_info!(rs, config) for rs in rss]
    return Int64[0; cumsum([length(x[1]) for x in pq])], vcat([x[1] for x in pq]...), vcat([x[2] for x in pq]...)
end

# EM starts: Julia's RNG (the reference's gen_ϕ0s draw, em_algorithm.jl:374) unless a device seed was given (phi0 = NULL)
starts(n_units, config) = isnothing(SEED[]) ? Float64.(vcat([vcat(gen_ϕ0s(config.max_em_init)...) for _ in 1:n_units]...)) : nothing
ptr_or_null(x) = isnothing(x) ? Ptr{Float64}(C_NULL) : pointer(x)

# ---- get_ϕhat!(rs, config)  em_algorithm.jl:371 ↔ cpn_fit_batch ---------------------------------------------------------------
function get_ϕhat_batch!(rss::Vector{RegStruct}, config::CpelNanoConfig)
    R = length(rss)
    R == 0 && return Int32[]
    grp_off, Nl, ρl, dl = pack_groups(rss)
    read_off, lu, lm = pack_calls(rss)
    ϕ0 = starts(R, config)
    ϕhat = Vector{Float64}(undef, 3R); Q = Vector{Float64}(undef, R); status = Vector{Int32}(undef, R)
    GC.@preserve grp_off Nl ρl dl read_off lu lm ϕ0 ϕhat Q status begin
        check(ccall((:cpn_fit_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Int32, Int32, Float64, Float64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], R, grp_off, Nl, ρl, dl, read_off, lu, lm, ptr_or_null(ϕ0), config.max_em_init, config.max_em_iters,
            CpelNano.AMAX, CpelNano.BMAX, CpelNano.CMAX, 1, ϕhat, Q, status, C_NULL, C_NULL, C_NULL))
    end
    for (r, rs) in enumerate(rss)                         # status word: ≥ 0 picked start; < 0 → rs.ϕhat = [] (em_algorithm.jl:385)
        rs.ϕhat = status[r] >= 0 ? ϕhat[(3r - 2):(3r)] : Vector{Float64}()
        status[r] == CPN_FIT_NUMERICAL && print_log("Numerical failure in the M-step of $(rs.chr):$(rs.chrst)-$(rs.chrend)")   # :168
    end
    return status
end
CpelNano.get_ϕhat!(rs::RegStruct, config::CpelNanoConfig) = (get_ϕhat_batch!([rs], config); nothing)

# ---- get_stat_sums!(rs, config)  statistical_summaries.jl:470 ↔ cpn_stat_sums_batch -------------------------------------------
function get_stat_sums_batch!(rss::Vector{RegStruct}, config::CpelNanoConfig)
    T = length(rss)
    T == 0 && return nothing
    sub_off, sub_p, sub_q = pack_subregs(rss, config)
    cpg_off = Int64[0; cumsum([rs.N for rs in rss])]
    ρn = Float64.(vcat([rs.ρl for rs in rss]...)); dn = Float64.(vcat([rs.dl for rs in rss]...))       # after rscle_grp_mod!
    ϕ = Float64.(vcat([rs.ϕhat for rs in rss]...))
    sN, sK = cpg_off[end], sub_off[end]
    logZ = Vector{Float64}(undef, T); ex = Vector{Float64}(undef, sN); exx = Vector{Float64}(undef, sN - T)
    lg = Vector{Float64}(undef, 4sK); e1 = Vector{Float64}(undef, sK); e2 = similar(e1); mml = similar(e1); nme = similar(e1)
    GC.@preserve ϕ cpg_off ρn dn sub_off sub_p sub_q logZ ex exx lg e1 e2 mml nme begin
        check(ccall((:cpn_stat_sums_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
             Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], T, ϕ, C_NULL, C_NULL, C_NULL, T, cpg_off, ρn, dn, sub_off, sub_p, sub_q, logZ, ex, exx, lg, e1, e2, mml, nme))
    end
    for (t, rs) in enumerate(rss)
        store_summaries!(rs, t, cpg_off, sub_off, logZ, ex, exx, lg, e1, e2, mml, nme)
    end
    return nothing
end
function store_summaries!(rs, t, cpg_off, sub_off, logZ, ex, exx, lg, e1, e2, mml, nme)
    n = (cpg_off[t] + 1):cpg_off[t + 1]; k = (sub_off[t] + 1):sub_off[t + 1]
    rs.logZ = logZ[t]; rs.exps.ex = ex[n]; rs.exps.exx = exx[(cpg_off[t] - t + 2):(cpg_off[t + 1] - t)]
    rs.exps.log_g1 = e1[k]; rs.exps.log_g2 = e2[k]; rs.mml = mml[k]; rs.nme = nme[k]
    rs.tm.log_gs.pm = lg[4 .* k .- 3]; rs.tm.log_gs.pp = lg[4 .* k .- 2]; rs.tm.log_gs.qm = lg[4 .* k .- 1]; rs.tm.log_gs.qp = lg[4 .* k]
end
CpelNano.get_stat_sums!(rs::RegStruct, config::CpelNanoConfig) = get_stat_sums_batch!([rs], config)

# ---- comp_cmd(rs1, rs2)  statistical_summaries.jl:512 ↔ cpn_cmd_batch ----------------------------------------------------------
# Both models carry their get_stat_sums! outputs (E[X], E[XX], NME) and live on the same region (same ρn, dn, sub-regions).
function CpelNano.comp_cmd(rs1::RegStruct, rs2::RegStruct)::Vector{Float64}
    K = rs1.nls_rgs.num
    N = rs1.N
    pair_a = Int64[0]; pair_b = Int64[1]; reg_of = Int64[0, 0]
    ϕ = Float64[rs1.ϕhat; rs2.ϕhat]
    ex_off = Int64[0, N, 2N]; ex = Float64[rs1.exps.ex; rs2.exps.ex]; exx = Float64[rs1.exps.exx; rs2.exps.exx]
    nme_off = Int64[0, K, 2K]; nme = Float64[rs1.nme; rs2.nme]
    cpg_off = Int64[0, N]; ρn = Float64.(rs1.ρn); dn = Float64.(rs1.dn)
    sub_off = Int64[0, K]; sub_p = Int64[first(r) for r in rs1.nls_rgs.cpg_ind]; sub_q = Int64[last(r) for r in rs1.nls_rgs.cpg_ind]
    cmd_off = Int64[0, K]; cmd = Vector{Float64}(undef, K)
    GC.@preserve pair_a pair_b reg_of ϕ ex_off ex exx nme_off nme cpg_off ρn dn sub_off sub_p sub_q cmd_off cmd begin
        check(ccall((:cpn_cmd_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
             Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
            CTX[], 1, pair_a, pair_b, 2, ϕ, reg_of, ex_off, ex, exx, nme_off, nme, 1, cpg_off, ρn, dn, sub_off, sub_p, sub_q, cmd_off, cmd))
    end
    return cmd
end

# ---- group tests  grp_perm_tests.jl:200 (unmatched) / :432 (matched) ↔ cpn_grp_test_batch --------------------------------------
# ms_g1[r], ms_g2[r]: the fitted models of the two groups for region r (ϕ̂ and region features; summaries are recomputed on the
# GPU).  All regions of a call must have the same (n1, n2); the drivers below group them.  Labelings are drawn per CALL where the
# reference draws them per region (grp_perm_tests.jl:242, :461) — only the non-exact case (> 1000 labelings) is affected.
function grp_test_batch(ms_g1::Vector{Vector{RegStruct}}, ms_g2::Vector{Vector{RegStruct}}, config::CpelNanoConfig; matched::Bool=false)
    R = length(ms_g1); n1 = length(ms_g1[1]); n2 = length(ms_g2[1]); S = n1 + n2
    reps = [ms_g1[r][1] for r in 1:R]
    sub_off, sub_p, sub_q = pack_subregs(reps, config)
    cpg_off, ρn, dn = pack_cpgs(reps)
    ϕ = Float64.(vcat([vcat([m.ϕhat for m in vcat(ms_g1[r], ms_g2[r])]...) for r in 1:R]...))
    if matched
        enough = 0.5^n1 < 0.05                                                          # :457
        exact = 2^n1 < CpelNano.LMAX
        lab = !enough ? Int64[] : (exact ? collect(Int64, 0:(2^n1 - 1)) : Int64.(rand(0:(2^n1 - 1), CpelNano.LMAX)))
        P = length(lab)
    else
        L = binomial(S, n1)
        exact = L < CpelNano.LMAX                                                       # :230
        combs = L > 20 ? (exact ? collect(combinations(1:S, n1)) : collect(combinations(1:S, n1))[unique(rand(1:L, CpelNano.LMAX))]) : Vector{Vector{Int}}()
        lab = Int32.(vcat(combs...))
        P = length(combs)
    end
    ns = matched ? 2 : 3; sK = sub_off[end]
    T = Vector{Float64}(undef, ns * sK); p = similar(T); cnt = Vector{Int64}(undef, ns * sK); tcmd = Vector{Float64}(undef, sK)
    GC.@preserve ϕ cpg_off ρn dn sub_off sub_p sub_q lab T cnt p tcmd begin
        check(ccall((:cpn_grp_test_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
             Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], R, n1, n2, matched ? 1 : 0, ϕ, cpg_off, ρn, dn, sub_off, sub_p, sub_q, lab, P, exact ? 1 : 0, T, cnt, p, tcmd))
    end
    # → RegStatTestStruct per region (structures.jl:222-236), tuples (T, p) per sub-region
    out = Vector{RegStatTestStruct}(undef, R)
    for r in 1:R
        ts = RegStatTestStruct(reps[r])
        K = sub_off[r + 1] - sub_off[r]; b = ns * sub_off[r]
        for k in 1:K
            ts.tests.tmml_test[k] = (T[b + k], p[b + k])
            ts.tests.tnme_test[k] = (T[b + K + k], p[b + K + k])
            ts.tests.tcmd_test[k] = matched ? (tcmd[sub_off[r] + k], NaN) : (T[b + 2K + k], p[b + 2K + k])      # :493
        end
        out[r] = ts
    end
    return out
end
CpelNano.unmat_est_reg_test(ms_g1::Vector{RegStruct}, ms_g2::Vector{RegStruct})::RegStatTestStruct =
    grp_test_batch([ms_g1], [ms_g2], CONFIG[]; matched=false)[1]
CpelNano.mat_est_reg_test(ms_g1::Vector{RegStruct}, ms_g2::Vector{RegStruct})::RegStatTestStruct =
    grp_test_batch([ms_g1], [ms_g2], CONFIG[]; matched=true)[1]
const CONFIG = Ref{CpelNanoConfig}(CpelNanoConfig())    # the per-region test functions of the reference take no config; the drivers set it

# ---- calls ingestion: get_calls!(rs, nano, config) (nanopore.jl:241-257) for all regions of a chromosome ↔ cpn_calls_* ------------
const CALLS = Dict{String,Ptr{Cvoid}}()
function calls_handle(nano::String)
    haskey(CALLS, nano) && return CALLS[nano]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cpn_calls_open, LIB[]), Cint, (Cstring, Int32, Ref{Ptr{Cvoid}}), nano, 0, h)
    rc == 0 || error(unsafe_string(ccall((:cpn_calls_open_error, LIB[]), Cstring, ())))
    CALLS[nano] = h[]
end
calls_err(h) = error(unsafe_string(ccall((:cpn_calls_error, LIB[]), Cstring, (Ptr{Cvoid},), h)))

# Emissions of `rss` (group-scale structs of one chromosome) as the CSR arrays cpn_fit_batch takes.  rs.m is set; rs.calls is NOT
# rebuilt: the boxed MethCallCpgGrp objects were only walked by get_ave_depth / perc_gprs_obs, which cpn_region_filter replaces.
function get_calls_csr!(rss::Vector{RegStruct}, nano::String, config::CpelNanoConfig)
    h = calls_handle(nano)
    R = length(rss)
    chrst = Int64[rs.chrst for rs in rss]; chrend = Int64[rs.chrend for rs in rss]
    grp_off = Int64[0; cumsum([length(rs.cpg_grps) for rs in rss])]
    grp_st = Int64[minimum(g.grp_int) for rs in rss for g in rs.cpg_grps]
    grp_end = Int64[maximum(g.grp_int) for rs in rss for g in rs.cpg_grps]
    n_reads = Vector{Int64}(undef, R)
    chr = rss[1].chr
    rc = ccall((:cpn_calls_count_reads, LIB[]), Cint, (Ptr{Cvoid}, Cstring, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int64}),
               h, chr, R, chrst, chrend, 0, n_reads)
    rc == 0 || calls_err(h)
    read_off = Int64[0; cumsum(n_reads)]
    n_obs = sum(n_reads .* diff(grp_off))
    lu = Vector{Float64}(undef, n_obs); lm = Vector{Float64}(undef, n_obs)
    rc = ccall((:cpn_calls_fill, LIB[]), Cint,
               (Ptr{Cvoid}, Cstring, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int64}, Int32,
                Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
               h, chr, R, chrst, chrend, grp_off, grp_st, grp_end, config.mod_nse ? 1 : 0, read_off, 0, lu, lm, C_NULL)
    rc == 0 || calls_err(h)
    for (i, rs) in enumerate(rss); rs.m = n_reads[i]; end
    return grp_off, read_off, lu, lm
end

# ---- analyze_nano(nano, fasta, config)  nanopore.jl:362-416: same signature, same files; the pmap at :406 is one call ---------
function analyze_chr(chr_part, chr, nano, fa_rec, config::CpelNanoConfig)
    rss = RegStruct[]
    for reg_int in chr_part                                   # host side of pmap_analyze_reg (nanopore.jl:296-310), unchanged
        rs = RegStruct(); rs.chr = chr; rs.chrst = minimum(reg_int); rs.chrend = maximum(reg_int)
        get_grp_info!(rs, fa_rec, config.min_grp_dist)
        push!(rss, rs)
    end
    cand = [rs for rs in rss if length(rs.cpg_pos) > 9]       # :309 (regions with fewer CpGs never read their calls)
    isempty(cand) && return rss
    grp_off, read_off, lu, lm = get_calls_csr!(cand, nano, config)
    # data filters :322-326 as status words (cpn_region_filter): 0 = fit the region
    status = Vector{Int32}(undef, length(cand))
    n_cpg = Int64[length(rs.cpg_pos) for rs in cand]
    ccall((:cpn_region_filter, LIB[]), Cint, (Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Float64, Int32, Ptr{Int32}),
          length(cand), n_cpg, grp_off, read_off, lu, Float64(config.min_cov), 0, status) == 0 || error("cpn_region_filter: bad arguments")
    keep = findall(==(0), status)
    isempty(keep) && return rss
    sel = cand[keep]
    # slice the CSR emissions to the kept regions and run fit → rscle_grp_mod! → get_stat_sums! in ONE call (cpn_analyze_batch)
    Ls = diff(grp_off)[keep]; Ms = diff(read_off)[keep]
    obs_off = Int64[0; cumsum(diff(grp_off) .* diff(read_off))]
    idx = vcat([(obs_off[i] + 1):obs_off[i + 1] for i in keep]...)
    g_off, Nl, ρl, dl = pack_groups(sel)
    r_off = Int64[0; cumsum(Ms)]
    lu_s = lu[idx]; lm_s = lm[idx]
    sub_off, sub_p, sub_q = pack_subregs(sel, config)
    cpg_off, ρn, dn = pack_cpgs(sel)
    R = length(sel); sN, sK = cpg_off[end], sub_off[end]
    ϕ0 = starts(R, config)
    ϕhat = Vector{Float64}(undef, 3R); Q = Vector{Float64}(undef, R); st = Vector{Int32}(undef, R)
    logZ = Vector{Float64}(undef, R); ex = Vector{Float64}(undef, sN); exx = Vector{Float64}(undef, sN - R)
    lg = Vector{Float64}(undef, 4sK); e1 = Vector{Float64}(undef, sK); e2 = similar(e1); mml = similar(e1); nme = similar(e1)
    GC.@preserve g_off Nl ρl dl r_off lu_s lm_s ϕ0 cpg_off ρn dn sub_off sub_p sub_q ϕhat Q st logZ ex exx lg e1 e2 mml nme begin
        check(ccall((:cpn_analyze_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Int32, Int32, Float64, Float64, Float64, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], R, g_off, Nl, ρl, dl, r_off, lu_s, lm_s, ptr_or_null(ϕ0), config.max_em_init, config.max_em_iters,
            CpelNano.AMAX, CpelNano.BMAX, CpelNano.CMAX, 1, cpg_off, ρn, dn, sub_off, sub_p, sub_q,
            ϕhat, Q, st, C_NULL, logZ, ex, exx, lg, e1, e2, mml, nme))
    end
    for (t, rs) in enumerate(sel)
        st[t] >= 0 || continue                                # rs.ϕhat stays [] and rs.proc false: the writers skip the region (:334)
        rs.ϕhat = ϕhat[(3t - 2):(3t)]
        rscle_grp_mod!(rs)                                    # :338
        store_summaries!(rs, t, cpg_off, sub_off, logZ, ex, exx, lg, e1, e2, mml, nme)
        rs.proc = true
    end
    return rss                                                # → write_output(rss, config), unchanged
end

function CpelNano.analyze_nano(nano::String, fasta::String, config::CpelNanoConfig)::Nothing
    chr_names, chr_sizes = get_genome_info(fasta)
    isdir(config.out_dir) || mkdir(config.out_dir)
    config.out_files = OutputFiles(config.out_dir, config.out_prefix)
    check_output_exists(config) && return nothing
    targ_regs = isempty(config.bed_reg) ? nothing : read_bed_reg(config)
    for i = 1:length(chr_names)
        chr = chr_names[i]
        if isnothing(targ_regs)
            chr_part = get_chr_part(chr, config, fasta)
        else
            haskey(targ_regs, chr) || continue
            chr_part = get_chr_part(chr, config, fasta, targ_regs[chr])
        end
        fa_rec = get_chr_fa_rec(chr, fasta)
        write_output(analyze_chr(chr_part, chr, nano, fa_rec, config), config)      # was: pmap(pmap_analyze_reg, chr_part) :406
    end
    return nothing
end

# ---- WGBS: get_calls_bam (wgbs.jl:222-323) ↔ cpn_bam_*, analyze_bam (wgbs.jl:401-464) ↔ cpn_analyze_wgbs_batch ------------------
# The BAM file is inflated and indexed once per path (no .bai needed); the calls of all regions of a chromosome come back as one
# Int8 table (−1 unmethylated, +1 methylated, 0 not covered), which is also what cpn_analyze_wgbs_batch takes.
const BAMS = Dict{String,Ptr{Cvoid}}()
function bam_handle(bam::String)
    get!(BAMS, bam) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:cpn_bam_open, LIB[]), Cint, (Cstring, Int32, Ref{Ptr{Cvoid}}), bam, 0, h)
        rc == 0 || error("cpn_bam_open: " * unsafe_string(ccall((:cpn_bam_open_error, LIB[]), Cstring, ())))
        h[]
    end
end
bam_check(h, rc) = rc == 0 || error(unsafe_string(ccall((:cpn_bam_error, LIB[]), Cstring, (Ptr{Cvoid},), h)))
# → (n_reads [R], cell_off [R+1], calls Int8 [Σ n_reads·N]) for regions (sts[r], ens[r]) with CpG sites cpg_pos[cpg_off[r]+1 : cpg_off[r+1]]
function bam_calls(bam::String, chr::String, sts::Vector{Int64}, ens::Vector{Int64}, cpg_off::Vector{Int64}, cpg_pos::Vector{Int64},
                   pe::Bool, trim::NTuple{4,Int64})
    h = bam_handle(bam); R = length(sts); tr = Int64[trim...]
    n = zeros(Int64, R)
    GC.@preserve sts ens cpg_off cpg_pos tr n begin
        bam_check(h, ccall((:cpn_bam_count_reads, LIB[]), Cint,
            (Ptr{Cvoid}, Cstring, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int64}, Int32, Ptr{Int64}),
            h, chr, R, sts, ens, cpg_off, cpg_pos, pe, tr, 0, n))
    end
    cell_off = Int64[0; cumsum(n .* diff(cpg_off))]
    calls = zeros(Int8, cell_off[end])
    GC.@preserve sts ens cpg_off cpg_pos tr cell_off calls begin
        bam_check(h, ccall((:cpn_bam_fill_calls, LIB[]), Cint,
            (Ptr{Cvoid}, Cstring, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int64}, Int32, Ptr{Int64}, Ptr{Int8}),
            h, chr, R, sts, ens, cpg_off, cpg_pos, pe, tr, 0, cell_off, calls))
    end
    return n, cell_off, calls
end
# same signature and return value as the reference's (one region, boxed calls)
function CpelNano.get_calls_bam(bam::String, chr::String, roi_st::Int64, roi_end::Int64, cpg_pos::Vector{Int64}, pe::Bool,
                                trim::NTuple{4,Int64})::Tuple{Int64,Vector{Vector{MethCallCpgGrp}}}
    N = length(cpg_pos)
    n, _, x = bam_calls(bam, chr, [roi_st], [roi_end], Int64[0, N], cpg_pos, pe, trim)
    calls = Vector{Vector{MethCallCpgGrp}}()
    for m = 1:n[1]
        call = Vector{MethCallCpgGrp}()
        for k = 1:N
            c = MethCallCpgGrp(); v = x[(m - 1) * N + k]
            if v != 0                                          # wgbs.jl:296-307
                c.obs = true
                c.log_pyx_u = v < 0 ? log_pyx_right_x : log_pyx_wrong_x
                c.log_pyx_m = v > 0 ? log_pyx_right_x : log_pyx_wrong_x
            end
            push!(call, c)
        end
        push!(calls, call)
    end
    return n[1], calls
end
function analyze_chr_bam(chr_part, chr, bam, fa_rec, config::CpelNanoConfig)
    rss = RegStruct[]
    for reg_int in chr_part                                   # host side of pmap_analyze_reg_bam (wgbs.jl:340-353), unchanged
        rs = RegStruct(); rs.chr = chr; rs.chrst = minimum(reg_int); rs.chrend = maximum(reg_int)
        get_grp_info!(rs, fa_rec, 1)                          # every CpG site its own group (:351)
        push!(rss, rs)
    end
    cand = [rs for rs in rss if length(rs.cpg_pos) > 9]       # :353
    isempty(cand) && return rss
    cpg_off = Int64[0; cumsum([length(rs.cpg_pos) for rs in cand])]
    n, cell_off, x = bam_calls(bam, chr, Int64[rs.chrst for rs in cand], Int64[rs.chrend for rs in cand], cpg_off,
                               Int64.(vcat([rs.cpg_pos for rs in cand]...)), config.pe, config.trim)
    keep = Int[]
    for (i, rs) in enumerate(cand)                            # get_ave_depth ≥ min_cov, perc_gprs_obs ≥ 0.66, L > 1 (:363-367)
        N = length(rs.cpg_pos); rs.m = n[i]
        n[i] > 0 || continue
        tbl = reshape(view(x, (cell_off[i] + 1):cell_off[i + 1]), N, n[i])       # column = read
        depth = count(!=(0), tbl) / N
        frac = count(k -> any(!=(0), view(tbl, k, :)), 1:N) / N
        (depth >= config.min_cov && frac >= 0.66 && rs.L > 1) && push!(keep, i)
    end
    isempty(keep) && return rss
    sel = cand[keep]; R = length(sel)
    c_off, ρn, dn = pack_cpgs(sel)
    r_off = Int64[0; cumsum(n[keep])]
    xs = vcat([x[(cell_off[i] + 1):cell_off[i + 1]] for i in keep]...)
    sub_off, sub_p, sub_q = pack_subregs(sel, config)
    sN, sK = c_off[end], sub_off[end]
    ϕ0 = starts(R, config)
    ϕhat = Vector{Float64}(undef, 3R); Q = Vector{Float64}(undef, R); st = Vector{Int32}(undef, R)
    logZ = Vector{Float64}(undef, R); ex = Vector{Float64}(undef, sN); exx = Vector{Float64}(undef, sN - R)
    lg = Vector{Float64}(undef, 4sK); e1 = Vector{Float64}(undef, sK); e2 = similar(e1); mml = similar(e1); nme = similar(e1)
    GC.@preserve c_off ρn dn r_off xs ϕ0 sub_off sub_p sub_q ϕhat Q st logZ ex exx lg e1 e2 mml nme begin
        check(ccall((:cpn_analyze_wgbs_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int8}, Ptr{Float64}, Int32, Int32,
             Float64, Float64, Float64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], R, c_off, ρn, dn, r_off, xs, ptr_or_null(ϕ0), config.max_em_init, config.max_em_iters,
            CpelNano.AMAX, CpelNano.BMAX, CpelNano.CMAX, 1, sub_off, sub_p, sub_q,
            ϕhat, Q, st, C_NULL, logZ, ex, exx, lg, e1, e2, mml, nme))
    end
    for (t, rs) in enumerate(sel)
        st[t] >= 0 || continue                                # rs.ϕhat stays []: the writers skip the region (:376)
        rs.ϕhat = ϕhat[(3t - 2):(3t)]
        store_summaries!(rs, t, c_off, sub_off, logZ, ex, exx, lg, e1, e2, mml, nme)
        rs.proc = true
    end
    return rss
end
function CpelNano.analyze_bam(bam::String, fasta::String, config::CpelNanoConfig)::Nothing
    chr_names, chr_sizes = get_genome_info(fasta)
    isdir(config.out_dir) || mkdir(config.out_dir)
    config.out_files = OutputFiles(config.out_dir, config.out_prefix)
    check_output_exists(config) && return nothing
    targ_regs = isempty(config.bed_reg) ? nothing : read_bed_reg(config)
    for i = 1:length(chr_names)
        chr = chr_names[i]
        if isnothing(targ_regs)
            chr_part = get_chr_part(chr, config, fasta)
        else
            haskey(targ_regs, chr) || continue
            chr_part = get_chr_part(chr, config, fasta, targ_regs[chr])
        end
        fa_rec = get_chr_fa_rec(chr, fasta)
        write_output(analyze_chr_bam(chr_part, chr, bam, fa_rec, config), config)   # was: pmap(pmap_analyze_reg_bam, chr_part) :455
    end
    return nothing
end

# ---- diff_grp_comp(mod_fls_g1, mod_fls_g2, fasta, config)  grp_perm_tests.jl:637-687 ---------------------------------------------
# The pmap over common regions (:674) becomes one cpn_grp_test_batch call per distinct (n1, n2) of the chromosome.
function CpelNano.diff_grp_comp(mod_fls_g1::Vector{String}, mod_fls_g2::Vector{String}, fasta::String, config::CpelNanoConfig)::Nothing
    chr_names, chr_sizes = get_genome_info(fasta)
    isdir(config.out_dir) || mkdir(config.out_dir)
    config.out_diff_files = OutputDiffFiles(config.out_dir, config.out_prefix)
    check_diff_output_exists(config) && return nothing
    CONFIG[] = config
    for chr in chr_names
        fa_rec = get_chr_fa_rec(chr, fasta)
        ms_g1 = read_in_grp_mods(mod_fls_g1, chr, fa_rec, config)
        ms_g2 = read_in_grp_mods(mod_fls_g2, chr, fa_rec, config)
        comm_regs = find_comm_regs(ms_g1, ms_g2)
        length(comm_regs) > 0 || continue
        # the models of every common region (pmap_diff_grp_comp :585-624), grouped by (n1, n2)
        byshape = Dict{Tuple{Int,Int},Vector{Tuple{Int,Vector{RegStruct},Vector{RegStruct}}}}()
        for (i, reg_id) in enumerate(comm_regs)
            if config.matched
                pairs = [r for r in 1:length(ms_g1) if haskey(ms_g1[r], reg_id) && haskey(ms_g2[r], reg_id)]
                g1 = [ms_g1[r][reg_id] for r in pairs]; g2 = [ms_g2[r][reg_id] for r in pairs]
            else
                g1 = [d[reg_id] for d in ms_g1 if haskey(d, reg_id)]; g2 = [d[reg_id] for d in ms_g2 if haskey(d, reg_id)]
            end
            push!(get!(byshape, (length(g1), length(g2)), []), (i, g1, g2))
        end
        pmap_out = Vector{RegStatTestStruct}(undef, length(comm_regs))
        for ((n1, n2), items) in byshape
            (n1 == 0 || n2 == 0) && continue
            tests = grp_test_batch([x[2] for x in items], [x[3] for x in items], config; matched=config.matched)
            for (x, t) in zip(items, tests); pmap_out[x[1]] = t; end
        end
        write_diff_output([pmap_out[i] for i in 1:length(comm_regs) if isassigned(pmap_out, i)], config)
    end
    print_log("Multiple hypothesis testing")
    mult_hyp_corr(config)
    return nothing
end

# ---- two-sample test  two_samp_perm_tests.jl ↔ cpn_two_samp_null_batch / cpn_two_samp_test_batch ------------------------------------
# pmap_two_samp_null_stats(rs_ref, calls_s1, calls_s2, perm_ids, config) :41 — ONE permutation, same signature (parity entry point)
function CpelNano.pmap_two_samp_null_stats(rs_ref::RegStruct, calls_s1::Vector{Vector{MethCallCpgGrp}}, calls_s2::Vector{Vector{MethCallCpgGrp}},
                                           perm_ids::Vector{Int64}, config::CpelNanoConfig)::NTuple{3,Vector{Float64}}
    K = rs_ref.nls_rgs.num
    T = two_samp_null_batch([rs_ref], [vcat(calls_s1, calls_s2)], [length(calls_s1)], [[perm_ids]], config)
    return T[1:K], T[(K + 1):2K], T[(2K + 1):3K]
end
# T_null of several regions: pooled[r] = vcat(calls_s1, calls_s2) of region r, perms[r] = its ascending index sets
function two_samp_null_batch(refs::Vector{RegStruct}, pooled, n1s::Vector{Int}, perms, config::CpelNanoConfig)
    R = length(refs)
    grp_off, Nl, ρl, dl = pack_groups(refs)                  # refs are at GROUP scale ("go back to group scale", :198-199)
    read_off, lu, lm = pack_calls(pooled, [rs.L for rs in refs])
    perm_off = Int64[0; cumsum([length(p) for p in perms])]
    perm_ids = Int32.(vcat([vcat(p...) for p in perms]...))
    n1 = Int64.(n1s)
    ϕ0 = starts(2 * perm_off[end], config)
    sub_off, sub_p, sub_q = pack_subregs(refs, config)
    cpg_off, ρn, dn = pack_cpgs(refs)
    Ks = diff(sub_off)
    T = Vector{Float64}(undef, 3 * sum(diff(perm_off) .* Ks))
    GC.@preserve grp_off Nl ρl dl read_off lu lm perm_off n1 perm_ids ϕ0 cpg_off ρn dn sub_off sub_p sub_q T begin
        check(ccall((:cpn_two_samp_null_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
             Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Int32, Int32, Float64, Float64, Float64, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int32}),
            CTX[], R, grp_off, Nl, ρl, dl, read_off, lu, lm, perm_off, n1, perm_ids, ptr_or_null(ϕ0), config.max_em_init, config.max_em_iters,
            CpelNano.AMAX, CpelNano.BMAX, CpelNano.CMAX, 1, cpg_off, ρn, dn, sub_off, sub_p, sub_q, T, C_NULL))
    end
    return T
end

# pmap_diff_two_samp_comp for ALL common regions of a chromosome in one cpn_two_samp_test_batch call (:157-273 over the map of :445)
function two_samp_test_chr(mods_s1::Vector{RegStruct}, mods_s2::Vector{RegStruct}, nano_s1, nano_s2, fa_rec, config::CpelNanoConfig)
    R = length(mods_s1)
    ϕ1 = Float64.(vcat([m.ϕhat for m in mods_s1]...)); ϕ2 = Float64.(vcat([m.ϕhat for m in mods_s2]...))
    # per-CpG tables and sub-regions from the fitted (CpG-scale) models, before going back to group scale
    cpg_off = Int64[0; cumsum([m.N for m in mods_s1])]
    ρn = Float64.(vcat([m.ρn for m in mods_s1]...)); dn = Float64.(vcat([m.dn for m in mods_s1]...))
    sub_off, sub_p, sub_q = pack_subregs(mods_s1, config)
    refs = RegStruct[]
    for m in mods_s1                                           # group-scale features of every region (:198)
        rs = RegStruct(); rs.chr = m.chr; rs.chrst = m.chrst; rs.chrend = m.chrend
        get_grp_info!(rs, fa_rec, config.min_grp_dist)
        push!(refs, rs)
    end
    grp_off, Nl, ρl, dl = pack_groups(refs)
    if config.pval_comp
        g1, r1, lu1, lm1 = get_calls_csr!(refs, nano_s1, config); m1 = diff(r1)
        g2, r2, lu2, lm2 = get_calls_csr!(refs, nano_s2, config); m2 = diff(r2)
    else
        m1 = ones(Int64, R); m2 = ones(Int64, R)              # one blank read per sample: the ABI wants a well-formed pooled sample
        lu1 = fill(NaN, Int(grp_off[end])); lm1 = copy(lu1); lu2 = copy(lu1); lm2 = copy(lu1)
        r1 = Int64[0; cumsum(m1)]; r2 = copy(r1)
    end
    # pooled reads, sample 1 first (vcat(calls_s1, calls_s2), :54)
    Ls = diff(grp_off)
    o1 = Int64[0; cumsum(m1 .* Ls)]; o2 = Int64[0; cumsum(m2 .* Ls)]
    lu = vcat([vcat(lu1[(o1[r] + 1):o1[r + 1]], lu2[(o2[r] + 1):o2[r + 1]]) for r in 1:R]...)
    lm = vcat([vcat(lm1[(o1[r] + 1):o1[r + 1]], lm2[(o2[r] + 1):o2[r + 1]]) for r in 1:R]...)
    read_off = Int64[0; cumsum(m1 .+ m2)]
    # permutations per region: none (≤ 20 splits or pval_comp off, :192), all splits (exact, :195-206) or LMAX_TWO_SAMP samples (:211-214)
    Ps = zeros(Int64, R); exact = zeros(Int32, R)
    perm_lists = Vector{Vector{Vector{Int}}}(undef, R)
    on_device = !isnothing(SEED[])
    for r in 1:R
        perm_lists[r] = Vector{Vector{Int}}()
        (config.pval_comp && m1[r] > 0 && m2[r] > 0) || continue
        L = binomial(BigInt(m1[r] + m2[r]), BigInt(m1[r]))
        L > 20 || continue
        if L < config.LMAX_TWO_SAMP
            exact[r] = 1; perm_lists[r] = collect(combinations(1:(m1[r] + m2[r]), m1[r])); Ps[r] = length(perm_lists[r])
        else
            Ps[r] = config.LMAX_TWO_SAMP
            on_device || (perm_lists[r] = [sort!(CpelNano.StatsBase.sample(1:(m1[r] + m2[r]), m1[r]; replace=false)) for _ in 1:Ps[r]])
        end
    end
    # with a device seed the sampled index sets are drawn on the GPU (perm_ids = NULL); exact regions still need their explicit lists,
    # so they go in a separate call — here, for brevity, a device seed is only used when no region of the chromosome is exact
    use_null = on_device && all(exact .== 0)
    perm_off = Int64[0; cumsum(Ps)]
    perm_ids = use_null ? nothing : Int32.(vcat([vcat((isempty(perm_lists[r]) && Ps[r] > 0 ?
                   [sort!(CpelNano.StatsBase.sample(1:(m1[r] + m2[r]), m1[r]; replace=false)) for _ in 1:Ps[r]] : perm_lists[r])...) for r in 1:R]...))
    ϕ0 = use_null ? nothing : starts(2 * perm_off[end], config)
    sK = sub_off[end]
    T = Vector{Float64}(undef, 3sK); cnt = Vector{Int64}(undef, 3sK); nv = similar(cnt); p = similar(T)
    n1 = Int64.(m1)
    GC.@preserve grp_off Nl ρl dl read_off lu lm n1 ϕ1 ϕ2 perm_off perm_ids exact ϕ0 cpg_off ρn dn sub_off sub_p sub_q T cnt nv p begin
        check(ccall((:cpn_two_samp_test_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int32, Int32, Float64, Float64, Float64, Int32,
             Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Int32}),
            CTX[], R, grp_off, Nl, ρl, dl, read_off, lu, lm, n1, ϕ1, ϕ2, perm_off, isnothing(perm_ids) ? Ptr{Int32}(C_NULL) : pointer(perm_ids),
            exact, ptr_or_null(ϕ0), config.max_em_init, config.max_em_iters, CpelNano.AMAX, CpelNano.BMAX, CpelNano.CMAX, 1,
            cpg_off, ρn, dn, sub_off, sub_p, sub_q, T, cnt, nv, p, C_NULL, C_NULL))
    end
    out = Vector{RegStatTestStruct}(undef, R)
    for r in 1:R
        ts = RegStatTestStruct(mods_s1[r])
        K = sub_off[r + 1] - sub_off[r]; b = 3 * sub_off[r]
        for k in 1:K
            ts.tests.tmml_test[k] = (T[b + k], p[b + k]); ts.tests.tnme_test[k] = (T[b + K + k], p[b + K + k])
            ts.tests.tcmd_test[k] = (T[b + 2K + k], p[b + 2K + k])
        end
        out[r] = ts
    end
    return out
end

function CpelNano.diff_two_samp_comp(mods_path_s1::String, mods_path_s2::String, nano_s1::String, nano_s2::String, fasta::String,
                                     config::CpelNanoConfig)::Nothing
    chr_names, chr_sizes = get_genome_info(fasta)
    isdir(config.out_dir) || mkdir(config.out_dir)
    config.out_diff_files = OutputDiffFiles(config.out_dir, config.out_prefix)
    check_diff_output_exists(config) && return nothing
    for chr in chr_names
        fa_rec = get_chr_fa_rec(chr, fasta)
        mods_chr_s1 = read_mod_file_chr(mods_path_s1, chr, fa_rec, config)
        mods_chr_s2 = read_mod_file_chr(mods_path_s2, chr, fa_rec, config)
        comm_regs = find_comm_regs([mods_chr_s1], [mods_chr_s2])
        length(comm_regs) > 0 || continue
        pmap_out = two_samp_test_chr([mods_chr_s1[x] for x in comm_regs], [mods_chr_s2[x] for x in comm_regs], nano_s1, nano_s2, fa_rec, config)
        write_diff_output(pmap_out, config)
    end
    if config.pval_comp
        print_log("Multiple hypothesis testing")
        mult_hyp_corr(config)
    end
    return nothing
end

end # module
