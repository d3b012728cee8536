# CpelNanoB200.jl — ccall shim that routes CpelNano's estimation hot path to libcpelnano_b200.so.
#
# Delivered UNTESTED BY EXECUTION: this build image has no Julia.  The executable proof of the ABI is the Python host
# (cpelnano.jl_b200/host.py) and tests/, which bind exactly the same symbols with the same argument order.
#
# Usage inside the package (after `include("nanopore.jl")` in src/CpelNano.jl):
#     include("CpelNanoB200.jl")
#     CpelNanoB200.enable!("/path/to/libcpelnano_b200.so"; device=0)
# which redefines get_ϕhat!, get_stat_sums!, comp_cmd and replaces the pmap body of analyze_nano by one
# cpn_analyze_batch call per chromosome.  Config struct, entry points and output files are unchanged.
module CpelNanoB200

using ..CpelNano: RegStruct, CpelNanoConfig, MethCallCpgGrp, AnalysisRegions, get_grp_info!, get_calls!, get_ave_depth,
                  perc_gprs_obs, rscle_grp_mod!, gen_ϕ0s, print_log
import ..CpelNano

const LIB = Ref{String}("libcpelnano_b200")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function enable!(lib::String; device::Int=0)
    LIB[] = lib
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cpn_create, LIB[]), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx)
    rc == 0 || error("cpn_create failed ($rc): no sm_100 GPU; there is no CPU fallback")
    CTX[] = ctx[]
    return nothing
end

check(rc) = rc == 0 || error(unsafe_string(ccall((:cpn_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[])))

# ---- packing -------------------------------------------------------------------------------------------------
# calls::Vector{Vector{MethCallCpgGrp}} (structures.jl:121-128,168) → read-major doubles, NaN = not observed
function pack_calls(rss::Vector{RegStruct})
    n = sum(rs -> rs.m * rs.L, rss)
    lu = Vector{Float64}(undef, n); lm = Vector{Float64}(undef, n)
    k = 0
    for rs in rss, call in rs.calls, c in call
        k += 1
        lu[k] = c.obs ? c.log_pyx_u : NaN
        lm[k] = c.obs ? c.log_pyx_m : NaN
    end
    read_off = Int64[0; cumsum([rs.m for rs in rss])]
    return read_off, lu, lm
end
pack_groups(rss) = (Int64[0; cumsum([rs.L for rs in rss])], Float64.(vcat([rs.Nl for rs in rss]...)),
                    Float64.(vcat([rs.ρl for rs in rss]...)), Float64.(vcat([rs.dl for rs in rss]...)))

# get_nls_reg_info!(rs, config)  genome_analysis.jl:472
function nls_reg_info!(rs::RegStruct, config::CpelNanoConfig)
    K = ccall((:cpn_nls_reg_info, LIB[]), Int64, (Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
              rs.chrst, rs.chrend, config.max_size_subreg, rs.cpg_pos, length(rs.cpg_pos), C_NULL, C_NULL, C_NULL, C_NULL)
    st = Vector{Int64}(undef, K); en = similar(st); p = similar(st); q = similar(st)
    ccall((:cpn_nls_reg_info, LIB[]), Int64, (Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
          rs.chrst, rs.chrend, config.max_size_subreg, rs.cpg_pos, length(rs.cpg_pos), st, en, p, q)
    rs.nls_rgs = AnalysisRegions([st[k]:en[k] for k = 1:K], [p[k]:q[k] for k = 1:K])
    return p, q
end

# ---- get_ϕhat!(rs, config)  em_algorithm.jl:371 — batched --------------------------------------------------------
function get_ϕhat_batch!(rss::Vector{RegStruct}, config::CpelNanoConfig)
    R = length(rss)
    grp_off, Nl, ρl, dl = pack_groups(rss)
    read_off, lu, lm = pack_calls(rss)
    ϕ0 = Float64.(vcat([vcat(gen_ϕ0s(config.max_em_init)...) for _ in 1:R]...))      # host RNG keeps ownership of the starts
    ϕhat = Vector{Float64}(undef, 3R); Q = Vector{Float64}(undef, R); status = Vector{Int32}(undef, R)
    GC.@preserve grp_off Nl ρl dl read_off lu lm ϕ0 ϕhat Q status begin
        check(ccall((:cpn_fit_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Int32, Int32, Float64, Float64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], R, grp_off, Nl, ρl, dl, read_off, lu, lm, ϕ0, config.max_em_init, config.max_em_iters,
            CpelNano.AMAX, CpelNano.BMAX, CpelNano.CMAX, 1, ϕhat, Q, status, C_NULL, C_NULL, C_NULL))
    end
    for (r, rs) in enumerate(rss)
        rs.ϕhat = status[r] >= 0 ? ϕhat[(3r - 2):(3r)] : Vector{Float64}()          # rs.ϕhat = [] when no start converged (:385)
    end
    return nothing
end
CpelNano.get_ϕhat!(rs::RegStruct, config::CpelNanoConfig) = get_ϕhat_batch!([rs], config)

# ---- get_stat_sums!(rs, config)  statistical_summaries.jl:470 — batched ---------------------------------------------
function get_stat_sums_batch!(rss::Vector{RegStruct}, config::CpelNanoConfig)
    T = length(rss)
    pq = [nls_reg_info!(rs, config) for rs in rss]
    cpg_off = Int64[0; cumsum([rs.N for rs in rss])]; sub_off = Int64[0; cumsum([length(x[1]) for x in pq])]
    ρn = Float64.(vcat([rs.ρl for rs in rss]...)); dn = Float64.(vcat([rs.dl for rs in rss]...))       # after rscle_grp_mod!
    sub_p = vcat([x[1] for x in pq]...); sub_q = vcat([x[2] for x in pq]...)
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
        n = (cpg_off[t] + 1):cpg_off[t + 1]; k = (sub_off[t] + 1):sub_off[t + 1]
        rs.logZ = logZ[t]; rs.exps.ex = ex[n]; rs.exps.exx = exx[(cpg_off[t] - t + 2):(cpg_off[t + 1] - t)]
        rs.exps.log_g1 = e1[k]; rs.exps.log_g2 = e2[k]; rs.mml = mml[k]; rs.nme = nme[k]
        rs.tm.log_gs.pm = lg[4 .* k .- 3]; rs.tm.log_gs.pp = lg[4 .* k .- 2]; rs.tm.log_gs.qm = lg[4 .* k .- 1]; rs.tm.log_gs.qp = lg[4 .* k]
    end
    return nothing
end
CpelNano.get_stat_sums!(rs::RegStruct, config::CpelNanoConfig) = get_stat_sums_batch!([rs], config)

# ---- region loop of the group comparison  grp_perm_tests.jl:599 (unmatched) / :618 (matched), pmap at :674 ---------
# ms_g1[r], ms_g2[r]: the fitted models (ϕ̂ and region features; no summaries needed) of the two groups for region r.
# Returns (T, cnt, p[, tcmd]) flat, sub-regions of region r at sub_off[r]+1:sub_off[r+1] inside blocks of 3 (or 2) statistics.
function grp_test_batch(ms_g1::Vector{Vector{RegStruct}}, ms_g2::Vector{Vector{RegStruct}}, config::CpelNanoConfig; matched::Bool=false)
    R = length(ms_g1); n1 = length(ms_g1[1]); n2 = length(ms_g2[1]); S = n1 + n2
    reps = [ms_g1[r][1] for r in 1:R]
    pq = [nls_reg_info!(rs, config) for rs in reps]
    cpg_off = Int64[0; cumsum([rs.N for rs in reps])]; sub_off = Int64[0; cumsum([length(x[1]) for x in pq])]
    ρn = Float64.(vcat([rs.ρn for rs in reps]...)); dn = Float64.(vcat([rs.dn for rs in reps]...))
    sub_p = vcat([x[1] for x in pq]...); sub_q = vcat([x[2] for x in pq]...)
    ϕ = Float64.(vcat([vcat([m.ϕhat for m in vcat(ms_g1[r], ms_g2[r])]...) for r in 1:R]...))
    exact = matched ? 2^n1 < CpelNano.LMAX : binomial(S, n1) < CpelNano.LMAX
    lab = matched ? (exact ? collect(Int64, 0:(2^n1 - 1)) : rand(0:(2^n1 - 1), CpelNano.LMAX)) :
                    Int32.(vcat((exact ? collect(combinations(1:S, n1)) : collect(combinations(1:S, n1))[unique(rand(1:binomial(S, n1), CpelNano.LMAX))])...))
    P = matched ? length(lab) : div(length(lab), n1)
    ns = matched ? 2 : 3; sK = sub_off[end]
    T = Vector{Float64}(undef, ns * sK); p = similar(T); cnt = Vector{Int64}(undef, ns * sK); tcmd = Vector{Float64}(undef, sK)
    GC.@preserve ϕ cpg_off ρn dn sub_off sub_p sub_q lab T cnt p tcmd begin
        check(ccall((:cpn_grp_test_batch, LIB[]), Cint,
            (Ptr{Cvoid}, Int64, Int32, Int32, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64},
             Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], R, n1, n2, matched ? 1 : 0, ϕ, cpg_off, ρn, dn, sub_off, sub_p, sub_q, lab, P, exact ? 1 : 0, T, cnt, p, tcmd))
    end
    return matched ? (T, cnt, p, tcmd, sub_off) : (T, cnt, p, sub_off)
end

# ---- calls ingestion: get_calls!(rs, nano, config) (nanopore.jl:241-257) for all regions of a chromosome --------------------
# get_dic_nanopolish re-reads the TSV for every region (input_output.jl:536-570); the native loader parses it once
# (cpn_calls_open, all threads) and fills every region's calls in two calls.  CALLS caches the handle per file.
const CALLS = Dict{String,Ptr{Cvoid}}()
function calls_handle(nano::String)
    haskey(CALLS, nano) && return CALLS[nano]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cpn_calls_open, LIB[]), Cint, (Cstring, Int32, Ref{Ptr{Cvoid}}), nano, 0, h)
    rc == 0 || error(unsafe_string(ccall((:cpn_calls_open_error, LIB[]), Cstring, ())))
    CALLS[nano] = h[]
end

function get_calls_batch!(rss::Vector{RegStruct}, nano::String, config::CpelNanoConfig)
    isempty(rss) && return nothing
    h = calls_handle(nano)
    R = length(rss)
    chrst = Int64[rs.chrst for rs in rss]; chrend = Int64[rs.chrend for rs in rss]
    grp_off = cumsum(vcat(0, [length(rs.cpg_grps) for rs in rss]))
    grp_st = Int64[minimum(g.grp_int) for rs in rss for g in rs.cpg_grps]
    grp_end = Int64[maximum(g.grp_int) for rs in rss for g in rs.cpg_grps]
    n_reads = Vector{Int64}(undef, R)
    chr = rss[1].chr
    rc = ccall((:cpn_calls_count_reads, LIB[]), Cint, (Ptr{Cvoid}, Cstring, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int64}),
               h, chr, R, chrst, chrend, 0, n_reads)
    rc == 0 || error(unsafe_string(ccall((:cpn_calls_error, LIB[]), Cstring, (Ptr{Cvoid},), h)))
    read_off = cumsum(vcat(0, n_reads))
    n_obs = sum(n_reads .* diff(grp_off))
    lu = Vector{Float64}(undef, n_obs); lm = Vector{Float64}(undef, n_obs)
    rc = ccall((:cpn_calls_fill, LIB[]), Cint,
               (Ptr{Cvoid}, Cstring, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int64}, Int32,
                Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
               h, chr, R, chrst, chrend, grp_off, grp_st, grp_end, config.mod_nse ? 1 : 0, read_off, 0, lu, lm, C_NULL)
    rc == 0 || error(unsafe_string(ccall((:cpn_calls_error, LIB[]), Cstring, (Ptr{Cvoid},), h)))
    # (lu, lm, read_off, grp_off) are already the arrays cpn_fit_batch takes; RegStruct.calls is rebuilt only for callers that
    # still walk the boxed objects (get_ave_depth, perc_gprs_obs: structures.jl:200-203)
    o = 0
    for (i, rs) in enumerate(rss)
        L = grp_off[i+1] - grp_off[i]; rs.m = n_reads[i]
        rs.calls = [[begin
                         u = lu[o + (m-1)*L + l]
                         isnan(u) ? MethCallCpgGrp() : MethCallCpgGrp(u, lm[o + (m-1)*L + l])     # (log_pyx_u, log_pyx_m) ctor, structures.jl:127
                     end for l in 1:L] for m in 1:rs.m]
        o += rs.m * L
    end
    return nothing
end

# ---- batch point: replaces `pmap(x -> pmap_analyze_reg(x, chr, nano, fa_rec, config), chr_part)`  nanopore.jl:406 -----
function analyze_chr(chr_part, chr, nano, fa_rec, config::CpelNanoConfig)
    rss = RegStruct[]
    for reg_int in chr_part                                   # host side of pmap_analyze_reg (nanopore.jl:296-326), unchanged
        rs = RegStruct(); rs.chr = chr; rs.chrst = minimum(reg_int); rs.chrend = maximum(reg_int)
        get_grp_info!(rs, fa_rec, config.min_grp_dist)
        push!(rss, rs)
    end
    get_calls_batch!([rs for rs in rss if length(rs.cpg_pos) > 9], nano, config)     # was: get_calls!(rs, nano, config) per region
    keep = [rs for rs in rss if length(rs.cpg_pos) > 9 && rs.m > 0 && get_ave_depth(rs) >= config.min_cov &&
                                perc_gprs_obs(rs) >= 0.66 && rs.L > 1]
    isempty(keep) && return rss
    get_ϕhat_batch!(keep, config)
    fitted = [rs for rs in keep if length(rs.ϕhat) > 0]
    foreach(rscle_grp_mod!, fitted)
    isempty(fitted) || get_stat_sums_batch!(fitted, config)
    for rs in fitted; rs.proc = true; end
    return rss                                                # → write_output(rss, config), unchanged
end

end # module
