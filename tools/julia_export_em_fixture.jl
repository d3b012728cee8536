# julia_export_em_fixture.jl — pins the EM / M-step path of CpelNano.jl itself (NLsolve, FiniteDiff, Calculus as resolved by
# the package's Manifest.toml) so that the C++ oracle's restatement of that solver, and through it the CUDA library, can be
# checked against the reference's own numbers.  This image has no Julia; a maintainer runs, from the CpelNano.jl checkout:
#
#     julia --project=. /path/to/tools/julia_export_em_fixture.jl  [out.txt]  [max_regions]
#
# and drops the output at  tests/golden/julia_em_fixture.txt ; tests/test_julia_fixture.py then stops skipping.
#
# What it does (reference entry points only; nothing is re-implemented here):
#   * the front end of pmap_analyze_reg (src/nanopore.jl:292-326) on examples/full_example: get_grp_info!, get_calls!, the
#     coverage filters — so the fixture carries the region's N_l, ρ_l, d_l and every read's (obs, log_pyx_u, log_pyx_m) in the
#     order rs.calls holds them (Dict order: only reproducible inside Julia, which is why it is exported);
#   * gen_ϕ0s (src/em_algorithm.jl:65-78) under a fixed Random.seed!, the starts written out;
#   * per start the loop of em_inst (src/em_algorithm.jl:304-350) run iteration by iteration through em_iter, so every EM
#     iterate is recorded, then conv / Q exactly as em_inst computes them;
#   * for the first EM iteration of every start, the M-step system of em_iter (src/em_algorithm.jl:152-167) solved once more with
#     store_trace / extended_trace, recording NLsolve's per-iteration x, f(x), delta and rho (ADVICE: pin trust_region_solve);
#   * pick_ϕ (src/em_algorithm.jl:35-53).
#
# Output: plain text, one record per line, `key<TAB>values…`, floats printed with repr (shortest round-trip digits).
using CpelNano
using Random
using NLsolve

out_path = length(ARGS) >= 1 ? ARGS[1] : "julia_em_fixture.txt"
max_regions = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 8

root = dirname(dirname(pathof(CpelNano)))
dir = joinpath(root, "examples", "full_example")
fasta = joinpath(dir, "reference", "hg38_chr17_43023997_43145780.fa")
nano = joinpath(dir, "nanopolish", "full_example_noise2.0_methylation_calls.sorted.tsv")
isfile(fasta) || (fasta = first(filter(f -> endswith(f, ".fa"), readdir(joinpath(dir, "reference"), join=true))))
isfile(nano) || (nano = first(filter(f -> endswith(f, ".tsv"), readdir(joinpath(dir, "nanopolish"), join=true))))

config = CpelNano.CpelNanoConfig()
config.max_em_init = 10
config.max_em_iters = 20
config.verbose = false

fl(x::Float64) = repr(x)
row(io, key, vals...) = println(io, join(vcat(String[key], String[string(v) for v in vals]), '\t'))

# M-step system of em_iter, rebuilt from the same package functions so that the trace can be switched on
function mstep_trace(io, rs, ϕinit)
    αs, βs = CpelNano.get_αβ_from_ϕ(ϕinit, rs)
    map_out = map(x -> CpelNano.get_cond_exs_log(αs, βs, x), rs.calls)
    ES = CpelNano.get_ess(map_out, rs)
    row(io, "ES", fl.(ES)...)
    function f!(U::Vector{Float64}, ϕ::Vector{Float64})
        ∇logZ = CpelNano.get_∇logZ(rs, ϕ)
        U[1] = ∇logZ[1] - ES[1] / length(rs.calls)
        U[2] = ∇logZ[2] - ES[2] / length(rs.calls)
        U[3] = ∇logZ[3] - ES[3] / length(rs.calls)
    end
    sol = try
        NLsolve.nlsolve(f!, copy(ϕinit); iterations=20, ftol=1e-3, store_trace=true, extended_trace=true)
    catch err
        nothing
    end
    if isnothing(sol)
        row(io, "nlsolve", "threw")
        return
    end
    row(io, "nlsolve", converged(sol) ? "converged" : "not_converged", sol.iterations, fl.(sol.zero)...)
    for st in sol.trace.states
        md = st.metadata
        x = haskey(md, "x") ? md["x"] : Float64[]
        fx = haskey(md, "f(x)") ? md["f(x)"] : Float64[]
        delta = haskey(md, "delta") ? md["delta"] : NaN
        rho = haskey(md, "rho") ? md["rho"] : NaN
        row(io, "nl_iter", st.iteration, fl(st.fnorm), fl(st.stepnorm), fl(Float64(delta)), fl(Float64(rho)), fl.(x)..., fl.(fx)...)
    end
end

open(out_path, "w") do io
    row(io, "fixture", "CpelNano.jl EM export", "julia", string(VERSION), "max_em_init", config.max_em_init, "max_em_iters", config.max_em_iters)
    chr_names, _ = CpelNano.get_genome_info(fasta)
    chr = chr_names[1]
    fa_rec = CpelNano.get_chr_fa_rec(chr, fasta)
    chr_part = CpelNano.get_chr_part(chr, config, fasta)
    n_done = 0
    for reg_int in chr_part
        n_done >= max_regions && break
        rs = CpelNano.RegStruct()
        rs.chr = chr
        rs.chrst = minimum(reg_int)
        rs.chrend = maximum(reg_int)
        CpelNano.get_grp_info!(rs, fa_rec, config.min_grp_dist)
        length(rs.cpg_pos) > 9 || continue
        CpelNano.get_calls!(rs, nano, config)
        CpelNano.get_ave_depth(rs) >= config.min_cov || continue
        CpelNano.perc_gprs_obs(rs) >= 0.66 || continue
        rs.L > 1 || continue
        n_done += 1
        row(io, "region", rs.chr, rs.chrst, rs.chrend, rs.L, length(rs.calls))
        row(io, "Nl", fl.(Float64.(rs.Nl))...)
        row(io, "rho_l", fl.(rs.ρl)...)
        row(io, "d_l", fl.(Float64.(rs.dl))...)
        for call in rs.calls
            row(io, "read_obs", [c.obs ? 1 : 0 for c in call]...)
            row(io, "read_u", [c.obs ? fl(c.log_pyx_u) : "NaN" for c in call]...)
            row(io, "read_m", [c.obs ? fl(c.log_pyx_m) : "NaN" for c in call]...)
        end
        Random.seed!(1234 + n_done)
        ϕ0s = CpelNano.gen_ϕ0s(config.max_em_init)
        em_out = Vector{Tuple{Vector{Float64},Float64}}()
        for (s, ϕ0) in enumerate(ϕ0s)
            row(io, "start", s, fl.(ϕ0)...)
            mstep_trace(io, rs, ϕ0)
            # the loop of em_inst, one em_iter at a time
            i = 1
            ϕt = ϕ0
            conv = false
            hit = false
            while !conv && !hit
                ϕtp1 = CpelNano.em_iter(ϕt, rs)
                conv = CpelNano.conv_check(ϕt, ϕtp1)
                hit = i >= config.max_em_iters
                ϕt = ϕtp1
                row(io, "em_iter", s, i, fl.(ϕt)...)
                i += 1
            end
            conv = conv && i > 2
            Q = conv ? CpelNano.comp_ave_log_py(rs, ϕt) : -Inf
            row(io, "em_end", s, conv ? 1 : 0, i, fl(Q), fl.(ϕt)...)
            push!(em_out, (ϕt, Q))
        end
        ϕhat, Qhat = CpelNano.pick_ϕ(em_out)
        row(io, "picked", fl(Qhat), (Qhat > -Inf ? fl.(ϕhat) : String[])...)
    end
    row(io, "end", n_done)
end
println("wrote ", out_path)
