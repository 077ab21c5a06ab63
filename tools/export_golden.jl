# export_golden.jl -- golden vectors from the REAL reference, for a machine that has Julia + Gen +
# GenParticleFilters (this repository's build image has no Julia, so this script could not be run here; the parity
# of ancestor indices is "unpinned" until its output is committed, see oracle/genpf_oracle.h and DESIGN.md section 2).
#
#   julia --project=/path/to/GenParticleFilters.jl tools/export_golden.jl out_dir
#   python tools/import_golden.py out_dir            # -> tests/golden/reference_julia.npz
#
# For every case it runs the reference's own functions on a seeded state and dumps inputs and outputs as raw
# little-endian arrays plus manifest.json:
#   lw                      log-weights (Float64)
#   lse, ess                Gen.logsumexp(lw), effective_sample_size(state)        (src/utils.jl:163-164)
#   norm_w                  get_norm_weights(state)                                (src/utils.jl:156)
#   strat[_sorted]/r        dense stratum uniforms: r[i] = the rand() the reference drew for stratum i
#                           (src/resample.jl:161-162); strata the lazy loop skipped get 0.5 (any value gives the
#                           same ancestor there)
#   strat[_sorted]/parents  state.parents after pf_stratified_resample! (1-based Int64)
#   strat[_sorted]/lml      get_lml_est(state) after the call
#   optimal_N/u, parents, lw_out, inv_w   pf_optimal_resize!(state, N): the single rand() (src/resize.jl:171),
#                           parents, new log-weights, find_inv_w_threshold
# Multinomial / residual ancestors are not exported: Distributions' alias sampler cannot be replayed from a stream
# of uniforms (SURVEY.md 8c), only distributional checks apply to them.
using Gen, GenParticleFilters, Random, Printf

@gen function dummy_model()
    x ~ normal(0, 1)
end

function fresh_state(lw::Vector{Float64})
    n = length(lw)
    state = pf_initialize(dummy_model, (), choicemap(), n)
    state.log_weights .= lw
    return state
end

# replay of the lazy loop's draw pattern (src/resample.jl:159-168) to place the recorded draws on their strata
function dense_uniforms(weights::Vector{Float64}, order::Vector{Int}, draws::Vector{Float64})
    n = length(weights)
    r = fill(0.5, n)
    i_old, weight_step, accum_weight, k = 0, 1 / n, 0.0, 0
    for (i_new, lower) in enumerate(0.0:weight_step:1.0-weight_step)
        if lower + weight_step > accum_weight
            k += 1
            r[i_new] = draws[k]
            u = draws[k] * weight_step + lower
            while accum_weight < u
                accum_weight += weights[order[i_old+1]]
                i_old += 1
            end
        end
    end
    return r, k
end

arrays = Tuple{String,String,Int}[]   # (name, dtype, length)
function dump_array(dir, name, x::Vector{Float64})
    open(joinpath(dir, replace(name, "/" => "__") * ".bin"), "w") do io
        write(io, htol.(x))
    end
    push!(arrays, (name, "f64", length(x)))
end
function dump_array(dir, name, x::Vector{Int64})
    open(joinpath(dir, replace(name, "/" => "__") * ".bin"), "w") do io
        write(io, htol.(x))
    end
    push!(arrays, (name, "i64", length(x)))
end
dump_array(dir, name, x::Real) = dump_array(dir, name, [Float64(x)])

function export_case(dir, case, n, seed, sigma)
    Random.seed!(seed)
    lw = sigma .* randn(n)
    dump_array(dir, "$case/lw", lw)
    state = fresh_state(lw)
    dump_array(dir, "$case/lse", Gen.logsumexp(lw))
    dump_array(dir, "$case/ess", effective_sample_size(state))
    dump_array(dir, "$case/norm_w", get_norm_weights(state))
    for (tag, sorted) in (("strat", false), ("strat_sorted", true))
        state = fresh_state(lw)
        Random.seed!(seed + 1000)
        pf_stratified_resample!(state; sort_particles=sorted)
        parents = copy(state.parents)
        Random.seed!(seed + 1000)
        draws = [rand() for _ in 1:n]                      # the same stream, more than enough draws
        weights, _ = GenParticleFilters.safe_softmax(lw)
        order = sorted ? sortperm(lw, rev=true) : collect(1:n)
        r, k = dense_uniforms(weights, order, draws)
        dump_array(dir, "$case/$tag/r", r)
        dump_array(dir, "$case/$tag/parents", Vector{Int64}(parents))
        dump_array(dir, "$case/$tag/lml", get_lml_est(state))
        dump_array(dir, "$case/$tag/n_draws", k)
    end
    for N in (max(1, n ÷ 4), n ÷ 2)
        state = fresh_state(lw)
        w = GenParticleFilters.softmax(lw)
        dump_array(dir, "$case/optimal_$N/inv_w", GenParticleFilters.find_inv_w_threshold(w, N))
        Random.seed!(seed + 2000)
        u = rand()
        Random.seed!(seed + 2000)
        pf_resize!(state, N, :optimal)
        dump_array(dir, "$case/optimal_$N/u", u)
        dump_array(dir, "$case/optimal_$N/parents", Vector{Int64}(state.parents))
        dump_array(dir, "$case/optimal_$N/lw_out", copy(state.log_weights))
    end
end

function main()
    dir = length(ARGS) >= 1 ? ARGS[1] : "golden_julia"
    mkpath(dir)
    for (case, n, seed, sigma) in (("n100_s1", 100, 1, 1.0), ("n1000_s5", 1000, 2, 5.0), ("n2048_s2", 2048, 3, 2.0),
                                   ("n3000_s1", 3000, 4, 1.0), ("n65536_s1", 65536, 5, 1.0))
        export_case(dir, case, n, seed, sigma)
    end
    open(joinpath(dir, "manifest.json"), "w") do io
        println(io, "{\"reference\": \"GenParticleFilters.jl\", \"julia\": \"$(VERSION)\", \"arrays\": [")
        for (k, (name, dt, len)) in enumerate(arrays)
            @printf(io, "  {\"name\": \"%s\", \"dtype\": \"%s\", \"length\": %d}%s\n", name, dt, len, k < length(arrays) ? "," : "")
        end
        println(io, "]}")
    end
    println("wrote $(length(arrays)) arrays to $dir")
end

main()
