# GenPFCuda.jl -- thin `ccall` shim that puts libgenpf_cuda.so behind GenParticleFilters.jl's own API.
#
# NOT RUNNABLE IN THE BUILD IMAGE (no julia there); it is mechanically derived from include/genpf.h and
# exercised through the identical C ABI by the Python ctypes binding (genparticlefilters.jl_b200/_lib.py).
# Two paths, as in the north star:
#   * arbitrary Gen models: `ParticleFilterState` keeps its traces in Julia; only `log_weights` goes to the
#     GPU and `parents` comes back; traces are gathered on the host exactly like the reference does
#     (`state.new_traces .= view(state.traces, state.parents)`, src/resample.jl:60).
#   * device plugins: `DevicePFState` wraps an opaque filter handle; every call is a kernel sequence.
module GenPFCuda

using Gen, GenParticleFilters
import GenParticleFilters: pf_resample!, pf_update!, pf_rejuvenate!, pf_replicate!, pf_dereplicate!
import GenParticleFilters: pf_move_reweight!, pf_optimal_resize!, pf_introduce!  # extended below, not shadowed
import GenParticleFilters: ParticleFilterView, ParticleFilterSubState, update_refs!
import Gen: effective_sample_size, log_ml_estimate
import Statistics: mean, var

const LIB = get(ENV, "GENPF_CUDA_LIB", "libgenpf_cuda")

# ---- enums (include/genpf.h)
const MULTINOMIAL, RESIDUAL, STRATIFIED = Int32(0), Int32(1), Int32(2)
const SORT_PARTICLES, SUBSTATE, INDEX_BASE1, DEVICE_PTRS, CHECK = UInt32(1), UInt32(2), UInt32(4), UInt32(8), UInt32(16)
const METHODS = Dict(:multinomial => MULTINOMIAL, :residual => RESIDUAL, :stratified => STRATIFIED)
const WARNINGS = Dict(  # the @warn texts of safe_softmax, src/utils.jl:120,124,132,135
    1 => "NaN found in input values. Returning NaN weights.",
    2 => "All input values are -Inf. Returning uniform weights.",
    3 => "All weights are zero. Returning uniform weights.",
    4 => "Total weight is NaN. Returning NaN weights.")

last_error() = unsafe_string(ccall((:genpf_last_error, LIB), Cstring, ()))
function check(status::Int32)
    status == 0 && return nothing
    status == -3 && error("Invalid weights.")                       # src/resample.jl:55,92,151
    error(last_error())
end

# ---- utils.jl:163-171 on the GPU (opt-in: `GenPFCuda.effective_sample_size(state)`)
function gpu_effective_sample_size(state::ParticleFilterView)
    out = Ref{Cdouble}()
    lw = state.log_weights isa Vector{Float64} ? state.log_weights : collect(state.log_weights)
    check(ccall((:genpf_ess, LIB), Int32, (Ptr{Cdouble}, Int64, UInt32, Ref{Cdouble}), lw, length(lw), 0, out))
    return out[]
end

# ---- pf_resample! (src/resample.jl:19-30): ancestors + new weights on the GPU, trace gather on the host.
# `uniforms` lets a test export the reference RNG's draws (one per output slot / stratum).
function gpu_resample!(state::ParticleFilterView, method::Symbol=:multinomial;
                       priority_fn=nothing, check_=:warn, sort_particles::Bool=true,
                       uniforms::Union{Nothing,Vector{Float64}}=nothing, seed::UInt64=rand(UInt64),
                       n_out::Int=length(state.traces))
    haskey(METHODS, method) || error("Resampling method $method not recognized.")   # src/resample.jl:28
    sub = state isa ParticleFilterSubState
    lw = sub ? collect(state.log_weights) : state.log_weights
    lp = priority_fn === nothing ? C_NULL : priority_fn.(lw)          # log_prio == NULL <=> `===` branch
    n_in = length(lw)
    parents = Vector{Int64}(undef, n_out)
    lw_out = Vector{Float64}(undef, n_out)
    lml_inc, kind = Ref{Cdouble}(0.0), Ref{Int32}(0)
    flags = INDEX_BASE1 | (sub ? SUBSTATE : UInt32(0)) | (check_ == true ? CHECK : UInt32(0)) |
            ((method == :stratified && sort_particles) ? SORT_PARTICLES : UInt32(0))
    check(ccall((:genpf_resample, LIB), Int32,
                (Int32, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, UInt64, UInt32,
                 Ptr{Int64}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Int32}),
                METHODS[method], lw, lp, n_in, n_out, uniforms === nothing ? C_NULL : uniforms, seed, flags,
                parents, lw_out, lml_inc, kind))
    kind[] != 0 && check_ != false && @warn(WARNINGS[Int(kind[])])
    if sub
        state.parents .= parents
        state.new_traces .= view(state.traces, parents)
        state.log_weights .= lw_out
        update_refs!(state)                                           # copy-back, src/utils.jl:17-20
    else
        state.log_ml_est += lml_inc[]                                 # update_lml_est!, src/resample.jl:178-182
        resize!(state.parents, n_out); resize!(state.new_traces, n_out)
        state.parents .= parents
        state.new_traces .= view(state.traces, parents)               # src/resample.jl:60
        resize!(state.log_weights, n_out); state.log_weights .= lw_out
        tmp = state.traces; state.traces = state.new_traces; state.new_traces = tmp
        resize!(state.new_traces, n_out)                              # src/resize.jl:441-449
    end
    return state
end

# ---- pf_optimal_resize! (src/resize.jl:149-196): threshold search, keep set and systematic draws on the GPU
function gpu_optimal_resize!(state::ParticleFilterState, n_particles::Int; check_=:warn,
                             uniform::Union{Nothing,Float64}=nothing, seed::UInt64=rand(UInt64))
    lw = state.log_weights
    parents = Vector{Int64}(undef, n_particles)
    lw_out = Vector{Float64}(undef, n_particles)
    n_keep, inv_w, kinds = Ref{Int64}(0), Ref{Cdouble}(0.0), zeros(Int32, 2)
    u = uniform === nothing ? C_NULL : Ref{Cdouble}(uniform)
    status = ccall((:genpf_optimal_resize, LIB), Int32,
                   (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, UInt64, UInt32, Ptr{Int64}, Ptr{Cdouble},
                    Ref{Int64}, Ref{Cdouble}, Ptr{Int32}),
                   lw, length(lw), n_particles, u, seed, INDEX_BASE1 | (check_ == true ? CHECK : UInt32(0)),
                   parents, lw_out, n_keep, inv_w, kinds)
    status == -8 && throw(AssertionError(last_error()))               # @assert, src/resize.jl:181,183
    check(status)
    kinds[1] != 0 && check_ != false && @warn(WARNINGS[Int(kinds[1])])
    resize!(state.parents, n_particles); resize!(state.new_traces, n_particles)
    state.parents .= parents
    state.new_traces .= view(state.traces, parents)                   # src/resize.jl:195
    resize!(state.log_weights, n_particles); state.log_weights .= lw_out
    tmp = state.traces; state.traces = state.new_traces; state.new_traces = tmp
    resize!(state.new_traces, n_particles)                            # update_refs!(state, n), src/resize.jl:441-449
    return state
end

# ---- statistics.jl:13-17,48-54
function gpu_mean_var(state::ParticleFilterView, addr)
    x = Float64.(getindex.(state.traces, addr))
    lw = collect(state.log_weights)
    m, v = Ref{Cdouble}(), Ref{Cdouble}()
    check(ccall((:genpf_weighted_mean_var, LIB), Int32,
                (Ptr{Cdouble}, Ptr{Cdouble}, Int64, UInt32, Ref{Cdouble}, Ref{Cdouble}), lw, x, length(lw), 0, m, v))
    return m[], v[]
end

# ======================================================================= device-resident plugin models
mutable struct DevicePFState
    handle::Ptr{Cvoid}
    model::Symbol
    n_filters::Int
    t::Int
    function DevicePFState(model::Symbol, n_particles::Int; n_filters::Int=1, seed::UInt64=UInt64(0),
                           params::Union{Nothing,Vector{Float64}}=nothing, keep_history::Bool=false)
        mid, h = Ref{Int32}(), Ref{Ptr{Cvoid}}()
        check(ccall((:genpf_model_builtin, LIB), Int32, (Cstring, Ref{Int32}), String(model), mid))
        check(ccall((:genpf_filter_create, LIB), Int32,
                    (Int32, Ptr{Cdouble}, Int32, Int64, Int64, UInt64, UInt32, Ref{Ptr{Cvoid}}),
                    mid[], params === nothing ? C_NULL : params, params === nothing ? 0 : length(params),
                    n_particles, n_filters, seed, keep_history ? UInt32(2) : UInt32(0), h))
        s = new(h[], model, n_filters, 0)
        finalizer(x -> ccall((:genpf_filter_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), s)
        return s
    end
end

aux(s::DevicePFState, t::Int) = s.model == :object_motion ? [sin(t)] : Float64[]   # README.md:48, Julia's sin

"pf_initialize(model, (1,), obs_1, n) -- src/initialize.jl:31-44"
function device_initialize(model::Symbol, y_obs::Vector{Float64}, n_particles::Int; kw...)
    s = DevicePFState(model, n_particles; n_filters=length(y_obs), kw...)
    check(ccall((:genpf_initialize, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, y_obs, aux(s, 1)))
    s.t = 1
    return s
end

"pf_update!(state, (t,), (UnknownChange(),), obs_t) -- src/update.jl:12-25"
function pf_update!(s::DevicePFState, t::Int, y_obs::Vector{Float64})
    check(ccall((:genpf_update, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, t, y_obs, aux(s, t)))
    s.t = t
    return s
end

function effective_sample_size(s::DevicePFState)
    out = Vector{Float64}(undef, s.n_filters)
    check(ccall((:genpf_ess_dev, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
    return s.n_filters == 1 ? out[1] : out
end

function log_ml_estimate(s::DevicePFState)
    out = Vector{Float64}(undef, s.n_filters)
    check(ccall((:genpf_lml_dev, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
    return s.n_filters == 1 ? out[1] : out
end

"pf_resample!(state, method; priority_fn = w -> alpha*w, check, sort_particles) -- src/resample.jl:19-30"
function pf_resample!(s::DevicePFState, method::Symbol=:multinomial; priority_scale=nothing, check_=:warn,
                      sort_particles::Bool=true, n_out::Int=0)
    haskey(METHODS, method) || error("Resampling method $method not recognized.")
    kinds = zeros(Int32, s.n_filters)
    flags = (check_ == true ? CHECK : UInt32(0)) | ((method == :stratified && sort_particles) ? SORT_PARTICLES : UInt32(0))
    check(ccall((:genpf_resample_dev, LIB), Int32,
                (Ptr{Cvoid}, Int32, Int32, Cdouble, Ptr{Cdouble}, Int64, UInt32, Ptr{Cdouble}, Ptr{Int32}),
                s.handle, METHODS[method], priority_scale === nothing ? 0 : 1,
                priority_scale === nothing ? 1.0 : Float64(priority_scale), C_NULL, n_out, flags, C_NULL, kinds))
    for k in kinds
        k != 0 && check_ != false && @warn(WARNINGS[Int(k)])
    end
    return s
end

"pf_rejuvenate!(state, mh, (select(tau => latents),), n_iters) -- src/rejuvenate.jl:18-27,40-53"
function pf_rejuvenate!(s::DevicePFState, tau::Int, y_obs::Vector{Float64}, n_iters::Int=1)
    acc = zeros(Int64, s.n_filters)
    check(ccall((:genpf_rejuvenate_mh, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ptr{Int64}),
                s.handle, tau, y_obs, aux(s, tau), n_iters, acc))
    @debug "Accepted: $(acc)"                                         # src/rejuvenate.jl:47
    return s
end

"One README loop iteration (README.md:66-77) in a single call"
function pf_step!(s::DevicePFState, t::Int, obs_prev::Vector{Float64}, obs_t::Vector{Float64};
                  method::Symbol=:stratified, ess_thresh::Float64=0.5, mh_iters::Int=1)
    ess = Vector{Float64}(undef, s.n_filters)
    check(ccall((:genpf_step, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Cdouble, Int32, Ptr{Cdouble}),
                s.handle, t, obs_prev, aux(s, t - 1), obs_t, aux(s, t), METHODS[method], ess_thresh, mh_iters, ess))
    s.t = t
    return ess
end

"mean(state, t => :field) / var(...) -- src/statistics.jl:13-17,48-54; field index: fp64 fields first, then Bool fields"
function mean_var(s::DevicePFState, t::Int, field::Int)
    m, v = Vector{Float64}(undef, s.n_filters), Vector{Float64}(undef, s.n_filters)
    check(ccall((:genpf_mean_var, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, field, t, m, v))
    return m, v
end

pf_replicate!(s::DevicePFState, k::Int; layout::Symbol=:contiguous) =
    (check(ccall((:genpf_replicate, LIB), Int32, (Ptr{Cvoid}, Int64, Int32), s.handle, k, layout == :contiguous ? 0 : 1)); s)
pf_dereplicate!(s::DevicePFState, k::Int; layout::Symbol=:contiguous, method::Symbol=:keepfirst) =
    (check(ccall((:genpf_dereplicate, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Int32, Ptr{Cdouble}),
                 s.handle, k, layout == :contiguous ? 0 : 1, method == :keepfirst ? 0 : 1, C_NULL)); s)
# pf_move_reweight!(state, move_reweight, (select(tau => ...),), n_iters), src/rejuvenate.jl:74-90,125-132
pf_move_reweight!(s::DevicePFState, tau::Int, y_obs::Vector{Float64}, n_iters::Int=1) =
    (check(ccall((:genpf_rejuvenate_reweight, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int32),
                 s.handle, tau, y_obs, aux(s, tau), n_iters)); s)
pf_optimal_resize!(s::DevicePFState, n_particles::Int; check_=:warn) =
    (check(ccall((:genpf_optimal_resize_dev, LIB), Int32,
                 (Ptr{Cvoid}, Int64, Ptr{Cdouble}, UInt32, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int32}),
                 s.handle, n_particles, C_NULL, check_ == true ? CHECK : UInt32(0), C_NULL, C_NULL, C_NULL)); s)

# pf_introduce!(state, observations, n_particles), src/resize.jl:351-421: y_hist[tau, filter] = the observation history
# (the reference's `observations` choicemap covers every time step of the new traces); use_proposal: the plugin's proposal
function pf_introduce!(s::DevicePFState, y_hist::Matrix{Float64}, n_particles::Int; use_proposal::Bool=false)
    obs = permutedims(y_hist)  # row-major [tau][filter] for the C ABI
    auxh = reduce(vcat, [aux(s, tau) for tau in 1:s.t]; init=Float64[])
    check(ccall((:genpf_introduce, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                s.handle, n_particles, obs, isempty(auxh) ? C_NULL : auxh, use_proposal ? 1 : 0, C_NULL, C_NULL))
    return s
end

"T README iterations enqueued by one asynchronous call (genpf_run_steps); graph=true replays steps 2.. as a CUDA graph"
function pf_run!(s::DevicePFState, t_first::Int, y_rows::Matrix{Float64}; ess_thresh::Float64=0.5, mh_iters::Int=1,
                 graph::Bool=false)
    T = size(y_rows, 1) - 1
    obs = permutedims(y_rows)
    auxh = reduce(vcat, [aux(s, t_first - 1 + r) for r in 0:T]; init=Float64[])
    check(ccall((:genpf_run_steps, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Cdouble, Int32, UInt32),
                s.handle, t_first, T, obs, isempty(auxh) ? C_NULL : auxh, STRATIFIED, ess_thresh, mh_iters, graph ? UInt32(1) : UInt32(0)))
    s.t = t_first + T - 1
    return s
end

# ---- drop-in switch for ordinary (host-trace) states: after `GenPFCuda.enable!()` the reference's OWN entry points
# run their array arithmetic on the GPU -- the same call sites, the same keyword arguments (`check` is spelled
# `check_` inside this module only because `check` is the status helper above).  It re-defines the reference's
# methods for `ParticleFilterView` in place (Julia prints "method overwritten"); `GenPFCuda.disable!()` needs a fresh
# session, as for any method overwrite.  Untested here (no Julia in the build image), mechanically derived from
# src/resample.jl:19-30, src/utils.jl:163-171, src/resize.jl:149-196.
function enable!()
    @eval GenParticleFilters begin
        function pf_resample!(state::ParticleFilterView, method::Symbol=:multinomial;
                              priority_fn=nothing, check=:warn, sort_particles::Bool=true)
            return $(gpu_resample!)(state, method; priority_fn=priority_fn, check_=check, sort_particles=sort_particles)
        end
        pf_optimal_resize!(state::ParticleFilterState, n_particles::Int; check=:warn) =
            $(gpu_optimal_resize!)(state, n_particles; check_=check)
        Gen.effective_sample_size(state::ParticleFilterView) = $(gpu_effective_sample_size)(state)
    end
    return nothing
end

end # module
