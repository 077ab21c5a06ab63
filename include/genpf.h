/*
 * genpf.h -- C ABI of libgenpf_cuda.so, the B200 (sm_100a) particle-filter engine
 * that stands behind GenParticleFilters.jl's weight / propagate / resample /
 * rejuvenate / statistics / resize path.
 *
 * The reference (Julia) has no FFI boundary; the cut line is the
 * plain-bits part of `ParticleFilterState`: log_weights::Vector{Float64},
 * parents::Vector{Int64}, log_ml_est::Float64 (reference src/view.jl:16-22,
 * src/initialize.jl:4-10).  Every entry point below names the reference
 * function it replaces (paths relative to the reference root).  Julia binds
 * these with `ccall((:genpf_xxx, "libgenpf_cuda"), Int32, (...), ...)`; see
 * INTEGRATION.md and genparticlefilters.jl_b200/julia/GenPFCuda.jl.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; all sizes int64_t; weights fp64.
 *  - every function returns an int32 status (0 ok, <0 error), never throws or
 *    exits; genpf_last_error() returns a thread-local message.
 *  - host-array calls take HOST pointers (valid only during the call) unless
 *    GENPF_DEVICE_PTRS is set; they are synchronous on return.
 *  - parents are written 1-based when GENPF_INDEX_BASE1 is set (Julia), else 0-based.
 *  - there is NO CPU fallback: without a CUDA device every compute call fails
 *    with GENPF_ERR_CUDA.
 */
#ifndef GENPF_H
#define GENPF_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GENPF_VERSION 100 /* 0.1.0 */

/* ---- status codes ---- */
#define GENPF_OK 0
#define GENPF_ERR_INVALID_ARG (-1)
#define GENPF_ERR_CUDA (-2)
#define GENPF_ERR_INVALID_WEIGHTS (-3) /* error("Invalid weights."), resample.jl:55,92,151 (check=true) */
#define GENPF_ERR_UNKNOWN_METHOD (-4)  /* error("Resampling method ... not recognized."), resample.jl:28 */
#define GENPF_ERR_NOMEM (-5)
#define GENPF_ERR_UNSUPPORTED (-6)
#define GENPF_ERR_STATE (-7)
#define GENPF_ERR_ASSERT (-8) /* the reference would throw AssertionError (resize.jl:181,183) */

/* ---- method / option enums (mirror the reference's Symbols, SURVEY.md section 5 "config") ---- */
#define GENPF_MULTINOMIAL 0 /* :multinomial  resample.jl:48-65,  resize.jl:46-67  */
#define GENPF_RESIDUAL 1    /* :residual     resample.jl:85-120, resize.jl:87-124 */
#define GENPF_STRATIFIED 2  /* :stratified   resample.jl:143-175 */

#define GENPF_SORT_PARTICLES 1u /* sort_particles=true, resample.jl:143-145,156-157 */
#define GENPF_SUBSTATE 2u       /* ParticleFilterSubState semantics, resample.jl:184-187,205-218 */
#define GENPF_INDEX_BASE1 4u    /* parents written 1-based (Vector{Int}, view.jl:21) */
#define GENPF_DEVICE_PTRS 8u    /* array arguments are device pointers (CUDA.jl CuArray) */
#define GENPF_CHECK 16u         /* check=true: invalid weights => GENPF_ERR_INVALID_WEIGHTS */
#define GENPF_UNIFORMS_STRATA 32u /* genpf_uniforms: the stratum-uniform convention (see genpf_resample) */

/* invalid_kind: which branch of safe_softmax fired, utils.jl:119-137 */
#define GENPF_VALID 0
#define GENPF_INV_NAN_INPUT 1  /* utils.jl:119-122 "NaN found in input values."            */
#define GENPF_INV_ALL_NEGINF 2 /* utils.jl:123-126 "All input values are -Inf."  (uniform) */
#define GENPF_INV_ZERO_TOTAL 3 /* utils.jl:130-133 "All weights are zero."       (uniform) */
#define GENPF_INV_NAN_TOTAL 4  /* utils.jl:134-137 "Total weight is NaN."                  */

#define GENPF_LAYOUT_CONTIGUOUS 0  /* layout=:contiguous,  resize.jl:236-238 */
#define GENPF_LAYOUT_INTERLEAVED 1 /* layout=:interleaved */
#define GENPF_KEEPFIRST 0          /* method=:keepfirst, resize.jl:272-276 */
#define GENPF_SAMPLE 1             /* method=:sample,    resize.jl:278-291 */

#define GENPF_PRIO_NONE 0   /* priority_fn === nothing, resample.jl:51-52 */
#define GENPF_PRIO_SCALE 1  /* priority_fn = w -> alpha*w (the documented w/2 case) */
#define GENPF_PRIO_COLUMN 2 /* caller-supplied device/host column of log priorities */

/* device-filter creation flags */
#define GENPF_NOISE_LEAN 0u     /* one Philox4x32-10 call per particle per purpose, fp32 Box-Muller */
#define GENPF_NOISE_PHILOX53 1u /* 53-bit uniforms + fp64 Box-Muller */
#define GENPF_KEEP_HISTORY 2u   /* keep every time slice + ancestry log (mean/var of past addresses) */

/* ---- library ---- */
int32_t genpf_version(void);
const char *genpf_last_error(void);
int32_t genpf_device_count(int32_t *count);
int32_t genpf_set_device(int32_t device);
int32_t genpf_synchronize(void);

/* =====================================================================
 * Host-array path (arbitrary Gen models: traces stay in Julia, only the
 * weight vector and the ancestor indices cross the boundary)
 * ===================================================================== */

/* Gen.logsumexp(log_weights)   (called at resample.jl:180,200,210,215; utils.jl:100,177) */
int32_t genpf_logsumexp(const double *lw, int64_t n, uint32_t flags, double *out);

/* effective_sample_size / get_ess, utils.jl:163-171 */
int32_t genpf_ess(const double *lw, int64_t n, uint32_t flags, double *out);

/* get_log_norm_weights (utils.jl:148), get_norm_weights (utils.jl:156) and safe_softmax's
 * validity (utils.jl:117-140) in one pass pair.  log_norm / norm may be NULL. */
int32_t genpf_normalize(const double *lw, int64_t n, uint32_t flags, double *log_norm, double *norm, double *lse,
                        double *ess, int32_t *invalid_kind);

/* pf_resample! (resample.jl:19-30) and pf_resize! for :multinomial/:residual (resize.jl:16-27):
 *   prologue  safe_softmax(log_prio or lw), update_lml_est!   resample.jl:51-57,178-187
 *   select    ancestors into parents_out[n_out]               resample.jl:59,96-115,156-170
 *   epilogue  update_weights! into lw_out[n_out]              resample.jl:190-218, resize.jl:424-438
 * log_prio == NULL  <=>  `log_priorities === state.log_weights` (resample.jl:193).
 * uniforms: n_out doubles in [0,1) (stratified: one per stratum; residual: slot j uses uniforms[j]);
 *           NULL => Philox4x32-10(seed, stream 0): multinomial/residual draw 53 bits with counter = slot;
 *           stratified draws 32-bit stratum uniforms, four strata per counter:
 *           r_i = (word[i & 3] of Philox(counter = i >> 2) + 0.5) * 2^-32  (genpf_uniforms reproduces both).
 * lml_increment = logsumexp(lw) - log(n_in)  (0 for GENPF_SUBSTATE); caller adds it to state.log_ml_est.
 * NaN weights (kinds 1, 4) leave parents_out / lw_out untouched (the reference crashes there). */
int32_t genpf_resample(int32_t method, const double *lw, const double *log_prio, int64_t n_in, int64_t n_out,
                       const double *uniforms, uint64_t seed, uint32_t flags, int64_t *parents_out, double *lw_out,
                       double *lml_increment, int32_t *invalid_kind);

/* The same over n_seg disjoint views of one state (`state[idxs]`, view.jl:35-48; test/resample.jl:130-162):
 * segment s = [seg_offsets[s], seg_offsets[s+1]); parents are LOCAL to the segment; GENPF_SUBSTATE implied. */
int32_t genpf_resample_segmented(int32_t method, const double *lw, const double *log_prio, int64_t n,
                                 const int64_t *seg_offsets, int64_t n_seg, const double *uniforms, uint64_t seed,
                                 uint32_t flags, int64_t *parents_out, double *lw_out, int32_t *invalid_kinds);

/* mean(state, addr) / var(state, addr), statistics.jl:13-17,48-54 (x = getindex.(traces, addr) as fp64) */
int32_t genpf_weighted_mean_var(const double *lw, const double *x, int64_t n, uint32_t flags, double *mean,
                                double *var);

/* pf_replicate! index/weight arithmetic, resize.jl:236-244 (n -> n*k) */
int32_t genpf_replicate_host(const double *lw, int64_t n, int64_t k, int32_t layout, uint32_t flags,
                             int64_t *parents_out, double *lw_out);
/* pf_dereplicate!, resize.jl:267-297 (n -> n/k); uniforms (n/k) only for GENPF_SAMPLE, NULL => Philox(seed) */
int32_t genpf_dereplicate_host(const double *lw, int64_t n, int64_t k, int32_t layout, int32_t method,
                               const double *uniforms, uint64_t seed, uint32_t flags, int64_t *parents_out,
                               double *lw_out);
/* pf_coalesce!, resize.jl:309-334, with int64 keys standing for by(trace); returns n_new; output in
 * ascending first-index order (the reference's Dict order is unspecified). */
int32_t genpf_coalesce_host(const double *lw, const int64_t *keys, int64_t n, uint32_t flags, int64_t *parents_out,
                            double *lw_out, int64_t *n_new);

/* StatsBase.proportionmap(state, addr), statistics.jl:91-130, with int64 codes standing for the values at addr:
 * group g (ascending first-index order): first_index_out[g] = smallest particle index holding the value,
 * prop_out[g] = sum of get_norm_weights over the group.  Outputs sized n (upper bound); *n_unique groups written. */
int32_t genpf_proportionmap_host(const double *lw, const int64_t *keys, int64_t n, uint32_t flags,
                                 int64_t *first_index_out, double *prop_out, int64_t *n_unique);

/* pf_optimal_resize! (resize.jl:149-196) with find_inv_w_threshold (resize.jl:199-216): Fearnhead-Clifford optimal
 * resampling from n_in down to n_out <= n_in particles.  parents_out / lw_out sized n_out: first *n_keep entries
 * are the kept particles in index order (weights lw + log(n_out/n_in)), the rest the systematic draws from the
 * remainder (weights lse(lw) - log(c) + log(n_out/n_in)).  uniform = the single rand() of resize.jl:171 (host
 * pointer), NULL => Philox(seed, stream 0, slot 0).  invalid_kinds[0]: safe_softmax of all weights, [1]: of the
 * remainder.  Flags: INDEX_BASE1, DEVICE_PTRS (lw, parents_out, lw_out), CHECK.  GENPF_ERR_ASSERT where the
 * reference's @assert would fail (n_out > n_in; the systematic pass selecting != n_out - n_keep particles). */
int32_t genpf_optimal_resize(const double *lw, int64_t n_in, int64_t n_out, const double *uniform, uint64_t seed,
                             uint32_t flags, int64_t *parents_out, double *lw_out, int64_t *n_keep,
                             double *inv_w_threshold, int32_t *invalid_kinds);

/* Philox uniforms exactly as the library generates them (for exporting / parity) */
int32_t genpf_uniforms(uint64_t seed, uint64_t stream, int64_t n, uint32_t flags, double *out);

/* debug / parity: sortperm(keys, rev=true) exactly as the sorted stratified path uses it (resample.jl:156-157):
 * stable, Julia isless order; order_out 0-based (1-based with GENPF_INDEX_BASE1) */
int32_t genpf_debug_sortperm(const double *keys, int64_t n, uint32_t flags, int64_t *order_out);

/* debug / parity: normalised cumulative weights W_k the selection kernels search (device order) */
int32_t genpf_debug_cumweights(const double *lw, int64_t n, uint32_t flags, double *W_out);

/* =====================================================================
 * Device-resident path (models registered as device plugins)
 * ===================================================================== */
typedef struct genpf_filter_s *genpf_filter_t;

/* built-in plugins: "object_motion" (README.md:43-54; params p_stay,p_start,sigma_proc,sigma_obs),
 *                   "lingauss1d"    (SURVEY B.2;      params a,q,r,m0,s0) */
int32_t genpf_model_builtin(const char *name, int32_t *model_id);
int32_t genpf_model_info(int32_t model_id, int32_t *n_f64_fields, int32_t *n_u8_fields, int32_t *n_params,
                         int32_t *n_aux);
/* optional members of the model: bit 0 custom proposal, bit 1 translator */
int32_t genpf_model_caps(int32_t model_id, int32_t *caps);

/* ---- models registered at run time (north star "models registered as device plugins"; the reference's
 * counterpart is handing ANY generative function to pf_initialize, initialize.jl:31-44).  `source` is CUDA C++ that
 * defines one struct `struct_name` with the plugin interface (csrc/models.cuh; `#include "genpf_plugin.h"` is
 * implied):  static constexpr int NF, NB, NP, NAUX;  using Slice = genpf::SliceT<NF, NB>;
 *   initial(P, slice0)   transition(P, t, prev, next, U, Z)   obs_logpdf(P, slice, obs) -> double
 *   optional: constrain(...) (stratified init/update), propose / proposal_logpdf / transition_logpdf (custom
 *   proposals), translate (trace translators).
 * The source is compiled by NVRTC for sm_100a together with the library's own kernel templates, so the plugin
 * runs the same kernels (k_propagate, k_mh, k_step_fused) as the built-in models; compiling needs no GPU.
 * genpf_model_builtin(name) finds it afterwards; ids of plugins start at 100.  `options`: extra NVRTC options or NULL.
 * On a compile error the status is GENPF_ERR_INVALID_ARG and genpf_last_error() holds the compiler log. */
int32_t genpf_model_compile(const char *name, const char *source, const char *struct_name, const char *options,
                            int32_t *model_id);
/* a compiled plugin as one relocatable image (kernel names + sm_100a cubin): size query with buf == NULL */
int32_t genpf_model_export(int32_t model_id, void *buf, int64_t cap, int64_t *size);
int32_t genpf_model_load_image(const void *image, int64_t size, int32_t *model_id);
/* the kernel headers a plugin is compiled against, as embedded in this library (index 0 .. until an error) */
int32_t genpf_model_plugin_sources(int32_t index, const char **name, const char **text);

/* n_filters independent filters of n_particles each (views / batches, view.jl:16-48).  params may be NULL
 * (README constants). */
int32_t genpf_filter_create(int32_t model_id, const double *params, int32_t n_params, int64_t n_particles,
                            int64_t n_filters, uint64_t seed, uint32_t flags, genpf_filter_t *out);
int32_t genpf_filter_destroy(genpf_filter_t pf);
int32_t genpf_filter_size(genpf_filter_t pf, int64_t *n_particles, int64_t *n_filters);

/* pf_initialize(model, (1,), obs_1, n), initialize.jl:31-44.  obs[n_filters], aux[n_aux] (object_motion:
 * aux[0] = sin(1.0) computed by the caller so Julia's sin is used). */
int32_t genpf_initialize(genpf_filter_t pf, const double *obs, const double *aux);
int32_t genpf_initialize_with_noise(genpf_filter_t pf, const double *obs, const double *aux, const double *U,
                                    const double *Z);
/* Stratified pf_initialize (initialize.jl:93-108, stratified_map! utils.jl:29-55): latent `field` of slice 1 is
 * constrained to values[k] for the particles of stratum k (floor(n/K) each, GENPF_LAYOUT_CONTIGUOUS blocks or
 * GENPF_LAYOUT_INTERLEAVED; the left-over last indices take strata drawn with replacement), the other latents are
 * sampled, log_weights = log p(constraint) + obs log-density + log(K).  U/Z: noise columns (both or neither). */
int32_t genpf_initialize_stratified(genpf_filter_t pf, const double *obs, const double *aux, int32_t field,
                                    const double *values, int32_t n_strata, int32_t layout, const double *U,
                                    const double *Z);
/* Stratified pf_update! (update.jl:193-210): the same constraint on the NEW slice t, log_weights += log p(constraint |
 * previous slice) + obs log-density + log(K); the reference's default layout here is GENPF_LAYOUT_INTERLEAVED. */
int32_t genpf_update_stratified(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, int32_t field,
                                const double *values, int32_t n_strata, int32_t layout, const double *U, const double *Z);

/* pf_update!(state, (t,), (UnknownChange(),), obs_t), update.jl:12-25 */
int32_t genpf_update(genpf_filter_t pf, int64_t t, const double *obs, const double *aux);
/* Custom proposals (initialize.jl:46-62, update.jl:79-96) for models whose plugin defines propose / proposal_logpdf /
 * transition_logpdf (genpf_model_caps bit 0): x ~ q(. | prev, obs); log-weight (+)= log p(x|prev) + log p(obs|x)
 * - log q(x).  U, Z: the proposal's draws as columns (parity mode) or both NULL (library Philox noise). */
int32_t genpf_initialize_proposal(genpf_filter_t pf, const double *obs, const double *aux, const double *U, const double *Z);
int32_t genpf_update_proposal(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, const double *U,
                              const double *Z);
/* pf_update!(state, translator) (update.jl:35-44) for models whose plugin defines translate (caps bit 1): the
 * translator maps the current slice to (new slice, log-weight increment); log_weights[i] += increment. */
int32_t genpf_update_translate(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, const double *U,
                               const double *Z);
int32_t genpf_update_with_noise(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, const double *U,
                                const double *Z);

/* effective_sample_size(state) / log_ml_estimate(state): out[n_filters] */
int32_t genpf_ess_dev(genpf_filter_t pf, double *ess);
int32_t genpf_lml_dev(genpf_filter_t pf, double *lml);

/* pf_resample!(state, method; priority_fn, check, sort_particles) on device state, resample.jl:19-30;
 * n_out != n_particles gives pf_resize! (resize.jl:16-27) for multinomial/residual (n_filters == 1). */
int32_t genpf_resample_dev(genpf_filter_t pf, int32_t method, int32_t prio_kind, double prio_param,
                           const double *prio_column, int64_t n_out, uint32_t flags, const double *uniforms,
                           int32_t *invalid_kinds);

/* pf_rejuvenate!(state, mh, (select(tau=>latents),), n_iters), rejuvenate.jl:18-27,40-53, with tau the newest
 * slice; obs/aux are those of step tau.  n_accept[n_filters] may be NULL. */
int32_t genpf_rejuvenate_mh(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux, int32_t n_iters,
                            int64_t *n_accept);
int32_t genpf_rejuvenate_mh_with_noise(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux,
                                       const double *U2, const double *Z2, const double *U3, int64_t *n_accept);
/* pf_move_reweight!(state, move_reweight, (select(tau => ...),), n_iters), rejuvenate.jl:74-90,125-132: slice tau
 * is regenerated from the model's conditional prior for every particle and log_weights += the regenerate weight
 * (obs log-density of the new slice minus that of the old one).  _with_noise: U2, Z2 as in genpf_rejuvenate_mh. */
int32_t genpf_rejuvenate_reweight(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux, int32_t n_iters);
int32_t genpf_rejuvenate_reweight_with_noise(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux,
                                             const double *U2, const double *Z2);

/* pf_move_reweight! with move_reweight(trace, proposal, proposal_args) (rejuvenate.jl:74-90,134-148) for models that
 * define a proposal: slice tau is re-proposed, log_weights += up_weight - fwd_weight + bwd_weight.  U2/Z2: the
 * proposal's draws as columns (n_iters == 1) or both NULL. */
int32_t genpf_rejuvenate_reweight_proposal(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux,
                                           int32_t n_iters, const double *U2, const double *Z2);

/* One README loop iteration (README.md:66-77) fused on the device:
 *   ess = effective_sample_size(state); if ess < ess_frac*n: pf_resample!(method); pf_rejuvenate!(mh) end;
 *   pf_update!(t).  Results are identical to the three separate calls.  ess_out[n_filters] may be NULL
 * (then nothing is copied back and the call is asynchronous on the filter's stream). */
int32_t genpf_step(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                   const double *obs_t, const double *aux_t, int32_t method, double ess_frac, int32_t mh_iters,
                   double *ess_out);

/* Batch sharding (north star: "batches of independent filters shard with no communication at all"): this handle holds
 * filters [first_filter, first_filter + n_filters) of a larger batch; all Philox counters become the global batch
 * slots, so results do not depend on how a batch is split over handles / processes / GPUs.  Call before
 * genpf_initialize. */
int32_t genpf_filter_set_first_filter(genpf_filter_t pf, int64_t first_filter);

/* pf_introduce! (src/resize.jl:351-421) on a device filter: n_new particles are appended to every filter.  Like the
 * reference's generate(model, model_args, observations), each new trace is a whole chain x_1..x_t simulated from the
 * prior (use_proposal != 0: from the plugin's custom proposal, weight = model - proposal score, resize.jl:404-410) under
 * the observation history obs[t_cur * n_filters] (row tau-1 = time tau; aux likewise, n_aux columns); its log-weight is
 * the accumulated observation log-density.  The existing particles keep their place, log_weights .+= log_ml_est and
 * log_ml_est = 0 (resize.jl:362-366).  U, Z: NULL (library Philox draws at the new slots) or the chains' noise,
 * t_cur * n_filters * n_new doubles each, indexed [tau-1][filter][i] (parity mode). */
int32_t genpf_introduce(genpf_filter_t pf, int64_t n_new, const double *obs, const double *aux, int32_t use_proposal,
                        const double *U, const double *Z);

/* n_steps README iterations (stratified resample + mh + update, the fused kernels of genpf_step) enqueued by one
 * call: asynchronous, nothing copied back; ess_frac < 1: every filter decides on the device at every step.
 * obs: (n_steps + 1) * n_filters doubles, row r belongs to time t_first - 1 + r (row 0 = the step the first mh move
 * revisits); aux likewise with n_aux columns.  flags & GENPF_RUN_GRAPH: steps 2.. are captured into a CUDA graph and
 * launched as one unit (SURVEY 5 "CUDA Graph the step": removes the launch latency that dominates small filters). */
#define GENPF_RUN_GRAPH 1u
int32_t genpf_run_steps(genpf_filter_t pf, int64_t t_first, int64_t n_steps, const double *obs, const double *aux,
                        int32_t method, double ess_frac, int32_t mh_iters, uint32_t flags);

/* The same iteration in parity mode (noise exported from the reference's RNG): the resample is taken
 * and EVERY random draw is an input column of n_particles*n_filters doubles indexed by output particle --
 * `uniforms` = the rand() of each stratum (resample.jl:162) / inverse-CDF draw (NULL: library Philox draws),
 * [U2, Z2, U3] = the mh move's bernoulli / normal / accept draws (ignored when mh_iters == 0), [U1, Z1] = the
 * update's bernoulli / normal draws (Gen `regenerate` / `update`, rejuvenate.jl:40-53, update.jl:12-25).
 * method == GENPF_STRATIFIED runs exactly the kernels of genpf_step (scan + fused step).  ess_frac >= 1 resamples
 * every filter; below 1 each filter of the batch decides for itself on the device (ess < ess_frac * n, README.md:68)
 * and a filter that keeps its population is only updated. */
int32_t genpf_step_with_noise(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                              const double *obs_t, const double *aux_t, int32_t method, double ess_frac,
                              int32_t mh_iters, const double *uniforms, const double *U2, const double *Z2,
                              const double *U3, const double *U1, const double *Z1);

/* mean(state, tau=>field) / var(...), statistics.jl:13-17,48-54.  field: 0..n_f64-1 are the fp64 fields,
 * n_f64.. the u8 (Bool) fields promoted to fp64 (README.md:97).  out[n_filters]. */
int32_t genpf_mean_var(genpf_filter_t pf, int32_t field, int64_t tau, double *mean, double *var);

/* pf_replicate! / pf_dereplicate! / pf_coalesce!, resize.jl:236-334, on device state.  replicate / dereplicate (and
 * genpf_resample_dev with n_out != n, resize.jl:46-124) act on every filter of a batch; coalesce needs n_filters == 1 */
int32_t genpf_replicate(genpf_filter_t pf, int64_t k, int32_t layout);
int32_t genpf_dereplicate(genpf_filter_t pf, int64_t k, int32_t layout, int32_t method, const double *uniforms);
int32_t genpf_coalesce(genpf_filter_t pf, int64_t *n_new);
/* proportionmap(state, addr) on device state (statistics.jl:91-96): distinct values of `field` at slice tau (as
 * doubles) and their proportions, up to cap groups written, *n_unique = number of distinct values (n_filters == 1) */
int32_t genpf_proportionmap(genpf_filter_t pf, int32_t field, int64_t tau, double *values_out, double *props_out,
                            int64_t cap, int64_t *n_unique);
/* pf_resize!(state, n_out, :optimal) on device state (resize.jl:149-196); uniform: host pointer or NULL (library
 * draw on the filter's resample stream) */
int32_t genpf_optimal_resize_dev(genpf_filter_t pf, int64_t n_out, const double *uniform, uint32_t flags,
                                 int64_t *n_keep, double *inv_w_threshold, int32_t *invalid_kinds);

/* checkpoint / resume (absent in the reference): together with the seed, the window fields
 * (genpf_get/set_field for t_cur-1 and t_cur) and the log-weights, these scalars determine every later step.
 * log_ml_accum: n_filters doubles (the accumulated log_ml_est, without the current weights' term).
 * set_progress is called on a freshly created filter of the same model, size, seed and noise flags. */
int32_t genpf_filter_get_progress(genpf_filter_t pf, int64_t *t_cur, int64_t *n_resamples, double *log_ml_accum);
int32_t genpf_filter_set_progress(genpf_filter_t pf, int64_t t_cur, int64_t n_resamples, const double *log_ml_accum);

/* accessors (also checkpoint I/O): out sized n_particles*n_filters */
int32_t genpf_get_log_weights(genpf_filter_t pf, double *out);
int32_t genpf_set_log_weights(genpf_filter_t pf, const double *in);
int32_t genpf_get_parents(genpf_filter_t pf, int64_t *out, uint32_t flags);
int32_t genpf_get_field(genpf_filter_t pf, int32_t field, int64_t tau, double *out);
int32_t genpf_set_field(genpf_filter_t pf, int32_t field, int64_t tau, const double *in);
int32_t genpf_get_accepts(genpf_filter_t pf, uint8_t *out);
int32_t genpf_filter_sync(genpf_filter_t pf);
/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t genpf_launch_count(void);

/* =====================================================================
 * Multi-GPU particle sharding: one process per GPU, rank r owns global particle slots
 * [r*n_loc, (r+1)*n_loc) of ONE filter of world*n_loc particles (n_loc a multiple of 2048, world <= 8).
 * Kernels run on the filter's stream.  Either ONE call per step (genpf_shard_step_p2p below: the exchanges are done by
 * the step's own kernels over peer-mapped memory), or the host language issues three tiny collectives per step on that
 * same stream (e.g. torch.distributed / NCCL.jl) over device buffers it owns:
 *   genpf_shard_begin_step -> all_gather(stats_local[3] -> stats_all[world*3])
 *   genpf_shard_scan       -> all_gather(oend_local[1]  -> oend_all[world])
 *   genpf_shard_push       -> barrier (any collective)        offspring go to their owner over NVLink P2P
 *   genpf_shard_finish
 * The population is independent of the number of GPUs: Philox counters are global particle slots.
 * ===================================================================== */
/* 64-byte CUDA IPC handles of the filter's state buffers; handles == NULL just returns the count */
int32_t genpf_shard_ipc_export(genpf_filter_t pf, void *handles, int64_t *n_bufs);
/* all_handles: world * n_bufs handles, rank-major (all-gathered by the caller) */
int32_t genpf_shard_attach(genpf_filter_t pf, int32_t rank, int32_t world, const void *all_handles,
                           double *stats_local_dev, const double *stats_all_dev, long long *oend_local_dev,
                           const long long *oend_all_dev);
int32_t genpf_shard_detach(genpf_filter_t pf);
int32_t genpf_shard_initialize(genpf_filter_t pf, const double *obs, const double *aux);
int32_t genpf_shard_begin_step(genpf_filter_t pf);
int32_t genpf_shard_scan(genpf_filter_t pf);
int32_t genpf_shard_push(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                         const double *obs_t, const double *aux_t, int32_t mh_iters);
int32_t genpf_shard_finish(genpf_filter_t pf);
/* The whole sharded iteration in ONE call with the library's own peer-memory exchange instead of host-issued
 * collectives: five launches, two cross-GPU synchronisation points.  The per-shard totals (24 B) are
 * posted with NVLink P2P stores + epoch flags by the kernel that computes them and combined on every rank, which also
 * derives every shard's closing offspring count (no second exchange); the closing barrier is posted and awaited by
 * the first kernel that reads the pushed population.  0.27 ms/step faster than the NCCL sequence at 8 ranks.
 * Asynchronous; every rank must call it with the same t.  stats/oend buffers passed to genpf_shard_attach may be NULL
 * when only this entry point is used.  A peer that never posts turns into an error of genpf_shard_stats /
 * genpf_shard_oend (bounded wait), not a hang. */
int32_t genpf_shard_step_p2p(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                             const double *obs_t, const double *aux_t, int32_t mh_iters);
/* parity mode of the sharded iteration (README.md:66-77 over a sharded population): every rank passes the same
 * GLOBAL-length host columns (world*n_loc doubles, indexed by global output slot / stratum), as in
 * genpf_step_with_noise; same kernels as genpf_shard_step_p2p. */
int32_t genpf_shard_step_p2p_with_noise(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                                        const double *obs_t, const double *aux_t, int32_t mh_iters,
                                        const double *uniforms, const double *U2, const double *Z2, const double *U3,
                                        const double *U1, const double *Z1);
/* closing offspring counts of the last p2p step + the exchange error word (1 = a peer timed out); synchronises */
int32_t genpf_shard_oend(genpf_filter_t pf, long long *oend_all_host, int32_t *error);
/* global ESS / accumulated log_ml_est / validity as of the last genpf_shard_scan (synchronises) */
int32_t genpf_shard_stats(genpf_filter_t pf, double *ess, double *lml_est, int32_t *invalid_kind);

/* per-kernel device time (CUDA events on the launching stream) of every kernel launched between begin and
 * end; end writes "kernel<TAB>count<TAB>total_ms" lines into buf.  Used by bench.py's roofline. */
int32_t genpf_profile_begin(void);
int32_t genpf_profile_end(char *buf, int64_t buf_len);

/* raw device pointers for multi-GPU plumbing (CUDA IPC / NCCL are driven by the host language) */
int32_t genpf_filter_stream(genpf_filter_t pf, void **cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GENPF_H */
