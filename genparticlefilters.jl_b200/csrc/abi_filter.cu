// abi_filter.cu -- C ABI, device-resident path: particle state lives in HBM as struct-of-arrays columns
// and pf_initialize / pf_update! / pf_resample! / pf_rejuvenate! / mean / var / resizing all run as CUDA
// kernels on the filter's stream.  See include/genpf.h for the reference function each call replaces.
#include <cmath>
#include <cstring>
#include <memory>


#include <type_traits>

#include "filter_state.hpp"
#include "plugin.hpp"

namespace genpf {

// a model registered at run time (abi_plugin.cu): its kernels come from the plugin image, their parameter lists
// are those of the built-in instantiations
struct PluginTag {};
template <class Model>
using SigModel = typename std::conditional<std::is_same<Model, PluginTag>::value, ObjectMotion, Model>::type;
#define GENPF_PLUGIN_OR_BUILTIN(Model, slot, builtin_fn, fn_out)                       \
    do {                                                                               \
        if constexpr (std::is_same<Model, PluginTag>::value) {                         \
            PluginModel *_pm = plugin_model(pf->model);                                \
            if (!_pm) return fail(GENPF_ERR_INVALID_ARG, "unknown model");             \
            GENPF_TRY(plugin_kernel(_pm, (slot), &(fn_out)));                          \
        } else {                                                                       \
            (fn_out) = (const void *)(builtin_fn);                                     \
        }                                                                              \
    } while (0)

int32_t check_filter(genpf_filter_t pf) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    GENPF_CUDA_TRY(cudaSetDevice(pf->device));
    return GENPF_OK;
}

static int32_t set_obs(genpf_filter_t pf, const double *obs, const double *aux, const double **obs_dev,
                       double *obs_val) {
    const ModelInfo &mi = *model_info(pf->model);
    if (!obs) return fail(GENPF_ERR_INVALID_ARG, "obs is NULL");
    if (mi.naux > 0 && !aux) return fail(GENPF_ERR_INVALID_ARG, "aux is NULL but the model needs per-step scalars");
    for (int i = 0; i < mi.naux; ++i) pf->P.aux[i] = aux[i];
    if (pf->nf == 1) {
        *obs_dev = nullptr;
        *obs_val = obs[0];
    } else {
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->obs_dev, obs, (size_t)pf->nf * 8, cudaMemcpyHostToDevice, pf->stream));
        *obs_dev = pf->obs_dev;
        *obs_val = 0.0;
    }
    return GENPF_OK;
}

int32_t stage_noise(genpf_filter_t pf, int which, const double *host, const double **dev, int64_t count) {
    if (!host) {
        *dev = nullptr;
        return GENPF_OK;
    }
    const size_t bytes = (size_t)(count > 0 ? count : pf->n * pf->nf) * 8;
    GENPF_TRY(pf->noise_buf[which].ensure(bytes));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->noise_buf[which].p, host, bytes, cudaMemcpyHostToDevice, pf->stream));
    *dev = pf->noise_buf[which].as<double>();
    return GENPF_OK;
}

template <class Model, class Noise>
static int32_t launch_propagate(genpf_filter_t pf, bool init, int64_t t, const double *obs_dev, double obs_val,
                                Noise noise, Strata strata) {
    const int64_t tpf = ceil_div(pf->n, kTile);
    const dim3 grid((unsigned)tpf, (unsigned)pf->nf);
    const void *fn = nullptr;
    constexpr int slot = kPlugProp + NoiseIndex<Noise>::value * 2;
    if (init) {
        GENPF_PLUGIN_OR_BUILTIN(Model, slot + 1, (&k_propagate<SigModel<Model>, Noise, true>), fn);
        return launch_typed("k_propagate", &k_propagate<SigModel<Model>, Noise, true>, fn, grid, dim3(kStateThreads), pf->stream,
                            pf->P, t, pf->slice(0), pf->slice(1), pf->lw, obs_dev, obs_val, pf->n, tpf, noise,
                            pf->sc.partials(0), pf->ew, strata);
    }
    GENPF_PLUGIN_OR_BUILTIN(Model, slot, (&k_propagate<SigModel<Model>, Noise, false>), fn);
    return launch_typed("k_propagate", &k_propagate<SigModel<Model>, Noise, false>, fn, grid, dim3(kStateThreads), pf->stream, pf->P,
                        t, pf->slice(t - 1), pf->slice(t), pf->lw, obs_dev, obs_val, pf->n, tpf, noise, pf->sc.partials(0),
                        pf->ew, strata);
}
template <class Model>
static int32_t propagate_model(genpf_filter_t pf, bool init, int64_t t, const double *obs_dev, double obs_val,
                               const double *U, const double *Z, Strata strata) {
    if (U || Z) {
        NoiseCols nz{U, Z, nullptr, nullptr, nullptr};
        return launch_propagate<Model, NoiseCols>(pf, init, t, obs_dev, obs_val, nz, strata);
    }
    if (pf->flags & GENPF_NOISE_PHILOX53) {
        NoisePhilox53 nz{pf->seed, (uint64_t)t, pf->rng_offset};
        return launch_propagate<Model, NoisePhilox53>(pf, init, t, obs_dev, obs_val, nz, strata);
    }
    NoiseLean nz{pf->seed, (uint64_t)t, pf->rng_offset};
    return launch_propagate<Model, NoiseLean>(pf, init, t, obs_dev, obs_val, nz, strata);
}

// freeze the slice that is about to leave the 2-slot window (GENPF_KEEP_HISTORY)
static int32_t archive_slice(genpf_filter_t pf, int64_t tau) {
    if (!(pf->flags & GENPF_KEEP_HISTORY) || tau < 1) return GENPF_OK;
    HistSlice h;
    h.tau = tau;
    h.n = pf->n;
    h.gen = pf->n_resamples;
    h.c = pf->slice(tau);  // steal the columns, give the window fresh ones
    GENPF_TRY(pf->alloc_cols(pf->win[pf->buf][tau & 1], pf->n * pf->nf));
    pf->hist.push_back(h);
    return GENPF_OK;
}

struct StrataHost {
    int32_t field;
    const double *values;
    int32_t K, layout;
};
static int32_t do_propagate(genpf_filter_t pf, bool init, int64_t t, const double *obs, const double *aux,
                            const double *U, const double *Z, const StrataHost *sh = nullptr, int mode = kPropPrior) {
    GENPF_TRY(check_filter(pf));
    if (mode != kPropPrior) {
        const int need = mode == kPropProposal ? 1 : 2;
        if (!(model_info(pf->model)->caps & need))
            return fail(GENPF_ERR_UNSUPPORTED, mode == kPropProposal
                                                   ? "this model defines no custom proposal (propose / proposal_logpdf / transition_logpdf)"
                                                   : "this model defines no translator (translate)");
        if ((U == nullptr) != (Z == nullptr)) return fail(GENPF_ERR_INVALID_ARG, "give both noise columns or neither");
    }
    if (init) {
        if (t != 1) return fail(GENPF_ERR_INVALID_ARG, "initialize must create time step 1");
    } else {
        if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
        if (t != pf->t_cur + 1) return fail(GENPF_ERR_INVALID_ARG, "pf_update! must advance to t_cur + 1");
    }
    const double *obs_dev, *dU, *dZ;
    double obs_val;
    GENPF_TRY(set_obs(pf, obs, aux, &obs_dev, &obs_val));
    GENPF_TRY(stage_noise(pf, 3, U, &dU, 0));
    GENPF_TRY(stage_noise(pf, 4, Z, &dZ, 0));
    if (init) {
        GENPF_CUDA_TRY(cudaMemsetAsync(pf->lml, 0, (size_t)pf->nf * 8, pf->stream));
        pf->n_resamples = 0;
        GENPF_LAUNCH(k_iota32, grid_1d(pf->n * pf->nf), 256, pf->stream, pf->parents, pf->n, pf->n * pf->nf);
    } else {
        GENPF_TRY(archive_slice(pf, t - 2));
    }
    Strata strata{};
    strata.mode = mode;
    if (sh) {
        if (!sh->values || sh->K < 1 || sh->K > pf->n) return fail(GENPF_ERR_INVALID_ARG, "bad strata (need 1 <= K <= n_particles)");
        if (sh->field < 0 || sh->field >= pf->NF + pf->NB) return fail(GENPF_ERR_INVALID_ARG, "strata field out of range");
        GENPF_TRY(pf->strata_buf.ensure((size_t)sh->K * 8));
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->strata_buf.p, sh->values, (size_t)sh->K * 8, cudaMemcpyHostToDevice, pf->stream));
        strata = Strata{pf->strata_buf.as<double>(), sh->K, sh->field, sh->layout == GENPF_LAYOUT_INTERLEAVED ? 1 : 0, pf->seed,
                        make_stream(kPurposeStrata, (uint64_t)t), kPropPrior};
    }
    int32_t st;
    switch (pf->model) {
        case kModelObjectMotion: st = propagate_model<ObjectMotion>(pf, init, t, obs_dev, obs_val, dU, dZ, strata); break;
        case kModelLinGauss1D: st = propagate_model<LinGauss1D>(pf, init, t, obs_dev, obs_val, dU, dZ, strata); break;
        default: st = propagate_model<PluginTag>(pf, init, t, obs_dev, obs_val, dU, dZ, strata); break;
    }
    GENPF_TRY(st);
    pf->t_cur = t;
    pf->part_valid = true;
    return GENPF_OK;
}

// make sure sc.partials(0) describes the current lw, then finalize into st(0)
int32_t ensure_stats(genpf_filter_t pf, double *tile_off, double ess_frac, double *lml_accum) {
    if (!pf->part_valid) {
        LwSrc src{pf->lw, 1.0};
        GENPF_TRY(launch_reduce(pf->stream, src, pf->n, pf->nf, pf->sc.partials(0), pf->ew));
        pf->part_valid = true;
    }
    return launch_finalize(pf->stream, pf->sc, pf->sc.partials(0), pf->n, pf->nf, pf->sc.st(0, pf->nf), tile_off, ess_frac,
                           lml_accum);
}

int32_t read_stats(genpf_filter_t pf, int which) {
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_stats, pf->sc.st(which, pf->nf), sizeof(Stats) * (size_t)pf->nf,
                                   cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}

template <class Model, class Noise>
static int32_t launch_mh(genpf_filter_t pf, int64_t tau, int iter, const double *obs_dev, double obs_val, Noise noise,
                         bool reweight, bool gated, int use_proposal = 0) {
    const int64_t tpf = ceil_div(pf->n, kTile);
    const dim3 grid((unsigned)tpf, (unsigned)pf->nf);
    const void *fn = nullptr;
    constexpr int slot = kPlugMh + NoiseIndex<Noise>::value * 2;
    const Stats *gst = gated ? (const Stats *)pf->sc.st(0, pf->nf) : (const Stats *)nullptr;
    if (reweight) {
        GENPF_PLUGIN_OR_BUILTIN(Model, slot + 1, (&k_mh<SigModel<Model>, Noise, true>), fn);
        return launch_typed("k_mh", &k_mh<SigModel<Model>, Noise, true>, fn, grid, dim3(kStateThreads), pf->stream, pf->P, tau, iter,
                            tau == 1 ? 1 : 0, pf->slice(tau - 1), pf->slice(tau), obs_dev, obs_val, pf->n, tpf, noise,
                            (uint8_t *)nullptr, (unsigned long long *)nullptr, pf->lw, gst, use_proposal);
    }
    GENPF_PLUGIN_OR_BUILTIN(Model, slot, (&k_mh<SigModel<Model>, Noise, false>), fn);
    return launch_typed("k_mh", &k_mh<SigModel<Model>, Noise, false>, fn, grid, dim3(kStateThreads), pf->stream, pf->P, tau, iter,
                        tau == 1 ? 1 : 0, pf->slice(tau - 1), pf->slice(tau), obs_dev, obs_val, pf->n, tpf, noise, pf->accepts,
                        pf->n_accept, (double *)nullptr, gst, 0);
}
template <class Model>
static int32_t mh_model(genpf_filter_t pf, int64_t tau, int iter, const double *obs_dev, double obs_val,
                        const double *U2, const double *Z2, const double *U3, bool reweight, bool gated,
                        int use_proposal = 0) {
    if (U2 || Z2 || U3) {
        NoiseCols nz{U2, Z2, U3, nullptr, nullptr};
        return launch_mh<Model, NoiseCols>(pf, tau, iter, obs_dev, obs_val, nz, reweight, gated, use_proposal);
    }
    // the mh move on slice tau belongs to README iteration s = tau + 1 (it runs right before pf_update!(s))
    if (pf->flags & GENPF_NOISE_PHILOX53) {
        NoisePhilox53 nz{pf->seed, (uint64_t)tau + 1, pf->rng_offset};
        return launch_mh<Model, NoisePhilox53>(pf, tau, iter, obs_dev, obs_val, nz, reweight, gated, use_proposal);
    }
    NoiseLean nz{pf->seed, (uint64_t)tau + 1, pf->rng_offset};
    return launch_mh<Model, NoiseLean>(pf, tau, iter, obs_dev, obs_val, nz, reweight, gated, use_proposal);
}

static int32_t do_mh(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux, int32_t n_iters,
                     const double *U2, const double *Z2, const double *U3, int64_t *n_accept, bool reweight = false,
                     bool gated = false, int use_proposal = 0) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    if (tau != pf->t_cur) return fail(GENPF_ERR_UNSUPPORTED, "mh rejuvenation is implemented for tau == newest time step");
    if (n_iters < 1) return GENPF_OK;
    const double *obs_dev, *dU2, *dZ2, *dU3;
    double obs_val;
    GENPF_TRY(set_obs(pf, obs, aux, &obs_dev, &obs_val));
    GENPF_TRY(stage_noise(pf, 0, U2, &dU2, 0));
    GENPF_TRY(stage_noise(pf, 1, Z2, &dZ2, 0));
    GENPF_TRY(stage_noise(pf, 2, U3, &dU3, 0));
    if (reweight) pf->part_valid = false;  // log-weights change
    GENPF_CUDA_TRY(cudaMemsetAsync(pf->n_accept, 0, (size_t)pf->nf * 8, pf->stream));
    for (int it = 0; it < n_iters; ++it) {
        int32_t st;
        switch (pf->model) {
            case kModelObjectMotion: st = mh_model<ObjectMotion>(pf, tau, it, obs_dev, obs_val, dU2, dZ2, dU3, reweight, gated, use_proposal); break;
            case kModelLinGauss1D: st = mh_model<LinGauss1D>(pf, tau, it, obs_dev, obs_val, dU2, dZ2, dU3, reweight, gated, use_proposal); break;
            default: st = mh_model<PluginTag>(pf, tau, it, obs_dev, obs_val, dU2, dZ2, dU3, reweight, gated, use_proposal); break;
        }
        GENPF_TRY(st);
    }
    if (n_accept) {
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, pf->n_accept, (size_t)pf->nf * 8, cudaMemcpyDeviceToHost, pf->stream));
        GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
        for (int64_t f = 0; f < pf->nf; ++f) n_accept[f] = (int64_t)reinterpret_cast<unsigned long long *>(pf->h_pinned)[f];
    }
    return GENPF_OK;
}

static GatherCols window_gather_cols(genpf_filter_t pf) {
    GatherCols g;
    memset(&g, 0, sizeof(g));
    int kf = 0, kb = 0;
    for (int sl = 0; sl < 2; ++sl) {
        for (int i = 0; i < pf->NF; ++i) {
            g.sf[kf] = pf->win[pf->buf][sl].f[i];
            g.df[kf] = pf->win[pf->buf ^ 1][sl].f[i];
            ++kf;
        }
        for (int i = 0; i < pf->NB; ++i) {
            g.sb[kb] = pf->win[pf->buf][sl].b[i];
            g.db[kb] = pf->win[pf->buf ^ 1][sl].b[i];
            ++kb;
        }
    }
    g.nf = kf;
    g.nb = kb;
    return g;
}

// grow / shrink the population buffers of the *other* buffer set + lw_alt + parents to n_out.  The old buffers
// are parked in pf->undo until the resize commits: a call that fails validation afterwards (check=true with
// invalid weights, the optimal resize's @assert) rolls back and leaves the filter exactly as it was.
static int32_t resize_target(genpf_filter_t pf, int64_t n_out) {
    if (n_out == pf->n) return GENPF_OK;
    const int64_t total = n_out * pf->nf;
    ResizeUndo &u = pf->undo;
    u.active = true;
    for (int sl = 0; sl < 2; ++sl) {
        u.win[sl] = pf->win[pf->buf ^ 1][sl];
        memset(&pf->win[pf->buf ^ 1][sl], 0, sizeof(Cols));
    }
    u.lw_alt = pf->lw_alt;
    u.parents = pf->parents;
    int32_t st = GENPF_OK;
    for (int sl = 0; sl < 2 && st == GENPF_OK; ++sl) st = pf->alloc_cols(pf->win[pf->buf ^ 1][sl], total);
    pf->lw_alt = nullptr;
    pf->parents = nullptr;
    if (st == GENPF_OK) st = pf->dalloc(&pf->lw_alt, (size_t)total);
    if (st == GENPF_OK) st = pf->dalloc(&pf->parents, (size_t)total);
    if (st == GENPF_OK) st = pf->sc.ensure(n_out > pf->n ? n_out : pf->n, pf->nf);
    if (st != GENPF_OK) resize_rollback(pf);
    return st;
}
void resize_rollback(genpf_filter_t pf) {
    ResizeUndo &u = pf->undo;
    if (!u.active) return;
    cudaStreamSynchronize(pf->stream);
    for (int sl = 0; sl < 2; ++sl) {
        pf->free_cols(pf->win[pf->buf ^ 1][sl]);
        pf->win[pf->buf ^ 1][sl] = u.win[sl];
    }
    pf->dfree(pf->lw_alt);
    pf->dfree(pf->parents);
    pf->lw_alt = u.lw_alt;
    pf->parents = u.parents;
    u.active = false;
}
// the resize went through: the parked buffers (old spare window, old lw_alt, old parents) are released
static void resize_commit(genpf_filter_t pf) {
    ResizeUndo &u = pf->undo;
    if (!u.active) return;
    for (int sl = 0; sl < 2; ++sl) pf->free_cols(u.win[sl]);
    pf->dfree(u.lw_alt);
    pf->dfree(u.parents);
    u.active = false;
}
#define GENPF_TRY_RB(pf, expr)             \
    do {                                   \
        int32_t _s = (expr);               \
        if (_s != GENPF_OK) {              \
            ::genpf::resize_rollback(pf);  \
            return _s;                     \
        }                                  \
    } while (0)
// after the swap: make the (now spare) old buffers match the new size (update_refs!, resize.jl:441-449)
static int32_t resize_spare(genpf_filter_t pf, int64_t n_new) {
    const int64_t total = n_new * pf->nf;
    for (int sl = 0; sl < 2; ++sl) {
        pf->free_cols(pf->win[pf->buf ^ 1][sl]);
        GENPF_TRY(pf->alloc_cols(pf->win[pf->buf ^ 1][sl], total));
    }
    pf->dfree(pf->lw_alt);
    GENPF_TRY(pf->dalloc(&pf->lw_alt, (size_t)total));
    pf->dfree(pf->accepts);
    GENPF_TRY(pf->dalloc(&pf->accepts, (size_t)total));
    pf->dfree(pf->ew);
    GENPF_TRY(pf->dalloc(&pf->ew, (size_t)total));
    return GENPF_OK;
}

int32_t log_parents(genpf_filter_t pf, int64_t n_prev, int64_t n_cur) {
    pf->n_resamples += 1;
    if (!(pf->flags & GENPF_KEEP_HISTORY)) return GENPF_OK;
    ParentLog pl;
    pl.n_prev = n_prev;
    pl.n_cur = n_cur;
    GENPF_TRY(pf->dalloc(&pl.parents, (size_t)(n_cur * pf->nf)));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pl.parents, pf->parents, (size_t)(n_cur * pf->nf) * 4, cudaMemcpyDeviceToDevice, pf->stream));
    pf->plog.push_back(pl);
    return GENPF_OK;
}

// pf_resample! / pf_resize! on device state.  gate: only filters whose device-side predicate fired are
// resampled (the others are copied through unchanged so the buffer swap stays unconditional).
static int32_t do_resample(genpf_filter_t pf, int32_t method, int32_t prio_kind, double prio_param,
                           const double *prio_column_host, int64_t n_out, uint32_t flags, const double *uniforms_host,
                           int32_t *invalid_kinds, double ess_frac) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    if (method != GENPF_MULTINOMIAL && method != GENPF_RESIDUAL && method != GENPF_STRATIFIED)
        return fail(GENPF_ERR_UNKNOWN_METHOD, "Resampling method not recognized.");
    if (n_out <= 0) n_out = pf->n;
    if (n_out != pf->n && method == GENPF_STRATIFIED)
        return fail(GENPF_ERR_INVALID_ARG, "stratified resampling cannot resize (resize.jl:16-27)");
    const int gate = ess_frac >= 0.0 ? 1 : 0;
    const int64_t n = pf->n, nf = pf->nf;
    cudaStream_t s = pf->stream;
    Scratch &sc = pf->sc;
    const bool substate = flags & GENPF_SUBSTATE;
    const bool want_host_check = (flags & GENPF_CHECK) || invalid_kinds;

    // selection source
    LwSrc lw_src{pf->lw, 1.0}, sel = lw_src;
    bool has_prio = false;
    if (prio_kind == GENPF_PRIO_SCALE) {
        sel = LwSrc{pf->lw, prio_param};
        has_prio = true;
    } else if (prio_kind == GENPF_PRIO_COLUMN) {
        if (!prio_column_host) return fail(GENPF_ERR_INVALID_ARG, "prio_column is NULL");
        GENPF_TRY(pf->prio_buf.ensure((size_t)(n * nf) * 8));
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->prio_buf.p, prio_column_host, (size_t)(n * nf) * 8, cudaMemcpyHostToDevice, s));
        sel = LwSrc{pf->prio_buf.as<double>(), 1.0};
        has_prio = true;
    } else if (prio_kind != GENPF_PRIO_NONE) {
        return fail(GENPF_ERR_INVALID_ARG, "unknown priority kind");
    }
    if (has_prio && gate) return fail(GENPF_ERR_UNSUPPORTED, "gated resample with priorities is not supported");
    Stats *st_lw = sc.st(0, nf), *st_sel = has_prio ? sc.st(1, nf) : st_lw, *st_d = sc.st(2, nf);
    double *lml_now = (substate || want_host_check) ? nullptr : pf->lml;
    GENPF_TRY(ensure_stats(pf, has_prio ? nullptr : sc.tile_off.as<double>(), ess_frac, lml_now));
    if (has_prio) {
        GENPF_TRY(launch_reduce(s, sel, n, nf, sc.partials(1)));
        GENPF_TRY(launch_finalize(s, sc, sc.partials(1), n, nf, st_sel, sc.tile_off.as<double>(), -1.0, nullptr));
    }
    if (want_host_check) {
        GENPF_TRY(read_stats(pf, has_prio ? 1 : 0));
        bool any_invalid = false, any_nan = false;
        for (int64_t f = 0; f < nf; ++f) {
            int k = pf->h_stats[f].invalid_kind;
            if (invalid_kinds) invalid_kinds[f] = k;
            any_invalid |= (k != GENPF_VALID);
            any_nan |= (k == GENPF_INV_NAN_INPUT || k == GENPF_INV_NAN_TOTAL);
        }
        if ((flags & GENPF_CHECK) && any_invalid) return fail(GENPF_ERR_INVALID_WEIGHTS, "Invalid weights.");
        if (!substate) {  // update_lml_est! (resample.jl:178-182), after the check like the reference
            GENPF_TRY(launch_finalize(s, sc, sc.partials(0), n, nf, st_lw, nullptr, ess_frac, pf->lml));
        }
        if (any_nan && nf == 1) return GENPF_OK;
    }
    // uniforms
    const double *d_u = nullptr;
    if (uniforms_host) {
        GENPF_TRY(pf->uni_buf.ensure((size_t)(n_out * nf) * 8));
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->uni_buf.p, uniforms_host, (size_t)(n_out * nf) * 8, cudaMemcpyHostToDevice, s));
        d_u = pf->uni_buf.as<double>();
    }
    // the selection draws are indexed by OUTPUT slot (filter * n_out + j): a batch shard offsets by its position in
    // output slots, which differs from rng_offset (input slots) when the call resizes
    const int64_t uni_off = pf->first_filter ? pf->first_filter * n_out : pf->rng_offset;
    UniSrc uni{d_u, pf->seed, make_stream(kPurposeResample, (uint64_t)pf->n_resamples + 1), uni_off};
    // every validation that can refuse the call is behind us: only now are the target buffers resized
    GENPF_TRY(resize_target(pf, n_out));
    GENPF_TRY_RB(pf, select_ancestors<int32_t>(s, sc, method, sel, n, n_out, nf, st_sel, uni, flags, pf->parents, 0, gate,
                                               has_prio ? nullptr : (const double *)pf->ew));
    // gather the window into the other buffer; no-priority reweight fused in
    const int64_t tpf_out = ceil_div(n_out, kTile);
    GatherCols g = window_gather_cols(pf);
    GENPF_LAUNCH(k_gather, dim3((unsigned)tpf_out, (unsigned)nf), kThreads, s, g, pf->parents, n, n_out, tpf_out, (const double *)pf->lw,
                 has_prio ? (double *)nullptr : pf->lw_alt, (const Stats *)st_lw, gate, substate ? 1 : 0);
    if (has_prio) {
        GENPF_LAUNCH((k_prio_ratio<int32_t>), dim3(grid_1d(n_out), (unsigned)nf), 256, s, pf->lw, sel, pf->parents,
                     (int64_t)0, n, n_out, pf->lw_alt);
        LwSrc dsrc{pf->lw_alt, 1.0};
        GENPF_TRY(launch_reduce(s, dsrc, n_out, nf, sc.partials(2)));
        GENPF_TRY(launch_finalize(s, sc, sc.partials(2), n_out, nf, st_d, nullptr, -1.0, nullptr));
        GENPF_LAUNCH(k_prio_shift, dim3(grid_1d(n_out), (unsigned)nf), 256, s, pf->lw_alt, n_out, st_d, st_lw, substate ? 1 : 0);
    }
    // update_refs!: swap (utils.jl:10-15)
    pf->buf ^= 1;
    std::swap(pf->lw, pf->lw_alt);
    pf->part_valid = false;
    resize_commit(pf);
    GENPF_TRY(log_parents(pf, n, n_out));
    if (n_out != n) {
        pf->n = n_out;
        if (pf->first_filter) pf->rng_offset = pf->first_filter * n_out;  // Philox counters stay global batch slots
        GENPF_TRY(resize_spare(pf, n_out));
    }
    return GENPF_OK;
}


// ---- fused README iteration (stratified, resample taken, Philox noise): finalize -> scan -> k_step_fused
template <class Model, class Noise>
static int32_t launch_step_fused(genpf_filter_t pf, const StepArgs &a, Noise noise, int gate) {
    const int64_t tpf = ceil_div(pf->n, kTile);
    const int64_t t = a.t;
    const Stats *gst = pf->sc.st(0, pf->nf);
    const double *lw_src = pf->lw;
    const dim3 grid((unsigned)tpf, (unsigned)pf->nf);
    const void *fn = nullptr;
    constexpr int slot = kPlugFused + NoiseIndex<Noise>::value * 3;
#define GENPF_FUSED_ARGS                                                                                                  \
    a, (const int32_t *)pf->sc.O.as<int32_t>(), (const int32_t *)pf->sc.tile_last.as<int32_t>(), pf->slice(t - 2),         \
        pf->slice(t - 1), pf->slice_alt(t - 1), pf->slice_alt(t), pf->parents, pf->lw_alt, pf->n, tpf, noise,             \
        (uint8_t *)nullptr, (unsigned long long *)nullptr, pf->sc.partials(0), pf->ew, gst, gate, lw_src
    if (a.mh_iters == 1) {
        GENPF_PLUGIN_OR_BUILTIN(Model, slot + 0, (&k_step_fused<SigModel<Model>, Noise, int32_t, 1>), fn);
        return launch_typed("k_step_fused", &k_step_fused<SigModel<Model>, Noise, int32_t, 1>, fn, grid, dim3(kStateThreads),
                            pf->stream, GENPF_FUSED_ARGS);
    }
    if (a.mh_iters == 0) {
        GENPF_PLUGIN_OR_BUILTIN(Model, slot + 1, (&k_step_fused<SigModel<Model>, Noise, int32_t, 0>), fn);
        return launch_typed("k_step_fused", &k_step_fused<SigModel<Model>, Noise, int32_t, 0>, fn, grid, dim3(kStateThreads),
                            pf->stream, GENPF_FUSED_ARGS);
    }
    GENPF_PLUGIN_OR_BUILTIN(Model, slot + 2, (&k_step_fused<SigModel<Model>, Noise, int32_t, -1>), fn);
    return launch_typed("k_step_fused", &k_step_fused<SigModel<Model>, Noise, int32_t, -1>, fn, grid, dim3(kStateThreads),
                        pf->stream, GENPF_FUSED_ARGS);
#undef GENPF_FUSED_ARGS
}
template <class Model>
static int32_t step_fused_model(genpf_filter_t pf, const StepArgs &a, const NoiseCols *cols, int gate) {
    if (cols) return launch_step_fused<Model, NoiseCols>(pf, a, *cols, gate);
    if (pf->flags & GENPF_NOISE_PHILOX53) {
        NoisePhilox53 nz{pf->seed, (uint64_t)a.t, pf->rng_offset};
        return launch_step_fused<Model, NoisePhilox53>(pf, a, nz, gate);
    }
    NoiseLean nz{pf->seed, (uint64_t)a.t, pf->rng_offset};
    return launch_step_fused<Model, NoiseLean>(pf, a, nz, gate);
}

static int32_t do_step_fused(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                             const double *obs_t, const double *aux_t, int32_t mh_iters, bool finalized,
                             const double *d_uniforms = nullptr, const NoiseCols *cols = nullptr, int gate = 0,
                             double ess_frac = -1.0, const double *d_obs_prev = nullptr, const double *d_obs_t = nullptr) {
    const ModelInfo &mi = *model_info(pf->model);
    const int64_t n = pf->n, nf = pf->nf;
    cudaStream_t s = pf->stream;
    Scratch &sc = pf->sc;
    if (!obs_prev || !obs_t) return fail(GENPF_ERR_INVALID_ARG, "obs is NULL");
    if (mi.naux > 0 && (!aux_prev || !aux_t)) return fail(GENPF_ERR_INVALID_ARG, "aux is NULL");
    StepArgs a;
    a.P_prev = pf->P;
    a.P_t = pf->P;
    for (int i = 0; i < mi.naux; ++i) {
        a.P_prev.aux[i] = aux_prev[i];
        a.P_t.aux[i] = aux_t[i];
    }
    a.t = t;
    a.mh_iters = mh_iters;
    if (nf == 1) {
        a.obs_prev_dev = a.obs_t_dev = nullptr;
        a.obs_prev = obs_prev[0];
        a.obs_t = obs_t[0];
    } else if (d_obs_prev && d_obs_t) {  // rows already staged on the device (genpf_run_steps)
        a.obs_prev_dev = d_obs_prev;
        a.obs_t_dev = d_obs_t;
        a.obs_prev = a.obs_t = 0.0;
    } else {
        GENPF_TRY(pf->step_obs.ensure((size_t)nf * 16));
        double *d = pf->step_obs.as<double>();
        GENPF_CUDA_TRY(cudaMemcpyAsync(d, obs_prev, (size_t)nf * 8, cudaMemcpyHostToDevice, s));
        GENPF_CUDA_TRY(cudaMemcpyAsync(d + nf, obs_t, (size_t)nf * 8, cudaMemcpyHostToDevice, s));
        a.obs_prev_dev = d;
        a.obs_t_dev = d + nf;
        a.obs_prev = a.obs_t = 0.0;
    }
    const int64_t tpf = ceil_div(n, kTile);
    GENPF_TRY(sc.O.ensure((size_t)(n * nf) * 4));
    GENPF_TRY(sc.tile_last.ensure((size_t)(tpf * nf) * 4));
    if (!finalized) GENPF_TRY(ensure_stats(pf, sc.tile_off.as<double>(), ess_frac, pf->lml));
    UniSrc uni{d_uniforms, pf->seed, make_stream(kPurposeResample, (uint64_t)pf->n_resamples + 1), pf->rng_offset};
    StratArgs strat = make_strat(uni, n);
    LwSrc lw_src{pf->lw, 1.0};
    GENPF_TRY(launch_scan_counts<int32_t>(s, lw_src, n, tpf, nf, sc.st(0, nf), sc.tile_off.as<double>(), sc.O.as<int32_t>(),
                                          sc.tile_last.as<int32_t>(), strat, gate, nullptr, 0, sc.chunk_info_ptr(n), pf->ew,
                                          sc.tile_scale.as<double>()));
    int32_t st;
    switch (pf->model) {
        case kModelObjectMotion: st = step_fused_model<ObjectMotion>(pf, a, cols, gate); break;
        case kModelLinGauss1D: st = step_fused_model<LinGauss1D>(pf, a, cols, gate); break;
        default: st = step_fused_model<PluginTag>(pf, a, cols, gate); break;
    }
    GENPF_TRY(st);
    pf->buf ^= 1;
    std::swap(pf->lw, pf->lw_alt);
    GENPF_TRY(log_parents(pf, n, n));
    if ((pf->flags & GENPF_KEEP_HISTORY) && t - 2 >= 1) {
        // slice t-2 left the window: it stays in the (now spare) old buffer, in pre-resample order
        HistSlice h;
        h.tau = t - 2;
        h.n = n;
        h.gen = pf->n_resamples - 1;
        h.c = pf->win[pf->buf ^ 1][(t - 2) & 1];
        GENPF_TRY(pf->alloc_cols(pf->win[pf->buf ^ 1][(t - 2) & 1], n * nf));
        pf->hist.push_back(h);
    }
    pf->t_cur = t;
    pf->part_valid = true;
    return GENPF_OK;
}

}  // namespace genpf

extern "C" {

int32_t genpf_model_builtin(const char *name, int32_t *model_id) {
    if (!name || !model_id) return fail(GENPF_ERR_INVALID_ARG, "genpf_model_builtin: NULL argument");
    const int32_t id = find_model(name);  // built-in models and plugins registered at run time
    if (id < 0) return fail(GENPF_ERR_INVALID_ARG, std::string("unknown model: ") + name);
    *model_id = id;
    return GENPF_OK;
}

int32_t genpf_model_info(int32_t model_id, int32_t *n_f64_fields, int32_t *n_u8_fields, int32_t *n_params,
                         int32_t *n_aux) {
    const ModelInfo *mi = model_info(model_id);
    if (!mi) return fail(GENPF_ERR_INVALID_ARG, "unknown model id");
    if (n_f64_fields) *n_f64_fields = mi->nf;
    if (n_u8_fields) *n_u8_fields = mi->nb;
    if (n_params) *n_params = mi->np;
    if (n_aux) *n_aux = mi->naux;
    return GENPF_OK;
}
int32_t genpf_model_caps(int32_t model_id, int32_t *caps) {
    const ModelInfo *mi = model_info(model_id);
    if (!mi || !caps) return fail(GENPF_ERR_INVALID_ARG, "unknown model id");
    *caps = mi->caps;
    return GENPF_OK;
}

int32_t genpf_filter_create(int32_t model_id, const double *params, int32_t n_params, int64_t n_particles,
                            int64_t n_filters, uint64_t seed, uint32_t flags, genpf_filter_t *out) {
    if (!out) return fail(GENPF_ERR_INVALID_ARG, "out is NULL");
    if (!model_info(model_id)) return fail(GENPF_ERR_INVALID_ARG, "unknown model id");
    if (n_particles <= 0 || n_filters <= 0) return fail(GENPF_ERR_INVALID_ARG, "n_particles and n_filters must be > 0");
    if (n_particles >= 0x7FFFFFF0ll) return fail(GENPF_ERR_UNSUPPORTED, "n_particles per filter must be < 2^31");
    if (n_filters > 65535) return fail(GENPF_ERR_UNSUPPORTED, "n_filters must be <= 65535 (grid.y)");
    const ModelInfo &mi = *model_info(model_id);
    if (params && n_params != mi.np) return fail(GENPF_ERR_INVALID_ARG, "wrong number of model parameters");
    std::unique_ptr<genpf_filter_s> pf(new genpf_filter_s());
    pf->model = model_id;
    pf->NF = mi.nf;
    pf->NB = mi.nb;
    pf->n = n_particles;
    pf->nf = n_filters;
    pf->seed = seed;
    pf->flags = flags;
    GENPF_CUDA_TRY(cudaGetDevice(&pf->device));
    memset(&pf->P, 0, sizeof(pf->P));
    if (model_id == kModelObjectMotion) {
        const double def[4] = {0.75, 0.25, 0.01, 0.25};  // README.md:47-50
        for (int i = 0; i < 4; ++i) pf->P.v[i] = params ? params[i] : def[i];
        pf->P.v[4] = log(pf->P.v[3]);
        pf->P.v[5] = 1.0 / pf->P.v[3];
    } else if (model_id >= kPluginIdBase) {
        for (int i = 0; i < mi.np; ++i) pf->P.v[i] = params ? params[i] : 0.0;  // a plugin derives what it needs itself
    } else {
        const double def[5] = {0.9, 1.0, 1.0, 0.0, 1.0};
        for (int i = 0; i < 5; ++i) pf->P.v[i] = params ? params[i] : def[i];
        pf->P.v[5] = log(pf->P.v[2]);
        pf->P.v[6] = sqrt(pf->P.v[0] * pf->P.v[0] * pf->P.v[4] * pf->P.v[4] + pf->P.v[1] * pf->P.v[1]);
        pf->P.v[7] = 1.0 / pf->P.v[2];
    }
    GENPF_CUDA_TRY(cudaStreamCreateWithFlags(&pf->stream, cudaStreamNonBlocking));
    int32_t st = pf->alloc_population(n_particles);
    if (st == GENPF_OK) st = pf->dalloc(&pf->lml, (size_t)n_filters);
    if (st == GENPF_OK) st = pf->dalloc(&pf->obs_dev, (size_t)n_filters);
    if (st == GENPF_OK) st = pf->dalloc(&pf->n_accept, (size_t)n_filters);
    if (st != GENPF_OK) {
        genpf_filter_destroy(pf.release());
        return st;
    }
    const size_t pin = (size_t)(n_filters > 16 ? n_filters : 16);
    GENPF_CUDA_TRY(cudaMallocHost(&pf->h_pinned, pin * 8 * 4));
    GENPF_CUDA_TRY(cudaMallocHost(&pf->h_stats, pin * sizeof(Stats)));
    GENPF_CUDA_TRY(cudaMemsetAsync(pf->lml, 0, (size_t)n_filters * 8, pf->stream));
    *out = pf.release();
    return GENPF_OK;
}

// Batch sharding ("batches of independent filters shard with no communication at all"): this handle holds
// filters [first_filter, first_filter + n_filters) of a larger batch.  Every Philox counter of the library is a global
// particle slot (filter * n_particles + i), so the shard draws exactly what the same filters would draw inside one
// big batch: the result does not depend on how the batch is split over handles, processes or GPUs.
int32_t genpf_filter_set_first_filter(genpf_filter_t pf, int64_t first_filter) {
    GENPF_TRY(check_filter(pf));
    if (first_filter < 0) return fail(GENPF_ERR_INVALID_ARG, "first_filter must be >= 0");
    if (pf->shard) return fail(GENPF_ERR_STATE, "a particle-sharded filter has its own slot offset");
    if (pf->t_cur != 0) return fail(GENPF_ERR_STATE, "set the batch position before the filter is initialised");
    pf->first_filter = first_filter;
    pf->rng_offset = first_filter * pf->n;
    return GENPF_OK;
}

int32_t genpf_shard_detach(genpf_filter_t pf);
int32_t genpf_filter_destroy(genpf_filter_t pf) {
    if (!pf) return GENPF_OK;
    genpf_shard_detach(pf);
    cudaSetDevice(pf->device);
    if (pf->stream) cudaStreamSynchronize(pf->stream);
    for (auto &b : pf->live) cudaFree(b.p);
    for (auto &b : pf->cache) cudaFree(b.p);
    pf->sc.release();
    pf->cb.release();
    pf->ob.release();
    if (pf->h_opt_ctrl) cudaFreeHost(pf->h_opt_ctrl);
    if (pf->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)pf->graph_exec);
    pf->run_obs.release();
    pf->key_buf.release();
    for (DevBuf *b : {&pf->noise_buf[0], &pf->noise_buf[1], &pf->noise_buf[2], &pf->noise_buf[3], &pf->noise_buf[4], &pf->step_obs, &pf->strata_buf, &pf->uni_buf, &pf->tmp_col, &pf->tmp_idx, &pf->prio_buf})
        b->release();
    if (pf->h_pinned) cudaFreeHost(pf->h_pinned);
    if (pf->h_stats) cudaFreeHost(pf->h_stats);
    if (pf->stream) cudaStreamDestroy(pf->stream);
    delete pf;
    return GENPF_OK;
}

int32_t genpf_filter_size(genpf_filter_t pf, int64_t *n_particles, int64_t *n_filters) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    if (n_particles) *n_particles = pf->n;
    if (n_filters) *n_filters = pf->nf;
    return GENPF_OK;
}

int32_t genpf_filter_sync(genpf_filter_t pf) {
    GENPF_TRY(check_filter(pf));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}
int32_t genpf_filter_stream(genpf_filter_t pf, void **cuda_stream) {
    if (!pf || !cuda_stream) return fail(GENPF_ERR_INVALID_ARG, "NULL argument");
    *cuda_stream = (void *)pf->stream;
    return GENPF_OK;
}

int32_t genpf_initialize(genpf_filter_t pf, const double *obs, const double *aux) {
    return do_propagate(pf, true, 1, obs, aux, nullptr, nullptr);
}
int32_t genpf_initialize_with_noise(genpf_filter_t pf, const double *obs, const double *aux, const double *U,
                                    const double *Z) {
    if (!U || !Z) return fail(GENPF_ERR_INVALID_ARG, "noise columns are NULL");
    return do_propagate(pf, true, 1, obs, aux, U, Z);
}
int32_t genpf_initialize_stratified(genpf_filter_t pf, const double *obs, const double *aux, int32_t field,
                                    const double *values, int32_t n_strata, int32_t layout, const double *U,
                                    const double *Z) {
    if ((U == nullptr) != (Z == nullptr)) return fail(GENPF_ERR_INVALID_ARG, "give both noise columns or neither");
    StrataHost sh{field, values, n_strata, layout};
    return do_propagate(pf, true, 1, obs, aux, U, Z, &sh);
}
int32_t genpf_update_stratified(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, int32_t field,
                                const double *values, int32_t n_strata, int32_t layout, const double *U, const double *Z) {
    if ((U == nullptr) != (Z == nullptr)) return fail(GENPF_ERR_INVALID_ARG, "give both noise columns or neither");
    StrataHost sh{field, values, n_strata, layout};
    return do_propagate(pf, false, t, obs, aux, U, Z, &sh);
}
int32_t genpf_update(genpf_filter_t pf, int64_t t, const double *obs, const double *aux) {
    return do_propagate(pf, false, t, obs, aux, nullptr, nullptr);
}
int32_t genpf_update_with_noise(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, const double *U,
                                const double *Z) {
    if (!U || !Z) return fail(GENPF_ERR_INVALID_ARG, "noise columns are NULL");
    return do_propagate(pf, false, t, obs, aux, U, Z);
}

// pf_initialize(model, args, obs, proposal, proposal_args, n) (initialize.jl:46-62) and
// pf_update!(state, new_args, argdiffs, obs, proposal, proposal_args) (update.jl:79-96) for plugins that define a
// proposal: x ~ q(. | prev, obs), log-weight (+)= log p(x | prev) + log p(obs | x) - log q(x).  U, Z: the proposal's
// draws as columns (parity mode) or both NULL (library Philox noise).
int32_t genpf_initialize_proposal(genpf_filter_t pf, const double *obs, const double *aux, const double *U, const double *Z) {
    return do_propagate(pf, true, 1, obs, aux, U, Z, nullptr, kPropProposal);
}
int32_t genpf_update_proposal(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, const double *U,
                              const double *Z) {
    return do_propagate(pf, false, t, obs, aux, U, Z, nullptr, kPropProposal);
}
// pf_update!(state, translator) (update.jl:35-44): the plugin's translate(current slice) returns the new slice and the
// log-weight increment, log_weights[i] += increment
int32_t genpf_update_translate(genpf_filter_t pf, int64_t t, const double *obs, const double *aux, const double *U,
                               const double *Z) {
    return do_propagate(pf, false, t, obs, aux, U, Z, nullptr, kPropTranslate);
}

int32_t genpf_ess_dev(genpf_filter_t pf, double *ess) {
    GENPF_TRY(check_filter(pf));
    if (!ess) return fail(GENPF_ERR_INVALID_ARG, "ess is NULL");
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    GENPF_TRY(ensure_stats(pf, nullptr, -1.0, nullptr));
    GENPF_TRY(read_stats(pf, 0));
    for (int64_t f = 0; f < pf->nf; ++f) ess[f] = pf->h_stats[f].invalid_kind == GENPF_VALID ? pf->h_stats[f].ess : NAN;
    return GENPF_OK;
}

// log_ml_estimate(state) = log_ml_est + logsumexp(lw) - log(n)  (Gen; utils.jl:174-178)
int32_t genpf_lml_dev(genpf_filter_t pf, double *lml) {
    GENPF_TRY(check_filter(pf));
    if (!lml) return fail(GENPF_ERR_INVALID_ARG, "lml is NULL");
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    GENPF_TRY(ensure_stats(pf, nullptr, -1.0, nullptr));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, pf->lml, (size_t)pf->nf * 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_TRY(read_stats(pf, 0));
    for (int64_t f = 0; f < pf->nf; ++f) lml[f] = pf->h_pinned[f] + pf->h_stats[f].lse - log((double)pf->n);
    return GENPF_OK;
}

int32_t genpf_resample_dev(genpf_filter_t pf, int32_t method, int32_t prio_kind, double prio_param,
                           const double *prio_column, int64_t n_out, uint32_t flags, const double *uniforms,
                           int32_t *invalid_kinds) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    return do_resample(pf, method, prio_kind, prio_param, prio_column, n_out, flags, uniforms, invalid_kinds, -1.0);
}

int32_t genpf_rejuvenate_mh(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux, int32_t n_iters,
                            int64_t *n_accept) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    return do_mh(pf, tau, obs, aux, n_iters, nullptr, nullptr, nullptr, n_accept);
}
int32_t genpf_rejuvenate_mh_with_noise(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux,
                                       const double *U2, const double *Z2, const double *U3, int64_t *n_accept) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    if (!U2 || !Z2 || !U3) return fail(GENPF_ERR_INVALID_ARG, "noise columns are NULL");
    return do_mh(pf, tau, obs, aux, 1, U2, Z2, U3, n_accept);
}

int32_t genpf_rejuvenate_reweight(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux, int32_t n_iters) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    return do_mh(pf, tau, obs, aux, n_iters, nullptr, nullptr, nullptr, nullptr, true);
}
int32_t genpf_rejuvenate_reweight_with_noise(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux,
                                             const double *U2, const double *Z2) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    if (!U2 || !Z2) return fail(GENPF_ERR_INVALID_ARG, "noise columns are NULL");
    return do_mh(pf, tau, obs, aux, 1, U2, Z2, Z2, nullptr, true);
}

// pf_move_reweight! with move_reweight(trace, proposal, proposal_args) (rejuvenate.jl:74-90,134-148): slice tau is
// re-proposed from the plugin's proposal and log_weights += up_weight - fwd_weight + bwd_weight.  U2, Z2: the
// proposal's draws (n_iters must be 1 then) or both NULL.
int32_t genpf_rejuvenate_reweight_proposal(genpf_filter_t pf, int64_t tau, const double *obs, const double *aux,
                                           int32_t n_iters, const double *U2, const double *Z2) {
    if (!pf) return fail(GENPF_ERR_INVALID_ARG, "filter handle is NULL");
    if (!(model_info(pf->model)->caps & 1)) return fail(GENPF_ERR_UNSUPPORTED, "this model defines no custom proposal");
    if ((U2 == nullptr) != (Z2 == nullptr)) return fail(GENPF_ERR_INVALID_ARG, "give both noise columns or neither");
    if (U2 && n_iters != 1) return fail(GENPF_ERR_INVALID_ARG, "noise columns serve one iteration");
    return do_mh(pf, tau, obs, aux, n_iters, U2, Z2, U2 ? Z2 : nullptr, nullptr, true, false, 1);
}

int32_t genpf_step(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                   const double *obs_t, const double *aux_t, int32_t method, double ess_frac, int32_t mh_iters,
                   double *ess_out) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    if (t != pf->t_cur + 1) return fail(GENPF_ERR_INVALID_ARG, "genpf_step must advance to t_cur + 1");
    // README.md:68-74: the resample + rejuvenation happen only when ESS < ess_frac * n.  ess_frac >= 1 means
    // "always" and needs no host round trip; otherwise the ESS decides on the host like the reference loop.
    const bool fusable = (method == GENPF_STRATIFIED) && mh_iters >= 0 && mh_iters < 256;
    bool any = false, every = true, finalized = false;
    if (ess_frac >= 1.0 && !ess_out) {
        any = true;
    } else if (!ess_out && fusable) {
        // asynchronous adaptive step: every filter decides ON THE DEVICE (ess < ess_frac * n, README.md:68); the
        // finalize sets the per-filter flag that gates the scan and the fused kernel -- no host round trip
        GENPF_TRY(pf->sc.O.ensure(4));
        return do_step_fused(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, false, nullptr, nullptr, 1, ess_frac);
    } else {
        // one finalize serves the decision, the scan's tile offsets and (if taken) update_lml_est!
        GENPF_TRY(pf->sc.O.ensure(4));
        GENPF_TRY(ensure_stats(pf, pf->sc.tile_off.as<double>(), ess_frac >= 1.0 ? -1.0 : ess_frac,
                               fusable ? pf->lml : nullptr));
        finalized = fusable;
        GENPF_TRY(read_stats(pf, 0));
        for (int64_t f = 0; f < pf->nf; ++f) {
            if (ess_out) ess_out[f] = pf->h_stats[f].ess;
            const bool r = pf->h_stats[f].do_resample != 0;
            any |= r;
            every &= r;
        }
    }
    // filters of one batch may disagree (every view decides for itself, README.md:68-74): the device-side
    // do_resample flag of each filter gates the kernels, the others are only updated
    const int gate = (any && !every) ? 1 : 0;
    if (any && fusable)
        return do_step_fused(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, finalized, nullptr, nullptr, gate);
    if (any) {
        GENPF_TRY(do_resample(pf, method, GENPF_PRIO_NONE, 1.0, nullptr, pf->n, 0, nullptr, nullptr, gate ? ess_frac : -1.0));
        GENPF_TRY(do_mh(pf, t - 1, obs_prev, aux_prev, mh_iters, nullptr, nullptr, nullptr, nullptr, false, gate != 0));
    }
    return do_propagate(pf, false, t, obs_t, aux_t, nullptr, nullptr);
}

// T README iterations enqueued by ONE call ("T = 1000 is a host loop ... CUDA Graph the step"): nothing is
// copied back and nothing synchronises; with GENPF_RUN_GRAPH steps 2..T are captured into a CUDA graph and launched
// as one unit, which removes the per-kernel launch cost that dominates small filters (config 1: n = 100).
int32_t genpf_run_steps(genpf_filter_t pf, int64_t t_first, int64_t n_steps, const double *obs, const double *aux,
                        int32_t method, double ess_frac, int32_t mh_iters, uint32_t flags) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    if (t_first != pf->t_cur + 1) return fail(GENPF_ERR_INVALID_ARG, "genpf_run_steps must start at t_cur + 1");
    if (n_steps < 1 || !obs) return fail(GENPF_ERR_INVALID_ARG, "genpf_run_steps: need n_steps >= 1 and observations");
    if (method != GENPF_STRATIFIED || mh_iters < 0 || mh_iters > 255)
        return fail(GENPF_ERR_UNSUPPORTED, "genpf_run_steps runs the fused stratified step (use genpf_step for the other methods)");
    const ModelInfo &mi = *model_info(pf->model);
    if (mi.naux > 0 && !aux) return fail(GENPF_ERR_INVALID_ARG, "aux is NULL");
    const int64_t nf = pf->nf;
    // row r of obs / aux belongs to time t_first - 1 + r: row 0 is the step the first mh move revisits
    const double *d_rows = nullptr;
    if (nf > 1) {
        GENPF_TRY(pf->run_obs.ensure((size_t)((n_steps + 1) * nf) * 8));
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->run_obs.p, obs, (size_t)((n_steps + 1) * nf) * 8, cudaMemcpyHostToDevice, pf->stream));
        d_rows = pf->run_obs.as<double>();
    }
    const int gate = ess_frac < 1.0 ? 1 : 0;
    auto one = [&](int64_t r) -> int32_t {
        const double *ap = aux ? aux + (r - 1) * mi.naux : nullptr, *at = aux ? aux + r * mi.naux : nullptr;
        return do_step_fused(pf, t_first + r - 1, obs + (r - 1) * nf, ap, obs + r * nf, at, mh_iters, false, nullptr, nullptr,
                             gate, gate ? ess_frac : -1.0, d_rows ? d_rows + (r - 1) * nf : nullptr,
                             d_rows ? d_rows + r * nf : nullptr);
    };
    GENPF_TRY(pf->sc.O.ensure(4));
    const bool graph = (flags & GENPF_RUN_GRAPH) && n_steps > 1;
    if (graph && (pf->flags & GENPF_KEEP_HISTORY)) return fail(GENPF_ERR_UNSUPPORTED, "graph capture with GENPF_KEEP_HISTORY");
    GENPF_TRY(one(1));  // eager: sizes every scratch buffer, so the captured steps allocate nothing
    if (!graph) {
        for (int64_t r = 2; r <= n_steps; ++r) GENPF_TRY(one(r));
        return GENPF_OK;
    }
    if (pf->graph_exec) {  // the previous run's graph: its launch has been ordered before us on the stream
        GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
        cudaGraphExecDestroy((cudaGraphExec_t)pf->graph_exec);
        pf->graph_exec = nullptr;
    }
    GENPF_CUDA_TRY(cudaStreamBeginCapture(pf->stream, cudaStreamCaptureModeThreadLocal));
    int32_t st = GENPF_OK;
    for (int64_t r = 2; r <= n_steps && st == GENPF_OK; ++r) st = one(r);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(pf->stream, &g);
    if (st != GENPF_OK) {
        if (g) cudaGraphDestroy(g);
        return st;
    }
    GENPF_CUDA_TRY(e);
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    GENPF_CUDA_TRY(e);
    pf->graph_exec = ex;
    GENPF_CUDA_TRY(cudaGraphLaunch(ex, pf->stream));
    return GENPF_OK;
}

// One README iteration in parity mode: the resample is taken (forced), every random draw is supplied
// by the caller -- stratum / inverse-CDF uniforms (NULL: the library's Philox draws), the mh move's [U2, Z2, U3]
// and the update's [U1, Z1], n_particles * n_filters each, indexed by output particle.  Stratified resampling runs
// the SAME kernels as genpf_step (k_scan + k_step_fused) with the column noise policy.
int32_t genpf_step_with_noise(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                              const double *obs_t, const double *aux_t, int32_t method, double ess_frac,
                              int32_t mh_iters, const double *uniforms, const double *U2, const double *Z2,
                              const double *U3, const double *U1, const double *Z1) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    if (t != pf->t_cur + 1) return fail(GENPF_ERR_INVALID_ARG, "genpf_step_with_noise must advance to t_cur + 1");
    if (mh_iters != 0 && mh_iters != 1) return fail(GENPF_ERR_INVALID_ARG, "noise columns serve mh_iters 0 or 1");
    if (!U1 || !Z1 || (mh_iters == 1 && (!U2 || !Z2 || !U3))) return fail(GENPF_ERR_INVALID_ARG, "noise columns are NULL");
    // ess_frac >= 1: every filter resamples; below, each filter decides on the device (ess < ess_frac * n)
    const int gate = ess_frac < 1.0 ? 1 : 0;
    if (method == GENPF_STRATIFIED) {
        const double *d_u = nullptr;
        NoiseCols nz{nullptr, nullptr, nullptr, nullptr, nullptr};
        if (uniforms) {
            GENPF_TRY(pf->uni_buf.ensure((size_t)(pf->n * pf->nf) * 8));
            GENPF_CUDA_TRY(cudaMemcpyAsync(pf->uni_buf.p, uniforms, (size_t)(pf->n * pf->nf) * 8, cudaMemcpyHostToDevice, pf->stream));
            d_u = pf->uni_buf.as<double>();
        }
        GENPF_TRY(stage_noise(pf, 0, U2, &nz.U, 0));
        GENPF_TRY(stage_noise(pf, 1, Z2, &nz.Z, 0));
        GENPF_TRY(stage_noise(pf, 2, U3, &nz.U3, 0));
        GENPF_TRY(stage_noise(pf, 3, U1, &nz.Uup, 0));
        GENPF_TRY(stage_noise(pf, 4, Z1, &nz.Zup, 0));
        GENPF_TRY(pf->sc.O.ensure(4));
        return do_step_fused(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, false, d_u, &nz, gate, gate ? ess_frac : -1.0);
    }
    GENPF_TRY(do_resample(pf, method, GENPF_PRIO_NONE, 1.0, nullptr, pf->n, 0, uniforms, nullptr, gate ? ess_frac : -1.0));
    if (mh_iters == 1) GENPF_TRY(do_mh(pf, t - 1, obs_prev, aux_prev, 1, U2, Z2, U3, nullptr, false, gate != 0));
    return do_propagate(pf, false, t, obs_t, aux_t, U1, Z1);
}

// resolve which column holds (tau, field) and, for archived slices, the lineage index into it
static int32_t locate_field(genpf_filter_t pf, int32_t field, int64_t tau, XSrc *x, const int32_t **idx,
                            int64_t *n_src) {
    if (field < 0 || field >= pf->NF + pf->NB) return fail(GENPF_ERR_INVALID_ARG, "field index out of range");
    if (tau < 0 || tau > pf->t_cur) return fail(GENPF_ERR_INVALID_ARG, "time step out of range");
    const Cols *c = nullptr;
    *idx = nullptr;
    *n_src = pf->n;
    if (tau >= pf->t_cur - 1 && tau >= 1) {
        c = &pf->slice(tau);
    } else {
        const HistSlice *h = nullptr;
        for (const auto &hs : pf->hist)
            if (hs.tau == tau) h = &hs;
        if (!h) return fail(GENPF_ERR_STATE, "time slice is not resident (create the filter with GENPF_KEEP_HISTORY)");
        c = &h->c;
        *n_src = h->n;
        if (h->gen < pf->n_resamples) {
            // compose ancestry backwards: anc_j = parents_{g+1}[ ... parents_R[j] ]
            const int64_t total = pf->n * pf->nf;
            GENPF_TRY(pf->tmp_idx.ensure((size_t)total * 4));
            int32_t *anc = pf->tmp_idx.as<int32_t>();
            GENPF_LAUNCH(k_iota32, grid_1d(total), 256, pf->stream, anc, pf->n, total);
            for (int64_t r = pf->n_resamples; r > h->gen; --r) {
                const ParentLog &pl = pf->plog[(size_t)r - 1];
                GENPF_LAUNCH(k_compose_lineage, dim3(grid_1d(pf->n), (unsigned)pf->nf), 256, pf->stream, anc, pl.parents,
                             pl.n_cur, pf->n);
            }
            *idx = anc;
        }
    }
    x->d = field < pf->NF ? c->f[field] : nullptr;
    x->b = field < pf->NF ? nullptr : c->b[field - pf->NF];
    return GENPF_OK;
}

int32_t genpf_mean_var(genpf_filter_t pf, int32_t field, int64_t tau, double *mean, double *var) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    XSrc x;
    const int32_t *idx;
    int64_t n_src;
    GENPF_TRY(locate_field(pf, field, tau, &x, &idx, &n_src));
    if (idx || n_src != pf->n) {  // materialise the lineage-resolved column
        const int64_t total = pf->n * pf->nf;
        GENPF_TRY(pf->tmp_col.ensure((size_t)total * 8));
        GENPF_LAUNCH(k_read_field, dim3(grid_1d(pf->n), (unsigned)pf->nf), 256, pf->stream, x, idx, n_src, pf->n,
                     pf->tmp_col.as<double>());
        x.d = pf->tmp_col.as<double>();
        x.b = nullptr;
    }
    GENPF_TRY(ensure_stats(pf, nullptr, -1.0, nullptr));
    GENPF_TRY(launch_mean_var(pf->stream, pf->sc, pf->lw, x, pf->n, pf->nf, pf->sc.st(0, pf->nf)));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, pf->sc.moment_out.p, (size_t)pf->nf * 16, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    for (int64_t f = 0; f < pf->nf; ++f) {
        if (mean) mean[f] = pf->h_pinned[f];
        if (var) var[f] = pf->h_pinned[pf->nf + f];
    }
    return GENPF_OK;
}

int32_t genpf_get_log_weights(genpf_filter_t pf, double *out) {
    GENPF_TRY(check_filter(pf));
    if (!out) return fail(GENPF_ERR_INVALID_ARG, "out is NULL");
    GENPF_CUDA_TRY(cudaMemcpyAsync(out, pf->lw, (size_t)(pf->n * pf->nf) * 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}
int32_t genpf_set_log_weights(genpf_filter_t pf, const double *in) {
    GENPF_TRY(check_filter(pf));
    if (!in) return fail(GENPF_ERR_INVALID_ARG, "in is NULL");
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->lw, in, (size_t)(pf->n * pf->nf) * 8, cudaMemcpyHostToDevice, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    pf->part_valid = false;
    return GENPF_OK;
}
int32_t genpf_get_parents(genpf_filter_t pf, int64_t *out, uint32_t flags) {
    GENPF_TRY(check_filter(pf));
    if (!out) return fail(GENPF_ERR_INVALID_ARG, "out is NULL");
    const int64_t total = pf->n * pf->nf;
    GENPF_TRY(pf->tmp_col.ensure((size_t)total * 8));
    GENPF_LAUNCH((k_convert_idx<int32_t, long long>), grid_1d(total), 256, pf->stream, pf->parents,
                 pf->tmp_col.as<long long>(), total, (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0));
    GENPF_CUDA_TRY(cudaMemcpyAsync(out, pf->tmp_col.p, (size_t)total * 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}
int32_t genpf_get_field(genpf_filter_t pf, int32_t field, int64_t tau, double *out) {
    GENPF_TRY(check_filter(pf));
    if (!out) return fail(GENPF_ERR_INVALID_ARG, "out is NULL");
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    XSrc x;
    const int32_t *idx;
    int64_t n_src;
    GENPF_TRY(locate_field(pf, field, tau, &x, &idx, &n_src));
    const int64_t total = pf->n * pf->nf;
    GENPF_TRY(pf->tmp_col.ensure((size_t)total * 8));
    GENPF_LAUNCH(k_read_field, dim3(grid_1d(pf->n), (unsigned)pf->nf), 256, pf->stream, x, idx, n_src, pf->n,
                 pf->tmp_col.as<double>());
    GENPF_CUDA_TRY(cudaMemcpyAsync(out, pf->tmp_col.p, (size_t)total * 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}
int32_t genpf_set_field(genpf_filter_t pf, int32_t field, int64_t tau, const double *in) {
    GENPF_TRY(check_filter(pf));
    if (!in) return fail(GENPF_ERR_INVALID_ARG, "in is NULL");
    if (field < 0 || field >= pf->NF + pf->NB) return fail(GENPF_ERR_INVALID_ARG, "field index out of range");
    if (pf->t_cur < 1 || tau < pf->t_cur - 1 || tau > pf->t_cur || tau < 0)
        return fail(GENPF_ERR_INVALID_ARG, "only the resident window {t_cur-1, t_cur} can be written");
    const int64_t total = pf->n * pf->nf;
    GENPF_TRY(pf->tmp_col.ensure((size_t)total * 8));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->tmp_col.p, in, (size_t)total * 8, cudaMemcpyHostToDevice, pf->stream));
    Cols &c = pf->slice(tau);
    GENPF_LAUNCH(k_write_field, grid_1d(total), 256, pf->stream, pf->tmp_col.as<double>(),
                 field < pf->NF ? c.f[field] : (double *)nullptr,
                 field < pf->NF ? (uint8_t *)nullptr : c.b[field - pf->NF], total);
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}
int32_t genpf_filter_get_progress(genpf_filter_t pf, int64_t *t_cur, int64_t *n_resamples, double *log_ml_accum) {
    GENPF_TRY(check_filter(pf));
    if (t_cur) *t_cur = pf->t_cur;
    if (n_resamples) *n_resamples = pf->n_resamples;
    if (log_ml_accum) {
        GENPF_CUDA_TRY(cudaMemcpyAsync(log_ml_accum, pf->lml, (size_t)pf->nf * 8, cudaMemcpyDeviceToHost, pf->stream));
        GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    }
    return GENPF_OK;
}
int32_t genpf_filter_set_progress(genpf_filter_t pf, int64_t t_cur, int64_t n_resamples, const double *log_ml_accum) {
    GENPF_TRY(check_filter(pf));
    if (t_cur < 1 || n_resamples < 0 || !log_ml_accum) return fail(GENPF_ERR_INVALID_ARG, "genpf_filter_set_progress: bad arguments");
    if (pf->flags & GENPF_KEEP_HISTORY) return fail(GENPF_ERR_UNSUPPORTED, "resume of a filter that keeps its history is not supported");
    pf->t_cur = t_cur;
    pf->n_resamples = n_resamples;
    pf->part_valid = false;
    GENPF_LAUNCH(k_iota32, grid_1d(pf->n * pf->nf), 256, pf->stream, pf->parents, pf->n, pf->n * pf->nf);
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->lml, log_ml_accum, (size_t)pf->nf * 8, cudaMemcpyHostToDevice, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}
int32_t genpf_get_accepts(genpf_filter_t pf, uint8_t *out) {
    GENPF_TRY(check_filter(pf));
    if (!out) return fail(GENPF_ERR_INVALID_ARG, "out is NULL");
    GENPF_CUDA_TRY(cudaMemcpyAsync(out, pf->accepts, (size_t)(pf->n * pf->nf), cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}

}  // extern "C"

// ---- resizing on device state (resize.jl:236-334); parents already written for n_out slots, new weights in lw_alt
namespace genpf {
static int32_t apply_parents_and_swap(genpf_filter_t pf, int64_t n_out) {
    const int64_t tpf_out = ceil_div(n_out, kTile);
    GatherCols g = window_gather_cols(pf);
    GENPF_LAUNCH(k_gather, dim3((unsigned)tpf_out, (unsigned)pf->nf), kThreads, pf->stream, g, pf->parents, pf->n, n_out, tpf_out,
                 (const double *)nullptr, (double *)nullptr, (const Stats *)nullptr, 0, 0);
    pf->buf ^= 1;
    std::swap(pf->lw, pf->lw_alt);
    pf->part_valid = false;
    resize_commit(pf);
    const int64_t n_prev = pf->n;
    GENPF_TRY(log_parents(pf, n_prev, n_out));
    if (n_out != n_prev) {
        pf->n = n_out;
        if (pf->first_filter) pf->rng_offset = pf->first_filter * n_out;  // Philox counters stay global batch slots
        GENPF_TRY(resize_spare(pf, n_out));
    }
    return GENPF_OK;
}
// batched = true: the operation keeps every filter of a batch at the same size (replicate, dereplicate, resize by
// resampling); coalesce / optimal resize / proportionmap give data-dependent sizes and need n_filters == 1
static int32_t check_resizable(genpf_filter_t pf, bool batched = false) {
    GENPF_TRY(check_filter(pf));
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    if (pf->nf != 1 && !batched) return fail(GENPF_ERR_UNSUPPORTED, "this resizing operation needs n_filters == 1");
    if (pf->shard) return fail(GENPF_ERR_UNSUPPORTED, "a sharded filter cannot be resized");
    return GENPF_OK;
}
// pf_introduce!, first half (resize.jl:362-371): the old particles keep their place, log_weights .+= log_ml_est and
// log_ml_est = 0; the new slots get a placeholder ancestor until k_introduce fills them
static __global__ void k_introduce_prepare(const double *lw, const double *lml, int64_t n_old, int64_t n_new, int32_t *parents,
                                           double *lw_out) {
    const int64_t f = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_new; j += (int64_t)gridDim.x * blockDim.x) {
        parents[f * n_new + j] = j < n_old ? (int32_t)j : 0;
        lw_out[f * n_new + j] = j < n_old ? lw[f * n_old + j] + lml[f] : 0.0;
    }
}
template <class Model, class Noise>
static int32_t launch_introduce(genpf_filter_t pf, int64_t n_old, int64_t m, const double *obs_hist, const double *aux_hist,
                                Noise noise, int use_proposal) {
    const ModelInfo &mi = *model_info(pf->model);
    const void *fn = nullptr;
    GENPF_PLUGIN_OR_BUILTIN(Model, kPlugIntro + NoiseIndex<Noise>::value, (&k_introduce<SigModel<Model>, Noise>), fn);
    const int64_t t = pf->t_cur;
    return launch_typed("k_introduce", &k_introduce<SigModel<Model>, Noise>, fn, dim3(grid_1d(m), (unsigned)pf->nf), dim3(256),
                        pf->stream, pf->P, t, pf->slice(t - 1), pf->slice(t), pf->lw, obs_hist, aux_hist, mi.naux, n_old, m,
                        noise, use_proposal);
}
template <class Model>
static int32_t introduce_model(genpf_filter_t pf, int64_t n_old, int64_t m, const double *obs_hist, const double *aux_hist,
                               const double *dU, const double *dZ, int use_proposal) {
    if (dU || dZ) {
        NoiseCols nz{nullptr, nullptr, nullptr, dU, dZ};
        return launch_introduce<Model, NoiseCols>(pf, n_old, m, obs_hist, aux_hist, nz, use_proposal);
    }
    if (pf->flags & GENPF_NOISE_PHILOX53) {
        NoisePhilox53 nz{pf->seed, 0, pf->rng_offset};
        return launch_introduce<Model, NoisePhilox53>(pf, n_old, m, obs_hist, aux_hist, nz, use_proposal);
    }
    NoiseLean nz{pf->seed, 0, pf->rng_offset};
    return launch_introduce<Model, NoiseLean>(pf, n_old, m, obs_hist, aux_hist, nz, use_proposal);
}
}  // namespace genpf

extern "C" {

int32_t genpf_replicate(genpf_filter_t pf, int64_t k, int32_t layout) {
    GENPF_TRY(check_resizable(pf, true));
    if (k < 1) return fail(GENPF_ERR_INVALID_ARG, "n_replicates must be >= 1");
    const int64_t n_out = pf->n * k;
    if (n_out >= 0x7FFFFFF0ll) return fail(GENPF_ERR_UNSUPPORTED, "replicated population must stay < 2^31");
    GENPF_TRY(resize_target(pf, n_out));
    GENPF_LAUNCH((k_replicate<int32_t>), dim3(grid_1d(n_out), (unsigned)pf->nf), 256, pf->stream, (const double *)pf->lw, pf->n, k,
                 layout == GENPF_LAYOUT_INTERLEAVED ? 1 : 0, pf->parents, (int64_t)0, pf->lw_alt);
    return apply_parents_and_swap(pf, n_out);
}

// pf_introduce! (resize.jl:351-421) for device plugins: n_new particles generated under the whole observation history
// obs[t_cur * n_filters] (row tau-1 = time tau; aux likewise with n_aux columns) are appended to every filter.
int32_t genpf_introduce(genpf_filter_t pf, int64_t n_new, const double *obs, const double *aux, int32_t use_proposal,
                        const double *U, const double *Z) {
    GENPF_TRY(check_resizable(pf, true));
    const ModelInfo &mi = *model_info(pf->model);
    if (n_new < 1) return fail(GENPF_ERR_INVALID_ARG, "pf_introduce!: n_particles must be >= 1");
    if (!obs || (mi.naux > 0 && !aux)) return fail(GENPF_ERR_INVALID_ARG, "pf_introduce!: the observation history is NULL");
    if (use_proposal && !(mi.caps & 1)) return fail(GENPF_ERR_UNSUPPORTED, "this model has no custom proposal");
    if (pf->flags & GENPF_KEEP_HISTORY) return fail(GENPF_ERR_UNSUPPORTED, "pf_introduce! with GENPF_KEEP_HISTORY");
    const int64_t n_old = pf->n, n_out = n_old + n_new, t = pf->t_cur, nf = pf->nf;
    if (n_out >= 0x7FFFFFF0ll) return fail(GENPF_ERR_UNSUPPORTED, "population must stay < 2^31");
    // observation / aux history and (parity mode) the chains' noise columns [tau-1][filter][i]
    const size_t hist = (size_t)(t * nf + t * mi.naux) * 8;
    GENPF_TRY(pf->run_obs.ensure(hist + 16));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->run_obs.p, obs, (size_t)(t * nf) * 8, cudaMemcpyHostToDevice, pf->stream));
    double *d_obs = pf->run_obs.as<double>(), *d_aux = d_obs + t * nf;
    if (mi.naux > 0)
        GENPF_CUDA_TRY(cudaMemcpyAsync(d_aux, aux, (size_t)(t * mi.naux) * 8, cudaMemcpyHostToDevice, pf->stream));
    const double *dU = nullptr, *dZ = nullptr;
    GENPF_TRY(stage_noise(pf, 3, U, &dU, t * nf * n_new));
    GENPF_TRY(stage_noise(pf, 4, Z, &dZ, t * nf * n_new));
    GENPF_TRY(resize_target(pf, n_out));
    GENPF_LAUNCH(k_introduce_prepare, dim3(grid_1d(n_out), (unsigned)nf), 256, pf->stream, (const double *)pf->lw,
                 (const double *)pf->lml, n_old, n_out, pf->parents, pf->lw_alt);
    GENPF_TRY(apply_parents_and_swap(pf, n_out));
    GENPF_CUDA_TRY(cudaMemsetAsync(pf->lml, 0, (size_t)nf * 8, pf->stream));
    int32_t st;
    switch (pf->model) {
        case kModelObjectMotion: st = introduce_model<ObjectMotion>(pf, n_old, n_new, d_obs, d_aux, dU, dZ, use_proposal); break;
        case kModelLinGauss1D: st = introduce_model<LinGauss1D>(pf, n_old, n_new, d_obs, d_aux, dU, dZ, use_proposal); break;
        default: st = introduce_model<PluginTag>(pf, n_old, n_new, d_obs, d_aux, dU, dZ, use_proposal); break;
    }
    pf->part_valid = false;
    return st;
}

int32_t genpf_dereplicate(genpf_filter_t pf, int64_t k, int32_t layout, int32_t method, const double *uniforms) {
    GENPF_TRY(check_resizable(pf, true));
    if (k < 1 || pf->n % k != 0) return fail(GENPF_ERR_INVALID_ARG, "n must be a multiple of n_replicates (resize.jl:270)");
    const int64_t n_out = pf->n / k;
    const double *d_u = nullptr;
    if (uniforms) {
        GENPF_TRY(pf->uni_buf.ensure((size_t)(n_out * pf->nf) * 8));
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->uni_buf.p, uniforms, (size_t)(n_out * pf->nf) * 8, cudaMemcpyHostToDevice, pf->stream));
        d_u = pf->uni_buf.as<double>();
    }
    GENPF_TRY(resize_target(pf, n_out));
    UniSrc uni{d_u, pf->seed, make_stream(kPurposeDerep, (uint64_t)pf->n_resamples + 1),
               pf->first_filter ? pf->first_filter * n_out : pf->rng_offset};  // indexed by output slot
    GENPF_LAUNCH((k_dereplicate<int32_t>), dim3(grid_1d(n_out), (unsigned)pf->nf), 256, pf->stream, (const double *)pf->lw, pf->n, k,
                 layout == GENPF_LAYOUT_INTERLEAVED ? 1 : 0, method == GENPF_SAMPLE ? 1 : 0, uni, pf->parents,
                 (int64_t)0, pf->lw_alt);
    return apply_parents_and_swap(pf, n_out);
}

int32_t genpf_coalesce(genpf_filter_t pf, int64_t *n_new) {
    GENPF_TRY(check_resizable(pf));
    const int64_t n = pf->n;
    GENPF_TRY(pf->key_buf.ensure((size_t)n * 8));
    HashCols hc;
    memset(&hc, 0, sizeof(hc));
    for (int sl = 0; sl < 2; ++sl) {
        if (pf->t_cur == 1 && sl == 0) continue;  // slice 0 is the constant initial slice
        for (int i = 0; i < pf->NF; ++i) hc.f[hc.nf++] = pf->win[pf->buf][sl].f[i];
        for (int i = 0; i < pf->NB; ++i) hc.b[hc.nb++] = pf->win[pf->buf][sl].b[i];
    }
    GENPF_LAUNCH(k_hash_window, grid_1d(n), 256, pf->stream, hc, n, pf->key_buf.as<int64_t>());
    long long *n_new_dev = nullptr;
    GENPF_TRY(launch_coalesce<int32_t>(pf->stream, pf->cb, pf->lw, pf->key_buf.as<int64_t>(), n, pf->parents, 0,
                                       pf->lw_alt, &n_new_dev));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, n_new_dev, 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    const int64_t n_out = (int64_t) * reinterpret_cast<long long *>(pf->h_pinned);
    if (n_new) *n_new = n_out;
    return apply_parents_and_swap(pf, n_out);
}

int32_t genpf_proportionmap(genpf_filter_t pf, int32_t field, int64_t tau, double *values_out, double *props_out,
                            int64_t cap, int64_t *n_unique) {
    GENPF_TRY(check_resizable(pf));
    if (!values_out || !props_out || cap <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_proportionmap: bad arguments");
    const int64_t n = pf->n;
    XSrc x;
    const int32_t *idx;
    int64_t n_src;
    GENPF_TRY(locate_field(pf, field, tau, &x, &idx, &n_src));
    // the value column as doubles; its bit patterns are the grouping keys (Julia's isequal on Float64 / Bool)
    GENPF_TRY(pf->tmp_col.ensure((size_t)n * 8));
    GENPF_LAUNCH(k_read_field, dim3(grid_1d(n), 1), 256, pf->stream, x, idx, n_src, n, pf->tmp_col.as<double>());
    GENPF_TRY(ensure_stats(pf, nullptr, -1.0, nullptr));
    GENPF_TRY(pf->prio_buf.ensure((size_t)n * 8));
    GENPF_TRY(pf->uni_buf.ensure((size_t)n * 8));
    long long *n_dev = nullptr;
    int32_t *first = reinterpret_cast<int32_t *>(pf->prio_buf.p);
    double *prop = pf->uni_buf.as<double>();
    GENPF_TRY(launch_coalesce<int32_t>(pf->stream, pf->cb, pf->lw, reinterpret_cast<const int64_t *>(pf->tmp_col.p), n, first,
                                       (int64_t)0, prop, &n_dev, (const Stats *)pf->sc.st(0, 1)));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, n_dev, 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    const int64_t g = (int64_t) * reinterpret_cast<long long *>(pf->h_pinned);
    if (n_unique) *n_unique = g;
    const int64_t w = g < cap ? g : cap;
    // values at the groups' first indices
    GENPF_TRY(pf->key_buf.ensure((size_t)w * 8));
    GENPF_LAUNCH(k_gather_f64, dim3(grid_1d(w), 1), 256, pf->stream, (const double *)pf->tmp_col.as<double>(),
                 (const int32_t *)first, n, w, pf->key_buf.as<double>());
    GENPF_CUDA_TRY(cudaMemcpyAsync(values_out, pf->key_buf.p, (size_t)w * 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaMemcpyAsync(props_out, prop, (size_t)w * 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    return GENPF_OK;
}

int32_t genpf_optimal_resize_dev(genpf_filter_t pf, int64_t n_out, const double *uniform, uint32_t flags,
                                 int64_t *n_keep, double *inv_w_threshold, int32_t *invalid_kinds) {
    GENPF_TRY(check_resizable(pf));
    if (n_out <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_optimal_resize_dev: n_out must be positive");
    if (n_out > pf->n) return fail(GENPF_ERR_ASSERT, "optimal resize cannot grow the filter (@assert n_particles <= n_old, resize.jl:183)");
    GENPF_TRY(resize_target(pf, n_out));
    OptResult res;
    UniSrc uni{nullptr, pf->seed, make_stream(kPurposeResample, (uint64_t)pf->n_resamples + 1), pf->rng_offset};
    if (!pf->h_opt_ctrl) GENPF_CUDA_TRY(cudaMallocHost(&pf->h_opt_ctrl, sizeof(OptCtrl)));
    const int32_t st = optimal_resize_core<int32_t>(pf->stream, pf->sc, pf->ob, (OptCtrl *)pf->h_opt_ctrl, pf->h_stats, pf->lw,
                                                    pf->n, n_out, uniform, uni, flags & GENPF_CHECK, pf->parents, (int64_t)0,
                                                    pf->lw_alt, &res);
    pf->part_valid = false;  // the core reused the partial/statistics scratch
    if (n_keep) *n_keep = res.n_keep;
    if (inv_w_threshold) *inv_w_threshold = res.inv_w;
    if (invalid_kinds) {
        invalid_kinds[0] = res.kind;
        invalid_kinds[1] = res.kind_strat;
    }
    if (st != GENPF_OK) {
        resize_rollback(pf);
        return st;
    }
    return apply_parents_and_swap(pf, n_out);
}

}  // extern "C"
