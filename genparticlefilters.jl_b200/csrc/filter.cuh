// filter.cuh -- device-resident particle population kernels (K8, K10, K11).
//
// HBM layout (struct of arrays, one column per field per time slice, particles of filter f at [f*n, (f+1)*n)):
//   f64 fields  : double  col[slot][field][n_filters*n]
//   u8  fields  : uint8_t col[slot][field][n_filters*n]
//   log_weights : double  lw[n_filters*n]
//   parents     : int32   parents[n_filters*n]      (ancestor of each particle in the last resample, 0-based local)
// `slot` is a ring over time (slot = t mod K, K = 2 unless GENPF_KEEP_HISTORY): the fixed-lag window the
// README loop touches is {t-1, t}.  Every column exists twice (buffer A/B): the ancestor gather writes the
// other buffer and the two swap, exactly like `traces`/`new_traces` (utils.jl:10-15).
#pragma once
#include "kernels.cuh"
#include "models.cuh"

namespace genpf {

struct Cols {
    double *f[kMaxF];
    uint8_t *b[kMaxB];
};

#ifndef GENPF_STATE_THREADS
#define GENPF_STATE_THREADS 256
#endif
constexpr int kStateThreads = GENPF_STATE_THREADS;  // block size of every kernel that produces particle state + K1 partials

template <int T = kThreads>
__device__ __forceinline__ void load_tile_u8(const uint8_t *col, int64_t base, int64_t valid,
                                             uint8_t (&v)[kTile / T]) {
    const uint8_t *p = col + base;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & 1) == 0);
#pragma unroll
    for (int j = 0; j < kTile / T / 2; ++j) {
        int64_t e = (int64_t)(j * T + threadIdx.x) * 2;
        if (vec_ok && e + 1 < valid) {
            uchar2 d = *reinterpret_cast<const uchar2 *>(p + e);
            v[2 * j] = d.x;
            v[2 * j + 1] = d.y;
        } else {
            v[2 * j] = e < valid ? p[e] : 0;
            v[2 * j + 1] = e + 1 < valid ? p[e + 1] : 0;
        }
    }
}
template <int T = kThreads>
__device__ __forceinline__ void store_tile_u8(uint8_t *col, int64_t base, int64_t valid,
                                              const uint8_t (&v)[kTile / T]) {
    uint8_t *p = col + base;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & 1) == 0);
#pragma unroll
    for (int j = 0; j < kTile / T / 2; ++j) {
        int64_t e = (int64_t)(j * T + threadIdx.x) * 2;
        if (vec_ok && e + 1 < valid) {
            *reinterpret_cast<uchar2 *>(p + e) = make_uchar2(v[2 * j], v[2 * j + 1]);
        } else {
            if (e < valid) p[e] = v[2 * j];
            if (e + 1 < valid) p[e + 1] = v[2 * j + 1];
        }
    }
}

template <class Model, int T = kThreads>
__device__ __forceinline__ void load_slices(const Cols &c, int64_t base, int64_t valid,
                                            typename Model::Slice (&s)[kTile / T]) {
#pragma unroll
    for (int fld = 0; fld < Model::NF; ++fld) {
        double v[kTile / T];
        LwSrc src{c.f[fld], 1.0};
        load_tile<T>(src, base, valid, v, 0.0);
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) s[k].f[fld] = v[k];
    }
#pragma unroll
    for (int fld = 0; fld < Model::NB; ++fld) {
        uint8_t v[kTile / T];
        load_tile_u8<T>(c.b[fld], base, valid, v);
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) s[k].b[fld] = v[k];
    }
}
template <class Model, int T = kThreads>
__device__ __forceinline__ void store_slices(const Cols &c, int64_t base, int64_t valid,
                                             const typename Model::Slice (&s)[kTile / T]) {
#pragma unroll
    for (int fld = 0; fld < Model::NF; ++fld) {
        double v[kTile / T];
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) v[k] = s[k].f[fld];
        store_tile<double, T>(c.f[fld], base, valid, v);
    }
#pragma unroll
    for (int fld = 0; fld < Model::NB; ++fld) {
        uint8_t v[kTile / T];
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) v[k] = s[k].b[fld];
        store_tile_u8<T>(c.b[fld], base, valid, v);
    }
}

// ------------------------------------------------------------------ K10 propagate + weight (+ K1 partials)
// INIT: pf_initialize (initialize.jl:39-41): slice_1 = transition(initial), lw = obs_logpdf
// else: pf_update!    (update.jl:15-21):     slice_t = transition(slice_{t-1}), lw += obs_logpdf
// Stratified initialisation / update (initialize.jl:93-108, update.jl:193-210 with stratified_map!,
// utils.jl:29-55): K strata constrain one latent of the new slice; each gets floor(n/K) particles in contiguous blocks or interleaved, the n - K*floor(n/K)
// left-over particles (the last indices) go to strata drawn with replacement; log-weights gain log(K).
enum { kPropPrior = 0, kPropProposal = 1, kPropTranslate = 2 };
struct Strata {
    const double *values;  // device, K entries; null: plain pf_initialize
    int K, field, interleaved;
    uint64_t seed, stream;  // Philox stream of the left-over particles' strata
    // how the new slice is produced: kPropPrior = Gen `generate`/`update` from the model's prior transition;
    // kPropProposal = custom proposal (initialize.jl:46-62, update.jl:79-96): x ~ q, lw += log p(x|prev) + log p(obs|x)
    // - log q(x); kPropTranslate = a trace translator (update.jl:35-44): (x, increment) = translate(current slice)
    int mode;
    __device__ __forceinline__ int stratum(int64_t i, int64_t n) const {
        const int64_t B = n / K;
        if (i < B * K) return (int)(interleaved ? i % K : i / B);
        const uint4 o = philox_at(seed, stream, (uint64_t)i);
        const int k = (int)(u53(o.x, o.y) * (double)K);
        return k < K ? k : K - 1;
    }
};

template <class Model, class Noise, bool INIT>
GENPF_KERNEL void __launch_bounds__(kStateThreads)
    k_propagate(ModelParams P, int64_t t, Cols prev, Cols next, double *lw, const double *obs_dev, double obs_val,
                int64_t n, int64_t tpf, Noise noise, Partials partials, double *ew, Strata strata = Strata{}) {
    constexpr int T = kStateThreads, I = kTile / T;
    __shared__ PartialSmem ps;
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    const int64_t base = f * n + start;
    const double obs = obs_dev ? obs_dev[f] : obs_val;
    typename Model::Slice sp[I], sn[I];
    double v[I];
    if (INIT) {
#pragma unroll
        for (int k = 0; k < I; ++k) Model::initial(P, sp[k]);
    } else {
        load_slices<Model, T>(prev, base, valid, sp);
        LwSrc src{lw, 1.0};
        load_tile<T>(src, base, valid, v, -INFINITY);
    }
#pragma unroll
    for (int k = 0; k < I; ++k) {
        int e = tile_elem<T>(k);
        double U = 0.5, Z = 0.0;
        if (e < valid) noise.up(base + e, U, Z);
        double l = 0.0;
        bool weighted = false;  // the translator's increment already contains every weight term
        if (strata.values) {  // stratified initialise / update: one latent constrained, + log p(constraint) + log K
            const double val = strata.values[strata.stratum(e < valid ? start + e : 0, n)];
            l = Model::constrain(P, t, sp[k], sn[k], U, Z, strata.field, val) + log((double)strata.K);
        } else if (strata.mode == kPropProposal) {
            if constexpr (has_proposal<Model>::value) {
                Model::propose(P, t, sp[k], obs, sn[k], U, Z);
                l = Model::transition_logpdf(P, t, sp[k], sn[k]) - Model::proposal_logpdf(P, t, sp[k], obs, sn[k]);
            }
        } else if (strata.mode == kPropTranslate) {
            if constexpr (has_translate<Model>::value) {
                l = Model::translate(P, t, sp[k], obs, sn[k], U, Z);
                weighted = true;
            }
        } else {
            Model::transition(P, t, sp[k], sn[k], U, Z);
        }
        if (!weighted) l += Model::obs_logpdf(P, sn[k], obs);
        if (e < valid) v[k] = INIT ? l : v[k] + l;
        else v[k] = -INFINITY;
    }
    store_slices<Model, T>(next, base, valid, sn);
    store_tile<double, T>(lw, base, valid, v);
    emit_partials<T>(v, partials, ps, -1, ew ? ew + base : nullptr, (int)valid);
}

// ------------------------------------------------------------------ K11 MH rejuvenation (move-accept)
// pf_move_accept! (rejuvenate.jl:40-53) with kern = Gen.mh on select(tau => latents), tau the newest
// slice: regenerate slice tau from the prior given slice tau-1; weight = obs log-density ratio;
// accept iff log(rand()) < weight.  Log-weights are untouched.
// REWEIGHT: pf_move_reweight! with move_reweight(trace, selection) (rejuvenate.jl:74-90,125-132): the regenerated
// slice always replaces the current one and the same weight is added to the particle's log-weight.
template <class Model, class Noise, bool REWEIGHT = false>
GENPF_KERNEL void __launch_bounds__(kStateThreads)
    k_mh(ModelParams P, int64_t tau, int iter, int first_step, Cols prevprev, Cols cur, const double *obs_dev,
         double obs_val, int64_t n, int64_t tpf, Noise noise, uint8_t *accepts, unsigned long long *n_accept,
         double *lw = nullptr, const Stats *gate_stats = nullptr, int use_proposal = 0) {
    constexpr int T = kStateThreads, I = kTile / T;
    __shared__ double sm[T / 32];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    // gated batch (README.md:68-74 per view): filters that did not resample are not rejuvenated either
    if (gate_stats && (!gate_stats[f].do_resample || gate_stats[f].invalid_kind == 1 || gate_stats[f].invalid_kind == 4)) return;
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    const int64_t base = f * n + start;
    const double obs = obs_dev ? obs_dev[f] : obs_val;
    typename Model::Slice sp[I], sc[I];
    if (first_step) {
#pragma unroll
        for (int k = 0; k < I; ++k) Model::initial(P, sp[k]);
    } else {
        load_slices<Model, T>(prevprev, base, valid, sp);
    }
    load_slices<Model, T>(cur, base, valid, sc);
    uint8_t acc[I];
    double cnt = 0.0;
#pragma unroll
    for (int k = 0; k < I; ++k) {
        int e = tile_elem<T>(k);
        double U = 0.5, Z = 0.0, U3 = 1.0;
        if (e < valid) noise.mh(base + e, iter, U, Z, U3);
        typename Model::Slice q;
        double alpha;
        bool proposed = false;
        if constexpr (REWEIGHT && has_proposal<Model>::value) {
            if (use_proposal) {
                // move_reweight(trace, proposal, proposal_args) (rejuvenate.jl:134-148): rel. weight = up_weight -
                // fwd_weight + bwd_weight = [p(q|pp) p(obs|q) / Q(q)] / [p(cur|pp) p(obs|cur) / Q(cur)]
                Model::propose(P, tau, sp[k], obs, q, U, Z);
                const double wn = Model::transition_logpdf(P, tau, sp[k], q) + Model::obs_logpdf(P, q, obs) -
                                  Model::proposal_logpdf(P, tau, sp[k], obs, q);
                const double wo = Model::transition_logpdf(P, tau, sp[k], sc[k]) + Model::obs_logpdf(P, sc[k], obs) -
                                  Model::proposal_logpdf(P, tau, sp[k], obs, sc[k]);
                alpha = wn - wo;
                proposed = true;
            }
        }
        if (!proposed) {
            Model::transition(P, tau, sp[k], q, U, Z);
            alpha = Model::obs_logpdf(P, q, obs) - Model::obs_logpdf(P, sc[k], obs);
        }
        bool a;
        if (REWEIGHT) {
            a = e < valid;
            if (a) lw[base + e] += alpha;
        } else {
            a = (e < valid) && mh_accept(U3, alpha);
        }
        if (a) sc[k] = q;
        acc[k] = a ? 1 : 0;
        cnt += a ? 1.0 : 0.0;
    }
    store_slices<Model, T>(cur, base, valid, sc);
    if (accepts) store_tile_u8<T>(accepts, base, valid, acc);
    if (n_accept) {
        cnt = block_sum<T>(cnt, sm);
        if (threadIdx.x == 0 && cnt > 0.0) atomicAdd(&n_accept[f], (unsigned long long)cnt);
    }
}

// ------------------------------------------------------------------ pf_introduce! (resize.jl:351-421)
// m fresh particles appended to a filter at time t: generate(model, args, observations) simulates the whole chain
// x_1 .. x_t from the prior (or the plugin's custom proposal) under the observations y_1 .. y_t; the new log-weight
// is the sum of the observation log-densities (+ log p(x|prev) - log q(x) per step with a proposal).  Only the
// fixed-lag window {t-1, t} is kept.  One thread per new particle; the chain's noise is the update noise of step tau
// at the particle's (never used before) slot, or supplied columns [tau-1][filter][i] in parity mode.
template <class N>
__device__ __forceinline__ N noise_of_step(N nz, int64_t tau, int64_t) {
    nz.step = (uint64_t)tau;
    return nz;
}
__device__ __forceinline__ NoiseCols noise_of_step(NoiseCols nz, int64_t tau, int64_t stride) {
    const double *pu = (nz.Uup || nz.Zup) ? nz.Uup : nz.U, *pz = (nz.Uup || nz.Zup) ? nz.Zup : nz.Z;
    NoiseCols o{nullptr, nullptr, nullptr, pu ? pu + (tau - 1) * stride : nullptr, pz ? pz + (tau - 1) * stride : nullptr};
    return o;
}
template <class Model, class Noise>
GENPF_KERNEL void __launch_bounds__(256)
    k_introduce(ModelParams P, int64_t t, Cols dst_prev, Cols dst_cur, double *lw, const double *obs_hist,
                const double *aux_hist, int naux, int64_t n_old, int64_t m, Noise noise, int use_proposal) {
    const int64_t f = blockIdx.y, nf = gridDim.y, n_new = n_old + m;
    const ModelParams P0 = P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {  // grid-stride: the launch caps its blocks
    P = P0;
    const int64_t slot = f * n_new + n_old + i;  // where the particle lives (and its Philox counter)
    typename Model::Slice prev, cur;
    Model::initial(P, cur);
    double w = 0.0;
    for (int64_t tau = 1; tau <= t; ++tau) {
        prev = cur;
        for (int a = 0; a < naux; ++a) P.aux[a] = aux_hist[(tau - 1) * naux + a];
        const double obs = obs_hist[(tau - 1) * nf + f];
        const Noise nz = noise_of_step(noise, tau, nf * m);
        double U = 0.5, Z = 0.0;
        nz.up(Noise::kIndexed ? f * m + i : slot, U, Z);
        bool proposed = false;
        if constexpr (has_proposal<Model>::value) {
            if (use_proposal) {
                Model::propose(P, tau, prev, obs, cur, U, Z);
                w += Model::transition_logpdf(P, tau, prev, cur) - Model::proposal_logpdf(P, tau, prev, obs, cur);
                proposed = true;
            }
        }
        if (!proposed) Model::transition(P, tau, prev, cur, U, Z);
        w += Model::obs_logpdf(P, cur, obs);
    }
#pragma unroll
    for (int c = 0; c < Model::NF; ++c) {
        dst_prev.f[c][slot] = prev.f[c];
        dst_cur.f[c][slot] = cur.f[c];
    }
#pragma unroll
    for (int c = 0; c < Model::NB; ++c) {
        dst_prev.b[c][slot] = prev.b[c];
        dst_cur.b[c][slot] = cur.b[c];
    }
    lw[slot] = w;
    }
}

#ifndef GENPF_PLUGIN_BUILD
// ------------------------------------------------------------------ K8 ancestor gather
// new_traces .= view(traces, parents) (resample.jl:60,102-104,114,169) over the window's columns.
// Output centric, coalesced 16-byte stores; reads follow the (monotone for stratified/residual) parents.
// Also applies update_weights! for the no-priority case (lw .= 0 | lse - log n, resample.jl:193-195,208-210).
struct GatherCols {
    const double *sf[2 * kMaxF];
    double *df[2 * kMaxF];
    const uint8_t *sb[2 * kMaxB];
    uint8_t *db[2 * kMaxB];
    int nf, nb;
};
static __global__ void __launch_bounds__(kThreads)
    k_gather(GatherCols c, int32_t *parents, int64_t n_in, int64_t n_out, int64_t tpf_out, const double *lw_src,
             double *lw_fill, const Stats *stats, int gate, int substate) {
    int64_t f, tile;
    blk_to_tile(tpf_out, f, tile);
    bool pass = false;  // gated and this filter keeps its population: identity parents, weights copied through
    if (stats) {
        const int kind = stats[f].invalid_kind;
        if (kind == 1 || kind == 4) pass = true;
        if (gate && !stats[f].do_resample) pass = true;
    }
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n_out - start);
    const int64_t obase = f * n_out + start;
    int64_t p[kItems];
    if (pass) {
        int32_t id[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            int e = tile_elem(k);
            id[k] = (int32_t)(start + e);
            p[k] = e < valid ? f * n_in + start + e : f * n_in;
        }
        store_tile<int32_t>(parents, obase, valid, id);
    } else {
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            int e = tile_elem(k);
            p[k] = e < valid ? f * n_in + (int64_t)parents[obase + e] : f * n_in;
        }
    }
    for (int col = 0; col < c.nf; ++col) {
        double v[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) v[k] = __ldg(c.sf[col] + p[k]);
        store_tile<double>(c.df[col], obase, valid, v);
    }
    for (int col = 0; col < c.nb; ++col) {
        uint8_t v[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) v[k] = __ldg(c.sb[col] + p[k]);
        store_tile_u8(c.db[col], obase, valid, v);
    }
    if (lw_fill) {
        const double val = substate ? stats[f].lse - log((double)n_in) : 0.0;
        double v[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) v[k] = pass ? __ldg(lw_src + p[k]) : val;
        store_tile<double>(lw_fill, obase, valid, v);
    }
}

// gather of one fp64 column through parents (priorities: lw[parents], and lineage resolution)
static __global__ void k_gather_f64(const double *src, const int32_t *idx, int64_t n_in, int64_t n_out, double *dst) {
    int64_t f = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x)
        dst[f * n_out + j] = src[f * n_in + idx[f * n_out + j]];
}
// lineage composition: anc[j] = parents_prev[anc[j]]
static __global__ void k_compose_lineage(int32_t *anc, const int32_t *parents_prev, int64_t n_prev, int64_t n_cur) {
    int64_t f = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_cur; j += (int64_t)gridDim.x * blockDim.x)
        anc[f * n_cur + j] = parents_prev[f * n_prev + anc[f * n_cur + j]];
}
static __global__ void k_iota32(int32_t *a, int64_t n, int64_t total) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x)
        a[j] = (int32_t)(j % n);
}
// read a field through an index column as fp64 (u8 promoted)
static __global__ void k_read_field(XSrc x, const int32_t *idx, int64_t n_src, int64_t n, double *out) {
    int64_t f = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = idx ? f * n_src + idx[f * n + j] : f * n + j;
        out[f * n + j] = x(s);
    }
}
static __global__ void k_write_field(const double *in, double *dcol, uint8_t *bcol, int64_t total) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
        if (dcol) dcol[j] = in[j];
        else bcol[j] = in[j] != 0.0 ? 1 : 0;
    }
}

#endif  // GENPF_PLUGIN_BUILD

}  // namespace genpf
