// fused.cuh -- one README loop iteration after the scan, in ONE kernel (stratified, resample taken):
//   ancestors (expand of the offspring counts) -> gather of the {t-2, t-1} window from the parent ->
//   mh rejuvenation of slice t-1 -> pf_update! to t -> new log-weight -> K1 partials of the new tile.
// Replaces, with identical results, the sequence k_expand + k_gather + k_mh + k_propagate
// (resample.jl:156-170,193-195; rejuvenate.jl:40-53; update.jl:12-25).  Identical because the same
// device functions, the same Philox streams keyed by the OUTPUT particle slot and the same tile
// partition are used; tests/test_gpu_filter.py::test_step_equals_separate_calls pins it.
//
// Traffic per particle: R O ~4, R window 18 (gathered, monotone), W parents 4, W slice t-1 9, W slice t 9,
// W lw 8  = 52 B against 117 B algorithmic (slice t-2 is never copied: it leaves the window at this step).
#pragma once
#include "filter.cuh"

namespace genpf {

struct StepArgs {
    ModelParams P_prev;  // aux of step t-1 (mh)
    ModelParams P_t;     // aux of step t   (update)
    const double *obs_prev_dev, *obs_t_dev;
    double obs_prev, obs_t;
    int64_t t;
    int mh_iters;
};

template <class Model, class Noise, typename IdxT>
static __global__ void __launch_bounds__(kThreads, 2)
    k_step_fused(StepArgs a, const IdxT *O, const IdxT *tile_last_O, Cols src_pp, Cols src_cur, Cols dst_cur,
                 Cols dst_new, int32_t *parents, double *lw_dst, int64_t n, int64_t tpf, Noise noise_mh,
                 Noise noise_up, uint8_t *accepts, unsigned long long *n_accept, Partials partials) {
    __shared__ ExpandSmem<IdxT> sm;
    __shared__ double smd[kWarps];
    __shared__ int smi[kWarps];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const int64_t i0 = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - i0);
    const int64_t obase = f * n + i0;
    int64_t p[kItems];
    block_expand<IdxT>(O + f * n, tile_last_O + f * tpf, n, tpf, i0, valid, sm, p);
    {
        int32_t q[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) q[k] = (int32_t)p[k];
        store_tile<int32_t>(parents, obase, valid, q);
    }
    const double obs_prev = a.obs_prev_dev ? a.obs_prev_dev[f] : a.obs_prev;
    const double obs_t = a.obs_t_dev ? a.obs_t_dev[f] : a.obs_t;
    const bool first = (a.t - 1) == 1;  // slice t-2 is the constant initial slice
    typename Model::Slice sc[kItems], sn[kItems];
    double v[kItems];
    uint8_t acc[kItems];
    double cnt = 0.0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int e = tile_elem(k);
        const bool live = e < valid;
        const int64_t s = f * n + p[k];
        typename Model::Slice pp, cur;
        if (first) {
            Model::initial(a.P_prev, pp);
        } else {
#pragma unroll
            for (int c = 0; c < Model::NF; ++c) pp.f[c] = __ldg(src_pp.f[c] + s);
#pragma unroll
            for (int c = 0; c < Model::NB; ++c) pp.b[c] = __ldg(src_pp.b[c] + s);
        }
#pragma unroll
        for (int c = 0; c < Model::NF; ++c) cur.f[c] = __ldg(src_cur.f[c] + s);
#pragma unroll
        for (int c = 0; c < Model::NB; ++c) cur.b[c] = __ldg(src_cur.b[c] + s);
        bool any_acc = false;
        for (int it = 0; it < a.mh_iters; ++it) {
            double U = 0.5, Z = 0.0, U3 = 1.0;
            Noise nz = noise_mh;
            nz.stream += (uint64_t)it;
            if (live) nz.get(obase + e, U, Z, U3);
            typename Model::Slice q;
            Model::transition(a.P_prev, a.t - 1, pp, q, U, Z);
            const double alpha = Model::obs_logpdf(a.P_prev, q, obs_prev) - Model::obs_logpdf(a.P_prev, cur, obs_prev);
            const bool ok = live && mh_accept(U3, alpha);
            if (ok) cur = q;
            any_acc = ok;  // flag of the last iteration, like k_mh launched once per iteration
            cnt += ok ? 1.0 : 0.0;
        }
        acc[k] = any_acc ? 1 : 0;
        double U = 0.5, Z = 0.0, U3;
        if (live) noise_up.get(obase + e, U, Z, U3);
        Model::transition(a.P_t, a.t, cur, sn[k], U, Z);
        sc[k] = cur;
        v[k] = live ? 0.0 + Model::obs_logpdf(a.P_t, sn[k], obs_t) : -INFINITY;
    }
    store_slices<Model>(dst_cur, obase, valid, sc);
    store_slices<Model>(dst_new, obase, valid, sn);
    store_tile<double>(lw_dst, obase, valid, v);
    if (accepts) store_tile_u8(accepts, obase, valid, acc);
    if (n_accept) {
        cnt = block_sum(cnt, smd);
        if (threadIdx.x == 0 && cnt > 0.0) atomicAdd(&n_accept[f], (unsigned long long)cnt);
    }
    emit_partials(v, partials, smd, smi);
}

}  // namespace genpf
