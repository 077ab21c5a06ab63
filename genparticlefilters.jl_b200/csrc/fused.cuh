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

template <typename T>
__device__ __forceinline__ void store_pair(T *col, int64_t base, int64_t valid, int j, T a, T b) {
    T *p = col + base;
    const int64_t e = (int64_t)(j * kThreads + threadIdx.x) * 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & (2 * sizeof(T) - 1)) == 0);
    if (vec_ok && e + 1 < valid) {
        struct alignas(2 * sizeof(T)) V2 { T a, b; };
        *reinterpret_cast<V2 *>(p + e) = V2{a, b};
    } else {
        if (e < valid) p[e] = a;
        if (e + 1 < valid) p[e + 1] = b;
    }
}

template <class Model, class Noise, typename IdxT>
static __global__ void __launch_bounds__(kThreads, 2)
    k_step_fused(StepArgs a, const IdxT *O, const IdxT *tile_last_O, Cols src_pp, Cols src_cur, Cols dst_cur,
                 Cols dst_new, int32_t *parents, double *lw_dst, int64_t n, int64_t tpf, Noise noise,
                 uint8_t *accepts, unsigned long long *n_accept, Partials partials) {
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
    const double obs_prev = a.obs_prev_dev ? a.obs_prev_dev[f] : a.obs_prev;
    const double obs_t = a.obs_t_dev ? a.obs_t_dev[f] : a.obs_t;
    const bool first = (a.t - 1) == 1;  // slice t-2 is the constant initial slice
    double v[kItems];
    double cnt = 0.0;
#pragma unroll
    for (int j = 0; j < kVecs; ++j) {
        typename Model::Slice sc[2], sn[2];
        uint8_t acc[2];
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            const int k = 2 * j + c2;
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
            double U_mh = 0.5, Z_mh = 0.0, U_acc = 1.0, U_up = 0.5, Z_up = 0.0;
            if (live) noise.both(obase + e, U_mh, Z_mh, U_acc, U_up, Z_up);
            bool ok = false;
            for (int it = 0; it < a.mh_iters; ++it) {
                if (it > 0 && live) noise.mh(obase + e, it, U_mh, Z_mh, U_acc);
                typename Model::Slice q;
                Model::transition(a.P_prev, a.t - 1, pp, q, U_mh, Z_mh);
                const double alpha =
                    Model::obs_logpdf(a.P_prev, q, obs_prev) - Model::obs_logpdf(a.P_prev, cur, obs_prev);
                ok = live && mh_accept(U_acc, alpha);
                if (ok) cur = q;
                cnt += ok ? 1.0 : 0.0;
            }
            acc[c2] = ok ? 1 : 0;  // flag of the last iteration, like k_mh launched once per iteration
            Model::transition(a.P_t, a.t, cur, sn[c2], U_up, Z_up);
            sc[c2] = cur;
            v[k] = live ? 0.0 + Model::obs_logpdf(a.P_t, sn[c2], obs_t) : -INFINITY;
        }
        store_pair<int32_t>(parents, obase, valid, j, (int32_t)p[2 * j], (int32_t)p[2 * j + 1]);
#pragma unroll
        for (int c = 0; c < Model::NF; ++c) {
            store_pair<double>(dst_cur.f[c], obase, valid, j, sc[0].f[c], sc[1].f[c]);
            store_pair<double>(dst_new.f[c], obase, valid, j, sn[0].f[c], sn[1].f[c]);
        }
#pragma unroll
        for (int c = 0; c < Model::NB; ++c) {
            store_pair<uint8_t>(dst_cur.b[c], obase, valid, j, sc[0].b[c], sc[1].b[c]);
            store_pair<uint8_t>(dst_new.b[c], obase, valid, j, sn[0].b[c], sn[1].b[c]);
        }
        store_pair<double>(lw_dst, obase, valid, j, v[2 * j], v[2 * j + 1]);
        if (accepts) store_pair<uint8_t>(accepts, obase, valid, j, acc[0], acc[1]);
    }
    if (n_accept) {
        cnt = block_sum(cnt, smd);
        if (threadIdx.x == 0 && cnt > 0.0) atomicAdd(&n_accept[f], (unsigned long long)cnt);
    }
    emit_partials(v, partials, smd, smi);
}

}  // namespace genpf
