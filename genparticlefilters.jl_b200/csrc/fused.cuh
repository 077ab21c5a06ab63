// fused.cuh -- one README loop iteration after the scan, in ONE kernel (stratified, resample taken):
//   ancestors (expand of the offspring counts) -> gather of the {t-2, t-1} window from the parent ->
//   mh rejuvenation of slice t-1 -> pf_update! to t -> new log-weight -> K1 partials of the new tile.
// Replaces, with identical results, the sequence k_expand + k_gather + k_mh + k_propagate
// (resample.jl:156-170,193-195; rejuvenate.jl:40-53; update.jl:12-25).  Identical because the same
// device functions, the same Philox streams keyed by the OUTPUT particle slot and the same tile
// partition / block size (kStateThreads) are used; tests/test_gpu_filter.py::test_step_equals_separate_calls
// pins it.
//
// Traffic per particle: R O ~4, R window 18 (gathered, monotone), W parents 4, W slice t-1 9, W slice t 9,
// W lw 8  = 52 B against 117 B algorithmic (slice t-2 is never copied: it leaves the window at this step).
// The kernel is instruction-issue bound, not DRAM bound (ncu, profiles/), hence 512 threads x 4 particles:
// small per-thread footprint for occupancy, Philox + Box-Muller shared between the mh move and the update.
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

// store the pair (a, b) at tile elements (e, e+1); `fast` is block-uniform (full, aligned tile)
template <typename X>
__device__ __forceinline__ void store_pair(X *p, int e, int valid, bool fast, X a, X b) {
    if (fast) {
        struct alignas(2 * sizeof(X)) V2 { X a, b; };
        *reinterpret_cast<V2 *>(p + e) = V2{a, b};
    } else {
        if (e < valid) p[e] = a;
        if (e + 1 < valid) p[e + 1] = b;
    }
}

// MH: number of mh iterations known at compile time (1 = the README configuration) or -1 = a.mh_iters
template <class Model, class Noise, typename IdxT, int MH>
static __global__ void __launch_bounds__(kStateThreads, 2)
    k_step_fused(StepArgs a, const IdxT *O, const IdxT *tile_last_O, Cols src_pp, Cols src_cur, Cols dst_cur,
                 Cols dst_new, int32_t *parents, double *lw_dst, int64_t n, int64_t tpf, Noise noise,
                 uint8_t *accepts, unsigned long long *n_accept, Partials partials) {
    constexpr int T = kStateThreads, I = kTile / T;
    __shared__ ExpandSmem<IdxT> sm;
    __shared__ double smd[2 * (T / 32)];
    __shared__ int smi[T / 32];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const int64_t i0 = tile * kTile;
    const int valid = (int)min((int64_t)kTile, n - i0);
    const int64_t obase = f * n + i0;
    int32_t rel[I];
    const int64_t s0 = block_expand<IdxT, T>(O + f * n, tile_last_O + f * tpf, n, tpf, i0, valid, sm, rel);
    const int64_t sbase = f * n + s0;
    const bool fast = (valid == kTile) && ((obase & 1) == 0);  // every column base is 256-B aligned
    const double obs_prev = a.obs_prev_dev ? a.obs_prev_dev[f] : a.obs_prev;
    const double obs_t = a.obs_t_dev ? a.obs_t_dev[f] : a.obs_t;
    const bool first = (a.t - 1) == 1;  // slice t-2 is the constant initial slice
    const int iters = MH >= 0 ? MH : a.mh_iters;
    double v[I];
    double cnt = 0.0;
#pragma unroll
    for (int j = 0; j < I / 2; ++j) {
        typename Model::Slice sc[2], sn[2];
        uint8_t acc[2];
        const int e0 = (j * T + threadIdx.x) * 2;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            const int k = 2 * j + c2;
            const int e = e0 + c2;
            const bool live = e < valid;
            const int64_t s = sbase + rel[k];  // rel is 0 for dead slots: always a valid address
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
            double U_mh, Z_mh, U_acc, U_up, Z_up;
            noise.both(obase + e, U_mh, Z_mh, U_acc, U_up, Z_up);  // dead slots draw too: no divergence
            bool ok = false;
            for (int it = 0; it < iters; ++it) {
                if (MH < 0 && it > 0) noise.mh(obase + e, it, U_mh, Z_mh, U_acc);
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
        store_pair<int32_t>(parents + obase, e0, valid, fast, (int32_t)(s0 + rel[2 * j]), (int32_t)(s0 + rel[2 * j + 1]));
#pragma unroll
        for (int c = 0; c < Model::NF; ++c) {
            store_pair<double>(dst_cur.f[c] + obase, e0, valid, fast, sc[0].f[c], sc[1].f[c]);
            store_pair<double>(dst_new.f[c] + obase, e0, valid, fast, sn[0].f[c], sn[1].f[c]);
        }
#pragma unroll
        for (int c = 0; c < Model::NB; ++c) {
            store_pair<uint8_t>(dst_cur.b[c] + obase, e0, valid, fast, sc[0].b[c], sc[1].b[c]);
            store_pair<uint8_t>(dst_new.b[c] + obase, e0, valid, fast, sn[0].b[c], sn[1].b[c]);
        }
        store_pair<double>(lw_dst + obase, e0, valid, fast, v[2 * j], v[2 * j + 1]);
        if (accepts) store_pair<uint8_t>(accepts + obase, e0, valid, fast, acc[0], acc[1]);
    }
    if (n_accept) {
        cnt = block_sum<T>(cnt, smd);
        if (threadIdx.x == 0 && cnt > 0.0) atomicAdd(&n_accept[f], (unsigned long long)cnt);
    }
    emit_partials<T>(v, partials, smd, smi);
}

}  // namespace genpf
