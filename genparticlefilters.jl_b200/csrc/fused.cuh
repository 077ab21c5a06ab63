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
// W lw 8, W e 8 = 60 B (ncu, final build: 57 B) against the 109 B the unfused kernels would move (slice t-2 is
// never copied: it leaves the window at this step).  The kernel is instruction-issue bound, not DRAM bound (ncu,
// profiles/), hence 256 threads x 8 particles at 64 registers (4 blocks/SM, no spills; per-thread overheads amortised): small
// per-thread footprint for occupancy, Philox + Box-Muller shared between the mh move and the update.
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
#ifndef GENPF_FUSED_MINB
#define GENPF_FUSED_MINB 4
#endif
#ifndef GENPF_FUSED_MINB256
#define GENPF_FUSED_MINB256 4  // 64 registers, no spills (3 -> 85 registers and 5 -> 51 were measured)
#endif
template <class Model, class Noise, typename IdxT, int MH>
GENPF_KERNEL void __launch_bounds__(kStateThreads, kStateThreads == 512 ? GENPF_FUSED_MINB : GENPF_FUSED_MINB256)
    k_step_fused(StepArgs a, const IdxT *O, const IdxT *tile_last_O, Cols src_pp, Cols src_cur, Cols dst_cur,
                 Cols dst_new, int32_t *parents, double *lw_dst, int64_t n, int64_t tpf, Noise noise,
                 uint8_t *accepts, unsigned long long *n_accept, Partials partials, double *ew,
                 const Stats *stats = nullptr, int gate = 0, const double *lw_src = nullptr) {
    pdl_enter();
    constexpr int T = kStateThreads, I = kTile / T;
    __shared__ ExpandSmem<IdxT> sm;
    __shared__ PartialSmem ps;
    __shared__ double smd[T / 32];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const int64_t i0 = tile * kTile;
    const int valid = (int)min((int64_t)kTile, n - i0);
    const int64_t obase = f * n + i0;
    // Per-filter decision of a batch (README.md:68-74 inside every view, test/resample.jl:130-162): a filter whose
    // ESS stayed above the threshold (or whose weights are NaN) is only updated -- identity ancestors, no mh move,
    // lw += increment -- while its neighbours in the same launch resample.  Block-uniform.
    bool pass = false;
    if (stats) {
        const int kind = stats[f].invalid_kind;
        pass = kind == 1 || kind == 4 || (gate && !stats[f].do_resample);
    }
    int32_t rel[I];
    int64_t s0 = i0;
    if (pass) {
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const int e = tile_elem<T>(k);
            rel[k] = e < valid ? e : 0;
        }
    } else {
        s0 = block_expand<IdxT, T>(O + f * n, tile_last_O + f * tpf, n, tpf, i0, valid, sm, rel);
    }
    const int64_t sbase = f * n + s0;
    const bool fast = (valid == kTile) && ((obase & 1) == 0);  // every column base is 256-B aligned
    const double obs_prev = a.obs_prev_dev ? a.obs_prev_dev[f] : a.obs_prev;
    const double obs_t = a.obs_t_dev ? a.obs_t_dev[f] : a.obs_t;
    const bool first = (a.t - 1) == 1;  // slice t-2 is the constant initial slice
    const int iters = pass ? 0 : (MH >= 0 ? MH : a.mh_iters);
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
            if (MH == 0 || first) {
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
            double U_mh = 0.5, Z_mh = 0.0, U_acc = 1.0, U_up, Z_up;
            // dead slots draw too (no divergence); noise columns are read at an in-range slot instead
            const int64_t ni = (Noise::kIndexed && !live) ? obase : obase + e;
            if (MH == 0) noise.up(ni, U_up, Z_up);  // no mh move: only the update's draws
            else noise.both(ni, U_mh, Z_mh, U_acc, U_up, Z_up);
            bool ok = false;
            for (int it = 0; it < iters; ++it) {
                if (MH < 0 && it > 0) noise.mh(ni, it, U_mh, Z_mh, U_acc);
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
            // update_weights! left lw = 0 on a resampled filter (resample.jl:193-195); a passing one keeps its weight
            const double lw0 = pass ? __ldg(lw_src + s) : 0.0;
            v[k] = live ? lw0 + Model::obs_logpdf(a.P_t, sn[c2], obs_t) : -INFINITY;
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
    emit_partials<T>(v, partials, ps, -1, ew ? ew + obase : nullptr, valid);
}

#ifndef GENPF_PLUGIN_BUILD
// ------------------------------------------------------------------ multi-GPU: source-side push
// Particle sharding: rank r owns global particle slots [r*n_loc, (r+1)*n_loc).  After the shard-aware scan the
// local O_k are GLOBAL cumulative offspring counts, so this rank's particles parent the contiguous global
// output range [out_begin, out_end).  The kernel computes those offspring (gather from the LOCAL parent, mh,
// update -- exactly k_step_fused's arithmetic, Philox keyed by the GLOBAL output slot, so the population is
// independent of the number of GPUs) and stores them straight into the OWNER's buffers through peer-mapped
// pointers (NVLink P2P stores; n_loc is a multiple of the 2048 tile so a tile never straddles two owners).
// No all-to-all, no staging: stores are fire-and-forget, the ranks meet at one stream-ordered barrier after.
struct PeerDst {
    Cols dst_cur[kMaxPeers], dst_new[kMaxPeers];
    double *lw[kMaxPeers];
    int32_t *parents[kMaxPeers];
    Partials part[kMaxPeers];  // the owner's K1 partial arrays (full tiles are reduced by their producer)
    double *ew[kMaxPeers];     // the owner's e_i = exp(lw_i - m_tile) column
};
// [begin, end) of the global outputs rank `rank` parents, from the all-gathered closing counts (monotone by
// construction: exact cover of [0, n_total), mirrored on the host in sharded.py::exchange_plan)
__device__ __forceinline__ void shard_range(const long long *oend_all, int world, int rank, long long n_total,
                                            long long &begin, long long &end) {
    begin = 0;
    for (int g = 0; g < rank; ++g) begin = max(begin, oend_all[g]);
    end = max(begin, oend_all[rank]);
    if (rank == world - 1) end = n_total;
}
template <class Model, class Noise, typename IdxT, int MH>
static __global__ void __launch_bounds__(kStateThreads, 4)
    k_step_push(StepArgs a, const IdxT *O, const IdxT *tile_last_O, Cols src_pp, Cols src_cur, PeerDst peer,
                const long long *oend_all, int world, int64_t n_loc, int64_t tpf_loc, int rank, Noise noise,
                const Stats *stats = nullptr) {
    pdl_enter();
    constexpr int T = kStateThreads, I = kTile / T;
    __shared__ ExpandSmem<IdxT> sm;
    __shared__ PartialSmem ps;
    if (stats) {  // invalid weights (NaN / +Inf): nothing is resampled, the population stays as it is
        const int kind = stats[0].invalid_kind;
        if (kind == 1 || kind == 4) return;
    }
    long long out_begin, out_end;
    shard_range(oend_all, world, rank, (long long)world * n_loc, out_begin, out_end);
    // grid ~ one block per local tile (+2): balanced shards do one tile per block, a shard that parents more
    // than its share loops (block-uniform trip count)
    for (int64_t tile = out_begin / kTile + blockIdx.x; tile * kTile < out_end; tile += gridDim.x) {
    const int64_t t0 = tile * kTile;
    const int64_t i0 = max(t0, (int64_t)out_begin);
    const int64_t i1 = min(t0 + (int64_t)kTile, (int64_t)out_end);
    if (i1 <= i0) continue;
    const int valid = (int)(i1 - i0);
    const int owner = (int)(t0 / n_loc);
    const int64_t lbase = i0 - (int64_t)owner * n_loc;  // slot of output i0 inside the owner's shard
    int32_t rel[I];
    const int64_t guess = max((int64_t)0, i0 / kTile - (int64_t)rank * tpf_loc);  // balanced shards: parent ~ output
    const int64_t s0 = block_expand<IdxT, T>(O, tile_last_O, n_loc, tpf_loc, i0, valid, sm, rel, (IdxT)out_begin, guess);
    const bool fast = (valid == kTile);  // interior tiles are full and 2048-aligned
    const double obs_prev = a.obs_prev, obs_t = a.obs_t;
    const bool first = (a.t - 1) == 1;
    const int iters = MH >= 0 ? MH : a.mh_iters;
    const Cols dst_cur = peer.dst_cur[owner], dst_new = peer.dst_new[owner];
    double *lw_dst = peer.lw[owner];
    int32_t *par_dst = peer.parents[owner];
    const int64_t gsrc = (int64_t)rank * n_loc + s0;  // global index of local source s0
    double vall[I];
#pragma unroll
    for (int j = 0; j < I / 2; ++j) {
        typename Model::Slice sc[2], sn[2];
        double v[2];
        const int e0 = (j * T + threadIdx.x) * 2;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            const int k = 2 * j + c2;
            const int e = e0 + c2;
            const bool live = e < valid;
            const int64_t s = s0 + rel[k];
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
            const int64_t ni = (Noise::kIndexed && !live) ? i0 : i0 + e;  // GLOBAL output slot
            noise.both(ni, U_mh, Z_mh, U_acc, U_up, Z_up);
            for (int it = 0; it < iters; ++it) {
                if (MH < 0 && it > 0) noise.mh(ni, it, U_mh, Z_mh, U_acc);
                typename Model::Slice q;
                Model::transition(a.P_prev, a.t - 1, pp, q, U_mh, Z_mh);
                const double alpha =
                    Model::obs_logpdf(a.P_prev, q, obs_prev) - Model::obs_logpdf(a.P_prev, cur, obs_prev);
                if (live && mh_accept(U_acc, alpha)) cur = q;
            }
            Model::transition(a.P_t, a.t, cur, sn[c2], U_up, Z_up);
            sc[c2] = cur;
            v[c2] = 0.0 + Model::obs_logpdf(a.P_t, sn[c2], obs_t);
        }
        store_pair<int32_t>(par_dst + lbase, e0, valid, fast, (int32_t)(gsrc + rel[2 * j]), (int32_t)(gsrc + rel[2 * j + 1]));
#pragma unroll
        for (int c = 0; c < Model::NF; ++c) {
            store_pair<double>(dst_cur.f[c] + lbase, e0, valid, fast, sc[0].f[c], sc[1].f[c]);
            store_pair<double>(dst_new.f[c] + lbase, e0, valid, fast, sn[0].f[c], sn[1].f[c]);
        }
#pragma unroll
        for (int c = 0; c < Model::NB; ++c) {
            store_pair<uint8_t>(dst_cur.b[c] + lbase, e0, valid, fast, sc[0].b[c], sc[1].b[c]);
            store_pair<uint8_t>(dst_new.b[c] + lbase, e0, valid, fast, sn[0].b[c], sn[1].b[c]);
        }
        store_pair<double>(lw_dst + lbase, e0, valid, fast, v[0], v[1]);
        vall[2 * j] = v[0];
        vall[2 * j + 1] = v[1];
    }
    // a full tile is reduced here and its K1 partial stored into the owner's arrays; the (at most world+1)
    // tiles split between two producers are reduced by their owner after the barrier (k_reduce_boundary)
    if (fast) emit_partials<T>(vall, peer.part[owner], ps, (t0 - (int64_t)owner * n_loc) / kTile,
                               peer.ew[owner] + (t0 - (int64_t)owner * n_loc), kTile);
    __syncthreads();  // shared staging is reused by the next tile of this block
    }
}

// small shards (one finalize block, no k_chunk_combine to ride on): the statistics exchange as a kernel of its own
static __global__ void k_xchg_stats_combine(const Stats *local, XchgLink link, int64_t n_total, Stats *stats,
                                            double *shard_info, double *lml_accum, StratArgs strat, long long *oend_out) {
    pdl_enter();
    const int kind = local->invalid_kind;
    xchg_stats_combine(link, local->M, (kind == 1 || kind == 4) ? NAN : local->S, local->S2, n_total, stats, shard_info,
                       lml_accum, &strat, oend_out);
}

// owner side: K1 partials of the tiles that two producers shared (global tile index = a range boundary)
static __global__ void __launch_bounds__(kReduceThreads)
    k_reduce_boundary(LwSrc src, const long long *oend_all, int world, int rank, int64_t n_loc, Partials out,
                      double *ew, XchgLink link) {
    pdl_enter();
    constexpr int T = kReduceThreads;
    __shared__ PartialSmem ps;
    // The step's closing barrier, taken by the first reader of the pushed population.  Stream order puts the push
    // kernel's P2P stores before this kernel, so block 0 first tells every peer "my offspring have left"; every
    // block then waits for all producers' flags.  Everything later on this stream (the next finalize reads the
    // pushed K1 partials) is ordered behind this kernel.
    if (link.world > 0 && (int)threadIdx.x < link.world) {
        Xchg *mine = link.peers.x[link.rank];
        if (blockIdx.x == 0) {
            __threadfence_system();
            *(volatile unsigned long long *)&link.peers.x[threadIdx.x]->flag_done[link.rank] = link.epoch;
        }
        xchg_wait(&mine->flag_done[threadIdx.x], link.epoch, &mine->error);
    }
    __syncthreads();
    long long b, e;
    shard_range(oend_all, world, (int)blockIdx.x, (long long)world * n_loc, b, e);
    if (b % kTile == 0) return;                       // boundary on a tile edge: nothing was split
    const int64_t gt = b / kTile;                     // the split tile
    if (gt / (n_loc / kTile) != rank) return;         // not mine
    for (int g = 0; g < (int)blockIdx.x; ++g) {       // several boundaries can fall into one tile: reduce it once
        long long bg, eg;
        shard_range(oend_all, world, g, (long long)world * n_loc, bg, eg);
        if (bg % kTile != 0 && bg / kTile == gt) return;
    }
    const int64_t lt = gt - (int64_t)rank * (n_loc / kTile);
    double v[kTile / T];
    load_tile<T>(src, lt * kTile, kTile, v, -INFINITY);
    emit_partials<T>(v, out, ps, lt, ew + lt * kTile, kTile);
}

// global statistics from the all-gathered per-shard (max, sum e, sum e^2); one thread (world <= 8)
static __global__ void k_shard_combine(const double *gathered, int world, int rank, int64_t n_total, Stats *stats,
                                       double *shard_info, double *lml_accum) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double M = -INFINITY;
    bool nan = false;
    for (int g = 0; g < world; ++g) {
        double m = gathered[3 * g];
        if (isnan(m) || isnan(gathered[3 * g + 1])) nan = true;
        M = fmax(M, m);
    }
    double S = 0.0, S2 = 0.0, prefix = 0.0, mine = 0.0;
    const bool finite = M > -INFINITY && M < INFINITY;
    for (int g = 0; g < world; ++g) {
        double m = gathered[3 * g];
        double sc = (finite && m > -INFINITY) ? exp(m - M) : 0.0;
        S += gathered[3 * g + 1] * sc;
        S2 += gathered[3 * g + 2] * (sc * sc);
    }
    for (int g = 0; g < world; ++g) {
        double m = gathered[3 * g];
        double sc = (finite && m > -INFINITY) ? exp(m - M) : 0.0;
        double share = gathered[3 * g + 1] * sc / S;
        if (g < rank) prefix += share;
        if (g == rank) mine = share;
    }
    int kind = 0;
    if (nan) kind = 1;
    else if (M == -INFINITY) kind = 2;
    else if (M == INFINITY || isnan(S)) kind = 4;
    else if (S == 0.0) kind = 3;
    Stats st;
    st.M = M; st.S = S; st.S2 = S2;
    st.lse = (M == -INFINITY) ? -INFINITY : M + log(S);
    st.ess = S * S / S2;
    st.invalid_kind = kind;
    st.do_resample = (kind == 1 || kind == 4) ? 0 : 1;
    stats[0] = st;
    if (kind == 2 || kind == 3) {  // uniform fallback (utils.jl:123-133): every shard carries 1/world
        prefix = (double)rank / (double)world;
        mine = 1.0 / (double)world;
    }
    shard_info[0] = prefix;
    shard_info[1] = mine;  // local offsets are normalised to 1 over the shard: scale by the shard's global share
    if (lml_accum && st.do_resample) lml_accum[0] += st.lse - log((double)n_total);
}

static __global__ void k_shard_oend(const int32_t *tile_last_O, int64_t tpf, long long *oend_local) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *oend_local = (long long)tile_last_O[tpf - 1];
}

#endif  // GENPF_PLUGIN_BUILD

}  // namespace genpf
