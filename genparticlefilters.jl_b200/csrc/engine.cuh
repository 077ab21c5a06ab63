// engine.cuh -- host-side launch sequences over the weight-vector kernels, shared by the host-array
// ABI (abi_host.cu) and the device-resident filter (abi_filter.cu).  Everything here takes DEVICE
// pointers and a stream and never synchronises.
#pragma once
#include <algorithm>

#include "filter.cuh"
#include "host.hpp"

namespace genpf {

// scratch owned by a workspace / filter; all grow-only
struct Scratch {
    DevBuf part[3][4];  // three partial sets (lw, selection source, ratio d): m, s, s2, flags
    DevBuf stats;       // Stats[4 * nf]: [0] lw  [1] selection  [2] ratio d  [3] sorted selection
    DevBuf tile_off, tile_scale, W, O, tile_last, guide, guide_O, guide_tile_last;
    DevBuf resid_c, resid_r, resid_coff, resid_roff, resid_rtot;
    DevBuf sort_tmp, sorted_keys, order, prio_col;
    DevBuf moment_partial, moment_out;
    DevBuf misc, chunk_stats, chunk_info;
    static constexpr int64_t kChunkTiles = genpf::kChunkTiles;  // tiles finalised by one block (2^20 particles)
    // {prefix, scale} pairs of the last finalize of a filter with more than kChunkTiles tiles (else null)
    const double *chunk_info_ptr(int64_t n) const { return ceil_div(n, kTile) > kChunkTiles ? chunk_info.as<double>() : nullptr; }
    int64_t cap_n = -1, cap_nf = -1;

    int32_t ensure(int64_t n, int64_t nf) {
        const int64_t tpf = ceil_div(n, kTile);
        const size_t np = (size_t)(tpf * nf);
        for (int k = 0; k < 3; ++k) {
            GENPF_TRY(part[k][0].ensure(np * 8));
            GENPF_TRY(part[k][1].ensure(np * 8));
            GENPF_TRY(part[k][2].ensure(np * 8));
            GENPF_TRY(part[k][3].ensure(np * 4));
        }
        GENPF_TRY(stats.ensure(sizeof(Stats) * 4 * (size_t)nf));
        GENPF_TRY(tile_off.ensure(np * 8));
        GENPF_TRY(tile_scale.ensure(np * 8));
        GENPF_TRY(moment_partial.ensure(np * 8));
        GENPF_TRY(moment_out.ensure((size_t)nf * 8 * 2));
        return GENPF_OK;
    }
    Partials partials(int k) { return Partials{part[k][0].as<double>(), part[k][1].as<double>(), part[k][2].as<double>(), part[k][3].as<int>()}; }
    Stats *st(int k, int64_t nf) { return stats.as<Stats>() + (size_t)k * nf; }
    void release() {
        for (auto &a : part) for (auto &b : a) b.release();
        for (DevBuf *b : {&stats, &tile_off, &tile_scale, &W, &O, &tile_last, &guide, &guide_O, &guide_tile_last, &resid_c, &resid_r, &resid_coff, &resid_roff, &resid_rtot,
                          &sort_tmp, &sorted_keys, &order, &prio_col, &moment_partial, &moment_out, &misc, &chunk_stats, &chunk_info})
            b->release();
    }
};

inline int32_t launch_reduce(cudaStream_t s, LwSrc src, int64_t n, int64_t nf, Partials part, double *ew = nullptr) {
    const int64_t tpf = ceil_div(n, kTile);
    GENPF_LAUNCH(k_reduce, dim3((unsigned)tpf, (unsigned)nf), kReduceThreads, s, src, n, tpf, part, ew);
    return GENPF_OK;
}
// link / n_total / shard_info (a shard of a multi-GPU population): the large-filter path's combine kernel also runs the
// statistics exchange (returns true through *exchanged); small shards exchange with a kernel of their own
inline int32_t launch_finalize(cudaStream_t s, Scratch &sc, Partials part, int64_t n, int64_t nf, Stats *st,
                               double *tile_off, double ess_frac, double *lml_accum,
                               const XchgLink *link = nullptr, int64_t n_total = 0, double *shard_info = nullptr,
                               bool *exchanged = nullptr, double *xchg_lml = nullptr, const StratArgs *strat = nullptr,
                               long long *oend_out = nullptr) {
    const int64_t tpf = ceil_div(n, kTile);
    if (exchanged) *exchanged = false;
    if (tpf <= 64) {
        GENPF_LAUNCH_PDL((k_finalize_fast<64, 1>), dim3(1, (unsigned)nf), 64, s, part, n, tpf, st, tile_off, ess_frac, lml_accum, tpf,
                     sc.tile_scale.as<double>());
    } else if (tpf <= Scratch::kChunkTiles) {
        GENPF_LAUNCH_PDL((k_finalize_fast<512, 1>), dim3(1, (unsigned)nf), 512, s, part, n, tpf, st, tile_off, ess_frac,
                     lml_accum, tpf, sc.tile_scale.as<double>());
    } else {
        // large filter: one block per chunk of 512 tiles (one tile per thread: the finalize is latency bound, so all
        // chunks run in parallel), then a one-warp combine (statistics, validity, lml, chunk {prefix, scale})
        const int64_t nchunks = ceil_div(tpf, Scratch::kChunkTiles);
        GENPF_TRY(sc.chunk_stats.ensure(sizeof(Stats) * (size_t)(nchunks * nf)));
        GENPF_TRY(sc.chunk_info.ensure(16 * (size_t)(nchunks * nf)));
        GENPF_LAUNCH_PDL((k_finalize_fast<512, 1>), dim3((unsigned)nchunks, (unsigned)nf), 512, s, part, n, tpf,
                     sc.chunk_stats.as<Stats>(), tile_off, -1.0, (double *)nullptr, Scratch::kChunkTiles,
                     sc.tile_scale.as<double>());
        const XchgLink lk = link ? *link : XchgLink{{}, 0, 0, 0};
        GENPF_LAUNCH_PDL(k_chunk_combine, (unsigned)nf, 32, s, (const Stats *)sc.chunk_stats.as<Stats>(), (int)nchunks, n,
                     Scratch::kChunkTiles * (int64_t)kTile, st, sc.chunk_info.as<double>(), ess_frac,
                     lk.world > 0 ? xchg_lml : lml_accum, lk, n_total, shard_info, strat ? *strat : StratArgs{}, oend_out);
        if (exchanged) *exchanged = lk.world > 0;
    }
    return GENPF_OK;
}

inline StratArgs make_strat(UniSrc uni, int64_t n) {
    StratArgs a;
    a.uni = uni;
    a.step = 1.0 / (double)n;
    a.n = n;
    a.pow2 = (n & (n - 1)) == 0;
    a.guide = 0;
    a.tol = (double)n * 0x1.0p-50;
    return a;
}

// cumulative offspring counts O_k of a stratified resample (no W output): the hot kernel when the strata are the
// library's Philox draws, the general k_scan for supplied uniform columns
template <typename IdxT>
inline int32_t launch_scan_counts(cudaStream_t s, LwSrc sel, int64_t n, int64_t tpf, int64_t nf, const Stats *st,
                                  const double *tile_off, IdxT *O, IdxT *tile_last, const StratArgs &strat, int gate,
                                  const double *shard_info, int64_t global_base, const double *chunk_info,
                                  const double *ew, const double *tile_scale, const long long *oend_pin = nullptr,
                                  int shard_rank = 0) {
    const dim3 grid((unsigned)tpf, (unsigned)nf);
    if (!strat.uni.col && !strat.guide && strat.pow2) {
        if (ew) {
            GENPF_LAUNCH_PDL((k_scan_hot<IdxT, true>), grid, kHotThreads, s, sel, n, tpf, st, tile_off, O, tile_last, strat, gate,
                         shard_info, global_base, chunk_info, ew, tile_scale, oend_pin, shard_rank);
        } else {
            GENPF_LAUNCH_PDL((k_scan_hot<IdxT, false>), grid, kHotThreads, s, sel, n, tpf, st, tile_off, O, tile_last, strat, gate,
                         shard_info, global_base, chunk_info, ew, tile_scale, oend_pin, shard_rank);
        }
        return GENPF_OK;
    }
    GENPF_LAUNCH_PDL((k_scan<IdxT>), grid, kScanThreads, s, sel, n, tpf, st, tile_off, WTables{nullptr}, O, tile_last, strat, gate,
                 shard_info, global_base, chunk_info, Scratch::kChunkTiles, ew, tile_scale, oend_pin, shard_rank);
    return GENPF_OK;
}

// guide-table size for the inverse-CDF lookups: two particles per bucket measured best at 2^24 (bucket sizes
// of 1..32 particles stay within 15 %: a bigger table misses L2 more often, a smaller one lengthens the bracket)
inline int64_t guide_buckets(int64_t n) {
    const int64_t b = n >> 1;
    return b > 0 ? b : 1;
}

// Ancestor selection given selection statistics `st_sel` + `tile_off` already computed for `sel`
// (unsorted order).  Writes parents (0-based + out_base), local to each filter.
//   stratified : scan -> offspring counts O -> expand                      (resample.jl:156-170)
//   multinomial: scan -> W -> inverse-CDF search per uniform                (resample.jl:59)
//   residual   : counts + residual scan -> expand + search on the tail      (resample.jl:96-115)
template <typename IdxT, typename OutT>
int32_t select_ancestors_t(cudaStream_t s, Scratch &sc, int method, LwSrc sel, int64_t n_in, int64_t n_out,
                           int64_t nf, Stats *st_sel, UniSrc uni, uint32_t flags, OutT *parents, int64_t out_base,
                           int gate, const double *ew = nullptr, LwFill fill = LwFill{nullptr, nullptr, 0}) {
    const int64_t tpf_in = ceil_div(n_in, kTile), tpf_out = ceil_div(n_out, kTile);
    GENPF_TRY(sc.O.ensure((size_t)(n_in * nf) * sizeof(IdxT)));
    GENPF_TRY(sc.tile_last.ensure((size_t)(tpf_in * nf) * sizeof(IdxT)));
    IdxT *O = sc.O.as<IdxT>();
    IdxT *tile_last = sc.tile_last.as<IdxT>();
    double *tile_off = sc.tile_off.as<double>();
    if (method == GENPF_STRATIFIED) {
        const int32_t *order = nullptr;
        if (flags & GENPF_SORT_PARTICLES) {
            // sortperm(log_priorities, rev=true): materialise keys, sort per filter, re-reduce in sorted order
            GENPF_TRY(sc.prio_col.ensure((size_t)(n_in * nf) * 8));
            GENPF_TRY(sc.sorted_keys.ensure((size_t)(n_in * nf) * 8));
            GENPF_TRY(sc.order.ensure((size_t)(n_in * nf) * 4));
            if (gate) return fail(GENPF_ERR_UNSUPPORTED, "gated resample with sort_particles is not supported");
            double *keys = sc.prio_col.as<double>();
            GENPF_LAUNCH(k_materialize, grid_1d(n_in * nf), 256, s, sel, n_in * nf, keys);
            GENPF_TRY(sort_desc_stable_batched(keys, n_in, nf, sc.sorted_keys.as<double>(), sc.order.as<int32_t>(),
                                               sc.sort_tmp, s));
            LwSrc sorted{sc.sorted_keys.as<double>(), 1.0};
            Stats *st_sorted = sc.st(3, nf);
            GENPF_TRY(launch_reduce(s, sorted, n_in, nf, sc.partials(1)));
            GENPF_TRY(launch_finalize(s, sc, sc.partials(1), n_in, nf, st_sorted, tile_off, -1.0, nullptr));
            sel = sorted;
            st_sel = st_sorted;
            order = sc.order.as<int32_t>();
        }
        StratArgs strat = make_strat(uni, n_in);
        const double *ew_use = order ? nullptr : ew;  // e_i is stored in particle order
        GENPF_TRY(launch_scan_counts<IdxT>(s, sel, n_in, tpf_in, nf, st_sel, tile_off, O, tile_last, strat, gate,
                                           nullptr, 0, sc.chunk_info_ptr(n_in), ew_use, sc.tile_scale.as<double>()));
        GENPF_LAUNCH((k_expand<IdxT, OutT>), dim3((unsigned)tpf_out, (unsigned)nf), kThreads, s, O, tile_last, n_in, n_out, tpf_out, order,
                     parents, out_base, st_sel, gate, 0, fill);
    } else if (method == GENPF_MULTINOMIAL) {
        GENPF_TRY(sc.W.ensure((size_t)(n_in * nf) * 8));
        const int64_t B = guide_buckets(n_in), tpf_b = ceil_div(B, kTile);
        GENPF_TRY(sc.guide.ensure((size_t)(B * nf) * (sizeof(IdxT) == 4 ? sizeof(GuideRec) : sizeof(IdxT))));
        WTables wt{sc.W.as<double>()};
        StratArgs gs = make_strat(uni, n_in);
        gs.guide = B;  // O = guide-table counts of the weight CDF
        IdxT *G = sc.guide.as<IdxT>();
        GENPF_LAUNCH((k_scan<IdxT>), dim3((unsigned)tpf_in, (unsigned)nf), kScanThreads, s, sel, n_in, tpf_in, st_sel, tile_off,
                     wt, O, tile_last, gs, gate, (const double *)nullptr, (int64_t)0,
                     sc.chunk_info_ptr(n_in), Scratch::kChunkTiles);
        const dim3 lgrid((unsigned)ceil_div(n_out, (int64_t)kThreads * kLookupItems), (unsigned)nf);
        if constexpr (sizeof(IdxT) == 4) {  // one 32-byte record per bucket: a draw is one sector
            GENPF_LAUNCH(k_guide_records, dim3((unsigned)tpf_b, (unsigned)nf), kThreads, s, (const int32_t *)O,
                         (const int32_t *)tile_last, (const double *)wt.W, n_in, B, tpf_b, sc.guide.as<GuideRec>(), st_sel, gate);
            GENPF_LAUNCH((k_lookup_rec<OutT>), lgrid, kThreads, s, (const double *)wt.W, (const GuideRec *)sc.guide.as<GuideRec>(),
                         B, n_in, n_out, uni, (const int32_t *)nullptr, parents, out_base, st_sel, gate, fill);
        } else {
            GENPF_LAUNCH((k_expand<IdxT, IdxT>), dim3((unsigned)tpf_b, (unsigned)nf), kThreads, s, O, tile_last, n_in, B, tpf_b,
                         (const int32_t *)nullptr, G, (int64_t)0, st_sel, gate, 0);
            GENPF_LAUNCH((k_lookup<IdxT, OutT>), lgrid, kThreads, s, (const double *)wt.W, (const IdxT *)G, B, n_in, n_out, uni,
                         (const IdxT *)nullptr, parents, out_base, st_sel, gate, fill);
        }
    } else if (method == GENPF_RESIDUAL) {
        const size_t np = (size_t)(tpf_in * nf);
        GENPF_TRY(sc.W.ensure((size_t)(n_in * nf) * 8));
        const int64_t B = guide_buckets(n_in), tpf_b = ceil_div(B, kTile);
        GENPF_TRY(sc.guide.ensure((size_t)(B * nf) * (sizeof(IdxT) == 4 ? sizeof(GuideRec) : sizeof(IdxT))));
        GENPF_TRY(sc.guide_O.ensure((size_t)(n_in * nf) * sizeof(IdxT)));
        GENPF_TRY(sc.guide_tile_last.ensure((size_t)(tpf_in * nf) * sizeof(IdxT)));
        WTables rt{sc.W.as<double>()};
        IdxT *G = sc.guide.as<IdxT>(), *GO = sc.guide_O.as<IdxT>(), *GTL = sc.guide_tile_last.as<IdxT>();
        GENPF_TRY(sc.resid_c.ensure(np * 8));
        GENPF_TRY(sc.resid_r.ensure(np * 8));
        GENPF_TRY(sc.resid_coff.ensure(np * 8));
        GENPF_TRY(sc.resid_roff.ensure(np * 8));
        GENPF_TRY(sc.resid_rtot.ensure((size_t)nf * 8));
        ResidPartials rp{sc.resid_c.as<long long>(), sc.resid_r.as<double>()};
        GENPF_LAUNCH(k_resid_partials, dim3((unsigned)tpf_in, (unsigned)nf), kThreads, s, sel, n_in, n_out, tpf_in, st_sel, rp);
        GENPF_LAUNCH(k_resid_finalize, (unsigned)nf, 1024, s, rp, tpf_in, sc.resid_rtot.as<double>(),
                     sc.resid_coff.as<long long>(), sc.resid_roff.as<double>());
        GENPF_LAUNCH((k_resid_scan<IdxT>), dim3((unsigned)tpf_in, (unsigned)nf), kThreads, s, sel, n_in, n_out, tpf_in, st_sel,
                     sc.resid_rtot.as<double>(), sc.resid_coff.as<long long>(), sc.resid_roff.as<double>(), O,
                     tile_last, rt, GO, GTL, B);
        GENPF_LAUNCH((k_expand<IdxT, OutT>), dim3((unsigned)tpf_out, (unsigned)nf), kThreads, s, O, tile_last, n_in, n_out, tpf_out,
                     (const int32_t *)nullptr, parents, out_base, st_sel, gate, 1, fill);
        const dim3 lgrid((unsigned)ceil_div(n_out, (int64_t)kThreads * kLookupItems), (unsigned)nf);
        if constexpr (sizeof(IdxT) == 4) {
            GENPF_LAUNCH(k_guide_records, dim3((unsigned)tpf_b, (unsigned)nf), kThreads, s, (const int32_t *)GO,
                         (const int32_t *)GTL, (const double *)rt.W, n_in, B, tpf_b, sc.guide.as<GuideRec>(), st_sel, 0);
            GENPF_LAUNCH((k_lookup_rec<OutT>), lgrid, kThreads, s, (const double *)rt.W, (const GuideRec *)sc.guide.as<GuideRec>(),
                         B, n_in, n_out, uni, (const int32_t *)O, parents, out_base, st_sel, gate, fill);
        } else {
            GENPF_LAUNCH((k_expand<IdxT, IdxT>), dim3((unsigned)tpf_b, (unsigned)nf), kThreads, s, GO, GTL, n_in, B, tpf_b,
                         (const int32_t *)nullptr, G, (int64_t)0, st_sel, 0, 0);
            GENPF_LAUNCH((k_lookup<IdxT, OutT>), lgrid, kThreads, s, (const double *)rt.W, (const IdxT *)G, B, n_in, n_out, uni,
                         (const IdxT *)O, parents, out_base, st_sel, gate, fill);
        }
    } else {
        return fail(GENPF_ERR_UNKNOWN_METHOD, "Resampling method not recognized.");
    }
    return GENPF_OK;
}

template <typename OutT>
int32_t select_ancestors(cudaStream_t s, Scratch &sc, int method, LwSrc sel, int64_t n_in, int64_t n_out, int64_t nf,
                         Stats *st_sel, UniSrc uni, uint32_t flags, OutT *parents, int64_t out_base, int gate,
                         const double *ew = nullptr, LwFill fill = LwFill{nullptr, nullptr, 0}) {
    if (n_in < 0x7FFFFFF0ll && n_out < 0x7FFFFFF0ll)
        return select_ancestors_t<int32_t, OutT>(s, sc, method, sel, n_in, n_out, nf, st_sel, uni, flags, parents,
                                                 out_base, gate, ew, fill);
    return select_ancestors_t<long long, OutT>(s, sc, method, sel, n_in, n_out, nf, st_sel, uni, flags, parents,
                                               out_base, gate, ew, fill);
}

// mean / var of column x under softmax(lw) (statistics.jl:13-17,48-54); results in sc.moment_out[0..nf) and [nf..2nf)
inline int32_t launch_mean_var(cudaStream_t s, Scratch &sc, const double *lw, XSrc x, int64_t n, int64_t nf,
                               Stats *st) {
    const int64_t tpf = ceil_div(n, kTile);
    LwSrc src{lw, 1.0};
    double *partial = sc.moment_partial.as<double>();
    double *mean = sc.moment_out.as<double>(), *var = mean + nf;
    GENPF_LAUNCH(k_weighted_moment, dim3((unsigned)tpf, (unsigned)nf), kThreads, s, src, x, n, tpf, st, (const double *)nullptr,
                 partial);
    GENPF_LAUNCH(k_sum_partials, (unsigned)nf, kThreads, s, partial, tpf, mean);
    GENPF_LAUNCH(k_weighted_moment, dim3((unsigned)tpf, (unsigned)nf), kThreads, s, src, x, n, tpf, st, (const double *)mean,
                 partial);
    GENPF_LAUNCH(k_sum_partials, (unsigned)nf, kThreads, s, partial, tpf, var);
    return GENPF_OK;
}

}  // namespace genpf
