// abi_shard.cu -- C ABI, multi-GPU particle sharding.  One process per GPU; rank r owns global
// particle slots [r*n_loc, (r+1)*n_loc) of ONE filter of world*n_loc particles.  Two ways to run a step, both
// without host synchronisation:
//
//   genpf_shard_step_p2p     the whole step in one call: the statistics exchange and the closing barrier are done by
//                            the step's own kernels over peer-mapped memory (NVLink stores + epoch flags), the closing
//                            counts are derived identically on every rank, offspring are pushed to their owners with
//                            P2P stores
//
//   genpf_shard_begin_step   K1 partials -> local (max, sum e, sum e^2)        -> stats_local   [allgather]
//   genpf_shard_scan         global (M,S,ESS,lml), shard prefix, scan -> O_k    -> oend_local    [allgather]
//   genpf_shard_push         offspring of local parents -> owner's buffers over NVLink P2P       [barrier]
//   genpf_shard_finish       swap buffers, K1 partials of the received population
//                            (the three tiny collectives issued by the host language -- torch.distributed / NCCL.jl --
//                            on the filter's stream, on device buffers it owns)
#include "filter_state.hpp"
#include "plugin.hpp"

namespace genpf {

struct ShardCtx {
    int rank = 0, world = 1;
    int64_t n_loc = 0, n_total = 0;
    PeerDst peer[2];  // destination tables for push into buffer set b (b = the owners' spare set)
    double *stats_local = nullptr;
    const double *stats_all = nullptr;
    long long *oend_local = nullptr;
    const long long *oend_all = nullptr;
    double *shard_info = nullptr;
    XchgPeers xp;                   // peer-mapped exchange blocks (xp.x[rank] is mine)
    long long *oend_p2p = nullptr;  // oend_all for the p2p protocol (derived on every rank by the statistics combine)
    unsigned long long epoch = 0;   // exchange epoch: grows by one per sharded step, never reset (re-initialising the
                                    // filter resets n_resamples, and stale flags must not satisfy a wait)
    XchgLink link(unsigned long long e) const { return XchgLink{xp, world, rank, e}; }
    std::vector<void *> opened;
};

static int n_export_bufs(genpf_filter_t pf) { return 4 * (pf->NF + pf->NB) + 3 + 4 + 1 + 1; }

// fixed export order: for buf in {0,1}: for slot in {0,1}: f64 fields, u8 fields; then lw[buf 0], lw[buf 1], parents
static void list_bufs(genpf_filter_t pf, std::vector<void *> &out) {
    for (int b = 0; b < 2; ++b)
        for (int sl = 0; sl < 2; ++sl) {
            for (int i = 0; i < pf->NF; ++i) out.push_back(pf->win[b][sl].f[i]);
            for (int i = 0; i < pf->NB; ++i) out.push_back(pf->win[b][sl].b[i]);
        }
    out.push_back(pf->lw_by_buf[0]);
    out.push_back(pf->lw_by_buf[1]);
    out.push_back(pf->parents);
    for (int k = 0; k < 4; ++k) out.push_back(pf->sc.part[0][k].p);  // K1 partial arrays (m, s, s2, flags)
    out.push_back(pf->shard_xchg);                                   // exchange block (flags + payload)
    out.push_back(pf->ew);                                           // e_i column
}

template <class Model, class Noise>
static int32_t launch_push(genpf_filter_t pf, ShardCtx *sh, const StepArgs &a, Noise noise, const long long *oend_all) {
    const int64_t tpf = ceil_div(pf->n, kTile);
    const int64_t t = a.t;
    // the range is device resident: one block per local tile (+2 for the split tiles); blocks stride
    const unsigned grid = (unsigned)(tpf + 2);
    const PeerDst &pd = sh->peer[pf->buf ^ 1];
    PeerDst d = pd;
    // the owner's spare set holds slice t-1 at parity (t-1)&1 and receives slice t at parity t&1
    for (int g = 0; g < sh->world; ++g) {
        if (((t - 1) & 1) == 1) std::swap(d.dst_cur[g], d.dst_new[g]);
    }
    if (a.mh_iters == 1) {
        GENPF_LAUNCH_PDL((k_step_push<Model, Noise, int32_t, 1>), grid, kStateThreads, pf->stream, a,
                     (const int32_t *)pf->sc.O.as<int32_t>(), (const int32_t *)pf->sc.tile_last.as<int32_t>(),
                     pf->slice(t - 2), pf->slice(t - 1), d, oend_all, sh->world, pf->n, tpf, sh->rank, noise,
                     (const Stats *)pf->sc.st(0, 1));
    } else {
        GENPF_LAUNCH_PDL((k_step_push<Model, Noise, int32_t, -1>), grid, kStateThreads, pf->stream, a,
                     (const int32_t *)pf->sc.O.as<int32_t>(), (const int32_t *)pf->sc.tile_last.as<int32_t>(),
                     pf->slice(t - 2), pf->slice(t - 1), d, oend_all, sh->world, pf->n, tpf, sh->rank, noise,
                     (const Stats *)pf->sc.st(0, 1));
    }
    return GENPF_OK;
}
template <class Model>
static int32_t push_model(genpf_filter_t pf, ShardCtx *sh, const StepArgs &a, const long long *oend_all,
                          const NoiseCols *cols = nullptr) {
    if (cols) return launch_push<Model, NoiseCols>(pf, sh, a, *cols, oend_all);
    if (pf->flags & GENPF_NOISE_PHILOX53) {
        NoisePhilox53 nz{pf->seed, (uint64_t)a.t, 0};
        return launch_push<Model, NoisePhilox53>(pf, sh, a, nz, oend_all);
    }
    NoiseLean nz{pf->seed, (uint64_t)a.t, 0};
    return launch_push<Model, NoiseLean>(pf, sh, a, nz, oend_all);
}

static int32_t make_step_args(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                              const double *obs_t, const double *aux_t, int32_t mh_iters, StepArgs *out) {
    if (t != pf->t_cur + 1) return fail(GENPF_ERR_INVALID_ARG, "the sharded step must advance to t_cur + 1");
    if (!obs_prev || !obs_t) return fail(GENPF_ERR_INVALID_ARG, "obs is NULL");
    if (mh_iters < 0 || mh_iters > 255) return fail(GENPF_ERR_INVALID_ARG, "mh_iters out of range");
    const ModelInfo &mi = *model_info(pf->model);
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
    a.obs_prev_dev = a.obs_t_dev = nullptr;
    a.obs_prev = obs_prev[0];
    a.obs_t = obs_t[0];
    *out = a;
    return GENPF_OK;
}

}  // namespace genpf

using namespace genpf;

extern "C" {

int32_t genpf_shard_ipc_export(genpf_filter_t pf, void *handles, int64_t *n_bufs) {
    GENPF_TRY(check_filter(pf));
    if (!n_bufs) return fail(GENPF_ERR_INVALID_ARG, "n_bufs is NULL");
    *n_bufs = n_export_bufs(pf);
    if (!handles) return GENPF_OK;
    if (!pf->shard_xchg) {
        Xchg *x = nullptr;
        GENPF_TRY(pf->dalloc(&x, 1));
        pf->shard_xchg = x;
    }
    // every (re-)attach starts from a zeroed block on every rank: the group's epochs restart at 1 together
    GENPF_CUDA_TRY(cudaMemsetAsync(pf->shard_xchg, 0, sizeof(Xchg), pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    std::vector<void *> bufs;
    list_bufs(pf, bufs);
    cudaIpcMemHandle_t *h = reinterpret_cast<cudaIpcMemHandle_t *>(handles);
    for (size_t i = 0; i < bufs.size(); ++i) GENPF_CUDA_TRY(cudaIpcGetMemHandle(&h[i], bufs[i]));
    return GENPF_OK;
}

int32_t genpf_shard_attach(genpf_filter_t pf, int32_t rank, int32_t world, const void *all_handles,
                           double *stats_local_dev, const double *stats_all_dev, long long *oend_local_dev,
                           const long long *oend_all_dev) {
    GENPF_TRY(check_filter(pf));
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(GENPF_ERR_INVALID_ARG, "genpf_shard_attach: need 1 <= world <= 8 and 0 <= rank < world");
    if (pf->nf != 1) return fail(GENPF_ERR_UNSUPPORTED, "sharding needs n_filters == 1");
    if (pf->model >= kPluginIdBase) return fail(GENPF_ERR_UNSUPPORTED, "run-time plugins cannot be sharded yet (no push kernel in the image)");
    if (pf->n % kTile != 0) return fail(GENPF_ERR_INVALID_ARG, "particles per shard must be a multiple of 2048");
    if (pf->flags & GENPF_KEEP_HISTORY) return fail(GENPF_ERR_UNSUPPORTED, "sharding with GENPF_KEEP_HISTORY");
    if ((int64_t)world * pf->n >= 0x7FFFFFF0ll) return fail(GENPF_ERR_UNSUPPORTED, "total population must be < 2^31");
    if (!pf->shard_xchg) return fail(GENPF_ERR_STATE, "call genpf_shard_ipc_export before genpf_shard_attach");
    if (world > 1 && !all_handles) return fail(GENPF_ERR_INVALID_ARG, "genpf_shard_attach: NULL handles");
    ShardCtx *sh = new ShardCtx();
    sh->rank = rank;
    sh->world = world;
    sh->n_loc = pf->n;
    sh->n_total = (int64_t)world * pf->n;
    sh->stats_local = stats_local_dev;
    sh->stats_all = stats_all_dev;
    sh->oend_local = oend_local_dev;
    sh->oend_all = oend_all_dev;
    const int nb = n_export_bufs(pf);
    std::vector<void *> mine;
    list_bufs(pf, mine);
    for (int g = 0; g < world; ++g) {
        std::vector<void *> ptrs(nb);
        if (g == rank) {
            ptrs = mine;
        } else {
            const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(all_handles) + (size_t)g * nb;
            for (int i = 0; i < nb; ++i) {
                GENPF_CUDA_TRY(cudaIpcOpenMemHandle(&ptrs[i], h[i], cudaIpcMemLazyEnablePeerAccess));
                sh->opened.push_back(ptrs[i]);
            }
        }
        int k = 0;
        for (int b = 0; b < 2; ++b) {
            // parity 0 -> "dst_cur" slot table, parity 1 -> "dst_new"; launch_push swaps per time step
            Cols c0, c1;
            memset(&c0, 0, sizeof(c0));
            memset(&c1, 0, sizeof(c1));
            for (int i = 0; i < pf->NF; ++i) c0.f[i] = (double *)ptrs[k++];
            for (int i = 0; i < pf->NB; ++i) c0.b[i] = (uint8_t *)ptrs[k++];
            for (int i = 0; i < pf->NF; ++i) c1.f[i] = (double *)ptrs[k++];
            for (int i = 0; i < pf->NB; ++i) c1.b[i] = (uint8_t *)ptrs[k++];
            sh->peer[b].dst_cur[g] = c0;
            sh->peer[b].dst_new[g] = c1;
        }
        sh->peer[0].lw[g] = (double *)ptrs[k++];
        sh->peer[1].lw[g] = (double *)ptrs[k++];
        sh->peer[0].parents[g] = sh->peer[1].parents[g] = (int32_t *)ptrs[k++];
        Partials pp{(double *)ptrs[k], (double *)ptrs[k + 1], (double *)ptrs[k + 2], (int *)ptrs[k + 3]};
        k += 4;
        sh->peer[0].part[g] = sh->peer[1].part[g] = pp;
        sh->xp.x[g] = (Xchg *)ptrs[k++];
        sh->peer[0].ew[g] = sh->peer[1].ew[g] = (double *)ptrs[k++];
    }
    GENPF_TRY(pf->dalloc(&sh->oend_p2p, kMaxPeers));
    GENPF_TRY(pf->dalloc(&sh->shard_info, 2));
    pf->rng_offset = 0;  // Philox counters are GLOBAL particle slots
    pf->shard = sh;
    return GENPF_OK;
}

int32_t genpf_shard_detach(genpf_filter_t pf) {
    if (!pf || !pf->shard) return GENPF_OK;
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    cudaSetDevice(pf->device);
    cudaStreamSynchronize(pf->stream);
    for (void *p : sh->opened) cudaIpcCloseMemHandle(p);
    delete sh;
    pf->shard = nullptr;
    return GENPF_OK;
}

// pf_initialize for a shard: Philox counters are global slots rank*n_loc + i
int32_t genpf_shard_initialize(genpf_filter_t pf, const double *obs, const double *aux) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    pf->rng_offset = (int64_t)sh->rank * sh->n_loc;
    return genpf_initialize(pf, obs, aux);
}

int32_t genpf_shard_begin_step(genpf_filter_t pf) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    if (!sh->stats_local || !sh->stats_all || !sh->oend_local || !sh->oend_all)
        return fail(GENPF_ERR_STATE, "no exchange buffers were attached (use genpf_shard_step_p2p)");
    // local statistics + locally normalised tile offsets; Stats starts with the three doubles (M, S, S2)
    GENPF_TRY(ensure_stats(pf, pf->sc.tile_off.as<double>(), -1.0, nullptr));
    GENPF_CUDA_TRY(cudaMemcpyAsync(sh->stats_local, pf->sc.st(0, 1), 3 * sizeof(double), cudaMemcpyDeviceToDevice,
                                   pf->stream));
    return GENPF_OK;
}

int32_t genpf_shard_scan(genpf_filter_t pf) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    const int64_t n = pf->n, tpf = ceil_div(n, kTile);
    Scratch &sc = pf->sc;
    GENPF_TRY(sc.O.ensure((size_t)n * 4));
    GENPF_TRY(sc.tile_last.ensure((size_t)tpf * 4));
    GENPF_LAUNCH(k_shard_combine, 1, 32, pf->stream, sh->stats_all, sh->world, sh->rank, sh->n_total, sc.st(0, 1),
                 sh->shard_info, pf->lml);
    UniSrc uni{nullptr, pf->seed, make_stream(kPurposeResample, (uint64_t)pf->n_resamples + 1), 0};
    StratArgs strat = make_strat(uni, sh->n_total);
    LwSrc lw_src{pf->lw, 1.0};
    GENPF_TRY(launch_scan_counts<int32_t>(pf->stream, lw_src, n, tpf, 1, sc.st(0, 1), sc.tile_off.as<double>(),
                                          sc.O.as<int32_t>(), sc.tile_last.as<int32_t>(), strat, 0, sh->shard_info,
                                          (int64_t)sh->rank * n, sc.chunk_info_ptr(n), pf->ew, sc.tile_scale.as<double>()));
    GENPF_LAUNCH(k_shard_oend, 1, 32, pf->stream, (const int32_t *)sc.tile_last.as<int32_t>(), tpf, sh->oend_local);
    return GENPF_OK;
}

int32_t genpf_shard_push(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                         const double *obs_t, const double *aux_t, int32_t mh_iters) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    if (!sh->oend_all) return fail(GENPF_ERR_STATE, "no exchange buffers were attached (use genpf_shard_step_p2p)");
    StepArgs a;
    GENPF_TRY(make_step_args(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, &a));
    int32_t st;
    switch (pf->model) {
        case kModelObjectMotion: st = push_model<ObjectMotion>(pf, sh, a, sh->oend_all); break;
        case kModelLinGauss1D: st = push_model<LinGauss1D>(pf, sh, a, sh->oend_all); break;
        default: return fail(GENPF_ERR_INVALID_ARG, "unknown model");
    }
    GENPF_TRY(st);
    pf->t_cur = t;  // slices now live in the spare buffer set; genpf_shard_finish swaps after the barrier
    return GENPF_OK;
}

// The whole sharded README iteration with the peer-memory exchange: no NCCL, no host synchronisation.
static int32_t shard_step_p2p(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                              const double *obs_t, const double *aux_t, int32_t mh_iters, const double *d_uniforms,
                              const NoiseCols *cols) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    if (pf->t_cur < 1) return fail(GENPF_ERR_STATE, "filter is not initialised");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    StepArgs a;
    GENPF_TRY(make_step_args(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, &a));
    const int64_t n = pf->n, tpf = ceil_div(n, kTile);
    Scratch &sc = pf->sc;
    cudaStream_t s = pf->stream;
    // No host synchronisation and TWO cross-GPU synchronisation points per step (statistics, done):
    //   finalize (16 chunk blocks) -> k_chunk_combine: local totals, POST stats, wait, population statistics and
    //               EVERY shard's closing offspring count (derived identically on every rank: no second exchange)
    //   k_scan_hot: shard-aware offspring counts pinned to the agreed closing counts
    //   k_step_push: offspring to their owners (NVLink P2P stores)
    //   k_reduce_boundary: POST "my offspring have left" (stream order behind the push), wait for every producer's
    //               flag, reduce the tiles two producers shared
    const unsigned long long epoch = ++sh->epoch;
    const XchgLink lk = sh->link(epoch);
    GENPF_TRY(sc.O.ensure((size_t)n * 4));
    GENPF_TRY(sc.tile_last.ensure((size_t)tpf * 4));
    // 1. local statistics (locally normalised tile offsets), post + gather + combine
    if (!pf->part_valid) {
        LwSrc src0{pf->lw, 1.0};
        GENPF_TRY(launch_reduce(s, src0, n, 1, sc.partials(0), pf->ew));
        pf->part_valid = true;
    }
    // the stratum uniforms of this resample: the combine derives every shard's closing count from them
    UniSrc uni{d_uniforms, pf->seed, make_stream(kPurposeResample, (uint64_t)pf->n_resamples + 1), 0};
    StratArgs strat = make_strat(uni, sh->n_total);
    bool exchanged = false;
    GENPF_TRY(launch_finalize(s, sc, sc.partials(0), n, 1, sc.st(0, 1), sc.tile_off.as<double>(), -1.0, nullptr, &lk,
                              sh->n_total, sh->shard_info, &exchanged, pf->lml, &strat, sh->oend_p2p));
    if (!exchanged)  // small shard: one finalize block wrote the local Stats
        GENPF_LAUNCH_PDL(k_xchg_stats_combine, 1, 32, s, (const Stats *)sc.st(0, 1), lk, sh->n_total, sc.st(0, 1), sh->shard_info,
                     pf->lml, strat, sh->oend_p2p);
    // 2. shard-aware scan, counts pinned to the agreed closing counts (no second exchange)
    LwSrc lw_src{pf->lw, 1.0};
    GENPF_TRY(launch_scan_counts<int32_t>(s, lw_src, n, tpf, 1, sc.st(0, 1), sc.tile_off.as<double>(), sc.O.as<int32_t>(),
                                          sc.tile_last.as<int32_t>(), strat, 0, sh->shard_info, (int64_t)sh->rank * n,
                                          sc.chunk_info_ptr(n), pf->ew, sc.tile_scale.as<double>(), sh->oend_p2p, sh->rank));
    // 3. offspring to their owners over NVLink
    int32_t st;
    switch (pf->model) {
        case kModelObjectMotion: st = push_model<ObjectMotion>(pf, sh, a, sh->oend_p2p, cols); break;
        case kModelLinGauss1D: st = push_model<LinGauss1D>(pf, sh, a, sh->oend_p2p, cols); break;
        default: return fail(GENPF_ERR_INVALID_ARG, "unknown model");
    }
    GENPF_TRY(st);
    // 4. swap; the closing barrier + K1 partials of the tiles two producers shared
    pf->t_cur = t;
    pf->buf ^= 1;
    std::swap(pf->lw, pf->lw_alt);
    pf->n_resamples += 1;
    LwSrc src{pf->lw, 1.0};
    GENPF_LAUNCH_PDL(k_reduce_boundary, (unsigned)sh->world, kReduceThreads, s, src, (const long long *)sh->oend_p2p,
                 sh->world, sh->rank, n, sc.partials(0), pf->ew, lk);
    pf->part_valid = true;
    return GENPF_OK;
}

int32_t genpf_shard_step_p2p(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                             const double *obs_t, const double *aux_t, int32_t mh_iters) {
    return shard_step_p2p(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, nullptr, nullptr);
}

// The sharded step in parity mode: every rank passes the SAME global-length host columns
// (world * n_local entries, indexed by GLOBAL output slot / stratum): "uniforms replicated first".  Runs the same
// kernels as genpf_shard_step_p2p (shard-aware k_scan, k_step_push, peer-memory exchange) with the column policy.
int32_t genpf_shard_step_p2p_with_noise(genpf_filter_t pf, int64_t t, const double *obs_prev, const double *aux_prev,
                                        const double *obs_t, const double *aux_t, int32_t mh_iters,
                                        const double *uniforms, const double *U2, const double *Z2, const double *U3,
                                        const double *U1, const double *Z1) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    if (mh_iters != 0 && mh_iters != 1) return fail(GENPF_ERR_INVALID_ARG, "noise columns serve mh_iters 0 or 1");
    if (!U1 || !Z1 || (mh_iters == 1 && (!U2 || !Z2 || !U3))) return fail(GENPF_ERR_INVALID_ARG, "noise columns are NULL");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    const int64_t nt = sh->n_total;
    const double *d_u = nullptr;
    if (uniforms) {
        GENPF_TRY(pf->uni_buf.ensure((size_t)nt * 8));
        GENPF_CUDA_TRY(cudaMemcpyAsync(pf->uni_buf.p, uniforms, (size_t)nt * 8, cudaMemcpyHostToDevice, pf->stream));
        d_u = pf->uni_buf.as<double>();
    }
    NoiseCols nz{nullptr, nullptr, nullptr, nullptr, nullptr};
    GENPF_TRY(stage_noise(pf, 0, U2, &nz.U, nt));
    GENPF_TRY(stage_noise(pf, 1, Z2, &nz.Z, nt));
    GENPF_TRY(stage_noise(pf, 2, U3, &nz.U3, nt));
    GENPF_TRY(stage_noise(pf, 3, U1, &nz.Uup, nt));
    GENPF_TRY(stage_noise(pf, 4, Z1, &nz.Zup, nt));
    return shard_step_p2p(pf, t, obs_prev, aux_prev, obs_t, aux_t, mh_iters, d_u, &nz);
}

// closing counts of the last step (host copy; synchronises) + the exchange error word
int32_t genpf_shard_oend(genpf_filter_t pf, long long *oend_all_host, int32_t *error) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, sh->oend_p2p, sizeof(long long) * kMaxPeers, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned + kMaxPeers, &sh->xp.x[sh->rank]->error, sizeof(int), cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(pf->stream));
    if (oend_all_host)
        for (int g = 0; g < sh->world; ++g) oend_all_host[g] = reinterpret_cast<long long *>(pf->h_pinned)[g];
    if (error) *error = *reinterpret_cast<int *>(pf->h_pinned + kMaxPeers);
    return GENPF_OK;
}

int32_t genpf_shard_finish(genpf_filter_t pf) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    pf->buf ^= 1;
    std::swap(pf->lw, pf->lw_alt);
    pf->n_resamples += 1;
    // full tiles arrived with their K1 partials; only tiles split between two producers are reduced here
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    LwSrc src{pf->lw, 1.0};
    GENPF_LAUNCH(k_reduce_boundary, (unsigned)sh->world, kReduceThreads, pf->stream, src, sh->oend_all, sh->world,
                 sh->rank, pf->n, pf->sc.partials(0), pf->ew, XchgLink{{}, 0, 0, 0});
    pf->part_valid = true;
    return GENPF_OK;
}

// global ESS / log_ml_estimate of the sharded population as of the last genpf_shard_scan
int32_t genpf_shard_stats(genpf_filter_t pf, double *ess, double *lml_est, int32_t *invalid_kind) {
    GENPF_TRY(check_filter(pf));
    if (!pf->shard) return fail(GENPF_ERR_STATE, "filter is not attached to a shard group");
    ShardCtx *sh = reinterpret_cast<ShardCtx *>(pf->shard);
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned, pf->lml, 8, cudaMemcpyDeviceToHost, pf->stream));
    GENPF_CUDA_TRY(cudaMemcpyAsync(pf->h_pinned + 1, &sh->xp.x[sh->rank]->error, sizeof(int), cudaMemcpyDeviceToHost, pf->stream));
    GENPF_TRY(read_stats(pf, 0));
    // a wait that ran into its bound (a lost or hung peer) left stale statistics / closing counts / payload behind
    if (*reinterpret_cast<int *>(pf->h_pinned + 1) != 0)
        return fail(GENPF_ERR_STATE, "peer exchange timed out: a rank of the shard group did not post (population is invalid)");
    if (ess) *ess = pf->h_stats[0].ess;
    if (lml_est) *lml_est = pf->h_pinned[0];  // log_ml_est accumulated by the resamples so far
    if (invalid_kind) *invalid_kind = pf->h_stats[0].invalid_kind;
    return GENPF_OK;
}

}  // extern "C"
