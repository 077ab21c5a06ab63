// filter_state.hpp -- the device-resident filter object behind genpf_filter_t (shared by abi_filter.cu
// and abi_shard.cu).
#pragma once
#include <cstring>
#include <vector>

#include "coalesce.cuh"
#include "optimal.cuh"
#include "engine.cuh"
#include "fused.cuh"

namespace genpf {

enum { kModelObjectMotion = 0, kModelLinGauss1D = 1, kNumModels = 2 };

struct ModelInfo {
    const char *name;
    int nf, nb, np, naux;
    int caps;  // bit 0: custom proposal (propose / proposal_logpdf / transition_logpdf), bit 1: translate
};
static const ModelInfo kModels[kNumModels] __attribute__((unused)) = {
    {"object_motion", ObjectMotion::NF, ObjectMotion::NB, ObjectMotion::NP, ObjectMotion::NAUX,
     (has_proposal<ObjectMotion>::value ? 1 : 0) | (has_translate<ObjectMotion>::value ? 2 : 0)},
    {"lingauss1d", LinGauss1D::NF, LinGauss1D::NB, LinGauss1D::NP, LinGauss1D::NAUX,
     (has_proposal<LinGauss1D>::value ? 1 : 0) | (has_translate<LinGauss1D>::value ? 2 : 0)},
};

struct Slab {  // one time slice's columns in one buffer
    Cols c;
};

struct HistSlice {  // a frozen slice (GENPF_KEEP_HISTORY): columns in the particle order of generation `gen`
    int64_t tau;
    Cols c;
    int64_t n;
    int64_t gen;
};
struct ParentLog {  // ancestry of resample number `gen` (1-based): population gen-1 -> gen
    int32_t *parents;
    int64_t n_prev, n_cur;
};

struct ResizeUndo {  // buffers parked by a resize that has not committed yet (abi_filter.cu::resize_target)
    bool active = false;
    Cols win[2];
    double *lw_alt = nullptr;
    int32_t *parents = nullptr;
};

}  // namespace genpf

using namespace genpf;

struct genpf_filter_s {
    int model = 0;
    int NF = 0, NB = 0;
    int64_t n = 0, nf = 0;
    uint64_t seed = 0;
    uint32_t flags = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    ModelParams P{};
    int64_t t_cur = 0;
    int64_t n_resamples = 0;
    int64_t first_filter = 0;  // index of filter 0 in a batch sharded over several handles / GPUs (rng_offset = first_filter * n)
    int64_t rng_offset = 0;
    Cols win[2][2];  // [buffer][slot parity]
    int buf = 0;
    double *lw = nullptr, *lw_alt = nullptr;
    double *ew = nullptr;  // e_i = exp(lw_i - m_tile), valid whenever part_valid (written with the K1 partials)
    double *lw_by_buf[2] = {nullptr, nullptr};  // identity of the two weight buffers (lw == lw_by_buf[buf])
    void *shard = nullptr;                      // ShardCtx (abi_shard.cu)
    void *shard_xchg = nullptr;                 // peer-mapped exchange block (abi_shard.cu)
    int32_t *parents = nullptr;
    uint8_t *accepts = nullptr;
    unsigned long long *n_accept = nullptr;
    double *lml = nullptr;
    double *obs_dev = nullptr;
    double *noise_cols[3] = {nullptr, nullptr, nullptr};
    DevBuf noise_buf[5], uni_buf, tmp_col, tmp_idx, prio_buf, key_buf, strata_buf, step_obs, run_obs;
    void *graph_exec = nullptr;  // cudaGraphExec_t of the last genpf_run_steps(GENPF_RUN_GRAPH)
    CoalesceBufs cb;
    OptimalBufs ob;
    void *h_opt_ctrl = nullptr;  // pinned OptCtrl, allocated on first optimal resize
    Scratch sc;
    bool part_valid = false;
    ResizeUndo undo;
    std::vector<HistSlice> hist;
    std::vector<ParentLog> plog;
    double *h_pinned = nullptr;  // pinned scratch: max(nf,16) doubles * 4
    Stats *h_stats = nullptr;

    // Device allocations of the filter.  Freed blocks are parked in a small cache instead of going back to the driver:
    // resizing operations (pf_replicate! -> resize back, every few steps in BASELINE config 5) alternate between two
    // population sizes, and cudaMalloc / cudaFree synchronise the device (measured: 16.7 ms per replicate + resize
    // cycle of 512 x 4096 particles, almost all of it allocator time).  A cached block serves a request of (nearly) its
    // own size only -- a looser match lets the small population squat in the large population's blocks and the
    // allocator thrash (measured) -- and the cache is bounded and released with the filter.
    struct Block {
        void *p;
        size_t bytes;
    };
    std::vector<Block> live, cache;
    static constexpr size_t kCacheBlocks = 64;
    template <typename T>
    int32_t dalloc(T **p, size_t count) {
        const size_t want = count * sizeof(T) + 16;
        for (size_t i = 0; i < cache.size(); ++i) {
            if (cache[i].bytes >= want && cache[i].bytes <= want + want / 8) {
                *p = reinterpret_cast<T *>(cache[i].p);
                live.push_back(cache[i]);
                cache.erase(cache.begin() + (long)i);
                return GENPF_OK;
            }
        }
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) {  // give the cached blocks back and try once more
            cudaGetLastError();
            for (auto &b : cache) cudaFree(b.p);
            cache.clear();
            e = cudaMalloc(&q, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(GENPF_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
        live.push_back(Block{q, want});
        *p = reinterpret_cast<T *>(q);
        return GENPF_OK;
    }
    // Stream-ordered reuse: everything the filter does runs on its one stream, so a block handed out again is only
    // touched by work enqueued after the work that last used it.
    void dfree(void *q) {
        if (!q) return;
        for (size_t i = 0; i < live.size(); ++i)
            if (live[i].p == q) {
                if (cache.size() < kCacheBlocks) cache.push_back(live[i]);
                else cudaFree(q);
                live.erase(live.begin() + (long)i);
                return;
            }
    }
    int32_t alloc_cols(Cols &c, int64_t total) {
        memset(&c, 0, sizeof(c));
        for (int i = 0; i < NF; ++i) GENPF_TRY(dalloc(&c.f[i], (size_t)total));
        for (int i = 0; i < NB; ++i) GENPF_TRY(dalloc(&c.b[i], (size_t)total));
        return GENPF_OK;
    }
    void free_cols(Cols &c) {
        for (int i = 0; i < NF; ++i) dfree(c.f[i]);
        for (int i = 0; i < NB; ++i) dfree(c.b[i]);
        memset(&c, 0, sizeof(c));
    }
    int32_t alloc_population(int64_t n_new) {
        const int64_t total = n_new * nf;
        for (int b = 0; b < 2; ++b)
            for (int sl = 0; sl < 2; ++sl) GENPF_TRY(alloc_cols(win[b][sl], total));
        GENPF_TRY(dalloc(&lw, (size_t)total));
        GENPF_TRY(dalloc(&lw_alt, (size_t)total));
        lw_by_buf[0] = lw;
        lw_by_buf[1] = lw_alt;
        GENPF_TRY(dalloc(&ew, (size_t)total));
        GENPF_TRY(dalloc(&parents, (size_t)total));
        GENPF_TRY(dalloc(&accepts, (size_t)total));
        GENPF_TRY(sc.ensure(n_new, nf));
        return GENPF_OK;
    }
    void free_population() {
        for (int b = 0; b < 2; ++b)
            for (int sl = 0; sl < 2; ++sl) free_cols(win[b][sl]);
        dfree(lw); dfree(lw_alt); dfree(parents); dfree(accepts); dfree(ew);
        ew = nullptr;
        lw = lw_alt = nullptr; parents = nullptr; accepts = nullptr;
    }
    Cols &slice(int64_t tau) { return win[buf][tau & 1]; }
    Cols &slice_alt(int64_t tau) { return win[buf ^ 1][tau & 1]; }
};


namespace genpf {
int32_t check_filter(genpf_filter_t pf);
int32_t log_parents(genpf_filter_t pf, int64_t n_prev, int64_t n_cur);
int32_t ensure_stats(genpf_filter_t pf, double *tile_off, double ess_frac, double *lml_accum);
int32_t read_stats(genpf_filter_t pf, int which);
void resize_rollback(genpf_filter_t pf);
int32_t stage_noise(genpf_filter_t pf, int which, const double *host, const double **dev, int64_t count);
}  // namespace genpf
