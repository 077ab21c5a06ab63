// optimal.cuh -- pf_optimal_resize! (reference src/resize.jl:149-196) and find_inv_w_threshold
// (resize.jl:199-216): Fearnhead-Clifford optimal resampling down to n_out <= n particles.
//
// The reference sorts the weights, walks them in ascending order with a running sum until
// B/kappa + A <= N, keeps every particle with c*w >= 1 and draws the rest with ONE uniform (systematic
// resampling over the re-normalised remainder).  Here:
//   threshold : ascending radix sort of the log-weights (sort.cu) -> k_reduce / finalize / k_scan give the
//               normalised prefix sums B_r of the sorted weights; the predicate is monotone in r, so every r is
//               tested in parallel and the smallest passing r wins (atomicMin) -- the loop's first hit
//   keep      : flags c*w_i >= 1, tile counts, one-block tile scan, stable compaction (keep indices ascending,
//               exactly findall's order) which also compacts the remainder (strat_idxs order)
//   remainder : k_reduce / finalize / k_scan over the compacted remainder = its safe_softmax + cumulative sums
//               C_q; particle q is drawn iff a threshold u0 + k/n_res lies in (C_{q-1}, C_q): counts
//               K(C) = #{k : u0 + k*step < C} are written once per particle and differenced, so neighbours agree
// The sequential fp64 recurrences of the reference (B += kappa; u -= w) are replaced by parallel sums: results
// differ only at cumulative-sum ties, the documented tolerance class.
#pragma once
#include "engine.cuh"

namespace genpf {

struct OptCtrl {
    unsigned long long r_star;  // smallest ascending rank passing the threshold test (n: none)
    double c;                   // inverse weight threshold
    double res_lw;              // log-weight of every resampled particle
    double log_n_ratio;
    long long n_keep;
    long long n_selected;
    double u_rand;  // the single rand() of resize.jl:171 when the library draws it
};

struct OptimalBufs {
    DevBuf sorted, order, flags, tile_cnt, strat_idx, strat_lw, Kc, ctrl, sort_tmp;
    void release() {
        for (DevBuf *b : {&sorted, &order, &flags, &tile_cnt, &strat_idx, &strat_lw, &Kc, &ctrl, &sort_tmp}) b->release();
    }
};

__device__ __forceinline__ double opt_weight(const Stats &st, double v, int64_t n) {
    if (st.invalid_kind == 2 || st.invalid_kind == 3) return 1.0 / (double)n;  // safe_softmax uniform fallback
    return exp_nonpos(v - st.M) * (1.0 / st.S);
}

// n_check = B/kappa + A <= N + eps(n_check) for ascending rank r (A = n-1-r weights above, B = sum up to r)
static __global__ void k_opt_threshold(const double *sorted_lw, const double *W_asc, const Stats *st_sorted, int64_t n,
                                       int64_t N, OptCtrl *ctrl) {
    const Stats st = st_sorted[0];
    unsigned long long best = ~0ull;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const double kappa = opt_weight(st, sorted_lw[r], n);
        const double n_check = W_asc[r] / kappa + (double)(n - 1 - r);
        const double a = fabs(n_check);
        const double eps = __longlong_as_double(__double_as_longlong(a) + 1) - a;  // Julia eps(x)
        if (n_check <= (double)N + eps) {  // false for NaN (0/0 at zero weights), like the reference
            best = (unsigned long long)r;
            break;  // r ascends within a thread
        }
    }
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 16));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 8));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 4));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(&ctrl->r_star, best);
}
static __global__ void k_opt_threshold_finish(const double *W_asc, const Stats *st_lw, int64_t n, int64_t N, OptCtrl *ctrl) {
    const unsigned long long r = ctrl->r_star;
    double c = (double)N;  // resize.jl:214
    if (r < (unsigned long long)n) c = ((double)N - (double)(n - 1 - (int64_t)r)) / W_asc[r];  // resize.jl:211
    ctrl->c = c;
    ctrl->log_n_ratio = log((double)N) - log((double)n);               // resize.jl:187
    ctrl->res_lw = st_lw[0].lse - log(c) + ctrl->log_n_ratio;          // resize.jl:190,193
}

constexpr int kOptThreads = 256, kOptItems = 8, kOptTile = kOptThreads * kOptItems;

// keep_i = c * w_i >= 1 (resize.jl:156); per-tile keep counts
static __global__ void __launch_bounds__(kOptThreads)
    k_opt_flags(const double *lw, const Stats *st_lw, const OptCtrl *ctrl, int64_t n, uint8_t *flags, int *tile_cnt) {
    __shared__ int sw[kOptThreads / 32];
    const Stats st = st_lw[0];
    const double c = ctrl->c;
    const int64_t base = (int64_t)blockIdx.x * kOptTile;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kOptItems; ++k) {
        const int64_t i = base + k * kOptThreads + threadIdx.x;
        if (i < n) {
            const int keep = c * opt_weight(st, lw[i], n) >= 1.0 ? 1 : 0;
            flags[i] = (uint8_t)keep;
            cnt += keep;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kOptThreads / 32; ++w) t += sw[w];
        tile_cnt[blockIdx.x] = t;
    }
}
// exclusive scan of the tile counts (one block, running carry), total -> ctrl->n_keep
static __global__ void __launch_bounds__(1024) k_opt_tile_scan(const int *tile_cnt, int64_t ntiles, long long *tile_off, OptCtrl *ctrl) {
    __shared__ long long sw[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const int64_t b = b0 + threadIdx.x;
        const long long v = b < ntiles ? tile_cnt[b] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        long long woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            if (w < warp) woff += sw[w];
            tot += sw[w];
        }
        const long long cr = carry;
        if (b < ntiles) tile_off[b] = cr + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = cr + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) ctrl->n_keep = carry;
}
// stable compaction: kept particles -> parents[0..n_keep) (+ their log-weights, resize.jl:192), the others ->
// strat_idx / strat_lw in index order (resize.jl:163,166)
template <typename OutT>
static __global__ void __launch_bounds__(kOptThreads)
    k_opt_compact(const double *lw, const uint8_t *flags, const long long *tile_off, const OptCtrl *ctrl, int64_t n,
                  int64_t N, OutT *parents, int64_t out_base, double *lw_out, long long *strat_idx, double *strat_lw) {
    __shared__ int sw[kOptThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kOptTile + (int64_t)threadIdx.x * kOptItems;  // blocked: 8 consecutive
    int f[kOptItems], run = 0;
#pragma unroll
    for (int k = 0; k < kOptItems; ++k) {
        f[k] = base + k < n ? flags[base + k] : 0;
        run += f[k];
    }
    int inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w)
        if (w < warp) woff += sw[w];
    long long p = tile_off[blockIdx.x] + woff + inc - run;  // kept particles before this thread's first element
    const double lnr = ctrl->log_n_ratio;
#pragma unroll
    for (int k = 0; k < kOptItems; ++k) {
        const int64_t i = base + k;
        if (i >= n) break;
        const double v = lw[i];
        if (f[k]) {
            if (p < N) {
                parents[p] = (OutT)(i + out_base);
                lw_out[p] = v + lnr;
            }
            ++p;
        } else {
            const int64_t q = i - p;
            strat_idx[q] = i;
            strat_lw[q] = v;
        }
    }
}
// K(C) = #{k in [0, n_res) : u0 + k*step < C}: how many systematic thresholds the cumulative weight has passed
// (the reference's `u -= w; if u < 0 ... u += step`, resize.jl:172-178, in closed form)
__device__ __forceinline__ long long opt_count(double C, double u0, double step, long long n_res) {
    const double x = (C - u0) * (double)n_res;
    long long k = x <= 0.0 ? 0 : (x >= (double)n_res ? n_res : (long long)ceil(x));
    while (k > 0 && !(__dadd_rn(u0, __dmul_rn((double)(k - 1), step)) < C)) --k;
    while (k < n_res && __dadd_rn(u0, __dmul_rn((double)k, step)) < C) ++k;
    return k;
}
static __global__ void k_opt_counts(const double *C, int64_t n_strat, long long n_res, double u0, double step,
                                    const Stats *st_strat, long long *Kc) {
    const int kind = st_strat[0].invalid_kind;
    if (kind == 1 || kind == 4) return;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_strat; q += (int64_t)gridDim.x * blockDim.x)
        Kc[q] = opt_count(C[q], u0, step, n_res);
}
template <typename OutT>
static __global__ void k_opt_select(const long long *Kc, const long long *strat_idx, int64_t n_strat, int64_t N,
                                    const Stats *st_strat, OptCtrl *ctrl, OutT *parents, int64_t out_base, double *lw_out) {
    const int kind = st_strat[0].invalid_kind;
    if (kind == 1 || kind == 4) return;
    const long long n_keep = ctrl->n_keep;
    const double res_lw = ctrl->res_lw;
    int mine = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_strat; q += (int64_t)gridDim.x * blockDim.x) {
        const long long prev = q ? Kc[q - 1] : 0;
        if (Kc[q] > prev) {  // one push per particle, like the reference loop
            const long long pos = n_keep + prev;
            if (pos < N) {
                parents[pos] = (OutT)(strat_idx[q] + out_base);
                lw_out[pos] = res_lw;
            }
            ++mine;
        }
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd((unsigned long long *)&ctrl->n_selected, (unsigned long long)mine);
}

struct OptResult {
    int64_t n_keep = 0;
    double inv_w = 0.0;
    int32_t kind = 0, kind_strat = 0;
};

int32_t sort_asc_f64(const double *keys, int64_t n, double *keys_sorted, int32_t *order32, DevBuf &tmp, cudaStream_t stream);

// h_ctrl: pinned OptCtrl, h_stats: pinned Stats[>=2].  d_lw_out must hold n_out entries, d_parents likewise.
template <typename OutT>
int32_t optimal_resize_core(cudaStream_t s, Scratch &sc, OptimalBufs &ob, OptCtrl *h_ctrl, Stats *h_stats,
                            const double *d_lw, int64_t n, int64_t N, const double *u_rand_or_null, UniSrc uni,
                            uint32_t flags, OutT *d_parents, int64_t out_base, double *d_lw_out, OptResult *res) {
    if (N < 1 || N > n) return fail(GENPF_ERR_INVALID_ARG, "optimal resize: need 1 <= n_particles <= current size (resize.jl:183)");
    GENPF_TRY(sc.ensure(n, 1));
    const int64_t ntiles = ceil_div(n, (int64_t)kOptTile), tpf = ceil_div(n, kTile);
    GENPF_TRY(ob.sorted.ensure((size_t)n * 8));
    GENPF_TRY(ob.order.ensure((size_t)n * 4));
    GENPF_TRY(ob.flags.ensure((size_t)n));
    GENPF_TRY(ob.tile_cnt.ensure((size_t)ntiles * 12 + 16));
    GENPF_TRY(ob.strat_idx.ensure((size_t)n * 8));
    GENPF_TRY(ob.strat_lw.ensure((size_t)n * 8));
    GENPF_TRY(ob.Kc.ensure((size_t)n * 8));
    GENPF_TRY(ob.ctrl.ensure(sizeof(OptCtrl)));
    GENPF_TRY(sc.W.ensure((size_t)n * 8));
    OptCtrl *ctrl = ob.ctrl.as<OptCtrl>();
    int *tile_cnt = ob.tile_cnt.as<int>();
    long long *tile_off = reinterpret_cast<long long *>(ob.tile_cnt.as<char>() + (((size_t)ntiles * 4 + 7) & ~(size_t)7));
    OptCtrl init{};
    init.r_star = (unsigned long long)n;
    *h_ctrl = init;
    GENPF_CUDA_TRY(cudaMemcpyAsync(ctrl, h_ctrl, sizeof(OptCtrl), cudaMemcpyHostToDevice, s));
    if (!u_rand_or_null) GENPF_LAUNCH(k_uniforms, 1, 32, s, uni, (int64_t)1, &ctrl->u_rand, 0);
    // safe_softmax(log_weights) (resize.jl:152): statistics of the weights in particle order
    LwSrc src{d_lw, 1.0};
    Stats *st_lw = sc.st(0, 1), *st_sorted = sc.st(1, 1), *st_strat = sc.st(2, 1);
    GENPF_TRY(launch_reduce(s, src, n, 1, sc.partials(0)));
    GENPF_TRY(launch_finalize(s, sc, sc.partials(0), n, 1, st_lw, nullptr, -1.0, nullptr));
    GENPF_CUDA_TRY(cudaMemcpyAsync(h_stats, st_lw, sizeof(Stats), cudaMemcpyDeviceToHost, s));
    GENPF_CUDA_TRY(cudaStreamSynchronize(s));
    res->kind = h_stats[0].invalid_kind;
    if ((flags & GENPF_CHECK) && res->kind != GENPF_VALID) return fail(GENPF_ERR_INVALID_WEIGHTS, "Invalid weights.");
    if (res->kind == GENPF_INV_NAN_INPUT || res->kind == GENPF_INV_NAN_TOTAL)
        return fail(GENPF_ERR_INVALID_WEIGHTS, "optimal resize: NaN weights select no particle (reference: AssertionError, resize.jl:181)");
    // find_inv_w_threshold: ascending sort, normalised prefix sums, first rank passing the test
    double *sorted = ob.sorted.as<double>();
    GENPF_TRY(sort_asc_f64(d_lw, n, sorted, ob.order.as<int32_t>(), ob.sort_tmp, s));
    LwSrc ssrc{sorted, 1.0};
    GENPF_TRY(launch_reduce(s, ssrc, n, 1, sc.partials(1)));
    GENPF_TRY(launch_finalize(s, sc, sc.partials(1), n, 1, st_sorted, sc.tile_off.as<double>(), -1.0, nullptr));
    StratArgs none = make_strat(UniSrc{nullptr, 0, 0, 0}, n);
    double *W = sc.W.as<double>();
    GENPF_LAUNCH((k_scan<int32_t>), dim3((unsigned)tpf, 1), kScanThreads, s, ssrc, n, tpf, (const Stats *)st_sorted,
                 (const double *)sc.tile_off.as<double>(), WTables{W}, (int32_t *)nullptr, (int32_t *)nullptr, none, 0,
                 (const double *)nullptr, (int64_t)0, sc.chunk_info_ptr(n), Scratch::kChunkTiles);
    GENPF_LAUNCH(k_opt_threshold, grid_1d(n), 256, s, (const double *)sorted, (const double *)W, (const Stats *)st_sorted, n, N, ctrl);
    GENPF_LAUNCH(k_opt_threshold_finish, 1, 1, s, (const double *)W, (const Stats *)st_lw, n, N, ctrl);
    // keep / remainder split
    uint8_t *fl = ob.flags.as<uint8_t>();
    long long *strat_idx = ob.strat_idx.as<long long>();
    double *strat_lw = ob.strat_lw.as<double>();
    GENPF_LAUNCH(k_opt_flags, (unsigned)ntiles, kOptThreads, s, d_lw, (const Stats *)st_lw, (const OptCtrl *)ctrl, n, fl, tile_cnt);
    GENPF_LAUNCH(k_opt_tile_scan, 1, 1024, s, (const int *)tile_cnt, ntiles, tile_off, ctrl);
    GENPF_LAUNCH((k_opt_compact<OutT>), (unsigned)ntiles, kOptThreads, s, d_lw, (const uint8_t *)fl, (const long long *)tile_off,
                 (const OptCtrl *)ctrl, n, N, d_parents, out_base, d_lw_out, strat_idx, strat_lw);
    GENPF_CUDA_TRY(cudaMemcpyAsync(h_ctrl, ctrl, sizeof(OptCtrl), cudaMemcpyDeviceToHost, s));
    GENPF_CUDA_TRY(cudaStreamSynchronize(s));
    res->n_keep = h_ctrl->n_keep;
    res->inv_w = h_ctrl->c;
    const int64_t n_keep = h_ctrl->n_keep, n_strat = n - n_keep, n_res = N - n_keep;
    if (n_res < 0) return fail(GENPF_ERR_ASSERT, "optimal resize: more particles kept than requested");
    if (n_res == 0) return GENPF_OK;
    if (n_strat <= 0) return fail(GENPF_ERR_ASSERT, "optimal resize: nothing left to draw from");
    // safe_softmax over the remainder (resize.jl:166) + its cumulative sums
    LwSrc rsrc{strat_lw, 1.0};
    const int64_t tpf_s = ceil_div(n_strat, kTile);
    GENPF_TRY(launch_reduce(s, rsrc, n_strat, 1, sc.partials(2)));
    GENPF_TRY(launch_finalize(s, sc, sc.partials(2), n_strat, 1, st_strat, sc.tile_off.as<double>(), -1.0, nullptr));
    StratArgs none_s = make_strat(UniSrc{nullptr, 0, 0, 0}, n_strat);
    GENPF_LAUNCH((k_scan<int32_t>), dim3((unsigned)tpf_s, 1), kScanThreads, s, rsrc, n_strat, tpf_s, (const Stats *)st_strat,
                 (const double *)sc.tile_off.as<double>(), WTables{W}, (int32_t *)nullptr, (int32_t *)nullptr, none_s, 0,
                 (const double *)nullptr, (int64_t)0, sc.chunk_info_ptr(n_strat), Scratch::kChunkTiles);
    const double step = 1.0 / (double)n_res;  // resize.jl:170
    const double u_rand = u_rand_or_null ? *u_rand_or_null : h_ctrl->u_rand;
    const double u0 = u_rand * step;          // resize.jl:171
    long long *Kc = ob.Kc.as<long long>();
    GENPF_LAUNCH(k_opt_counts, grid_1d(n_strat), 256, s, (const double *)W, n_strat, (long long)n_res, u0, step,
                 (const Stats *)st_strat, Kc);
    GENPF_LAUNCH((k_opt_select<OutT>), grid_1d(n_strat), 256, s, (const long long *)Kc, (const long long *)strat_idx, n_strat, N,
                 (const Stats *)st_strat, ctrl, d_parents, out_base, d_lw_out);
    GENPF_CUDA_TRY(cudaMemcpyAsync(h_ctrl, ctrl, sizeof(OptCtrl), cudaMemcpyDeviceToHost, s));
    GENPF_CUDA_TRY(cudaMemcpyAsync(h_stats + 1, st_strat, sizeof(Stats), cudaMemcpyDeviceToHost, s));
    GENPF_CUDA_TRY(cudaStreamSynchronize(s));
    res->kind_strat = h_stats[1].invalid_kind;
    if ((flags & GENPF_CHECK) && res->kind_strat != GENPF_VALID) return fail(GENPF_ERR_INVALID_WEIGHTS, "Invalid weights.");
    if (h_ctrl->n_selected != n_res)
        return fail(GENPF_ERR_ASSERT, "optimal resize: systematic pass selected a different number of particles than "
                                        "n_resample (reference: AssertionError, resize.jl:181)");
    return GENPF_OK;
}

}  // namespace genpf
