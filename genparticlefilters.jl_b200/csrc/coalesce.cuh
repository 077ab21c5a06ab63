// coalesce.cuh -- pf_coalesce! (reference src/resize.jl:309-334) as sort-by-key + segment sums.
//   key_i   : int64 standing for by(trace_i) (host path: supplied; device path: hash of the resident window)
//   first_i : smallest index holding the same key (the Dict's first insertion, resize.jl:317)
//   S_first : sum of exp(lw_i) over the group, un-shifted like the reference (resize.jl:318), added in index order
//   output  : one particle per group in ascending first-index order,
//             lw = log(S_first) + log(n_new) - log(n_old) (resize.jl:327)
// The reference emits groups in (unspecified) Dict order; callers compare as sets.
// The same grouping with NORMALISED weights summed instead is StatsBase.proportionmap (statistics.jl:91-130):
// proportion of each distinct value = sum of get_norm_weights over the group.
#pragma once
#include "engine.cuh"

namespace genpf {

// Group sums without atomics, in a fixed order: the sorted sequence is cut into chunks of kGroupChunk positions
// counted from each group's start; a chunk is summed sequentially in original-index order (the stable sort keeps
// it) by the thread sitting on its first position, and the thread on the group's first position adds the chunk
// sums in order.  A group of at most kGroupChunk particles is therefore summed exactly like the reference's
// `weights[key] += w` loop (resize.jl:315-320); larger groups differ from it by the association of the chunks only.
constexpr int64_t kGroupChunk = 4096;
struct GroupVal {
    const double *lw;
    const Stats *st_norm;  // null: exp(lw) un-shifted (coalesce); else safe_softmax weights (proportionmap)
    int64_t n;
    __device__ __forceinline__ double operator()(int64_t i, const Stats &st, bool uniform, double inv_S) const {
        const double v = lw[i];
        return st_norm ? (uniform ? 1.0 / (double)n : exp(v - st.M) * inv_S) : exp(v);
    }
};
__device__ __forceinline__ int64_t group_start(const int64_t *keys_sorted, int64_t s) {
    const int64_t key = keys_sorted[s];
    int64_t lo = 0, hi = s;  // first position with keys_sorted[pos] == key
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (keys_sorted[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
static __global__ void k_coalesce_chunks(const int64_t *keys_sorted, const int32_t *order, GroupVal val, int64_t n,
                                         double *chunk_sum) {
    Stats st{};
    bool uniform = false;
    double inv_S = 0.0;
    if (val.st_norm) {
        st = val.st_norm[0];
        uniform = st.invalid_kind == 2 || st.invalid_kind == 3;
        inv_S = 1.0 / st.S;
    }
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = group_start(keys_sorted, s);
        if ((s - lo) % kGroupChunk != 0) continue;
        const int64_t key = keys_sorted[s];
        double sum = 0.0;
        for (int64_t q = s; q < n && q < s + kGroupChunk && keys_sorted[q] == key; ++q) sum += val(order[q], st, uniform, inv_S);
        chunk_sum[s] = sum;
    }
}
static __global__ void k_coalesce_groups(const int64_t *keys_sorted, const int32_t *order, const double *chunk_sum,
                                         int64_t n, double *acc, int32_t *is_first) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        if (s > 0 && keys_sorted[s - 1] == keys_sorted[s]) continue;  // not a group's first position
        const int64_t key = keys_sorted[s];
        double total = 0.0;
        for (int64_t q = s; q < n && keys_sorted[q] == key; q += kGroupChunk) total += chunk_sum[q];
        const int32_t first = order[s];  // stable sort => smallest original index of the group
        acc[first] = total;
        is_first[first] = 1;
    }
}

// exclusive scan of int32 flags -> int64 positions (three phases over the shared 2048-tile partition)
static __global__ void __launch_bounds__(kThreads) k_flag_tile_sums(const int32_t *flags, int64_t n, long long *tile_sum) {
    __shared__ double sm[kWarps];
    const int64_t start = (int64_t)blockIdx.x * kTile;
    double c = 0.0;
    for (int e = threadIdx.x; e < kTile; e += kThreads)
        if (start + e < n) c += (double)flags[start + e];
    c = block_sum(c, sm);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = (long long)c;
}
static __global__ void k_flag_tile_offsets(long long *tile_sum, int64_t n_tiles, long long *total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long c = 0;
        for (int64_t b = 0; b < n_tiles; ++b) {
            long long t = tile_sum[b];
            tile_sum[b] = c;
            c += t;
        }
        *total = c;
    }
}
template <typename OutT>
static __global__ void __launch_bounds__(kThreads)
    k_coalesce_write(const int32_t *is_first, const long long *tile_off, const long long *total, const double *acc,
                     int64_t n, OutT *parents, int64_t out_base, double *lw_out, int proportions) {
    __shared__ long long smi[32];
    const int64_t start = (int64_t)blockIdx.x * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    long long fl[kItems], inc[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        int e = tile_elem(k);
        fl[k] = e < valid ? (long long)is_first[start + e] : 0;
    }
    tile_scan<long long>(fl, inc, smi);
    const long long off = tile_off[blockIdx.x];
    const double log_n_ratio = log((double)(*total)) - log((double)n);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        int e = tile_elem(k);
        if (e < valid && fl[k]) {
            long long pos = off + inc[k] - 1;
            parents[pos] = (OutT)(start + e + out_base);
            lw_out[pos] = proportions ? acc[start + e] : log(acc[start + e]) + log_n_ratio;
        }
    }
}

// 64-bit mix of the resident window's bits (device-path stand-in for by = get_choices)
struct HashCols {
    const double *f[2 * kMaxF];
    const uint8_t *b[2 * kMaxB];
    int nf, nb;
};
__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}
static __global__ void k_hash_window(HashCols c, int64_t n, int64_t *keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t h = 0x243F6A8885A308D3ull;
        for (int k = 0; k < c.nf; ++k) h = mix64(h, (uint64_t)__double_as_longlong(c.f[k][i]));
        for (int k = 0; k < c.nb; ++k) h = mix64(h, (uint64_t)c.b[k][i]);
        keys[i] = (int64_t)h;
    }
}

struct CoalesceBufs {
    DevBuf keys_sorted, order, acc, is_first, tile_sum, total, sort_tmp, chunk_sum;
    void release() {
        for (DevBuf *b : {&keys_sorted, &order, &acc, &is_first, &tile_sum, &total, &sort_tmp, &chunk_sum}) b->release();
    }
};

// device pointers in, device pointers out; *n_new_dev is a device long long
template <typename OutT>
int32_t launch_coalesce(cudaStream_t s, CoalesceBufs &cb, const double *lw, const int64_t *keys, int64_t n,
                        OutT *parents, int64_t out_base, double *lw_out, long long **n_new_dev,
                        const Stats *st_norm = nullptr) {
    const int64_t n_tiles = ceil_div(n, kTile);
    GENPF_TRY(cb.keys_sorted.ensure((size_t)n * 8));
    GENPF_TRY(cb.order.ensure((size_t)n * 4));
    GENPF_TRY(cb.acc.ensure((size_t)n * 8));
    GENPF_TRY(cb.is_first.ensure((size_t)n * 4));
    GENPF_TRY(cb.tile_sum.ensure((size_t)n_tiles * 8));
    GENPF_TRY(cb.total.ensure(8));
    GENPF_TRY(sort_keys_i64(keys, n, cb.keys_sorted.as<int64_t>(), cb.order.as<int32_t>(), cb.sort_tmp, s));
    GENPF_TRY(cb.chunk_sum.ensure((size_t)n * 8));
    GENPF_CUDA_TRY(cudaMemsetAsync(cb.is_first.p, 0, (size_t)n * 4, s));
    GENPF_LAUNCH(k_coalesce_chunks, grid_1d(n), 256, s, (const int64_t *)cb.keys_sorted.as<int64_t>(),
                 (const int32_t *)cb.order.as<int32_t>(), GroupVal{lw, st_norm, n}, n, cb.chunk_sum.as<double>());
    GENPF_LAUNCH(k_coalesce_groups, grid_1d(n), 256, s, (const int64_t *)cb.keys_sorted.as<int64_t>(),
                 (const int32_t *)cb.order.as<int32_t>(), (const double *)cb.chunk_sum.as<double>(), n, cb.acc.as<double>(),
                 cb.is_first.as<int32_t>());
    GENPF_LAUNCH(k_flag_tile_sums, (unsigned)n_tiles, kThreads, s, cb.is_first.as<int32_t>(), n, cb.tile_sum.as<long long>());
    GENPF_LAUNCH(k_flag_tile_offsets, 1, 32, s, cb.tile_sum.as<long long>(), n_tiles, cb.total.as<long long>());
    GENPF_LAUNCH((k_coalesce_write<OutT>), (unsigned)n_tiles, kThreads, s, cb.is_first.as<int32_t>(),
                 cb.tile_sum.as<long long>(), cb.total.as<long long>(), cb.acc.as<double>(), n, parents, out_base, lw_out,
                 st_norm ? 1 : 0);
    *n_new_dev = cb.total.as<long long>();
    return GENPF_OK;
}

}  // namespace genpf
