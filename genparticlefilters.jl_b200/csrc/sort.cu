// sort.cu -- K7: stable descending argsort of fp64 keys (sortperm(log_priorities, rev=true),
// reference src/resample.jl:156-157) and the key sort behind pf_coalesce! (resize.jl:309-334).
//
// Hand-written LSD radix sort, 8 passes of 8 bits over order-preserving 64-bit keys with a 32-bit index
// payload.  Per pass (tile = 2048 keys, the same tile as everywhere else):
//   k_radix_hist    per-tile 256-bin digit histogram            -> hist[digit][tile]   (digit-major)
//   k_radix_scan*   exclusive scan of the flattened histogram   -> global offset of every (digit, tile)
//   k_radix_scatter stable in-tile ranks (each warp owns a contiguous 256-key run; __match_any_sync ranks
//                   a 32-key chunk, per-warp digit counters carry the order across chunks and warps)
// Stability across tiles comes from the digit-major scan, inside a tile from the in-order walk, so equal keys
// keep ascending original index -- Julia's sortperm tie rule.  The key transform reproduces Julia's `isless`
// total order (-0.0 < 0.0).
#include "host.hpp"

namespace genpf {

constexpr int kSortTile = 2048;
constexpr int kSortThreads = 256;

__device__ __forceinline__ uint64_t order_bits(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // monotone fp64 -> u64, -0.0 < +0.0
}
__device__ __forceinline__ double order_bits_inv(uint64_t t) {
    uint64_t b = (t >> 63) ? (t & 0x7FFFFFFFFFFFFFFFull) : ~t;
    return __longlong_as_double((long long)b);
}
static __global__ void k_sort_prepare(const double *keys, int64_t n, uint64_t *k_out, int32_t *idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        k_out[i] = ~order_bits(keys[i]);  // ascending radix order == descending key order
        idx[i] = (int32_t)i;
    }
}
static __global__ void k_sort_finish(const uint64_t *k_sorted, int64_t n, double *keys_sorted) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys_sorted[i] = order_bits_inv(~k_sorted[i]);
}
static __global__ void k_sort_prepare_i64(const int64_t *keys, int64_t n, uint64_t *k_out, int32_t *idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        k_out[i] = (uint64_t)keys[i] ^ 0x8000000000000000ull;
        idx[i] = (int32_t)i;
    }
}
static __global__ void k_sort_finish_i64(const uint64_t *k_sorted, int64_t n, int64_t *keys_sorted) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys_sorted[i] = (int64_t)(k_sorted[i] ^ 0x8000000000000000ull);
}

// ---- pass kernels
static __global__ void __launch_bounds__(kSortThreads)
    k_radix_hist(const uint64_t *keys, int64_t n, int shift, int64_t ntiles, uint32_t *hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    // plain shared-memory atomics: measured 2.3x faster than warp-aggregating with __match_any_sync
    // (MATCH costs ~18 issue slots on sm_100a)
#pragma unroll
    for (int c = 0; c < kSortTile / kSortThreads; ++c) {
        const int64_t i = base + c * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of m uint32 values in three phases (2048 per block)
static __global__ void __launch_bounds__(kSortThreads) k_radix_scan_sums(const uint32_t *a, int64_t m, uint32_t *sums) {
    __shared__ uint32_t sw[kSortThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < kSortTile / kSortThreads; ++c) {
        const int64_t i = base + c * kSortThreads + threadIdx.x;
        if (i < m) s += a[i];
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kSortThreads / 32; ++w) t += sw[w];
        sums[blockIdx.x] = t;
    }
}
static __global__ void __launch_bounds__(1024) k_radix_scan_offsets(uint32_t *sums, int64_t nb) {
    // one block, 1024 threads, 8 consecutive entries each per round, running carry
    __shared__ uint32_t sw[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < nb; base += 1024 * 8) {
        uint32_t v[8], run = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t i = base + (int64_t)threadIdx.x * 8 + c;
            v[c] = run;
            run += i < nb ? sums[i] : 0u;
        }
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 0;
        __syncthreads();
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            if (w < warp) woff += sw[w];
            tot += sw[w];
        }
        const uint32_t cr = carry;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int64_t i = base + (int64_t)threadIdx.x * 8 + c;
            if (i < nb) sums[i] = cr + woff + ex + v[c];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry = cr + tot;
        __syncthreads();
    }
}
static __global__ void __launch_bounds__(kSortThreads)
    k_radix_scan_apply(uint32_t *a, int64_t m, const uint32_t *offsets) {
    // in-place exclusive scan of one 2048-entry block: thread t owns 8 consecutive entries
    __shared__ uint32_t sw[kSortThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)threadIdx.x * 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v[8], run = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        v[c] = run;
        run += base + c < m ? a[base + c] : 0u;
    }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = 0;
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w)
        if (w < warp) woff += sw[w];
    const uint32_t off = offsets[blockIdx.x] + woff + ex;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (base + c < m) a[base + c] = off + v[c];
}

static __global__ void __launch_bounds__(kSortThreads)
    k_radix_scatter(const uint64_t *keys, const int32_t *vals, int64_t n, int shift, int64_t ntiles,
                    const uint32_t *offsets, uint64_t *keys_out, int32_t *vals_out) {
    constexpr int NW = kSortThreads / 32, CH = kSortTile / kSortThreads;  // 8 warps, 8 chunks of 32 per warp
    __shared__ uint32_t whist[NW][256];   // per-warp digit counts -> tile-local start of (warp, digit)
    __shared__ uint32_t gdelta[256];      // global offset of (digit, tile) minus the digit's tile-local start
    __shared__ uint32_t scan_tmp[NW];
    __shared__ uint64_t skey[kSortTile];  // the tile in digit order: runs of equal digit leave coalesced
    __shared__ int32_t sval[kSortTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int d = threadIdx.x; d < NW * 256; d += kSortThreads) (&whist[0][0])[d] = 0;
    const int64_t tile_base = (int64_t)blockIdx.x * kSortTile;
    const int64_t base = tile_base + (int64_t)warp * (32 * CH);
    const int valid = (int)min((int64_t)kSortTile, n - tile_base);
    uint64_t k[CH];
    int32_t v[CH];
    int dgt[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int64_t i = base + c * 32 + lane;
        const bool ok = i < n;
        k[c] = ok ? keys[i] : 0;
        v[c] = ok ? vals[i] : 0;
        dgt[c] = ok ? (int)((k[c] >> shift) & 255u) : 256;  // 256 = out of range: ranked among themselves, dropped
    }
    __syncthreads();
    // 1. per-warp digit counts of its contiguous 256-key run (the peer masks are kept for step 3: MATCH is
    //    the expensive instruction here)
    unsigned peers[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        peers[c] = __match_any_sync(0xffffffffu, dgt[c]);
        if (dgt[c] < 256 && (peers[c] & lt_mask) == 0) whist[warp][dgt[c]] += __popc(peers[c]);
        __syncwarp();
    }
    __syncthreads();
    // 2. thread d owns digit d: tile count, exclusive scan over digits (tile-local start), then over warps
    {
        const int d = threadIdx.x;  // kSortThreads == 256 digits
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) cnt += whist[w][d];
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (w < warp) woff += scan_tmp[w];
        uint32_t b = woff + inc - cnt;  // tile-local start of digit d
        gdelta[d] = offsets[(int64_t)d * ntiles + blockIdx.x] - b;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t t = whist[w][d];
            whist[w][d] = b;
            b += t;
        }
    }
    __syncthreads();
    // 3. in-order walk: rank inside the chunk + the warp's running counter of the digit -> slot in the tile
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        uint32_t pos = 0;
        if (dgt[c] < 256) pos = whist[warp][dgt[c]] + __popc(peers[c] & lt_mask);
        __syncwarp();
        if (dgt[c] < 256 && (peers[c] & lt_mask) == 0) whist[warp][dgt[c]] += __popc(peers[c]);
        __syncwarp();
        if (dgt[c] < 256) {
            skey[pos] = k[c];
            sval[pos] = v[c];
        }
    }
    __syncthreads();
    // 4. stream the tile out in digit order: slot s of digit d goes to gdelta[d] + s
    for (int s = threadIdx.x; s < valid; s += kSortThreads) {
        const uint64_t kk = skey[s];
        const uint32_t pos = gdelta[(kk >> shift) & 255u] + (uint32_t)s;
        keys_out[pos] = kk;
        vals_out[pos] = sval[s];
    }
}

// sorts (kA, vA) ascending by key, stably; kB/vB are ping-pong buffers; the result ends in (kA, vA)
static int32_t radix_sort_pairs(uint64_t *kA, uint64_t *kB, int32_t *vA, int32_t *vB, int64_t n, uint32_t *hist,
                                uint32_t *sums, cudaStream_t stream) {
    const int64_t ntiles = ceil_div(n, kSortTile);
    const int64_t m = 256 * ntiles, nb = ceil_div(m, kSortTile);
    uint64_t *ki = kA, *ko = kB;
    int32_t *vi = vA, *vo = vB;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 8 * pass;
        GENPF_LAUNCH(k_radix_hist, (unsigned)ntiles, kSortThreads, stream, (const uint64_t *)ki, n, shift, ntiles, hist);
        GENPF_LAUNCH(k_radix_scan_sums, (unsigned)nb, kSortThreads, stream, (const uint32_t *)hist, m, sums);
        GENPF_LAUNCH(k_radix_scan_offsets, 1, 1024, stream, sums, nb);
        GENPF_LAUNCH(k_radix_scan_apply, (unsigned)nb, kSortThreads, stream, hist, m, (const uint32_t *)sums);
        GENPF_LAUNCH(k_radix_scatter, (unsigned)ntiles, kSortThreads, stream, (const uint64_t *)ki,
                     (const int32_t *)vi, n, shift, ntiles, (const uint32_t *)hist, ko, vo);
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    return GENPF_OK;  // 8 passes: back in (kA, vA)
}

static int32_t layout_tmp(int64_t n, DevBuf &tmp, uint64_t *&kA, uint64_t *&kB, int32_t *&vB, uint32_t *&hist,
                          uint32_t *&sums) {
    if (n > 0x7FFFFFFFll) return fail(GENPF_ERR_UNSUPPORTED, "sort: n must be < 2^31");
    const int64_t ntiles = ceil_div(n, kSortTile);
    const int64_t m = 256 * ntiles, nb = ceil_div(m, kSortTile);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t a = al((size_t)n * 8), b = al((size_t)n * 4), h = al((size_t)m * 4), s = al((size_t)nb * 4);
    GENPF_TRY(tmp.ensure(2 * a + b + h + s + 256));
    char *base = tmp.as<char>();
    kA = (uint64_t *)base;
    kB = (uint64_t *)(base + a);
    vB = (int32_t *)(base + 2 * a);
    hist = (uint32_t *)(base + 2 * a + b);
    sums = (uint32_t *)(base + 2 * a + b + h);
    return GENPF_OK;
}

int32_t sort_desc_stable(const double *keys, int64_t n, double *keys_sorted, int32_t *order32, DevBuf &tmp,
                         cudaStream_t stream) {
    uint64_t *kA, *kB;
    int32_t *vB;
    uint32_t *hist, *sums;
    GENPF_TRY(layout_tmp(n, tmp, kA, kB, vB, hist, sums));
    GENPF_LAUNCH(k_sort_prepare, grid_1d(n), 256, stream, keys, n, kA, order32);
    GENPF_TRY(radix_sort_pairs(kA, kB, order32, vB, n, hist, sums, stream));
    if (keys_sorted) GENPF_LAUNCH(k_sort_finish, grid_1d(n), 256, stream, (const uint64_t *)kA, n, keys_sorted);
    return GENPF_OK;
}

int32_t sort_keys_i64(const int64_t *keys, int64_t n, int64_t *keys_sorted, int32_t *order32, DevBuf &tmp,
                      cudaStream_t stream) {
    uint64_t *kA, *kB;
    int32_t *vB;
    uint32_t *hist, *sums;
    GENPF_TRY(layout_tmp(n, tmp, kA, kB, vB, hist, sums));
    GENPF_LAUNCH(k_sort_prepare_i64, grid_1d(n), 256, stream, keys, n, kA, order32);
    GENPF_TRY(radix_sort_pairs(kA, kB, order32, vB, n, hist, sums, stream));
    if (keys_sorted) GENPF_LAUNCH(k_sort_finish_i64, grid_1d(n), 256, stream, (const uint64_t *)kA, n, keys_sorted);
    return GENPF_OK;
}

}  // namespace genpf
