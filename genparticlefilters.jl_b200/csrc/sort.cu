// sort.cu -- K7: stable descending argsort of fp64 keys (sortperm(log_priorities, rev=true),
// reference src/resample.jl:156-157) and the key sort behind pf_coalesce! (resize.jl:309-334).
//
// Hand-written LSD radix sort ("onesweep" organisation), 8 digits of 8 bits over order-preserving 64-bit keys
// with a 32-bit index payload:
//   k_radix_prepare  key transform + index payload + ALL eight global digit histograms in one read of the keys
//   k_radix_plan     exclusive scan of each histogram; a digit on which every key agrees is marked "skip"
//                    (equal weights after a resample: all eight are skipped and the sort is one read)
//   k_radix_onesweep one kernel per remaining digit: stable in-tile ranks (each warp owns a contiguous
//                    256-key run; __match_any_sync ranks a 32-key chunk, per-warp digit counters carry the
//                    order across chunks and warps), the tile's global offsets come from a decoupled look-back
//                    over the per-tile digit counts (tiles take tickets, so every predecessor is resident), and
//                    the tile leaves in digit order through shared memory
//   k_fix_*          (>= 2^18 keys) only the five top digits are sorted; keys agreeing on them (within 4e-9
//                    relative) form runs in original-index order, runs containing an inversion are insertion-sorted in
//                    place, and a run too long for that un-gates a complete 8-digit sort whose kernels otherwise
//                    return at once (no host round trip)
//   k_sort_finish    inverse key transform, order -> caller's buffers (the ping-pong parity is device resident)
// Stability across tiles comes from the ticket order, inside a tile from the in-order walk, so equal keys keep
// ascending original index -- Julia's sortperm tie rule.  The key transform reproduces Julia's `isless` total
// order (-0.0 < 0.0).  Counts are integers, so the look-back changes no result: bit-deterministic.
#include "host.hpp"

namespace genpf {

constexpr int kSortThreads = 256;
constexpr int kSortPasses = 8;
constexpr uint32_t kFlagAgg = 0x40000000u, kFlagPrefix = 0x80000000u, kCountMask = 0x3FFFFFFFu;

__device__ __forceinline__ uint64_t order_bits(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // monotone fp64 -> u64, -0.0 < +0.0
}
__device__ __forceinline__ double order_bits_inv(uint64_t t) {
    uint64_t b = (t >> 63) ? (t & 0x7FFFFFFFFFFFFFFFull) : ~t;
    return __longlong_as_double((long long)b);
}
// MODE 0: fp64 keys, descending (ascending radix order of the complemented key); MODE 1: int64 keys, ascending;
// MODE 2: fp64 keys, ascending
template <int MODE>
__device__ __forceinline__ uint64_t key_encode(const void *keys, int64_t i) {
    if (MODE == 0) return ~order_bits(reinterpret_cast<const double *>(keys)[i]);
    if (MODE == 2) return order_bits(reinterpret_cast<const double *>(keys)[i]);
    return (uint64_t)reinterpret_cast<const int64_t *>(keys)[i] ^ 0x8000000000000000ull;
}

struct SortCtrl {
    uint32_t hist[kSortPasses][256];   // global digit histograms, then their exclusive scans
    uint32_t ticket[kSortPasses];      // next tile of each pass
    uint32_t skip[kSortPasses];        // every key has the same digit: the pass would be the identity
    uint32_t src_parity[kSortPasses];  // which ping-pong buffer holds the input of the pass
    uint32_t final_parity;
    uint32_t go;  // fallback sequence only: set by the fix-up when a run is too long to repair locally
    uint32_t low_const;  // every key has the same low digits: runs cannot contain inversions, no fix-up
};
constexpr int kLowDigits = 3;     // the hybrid sort orders by the top 64 - 8*kLowDigits bits, then repairs runs
constexpr int kFixWalk = 64;      // longest walk back to a run's start / longest run repaired in place: 2*kFixWalk

template <int MODE>
static __global__ void __launch_bounds__(kSortThreads)
    k_radix_prepare(const void *keys, int64_t n, uint64_t *k_out, int32_t *idx, SortCtrl *ctrl, int gated) {
    if (gated && !ctrl->go) return;
    __shared__ uint32_t h[kSortPasses][256];
    for (int d = threadIdx.x; d < kSortPasses * 256; d += kSortThreads) (&h[0][0])[d] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * kSortThreads; base < n; base += (int64_t)gridDim.x * kSortThreads) {
        const int64_t i = base + threadIdx.x;
        const bool ok = i < n;
        uint64_t k = 0;
        if (ok) {
            k = key_encode<MODE>(keys, i);
            k_out[i] = k;
            idx[i] = (int32_t)i;
        }
        // a warp whose keys all agree (equal weights) adds once per digit instead of 32 colliding atomics
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        if (act == 0xffffffffu && __all_sync(0xffffffffu, k == __shfl_sync(0xffffffffu, k, 0))) {
            if (lane < kSortPasses) atomicAdd(&h[lane][(k >> (8 * lane)) & 255u], 32u);
        } else if (ok) {
#pragma unroll
            for (int p = 0; p < kSortPasses; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kSortPasses * 256; d += kSortThreads) {
        const uint32_t c = (&h[0][0])[d];
        if (c) atomicAdd(&ctrl->hist[0][0] + d, c);
    }
}

static __global__ void __launch_bounds__(256) k_radix_plan(SortCtrl *ctrl, int64_t n, int first_digit, int gated) {
    if (gated && !ctrl->go) return;
    __shared__ uint32_t sw[8];
    __shared__ uint32_t s_skip[kSortPasses];
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    __shared__ uint32_t s_const[kSortPasses];
    if (d < kSortPasses) s_skip[d] = s_const[d] = 0;
    __syncthreads();
    for (int p = 0; p < kSortPasses; ++p) {
        const uint32_t c = ctrl->hist[p][d];
        if ((int64_t)c == n) s_const[p] = 1;
        if ((int64_t)c == n || p < first_digit) s_skip[p] = 1;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w)
            if (w < warp) woff += sw[w];
        ctrl->hist[p][d] = woff + inc - c;
        __syncthreads();
    }
    if (d == 0) {
        uint32_t parity = 0;
        for (int p = 0; p < kSortPasses; ++p) {
            ctrl->skip[p] = s_skip[p];
            ctrl->src_parity[p] = parity;
            if (!s_skip[p]) parity ^= 1u;
        }
        ctrl->final_parity = parity;
        uint32_t lc = 1;
        for (int p = 0; p < first_digit; ++p) lc &= s_const[p];
        ctrl->low_const = lc;
    }
}

// THREADS = 256 (tile 2048) or 512 (tile 4096): larger tiles halve the look-back work per key and space the
// tiles further apart in time, which is what keeps the look-back window short.
template <int THREADS, int MINB>
static __global__ void __launch_bounds__(THREADS, MINB)
    k_radix_onesweep(SortCtrl *ctrl, int pass, uint64_t *kA, uint64_t *kB, int32_t *vA, int32_t *vB, int64_t n,
                     uint32_t *state, int gated) {
    constexpr int NW = THREADS / 32, CH = 8, TILE = THREADS * CH;  // each warp owns a contiguous run of 256 keys
    if (gated && !ctrl->go) return;
    if (ctrl->skip[pass]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *skey = reinterpret_cast<uint64_t *>(smem_raw);           // [TILE] the tile in digit order
    int32_t *sval = reinterpret_cast<int32_t *>(skey + TILE);           // [TILE]
    uint32_t(*whist)[256] = reinterpret_cast<uint32_t(*)[256]>(sval + TILE);  // [NW][256] per-warp digit counts
    uint32_t *gdelta = &whist[NW][0];                                   // [256] global offset - tile-local start
    uint32_t *scan_tmp = gdelta + 256;                                  // [8]
    uint32_t *s_tile = scan_tmp + 8;
    const bool from_b = ctrl->src_parity[pass] != 0;
    const uint64_t *keys = from_b ? kB : kA;
    const int32_t *vals = from_b ? vB : vA;
    uint64_t *keys_out = from_b ? kA : kB;
    int32_t *vals_out = from_b ? vA : vB;
    const int shift = 8 * pass;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (threadIdx.x == 0) *s_tile = atomicAdd(&ctrl->ticket[pass], 1u);
    for (int d = threadIdx.x; d < NW * 256; d += THREADS) (&whist[0][0])[d] = 0;
    __syncthreads();
    const uint32_t tile = *s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int64_t base = tile_base + (int64_t)warp * (32 * CH);
    const int valid = (int)min((int64_t)TILE, n - tile_base);
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
    // 2. thread d < 256 owns digit d: tile count, exclusive scan over digits (tile-local start), then over
    //    warps; publish the count, look back over the preceding tiles for the digit's global offset
    {
        const int d = threadIdx.x;
        const bool own = d < 256;
        uint32_t cnt = 0;
        if (own) {
#pragma unroll
            for (int w = 0; w < NW; ++w) cnt += whist[w][d];
            *(volatile uint32_t *)(state + (size_t)tile * 256 + d) = (tile == 0 ? kFlagPrefix : kFlagAgg) | cnt;
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (own && lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        if (own) {
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (w < warp) woff += scan_tmp[w];
            uint32_t b = woff + inc - cnt;  // tile-local start of digit d
            // look-back, LB predecessors per round trip (independent loads); an unpublished entry restarts there
            uint32_t excl = 0;
            constexpr int LB = 8;
            int64_t t = (int64_t)tile - 1;
            bool done = t < 0;
            while (!done) {
                uint32_t pv[LB];
#pragma unroll
                for (int i = 0; i < LB; ++i)
                    pv[i] = t - i >= 0 ? *(volatile uint32_t *)(state + (size_t)(t - i) * 256 + d) : (uint32_t)0x80000000u;
                int used = 0;
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    if (done || used != i) continue;
                    if (pv[i] == 0u) continue;  // not published yet: re-read from t - i
                    excl += pv[i] & kCountMask;
                    used = i + 1;
                    if (pv[i] & kFlagPrefix) done = true;
                }
                t -= used;
            }
            if (tile != 0) *(volatile uint32_t *)(state + (size_t)tile * 256 + d) = kFlagPrefix | (excl + cnt);
            gdelta[d] = ctrl->hist[pass][d] + excl - b;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t t2 = whist[w][d];
                whist[w][d] = b;
                b += t2;
            }
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
    for (int s = threadIdx.x; s < valid; s += THREADS) {
        const uint64_t kk = skey[s];
        const uint32_t pos = gdelta[(kk >> shift) & 255u] + (uint32_t)s;
        keys_out[pos] = kk;
        vals_out[pos] = sval[s];
    }
}
template <int THREADS>
constexpr size_t onesweep_smem() {
    return (size_t)THREADS * 8 * 12 + (size_t)(THREADS / 32) * 1024 + 1024 + 64;
}

// ---- hybrid sort: after the passes over the top digits, keys that agree on those digits sit in one run in
// original-index order.  An inversion (key[s] < key[s-1] inside a run) marks the run's start dirty; dirty runs are
// insertion-sorted in place (stable).  Runs of equal keys -- replicated particles, equal weights, -Inf -- have no
// inversion and cost nothing.  A run longer than the walk limits raises `go`, which un-gates a full 8-digit sort.
static __global__ void k_fix_detect(const SortCtrl *ctrl, const uint64_t *kA, const uint64_t *kB, int64_t n, uint8_t *dirty,
                                    SortCtrl *fallback) {
    if (ctrl->low_const) return;
    const uint64_t *ks = ctrl->final_parity ? kB : kA;
    constexpr int sh = 8 * kLowDigits;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t a = ks[s - 1], b = ks[s];
        if ((a >> sh) != (b >> sh) || b >= a) continue;
        int64_t t = s - 1;
        int steps = 0;
        while (t > 0 && (ks[t - 1] >> sh) == (b >> sh)) {
            --t;
            if (++steps > kFixWalk) {
                fallback->go = 1;
                break;
            }
        }
        dirty[t] = 1;
    }
}
static __global__ void k_fix_sort(const SortCtrl *ctrl, uint64_t *kA, uint64_t *kB, int32_t *vA, int32_t *vB, int64_t n,
                                  const uint8_t *dirty, SortCtrl *fallback) {
    if (ctrl->low_const) return;
    uint64_t *ks = ctrl->final_parity ? kB : kA;
    int32_t *vs = ctrl->final_parity ? vB : vA;
    constexpr int sh = 8 * kLowDigits;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        if (!dirty[s]) continue;
        const uint64_t top = ks[s] >> sh;
        int64_t e = s + 1;
        while (e < n && (ks[e] >> sh) == top) {
            ++e;
            if (e - s > 2 * kFixWalk) break;
        }
        if (e - s > 2 * kFixWalk) {
            fallback->go = 1;
            continue;
        }
        for (int64_t i = s + 1; i < e; ++i) {  // stable insertion sort of the run by the full key
            const uint64_t k = ks[i];
            const int32_t v = vs[i];
            int64_t j = i;
            while (j > s && ks[j - 1] > k) {
                ks[j] = ks[j - 1];
                vs[j] = vs[j - 1];
                --j;
            }
            ks[j] = k;
            vs[j] = v;
        }
    }
}

template <int MODE>
static __global__ void k_sort_finish(const SortCtrl *ctrl, const SortCtrl *fallback, const uint64_t *kA, const uint64_t *kB,
                                     const int32_t *vA, const int32_t *vB, int64_t n, void *keys_sorted, int32_t *order32) {
    const SortCtrl *c = (fallback && fallback->go) ? fallback : ctrl;
    const bool in_b = c->final_parity != 0;
    const uint64_t *ks = in_b ? kB : kA;
    const int32_t *vs = in_b ? vB : vA;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (keys_sorted) {
            if (MODE == 0) reinterpret_cast<double *>(keys_sorted)[i] = order_bits_inv(~ks[i]);
            else if (MODE == 2) reinterpret_cast<double *>(keys_sorted)[i] = order_bits_inv(ks[i]);
            else reinterpret_cast<int64_t *>(keys_sorted)[i] = (int64_t)(ks[i] ^ 0x8000000000000000ull);
        }
        order32[i] = vs[i];
    }
}

template <int MODE>
static int32_t radix_sort(const void *keys, int64_t n, void *keys_sorted, int32_t *order32, DevBuf &tmp,
                          cudaStream_t stream) {
    if (n >= 0x40000000ll) return fail(GENPF_ERR_UNSUPPORTED, "sort: n must be < 2^30");
    // measured (2^22 / 2^24 / 2^26 keys): 2048-key tiles 0.49 / 1.66 / 6.66 ms, 4096-key tiles 0.53 / 1.63 / 5.92 ms
    const bool use_big = n >= (1 << 25);
    // hybrid (top digits + run repair, full sort only as a device-gated fallback) pays off once a pass costs more
    // than the dozen extra (mostly empty) launches
    const bool hybrid = n >= (1 << 18);
    const int64_t tile_keys = use_big ? 4096 : 2048;
    const int64_t ntiles = ceil_div(n, tile_keys);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const int n_seq = hybrid ? 2 : 1;
    const size_t a = al((size_t)n * 8), b = al((size_t)n * 4), c = al(sizeof(SortCtrl)),
                 st = al((size_t)ntiles * 256 * 4), dz = hybrid ? al((size_t)n) : 0;
    const size_t zero_bytes = n_seq * (c + kSortPasses * st) + dz;
    GENPF_TRY(tmp.ensure(2 * a + 2 * b + zero_bytes + 256));
    char *base = tmp.as<char>();
    uint64_t *kA = (uint64_t *)base, *kB = (uint64_t *)(base + a);
    int32_t *vA = (int32_t *)(base + 2 * a), *vB = (int32_t *)(base + 2 * a + b);
    char *z = base + 2 * a + 2 * b;
    SortCtrl *ctrl[2] = {(SortCtrl *)z, (SortCtrl *)(z + c + kSortPasses * st)};
    uint32_t *state[2] = {(uint32_t *)(z + c), (uint32_t *)(z + c + kSortPasses * st + c)};
    uint8_t *dirty = (uint8_t *)(z + n_seq * (c + kSortPasses * st));
    GENPF_CUDA_TRY(cudaMemsetAsync(z, 0, zero_bytes, stream));
    const unsigned gprep = (unsigned)std::min<int64_t>(ceil_div(n, kSortThreads), 148 * 16);
    // per device and cheap: set on every call (a process may drive several devices)
    if (use_big)
        GENPF_CUDA_TRY(cudaFuncSetAttribute(k_radix_onesweep<512, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)onesweep_smem<512>()));
    for (int seq = 0; seq < n_seq; ++seq) {
        // seq 0: the sort proper (hybrid: top digits only); seq 1: full sort, every kernel gated on ctrl[1]->go
        const int gated = seq, first_digit = (hybrid && seq == 0) ? kLowDigits : 0;
        GENPF_LAUNCH((k_radix_prepare<MODE>), gprep, kSortThreads, stream, keys, n, kA, vA, ctrl[seq], gated);
        GENPF_LAUNCH(k_radix_plan, 1, 256, stream, ctrl[seq], n, first_digit, gated);
        for (int pass = first_digit; pass < kSortPasses; ++pass) {
            uint32_t *stp = state[seq] + (size_t)pass * (st / 4);
            if (use_big)
                GENPF_LAUNCH_SMEM((k_radix_onesweep<512, 3>), (unsigned)ntiles, 512, onesweep_smem<512>(), stream, ctrl[seq], pass,
                                  kA, kB, vA, vB, n, stp, gated);
            else
                GENPF_LAUNCH_SMEM((k_radix_onesweep<256, 4>), (unsigned)ntiles, 256, onesweep_smem<256>(), stream, ctrl[seq], pass,
                                  kA, kB, vA, vB, n, stp, gated);
        }
        if (hybrid && seq == 0) {
            GENPF_LAUNCH(k_fix_detect, grid_1d(n), 256, stream, (const SortCtrl *)ctrl[0], (const uint64_t *)kA,
                         (const uint64_t *)kB, n, dirty, ctrl[1]);
            GENPF_LAUNCH(k_fix_sort, grid_1d(n), 256, stream, (const SortCtrl *)ctrl[0], kA, kB, vA, vB, n,
                         (const uint8_t *)dirty, ctrl[1]);
        }
    }
    GENPF_LAUNCH((k_sort_finish<MODE>), grid_1d(n), 256, stream, (const SortCtrl *)ctrl[0],
                 (const SortCtrl *)(hybrid ? ctrl[1] : nullptr), (const uint64_t *)kA, (const uint64_t *)kB, (const int32_t *)vA,
                 (const int32_t *)vB, n, keys_sorted, order32);
    return GENPF_OK;
}

// ------------------------------------------------------------------ batches of small filters (config 5)
// sortperm(log_priorities, rev=true) for EVERY filter of a batch in one launch: one block per filter sorts its
// n <= kSegSortMax keys in shared memory with a bitonic network over (encoded key, index) pairs.  Comparing the
// index on equal keys makes the network's result the stable order (Julia's sortperm tie rule) although the
// network itself is not stable.  Padding slots carry the largest pair and stay at the end.
constexpr int kSegSortMax = 4096;
constexpr int kSegSortThreads = 512;
static __global__ void __launch_bounds__(kSegSortThreads)
    k_segsort_desc(const double *keys, int n, int N, double *keys_sorted, int32_t *order32) {
    __shared__ uint64_t sk[kSegSortMax];
    __shared__ uint16_t si[kSegSortMax];
    const int64_t f = blockIdx.x;
    keys += f * n;
    for (int i = threadIdx.x; i < N; i += kSegSortThreads) {
        sk[i] = i < n ? key_encode<0>(keys, i) : ~0ull;
        si[i] = (uint16_t)(i < n ? i : 0xFFFF);
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += kSegSortThreads) {
                const int i = 2 * t - (t & (j - 1)), p = i + j;
                const uint64_t ka = sk[i], kb = sk[p];
                const uint16_t ia = si[i], ib = si[p];
                const bool a_first = ka < kb || (ka == kb && ia < ib);  // (a, b) already ascending
                const bool asc = (i & k) == 0;
                if (a_first != asc) {
                    sk[i] = kb; sk[p] = ka;
                    si[i] = ib; si[p] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += kSegSortThreads) {
        if (keys_sorted) keys_sorted[f * n + i] = order_bits_inv(~sk[i]);
        order32[f * n + i] = (int32_t)si[i];
    }
}

int32_t sort_desc_stable_batched(const double *keys, int64_t n, int64_t nf, double *keys_sorted, int32_t *order32,
                                 DevBuf &tmp, cudaStream_t stream) {
    if (nf > 1 && n <= kSegSortMax) {
        int N = 2;
        while (N < n) N <<= 1;
        GENPF_LAUNCH(k_segsort_desc, (unsigned)nf, kSegSortThreads, stream, keys, (int)n, N, keys_sorted, order32);
        return GENPF_OK;
    }
    for (int64_t f = 0; f < nf; ++f)
        GENPF_TRY(sort_desc_stable(keys + f * n, n, keys_sorted ? keys_sorted + f * n : nullptr, order32 + f * n, tmp, stream));
    return GENPF_OK;
}

int32_t sort_desc_stable(const double *keys, int64_t n, double *keys_sorted, int32_t *order32, DevBuf &tmp,
                         cudaStream_t stream) {
    return radix_sort<0>(keys, n, keys_sorted, order32, tmp, stream);
}

int32_t sort_asc_f64(const double *keys, int64_t n, double *keys_sorted, int32_t *order32, DevBuf &tmp,
                     cudaStream_t stream) {
    return radix_sort<2>(keys, n, keys_sorted, order32, tmp, stream);
}

int32_t sort_keys_i64(const int64_t *keys, int64_t n, int64_t *keys_sorted, int32_t *order32, DevBuf &tmp,
                      cudaStream_t stream) {
    return radix_sort<1>(keys, n, keys_sorted, order32, tmp, stream);
}

}  // namespace genpf
