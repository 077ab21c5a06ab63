// kernels.cuh -- weight-vector kernels (K1-K6, K9, K12, K13).
//
// Pipeline of one resample over n particles (per filter f; tile = 2048 particles):
//   k_reduce    : tile partials (max, sum e^{v-max}, sum e^{2(v-max)}, NaN/+Inf flags)     [R 8 B/particle]
//   k_finalize_fast: per filter M, S, S2, lse, ESS, invalid kind, exclusive tile offsets      [tiny]
//   k_scan      : w_i = e^{v_i-M}/S, in-tile fp64 inclusive scan + tile offset -> W_k,
//                 and/or cumulative offspring counts O_k = #{i : u_i <= W_k}                [R 8, W 8 or 4]
//   k_expand    : parent_i = min{k : O_k > i}  (load-balanced search, output centric)       [R ~4, W 4|8]
//   k_lookup    : parent_j = min{k : W_k > u_j} (multinomial, residual tail), guide table   [R ~3 sectors, W 4|8]
// The normalisation needs the global (M, S) before any cumulative weight exists, so a reduce pass is
// unavoidable; given that pass, "reduce-then-scan" with a shared tile partition gives every tile its
// exclusive prefix without the spinning of a decoupled look-back and is bit-deterministic.
#pragma once
#include "common.cuh"

namespace genpf {

struct Partials {
    double *m, *s, *s2;
    int *flags;  // bit0: NaN seen, bit1: +Inf seen
};

// grid = (tiles per filter, filters): no 64-bit division in any prologue
__device__ __forceinline__ void blk_to_tile(int64_t, int64_t &f, int64_t &tile) {
    f = blockIdx.y;
    tile = blockIdx.x;
}

// tile epilogue shared by every kernel that produces log-weights: the K1 partials of the tile.
// The maximum is reduced exactly in the integer key domain (REDUX), the sums with fixed shuffle trees in one
// warp: deterministic.  (A barrier-free variant -- per-warp maxima, last-arriving warp combines -- was
// measured and is slightly slower: the kernels are instruction-issue bound, not barrier bound.)
struct PartialSmem {
    double m[32], s[32], s2[32];
    int fl[32];
    unsigned count;
};
template <int T = kThreads>
__device__ __forceinline__ void emit_partials(const double (&v)[kTile / T], const Partials &out, PartialSmem &ps,
                                              int64_t slot = -1, double *e_tile = nullptr, int e_valid = kTile) {
    // e_tile (optional): this tile's slice of the filter's `ew` column; receives e_i = exp(v_i - m_tile), so the
    // next scan forms w_i = e_i * (exp(m_tile - M) / S) with one multiply instead of a second fp64 exp
    if (slot < 0) slot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;  // grid = (tiles, filters)
    constexpr int NW = T / 32;
    constexpr long long kMinKey = (long long)0x8000000000000000ull;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread maximum with fmax (NaN never wins), ONE key conversion per thread, REDUX across the warp
    double mt = v[0];
#pragma unroll
    for (int k = 1; k < kTile / T; ++k) mt = fmax(mt, v[k]);
    long long key = warp_max_key(f64_key(mt));
    long long *smk = reinterpret_cast<long long *>(ps.m);
    __syncthreads();
    if (lane == 0) smk[warp] = key;
    __syncthreads();
    // every thread needs the block max: lane l reads cell l mod NW, REDUX over the warp
    key = warp_max_key(smk[lane & (NW - 1)]);
    // all NaN / empty maps to the minimum key: treat as -Inf
    const double m = key == kMinKey ? -INFINITY : f64_from_key(key);
    double s = 0.0, s2 = 0.0;
    double ev[kTile / T];
#pragma unroll
    for (int k = 0; k < kTile / T; ++k) ev[k] = 0.0;
    int fl = 0;
    if (m == INFINITY) {
        fl = 2;
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) fl |= (v[k] != v[k]) ? 1 : 0;
    } else if (m > -INFINITY) {
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) {
            ev[k] = exp_nonpos(v[k] - m);  // NaN in, NaN out: the NaN flag is read off the sum (no per-element test)
            s += ev[k];
            s2 += ev[k] * ev[k];
        }
        fl = (s != s) ? 1 : 0;
        if (fl) {  // keep the sums finite like the per-element test did: NaN weights contribute nothing
            s = 0.0;
            s2 = 0.0;
#pragma unroll
            for (int k = 0; k < kTile / T; ++k) {
                if (ev[k] != ev[k]) ev[k] = 5e-308;
                s += ev[k];
                s2 += ev[k] * ev[k];
            }
        }
    } else {  // every weight of the tile is -Inf or NaN (block-uniform, rare): look for the NaNs explicitly
#pragma unroll
        for (int k = 0; k < kTile / T; ++k) fl |= (v[k] != v[k]) ? 1 : 0;
    }
    fl = __reduce_or_sync(0xffffffffu, (unsigned)fl);
    if (lane == 0) ps.fl[warp] = fl;
    if (e_tile) store_tile<double, T>(e_tile, 0, e_valid, ev);
    s = warp_sum(s);
    s2 = warp_sum(s2);
    if (lane == 0) {
        ps.s[warp] = s;
        ps.s2[warp] = s2;
    }
    __syncthreads();
    if (warp == 0) {
        s = ps.s[lane & (NW - 1)];
        s2 = ps.s2[lane & (NW - 1)];
        fl = __reduce_or_sync(0xffffffffu, (unsigned)ps.fl[lane & (NW - 1)]);
#pragma unroll
        for (int o = NW / 2; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            out.m[slot] = m;
            out.s[slot] = s;
            out.s2[slot] = s2;
            out.flags[slot] = fl;
        }
    }
}

#ifndef GENPF_PLUGIN_BUILD  // weight-vector kernels: compiled into the library only
// ------------------------------------------------------------------ K1/K2 reduce
// Replaces: Gen.logsumexp, lognorm/softmax (utils.jl:100-107), safe_softmax's validity scan
// (utils.jl:119-137), effective_sample_size (utils.jl:163-164).  Same block size and epilogue as the
// state-producing kernels, so the partials are bit-identical to theirs.
constexpr int kReduceThreads = 256;  // 8 particles per thread: four 16-byte loads in flight each, reductions amortised
#ifndef GENPF_REDUCE_MINB
#define GENPF_REDUCE_MINB 8  // 32 registers (a few spilled words), 8 blocks/SM: 0.164 ms at 2^26 against 0.175 at 40 registers
#endif                       // and 0.273 at 66 (the kernel lives on latency hiding: fp64 pipe + HBM both near half used)
static __global__ void __launch_bounds__(kReduceThreads, GENPF_REDUCE_MINB)
    k_reduce(LwSrc src, int64_t n, int64_t tpf, Partials out, double *ew = nullptr) {
    constexpr int T = kReduceThreads;
    __shared__ PartialSmem ps;
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    double v[kTile / T];
    load_tile<T>(src, f * n + start, valid, v, -INFINITY);
    emit_partials<T>(v, out, ps, -1, ew ? ew + f * n + start : nullptr, (int)valid);
}

// Finalize: one block per filter (or per chunk of a large filter).  Combines tile partials, classifies validity
// (utils.jl:119-137) and writes the exclusive tile offsets of the NORMALISED weights (the scan's carry-in).
//   lml_accum != null: log_ml_est[f] += lse - log(n)  (update_lml_est!, resample.jl:178-182), gated by do_resample
//   ess_frac < 0 => do_resample = 1, else do_resample = (ess < ess_frac * n)   (README.md:68)
// Register-resident finalize for tpf <= 8*THREADS tiles per filter: thread t owns tiles {c*THREADS + t}
// (coalesced loads, all in flight together); the three phases (max, rescaled sums, exclusive tile offsets)
// need only a handful of block-level combines.
template <int THREADS, int C = 8>
static __global__ void __launch_bounds__(THREADS)
    k_finalize_fast(Partials in, int64_t n_all, int64_t tpf_all, Stats *stats, double *tile_off, double ess_frac,
                    double *lml_accum, int64_t chunk_tiles, double *tile_scale = nullptr) {
    // grid = (chunks, filters).  chunk_tiles == tpf_all: the block finalises a whole filter.  Otherwise it
    // finalises one CHUNK of a large filter as if it were a filter of its own (Stats at [f*chunks + c], tile
    // offsets normalised within the chunk); k_chunk_combine then produces the filter's statistics and each
    // chunk's {prefix, scale}, which k_scan composes -- the same mechanism as a multi-GPU shard.
    pdl_enter();
    constexpr int NW = THREADS / 32;
    __shared__ double sm[3][NW];
    __shared__ int smi[NW];
    __shared__ double row_cell[C][NW];  // inclusive warp totals per row -> exclusive offsets
    __shared__ double row_total[C];
    const int64_t tile0 = (int64_t)blockIdx.x * chunk_tiles;
    const int64_t tpf = min(chunk_tiles, tpf_all - tile0);
    const int64_t n = min(n_all - tile0 * kTile, tpf * (int64_t)kTile);
    const int64_t f = blockIdx.y;
    in.m += f * tpf_all + tile0 - f * tpf;  // so that in.X[f*tpf + b] addresses tile (tile0 + b) of filter f
    in.s += f * tpf_all + tile0 - f * tpf;
    in.s2 += f * tpf_all + tile0 - f * tpf;
    in.flags += f * tpf_all + tile0 - f * tpf;
    if (tile_off) tile_off += f * tpf_all + tile0 - f * tpf;
    if (tile_scale) tile_scale += f * tpf_all + tile0 - f * tpf;
    stats += (int64_t)blockIdx.y * gridDim.x + blockIdx.x - f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double pm[C], ps[C], ps2[C];
    int fl = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int64_t b = (int64_t)c * THREADS + threadIdx.x;
        const bool ok = b < tpf;
        pm[c] = ok ? in.m[f * tpf + b] : -INFINITY;
        ps[c] = ok ? in.s[f * tpf + b] : 0.0;
        ps2[c] = ok ? in.s2[f * tpf + b] : 0.0;
        fl |= ok ? in.flags[f * tpf + b] : 0;
    }
    double m = pm[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmax(m, pm[c]);
    m = warp_max(m);
    fl = __reduce_or_sync(0xffffffffu, (unsigned)fl);
    if (lane == 0) {
        sm[0][warp] = m;
        smi[warp] = fl;
    }
    __syncthreads();
    double M = sm[0][0];
    fl = smi[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
        M = fmax(M, sm[0][w]);
        fl |= smi[w];
    }
    __syncthreads();
    double sc[C], s = 0.0, s2 = 0.0;
    const bool finite_max = (M > -INFINITY && M < INFINITY);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        sc[c] = (finite_max && pm[c] > -INFINITY) ? exp(pm[c] - M) : 0.0;
        s += ps[c] * sc[c];
        s2 += ps2[c] * (sc[c] * sc[c]);
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    if (lane == 0) {
        sm[0][warp] = s;
        sm[1][warp] = s2;
    }
    __syncthreads();
    double S = 0.0, S2 = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        S += sm[0][w];
        S2 += sm[1][w];
    }
    int kind = 0;
    if (fl & 1) kind = 1;
    else if (M == -INFINITY) kind = 2;
    else if ((fl & 2) || isnan(S)) kind = 4;
    else if (S == 0.0) kind = 3;
    const double lse = (M == -INFINITY) ? -INFINITY : M + log(S);
    const double ess = S * S / S2;
    int do_rs = 1;
    if (ess_frac >= 0.0) do_rs = (ess < ess_frac * (double)n) ? 1 : 0;
    if (kind == 1 || kind == 4) do_rs = 0;
    if (threadIdx.x == 0) {
        Stats st;
        st.M = M; st.S = S; st.S2 = S2; st.lse = lse; st.ess = ess;
        st.invalid_kind = kind; st.do_resample = do_rs;
        stats[f] = st;
        if (lml_accum && do_rs) lml_accum[f] += lse - log((double)n);
    }
    if (!tile_off) return;
    // exclusive scan over tiles in index order b = c*THREADS + t: rows c are contiguous runs of THREADS tiles
    const bool uniform = (kind == 2 || kind == 3);
    const double inv_n = 1.0 / (double)n;
    double ex[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int64_t b = (int64_t)c * THREADS + threadIdx.x;
        double v = 0.0;
        if (b < tpf) {
            if (uniform) v = (double)min((int64_t)kTile, n - b * kTile) * inv_n;
            else if (kind == 0) v = ps[c] * sc[c] / S;
            // per-tile factor turning e_i = exp(lw_i - m_tile) into the normalised weight
            if (tile_scale) tile_scale[f * tpf + b] = kind == 0 ? sc[c] / S : 0.0;
        }
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        double e = __shfl_up_sync(0xffffffffu, inc, 1);
        ex[c] = lane == 0 ? 0.0 : e;
        if (lane == 31) row_cell[c][warp] = inc;
    }
    __syncthreads();
    for (int row = warp; row < C; row += NW) {  // one warp scans a row's NW warp totals
        double t = lane < NW ? row_cell[row][lane] : 0.0;
        double inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        double e = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane < NW) row_cell[row][lane] = lane == 0 ? 0.0 : e;
        if (lane == 31) row_total[row] = inc;
    }
    __syncthreads();
    double roff = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int64_t b = (int64_t)c * THREADS + threadIdx.x;
        if (b < tpf) tile_off[f * tpf + b] = roff + (row_cell[c][warp] + ex[c]);
        roff += row_total[c];
    }
}

// ------------------------------------------------------------------ stratified thresholds
// u_i = r_i*(1/n) + lower_i with two roundings and no FMA (resample.jl:162); lower_i = element i of
// 0.0:1/n:1.0-1/n, i.e. (i-1)/n (exact for power-of-two n).
struct StratArgs {
    UniSrc uni;
    double step;  // 1/n
    int64_t n;
    int pow2;
    double tol;     // rounding slack of the stratum-count fast path, in stratum units
    int64_t guide;  // B > 0: no strata; O_k = guide-table counts ceil(B * W_k) for the inverse-CDF lookup (k_lookup)
};
// Guide table for the inverse-CDF draws (multinomial, residual tail): bucket b of B covers u in [b/B, (b+1)/B);
// G[b] = min{k : ceil(fl(B*W_k)) > b} is where a lookup with floor(fl(B*u)) == b starts.  The counts only need
// to be monotone: k_lookup verifies every answer against W itself.
template <typename J>
__device__ __forceinline__ J guide_count(double W, int64_t B) {
    const double nd = (double)B;
    const double c = ceil(W * nd);
    return c >= nd ? (J)B : (c <= 0.0 ? (J)0 : (J)c);
}
__device__ __forceinline__ double strat_lower(const StratArgs &a, int64_t i1) {
    const double im1 = a.n < 0x7FFFFFFFll ? (double)(int)(i1 - 1) : (double)(i1 - 1);
    return a.pow2 ? im1 * a.step : im1 / (double)a.n;
}
__device__ __forceinline__ double strat_u(const StratArgs &a, int64_t f, int64_t i1) {
    const int64_t slot = f * a.n + i1 - 1;
    double r;
    if (a.uni.col) r = a.uni.col[slot];
    else r = strata_word(philox_at(a.uni.seed, a.uni.stream, (uint64_t)(slot + a.uni.offset) >> 2), slot + a.uni.offset);
    return __dadd_rn(__dmul_rn(r, a.step), strat_lower(a, i1));
}
// C(W) = #{i in 1..n : u_i <= W}.  Because u is non-decreasing in i this is a prefix count, and
// parent_i = min{k : W_k >= u_i} (resample.jl:163-168) == min{k : C(W_k) >= i}.
// J = int32 when n < 2^31 (single-instruction fp64<->int conversions), else int64.
// The Philox rounds live in ONE out-of-line function (inlined at every call site k_scan had grown past
// 100 KB of SASS and thrashed the I-cache); a per-thread cache of the last block lets the four consecutive
// particles of a thread share draws.
static __device__ __noinline__ uint4 strata_block(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return philox_at(seed, stream, ctr);
}
// Warp-cooperative stratum draws: the 128 consecutive particles of a warp query ~128 consecutive strata, so
// each lane draws the Philox blocks of 8 strata of a 256-stratum window (2 calls per lane instead of one per
// particle, and no divergence); a query outside the window falls back to a direct draw.
constexpr int kStrataWindow = 256;
struct StrataWindow {
    const uint32_t *words;  // shared memory: `len` stratum words (a warp's window in k_scan, the tile's in k_scan_hot)
    int64_t base;           // global stratum slot of words[0]; < 0: no window (column uniforms)
    int len;
};
// General path: exact for every input (column uniforms, window misses, W within rounding of a stratum edge).
template <typename J>
static __device__ __noinline__ J strat_count_slow(const StratArgs &a, int64_t slot0, double W, const uint32_t *words,
                                                  int64_t win_base, int win_len) {
    const J n = (J)a.n;
    const double nd = (double)a.n;
    const double x = W * nd;
    J j = x >= nd ? n : (x <= 0.0 ? (J)0 : (J)x);
    const bool near_edge = (x - (double)j) < 1e-6;
    auto lower_of = [&](J i1) { return a.pow2 ? (double)(i1 - 1) * a.step : (double)(i1 - 1) / nd; };
    auto u_of = [&](J i1) {
        const int64_t slot = slot0 + (int64_t)i1 - 1;
        double r;
        if (a.uni.col) {
            r = a.uni.col[slot];
        } else {
            const int64_t gs = slot + a.uni.offset;
            const uint64_t rel = (uint64_t)(gs - win_base);
            if (win_base >= 0 && rel < (uint64_t)win_len) {
                r = ((double)words[rel] + 0.5) * 0x1.0p-32;
            } else {
                r = strata_word(strata_block(a.uni.seed, a.uni.stream, (uint64_t)gs >> 2), gs);
            }
        }
        return __dadd_rn(__dmul_rn(r, a.step), lower_of(i1));
    };
    // u_i >= lower_i, so a stratum whose lower bound already exceeds W needs no draw
    while (j < n && lower_of(j + 1) <= W && u_of(j + 1) <= W) ++j;
    if (near_edge)
        while (j > 0 && u_of(j) > W) --j;
    return j;
}
// Fast path (library-drawn strata inside the warp's window): with x = n*W, j = floor(x) and frac = x - j well
// inside (0, 1), strata 1..j lie below W, stratum j+2 above, and stratum j+1 counts iff r_{j+1} < frac -- decided
// without forming u when |frac - r| exceeds the rounding slack a.tol (4 x the worst-case error of n*W and of u);
// anything closer goes to the exact general path, so the count is identical to it in every case.
template <typename J>
__device__ __forceinline__ J strat_count(const StratArgs &a, int64_t slot0, double W, const StrataWindow &win) {
    const double nd = (double)a.n;
    const double x = W * nd;
    if (win.base >= 0 && x > 0.0 && x < nd) {
        const J j = (J)x;
        const double frac = x - (double)j;
        const uint64_t rel = (uint64_t)(slot0 + (int64_t)j + a.uni.offset - win.base);
        if (rel < (uint64_t)win.len && frac > 1e-6 && frac < 1.0 - 1e-6) {
            const double r = fma((double)win.words[rel], 0x1.0p-32, 0x1.0p-33);  // (word + 0.5) * 2^-32, exact
            const double d = frac - r;
            if (fabs(d) > a.tol) return j + (J)(d > 0.0 ? 1 : 0);
        }
        // n*W within 1e-6 of an integer (equal weights put EVERY particle here): the general path's two loops,
        // evaluated from the window with a bounded trip count; anything unusual falls through to it
        const J n = (J)a.n;
        auto lower_of = [&](J i1) { return a.pow2 ? (double)(i1 - 1) * a.step : (double)(i1 - 1) / nd; };
        bool ok = true;
        auto u_win = [&](J i1) {
            const uint64_t rr = (uint64_t)(slot0 + (int64_t)i1 - 1 + a.uni.offset - win.base);
            if (rr >= (uint64_t)win.len) {
                ok = false;
                return 0.0;
            }
            const double r = fma((double)win.words[rr], 0x1.0p-32, 0x1.0p-33);
            return __dadd_rn(__dmul_rn(r, a.step), lower_of(i1));
        };
        J c = j;
        int trips = 0;
        while (ok && c < n && lower_of(c + 1) <= W) {
            const double u = u_win(c + 1);
            if (!ok || !(u <= W)) break;
            ++c;
            if (++trips > 3) ok = false;
        }
        if (ok && frac < 1e-6) {
            trips = 0;
            while (ok && c > 0) {
                const double u = u_win(c);
                if (!ok || !(u > W)) break;
                --c;
                if (++trips > 3) ok = false;
            }
        }
        if (ok) return c;
    }
    return strat_count_slow<J>(a, slot0, W, win.words, win.base, win.len);
}

// the whole exact count behind one call: k_scan_hot keeps only the common case inline
template <typename J>
static __device__ __noinline__ J strat_count_outline(const StratArgs &a, int64_t slot0, double W, const uint32_t *words,
                                                     int64_t win_base, int win_len) {
    return strat_count<J>(a, slot0, W, StrataWindow{words, win_base, win_len});
}

// ------------------------------------------------------------------ multi-GPU statistics exchange
// Post this shard's (M, S, S2) into every peer's exchange block, wait for everybody's, derive the population's
// statistics and this shard's {prefix, share} of the normalised mass.  Called by ONE warp: lane g serves peer g,
// lane 0 writes the results.  The posting rides on the kernel that produced the local statistics (k_chunk_combine),
// so the exchange costs no launch of its own.
// Closing offspring counts without an exchange: every rank derives ALL shards' cumulative mass W_end(g) = prefix(g+1)
// from the same gathered totals with the same arithmetic, so O_end(g) = C(W_end(g)) (the exact stratified count; the
// stratum uniforms are counter-based) is identical on every rank.  The scan then pins its shard to [O_end(rank-1),
// O_end(rank)] (clamp + closing particle), which makes the ranks' output ranges an exact cover by construction; the
// pinned value differs from the scan's own last count only when W_end computed by the two summation orders
// straddles a stratum threshold -- inside the documented fp64 cumulative-sum tie class.
__device__ __forceinline__ void xchg_stats_combine(const XchgLink &lk, double M_loc, double S_loc, double S2_loc,
                                                   int64_t n_total, Stats *stats, double *shard_info, double *lml_accum,
                                                   const StratArgs *strat = nullptr, long long *oend_out = nullptr) {
    const int g = threadIdx.x & 31;
    const int world = lk.world, rank = lk.rank;
    Xchg *mine = lk.peers.x[rank];
    if (g < world) {
        Xchg *dst = lk.peers.x[g];
        dst->stats[rank][0] = M_loc;
        dst->stats[rank][1] = S_loc;
        dst->stats[rank][2] = S2_loc;
        __threadfence_system();
        *(volatile unsigned long long *)&dst->flag_stats[rank] = lk.epoch;
        xchg_wait(&mine->flag_stats[g], lk.epoch, &mine->error);
    }
    __syncwarp();
    // Lane r holds shard r; the exps, divisions and closing counts run in parallel lanes (a single thread doing them
    // one after the other was ~9 us of dependent fp64 latency at 8 ranks), while every SUM is taken in rank order
    // through shuffles, so all ranks (and the NCCL variant's one-thread combine) get bit-identical totals.
    const volatile double *gathered = &mine->stats[0][0];
    const bool have = g < world;
    const double m_r = have ? gathered[3 * g] : -INFINITY;
    const double s_r = have ? gathered[3 * g + 1] : 0.0, s2_r = have ? gathered[3 * g + 2] : 0.0;
    const bool nan = __any_sync(0xffffffffu, have && (isnan(m_r) || isnan(s_r)));
    const double M = warp_max(m_r);
    const bool finite = M > -INFINITY && M < INFINITY;
    const double sc = (have && finite && m_r > -INFINITY) ? exp(m_r - M) : 0.0;
    const double a_r = s_r * sc, b_r = s2_r * (sc * sc);
    double S = 0.0, S2 = 0.0;
    for (int r = 0; r < world; ++r) {
        S += __shfl_sync(0xffffffffu, a_r, r);
        S2 += __shfl_sync(0xffffffffu, b_r, r);
    }
    int kind = 0;
    if (nan) kind = 1;
    else if (M == -INFINITY) kind = 2;
    else if (M == INFINITY || isnan(S)) kind = 4;
    else if (S == 0.0) kind = 3;
    const bool uniform = kind == 2 || kind == 3;  // uniform fallback (utils.jl:123-133): every shard carries 1/world
    const double share = uniform ? 1.0 / (double)world : a_r / S;
    double prefix = 0.0, run = 0.0, run_mine = 0.0;  // prefix(r), run = prefix(r + 1): left-to-right sums
    for (int r = 0; r < world; ++r) {
        const double sh = __shfl_sync(0xffffffffu, share, r);
        if (r == g) prefix = run;
        run = uniform ? (double)(r + 1) / (double)world : run + sh;
        if (r == g) run_mine = run;
    }
    if (uniform) prefix = (double)g / (double)world;
    if (oend_out && have) {
        long long c;
        if (g == world - 1 || kind == 1 || kind == 4) c = (n_total / world) * (long long)(g + 1);
        else if (strat->n < 0x7FFFFFFFll) c = (long long)strat_count_slow<int32_t>(*strat, 0, run_mine, nullptr, -1, 0);
        else c = strat_count_slow<long long>(*strat, 0, run_mine, nullptr, -1, 0);
        oend_out[g] = c;
    }
    if (g == rank) {
        shard_info[0] = prefix;
        shard_info[1] = share;
    }
    if (g != 0) return;
    Stats st;
    st.M = M; st.S = S; st.S2 = S2;
    st.lse = (M == -INFINITY) ? -INFINITY : M + log(S);
    st.ess = S * S / S2;
    st.invalid_kind = kind;
    st.do_resample = (kind == 1 || kind == 4) ? 0 : 1;
    stats[0] = st;
    if (lml_accum && st.do_resample) lml_accum[0] += st.lse - log((double)n_total);
}
// Large filters (more than kChunkTiles tiles): combine the per-chunk statistics into the filter's, and give every
// chunk its {prefix, scale} (exclusive normalised mass before the chunk, the chunk's share).  One warp per filter;
// lane l owns chunks l, l+32, ...; sums use fixed shuffle trees, the prefix is an ordered warp scan with a carry.
static __global__ void __launch_bounds__(32)
    k_chunk_combine(const Stats *chunk_stats, int nchunks, int64_t n, int64_t chunk_particles, Stats *stats,
                    double *chunk_info, double ess_frac, double *lml_accum, XchgLink link = XchgLink{{}, 0, 0, 0},
                    int64_t n_total = 0, double *shard_info = nullptr, StratArgs strat = StratArgs{},
                    long long *oend_out = nullptr) {
    pdl_enter();
    const int64_t f = blockIdx.x;
    const int lane = threadIdx.x;
    const Stats *cs = chunk_stats + f * nchunks;
    double M = -INFINITY;
    int bits = 0;  // 1: NaN input somewhere, 2: NaN total somewhere, 4: some chunk is not all -Inf
    for (int c = lane; c < nchunks; c += 32) {
        M = fmax(M, cs[c].M);
        const int k = cs[c].invalid_kind;
        bits |= (k == 1 ? 1 : 0) | (k == 4 ? 2 : 0) | (k != 2 ? 4 : 0);
    }
    M = warp_max(M);
    bits = __reduce_or_sync(0xffffffffu, (unsigned)bits);
    const bool finite = M > -INFINITY && M < INFINITY;
    double S = 0.0, S2 = 0.0;
    for (int c = lane; c < nchunks; c += 32) {
        const double sc = (finite && cs[c].M > -INFINITY) ? exp(cs[c].M - M) : 0.0;
        S += cs[c].S * sc;
        S2 += cs[c].S2 * (sc * sc);
    }
    S = warp_sum(S);
    S2 = warp_sum(S2);
    int kind = 0;
    if (bits & 1) kind = 1;
    else if (!(bits & 4)) kind = 2;
    else if ((bits & 2) || isnan(S)) kind = 4;
    else if (S == 0.0) kind = 3;
    double carry = 0.0;
    for (int c0 = 0; c0 < nchunks; c0 += 32) {
        const int c = c0 + lane;
        double share = 0.0;
        if (c < nchunks) {
            if (kind == 2 || kind == 3) {
                const int64_t cnt = min(chunk_particles, n - (int64_t)c * chunk_particles);
                share = (double)cnt / (double)n;
            } else {
                const double sc = (finite && cs[c].M > -INFINITY) ? exp(cs[c].M - M) : 0.0;
                share = cs[c].S * sc / S;
            }
        }
        double inc = share;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        double ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 0.0;
        if (c < nchunks) {
            chunk_info[2 * (f * nchunks + c)] = carry + ex;
            chunk_info[2 * (f * nchunks + c) + 1] = share;
        }
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (link.world > 0) {  // a shard of a multi-GPU population: the local totals go straight into the exchange
        // NaN inputs only set a flag locally (fmax semantics): a NaN total carries the diagnosis to every rank
        xchg_stats_combine(link, M, (kind == 1 || kind == 4) ? NAN : S, S2, n_total, stats, shard_info, lml_accum, &strat,
                           oend_out);
        return;
    }
    if (lane != 0) return;
    Stats st;
    st.M = M; st.S = S; st.S2 = S2;
    st.lse = (M == -INFINITY) ? -INFINITY : M + log(S);
    st.ess = S * S / S2;
    st.invalid_kind = kind;
    int do_rs = 1;
    if (ess_frac >= 0.0) do_rs = (st.ess < ess_frac * (double)n) ? 1 : 0;
    if (kind == 1 || kind == 4) do_rs = 0;
    st.do_resample = do_rs;
    stats[f] = st;
    if (lml_accum && do_rs) lml_accum[f] += st.lse - log((double)n);
}

// ------------------------------------------------------------------ K3 normalise + scan
// Replaces safe_softmax line utils.jl:139 and the running accum_weight of resample.jl:163-166.
// Writes W (normalised inclusive cumulative weights) and/or O (cumulative offspring counts, stratified).
constexpr int kScanThreads = 512;  // every launch of k_scan uses this block size (fixed summation order)
// cumulative-weight output of the scan (inverse-CDF draws, genpf_debug_cumweights)
struct WTables {
    double *W;
};

// k_scan uses a BLOCKED layout: thread t owns the four consecutive particles 4t..4t+3 of the tile (two adjacent
// 16-byte loads, one 16-byte store of the counts).  One sequential 4-add chain + one warp scan per thread
// instead of two pair scans, and neighbouring particles query neighbouring strata, so a thread's four
// threshold draws usually come from one cached Philox block.
template <typename IdxT>
static __global__ void __launch_bounds__(kScanThreads, 4)
    k_scan(LwSrc src, int64_t n, int64_t tpf, const Stats *stats, const double *tile_off, WTables wt, IdxT *O_out,
           IdxT *tile_last_O, StratArgs strat, int gate, const double *shard_info = nullptr,
           int64_t global_base = 0, const double *chunk_info = nullptr, int64_t chunk_tiles = 0,
           const double *ew = nullptr, const double *tile_scale = nullptr, const long long *oend_pin = nullptr,
           int shard_rank = 0) {
    // ew / tile_scale (device filters): e_i = exp(lw_i - m_tile) stored by the kernel that produced lw and the
    // per-tile factor exp(m_tile - M)/S from the finalize: w_i = e_i * factor, no exp in this kernel.
    // shard_info (multi-GPU particle sharding): {prefix, scale}: this shard's cumulative weights are
    // prefix + scale * (locally normalised tile offsets) + in-tile sums of globally normalised weights.
    pdl_enter();
    constexpr int T = kScanThreads, I = 4, NW = T / 32;
    static_assert(T * I == kTile, "blocked scan layout");
    __shared__ double sm[NW];
    __shared__ alignas(16) uint32_t swin[NW][kStrataWindow];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const Stats st = stats[f];
    if (st.invalid_kind == 1 || st.invalid_kind == 4) return;
    if (gate && !st.do_resample) return;
    const int64_t start = tile * kTile;
    const int valid = (int)min((int64_t)kTile, n - start);
    const int e0 = 4 * (int)threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[I];
    const bool use_e = ew != nullptr;
    {
        const double *p = (use_e ? ew : src.p) + f * n + start;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && e0 + 3 < valid;
        if (vec_ok) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(p + e0));
            const double2 b = __ldg(reinterpret_cast<const double2 *>(p + e0 + 2));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (int k = 0; k < I; ++k) v[k] = e0 + k < valid ? __ldg(p + e0 + k) : (use_e ? 0.0 : -INFINITY);
        }
        if (!use_e) {
#pragma unroll
            for (int k = 0; k < I; ++k) v[k] = src.fix(v[k]);
        }
    }
    const bool uniform = (st.invalid_kind == 2 || st.invalid_kind == 3);
    const double inv_n = 1.0 / (double)strat.n;  // == n except for a shard of a multi-GPU population
    // w_i = e_i / S evaluated as e_i * (1/S): at most 1 ulp from the reference's division, far inside the
    // sequential-vs-parallel cumulative-sum noise that defines the documented tie class
    const double inv_S = 1.0 / st.S;
    double escale = 0.0;
    if (use_e) {
        escale = tile_scale[f * tpf + tile];
        if (chunk_info) escale *= chunk_info[2 * (f * ((tpf + kChunkTiles - 1) >> kChunkLog2) + (tile >> kChunkLog2)) + 1];
        if (shard_info) escale *= shard_info[1];
    }
    double W[I];
    {
        double run = 0.0;
#pragma unroll
        for (int k = 0; k < I; ++k) {
            double w;
            if (uniform) w = e0 + k < valid ? inv_n : 0.0;
            else if (use_e) w = v[k] * escale;
            else w = exp_nonpos(v[k] - st.M) * inv_S;
            run += w;
            W[k] = run;  // inclusive within the thread
        }
    }
    double inc = W[I - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    double lane_excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) lane_excl = 0.0;
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 16 warp totals
        double t = lane < NW ? sm[lane] : 0.0;
        double s = t;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        const double ex = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane < NW) sm[lane] = lane == 0 ? 0.0 : ex;
    }
    __syncthreads();
    double off = tile_off[f * tpf + tile];
    if (chunk_info) {  // large filter: offsets were normalised per chunk of chunk_tiles tiles
        const int64_t nchunks = (tpf + kChunkTiles - 1) >> kChunkLog2;
        const double *ci = chunk_info + 2 * (f * nchunks + (tile >> kChunkLog2));
        off = ci[0] + ci[1] * off;
    }
    if (shard_info) off = shard_info[0] + shard_info[1] * off;
    const double base_w = off + (sm[warp] + lane_excl);
#pragma unroll
    for (int k = 0; k < I; ++k) W[k] = base_w + W[k];
    if (wt.W) {
        double *pw = wt.W + f * n + start;
#pragma unroll
        for (int k = 0; k < I; ++k)
            if (e0 + k < valid) pw[e0 + k] = W[k];
    }
    if (O_out) {
        IdxT O[I];
        const int64_t slot0 = f * strat.n;  // stratum slot of this filter's first stratum (nf > 1: n == strat.n)
        StrataWindow win{swin[warp], -1, kStrataWindow};
        if (!strat.uni.col && !strat.guide) {
            // the warp's first query is at stratum floor(n * W_excl(lane 0)) + 1 or later
            const double w_first = __shfl_sync(0xffffffffu, base_w, 0);
            const double xf = w_first * (double)strat.n;
            const int64_t jf = xf <= 0.0 ? 0 : (xf >= (double)strat.n ? strat.n : (int64_t)xf);
            win.base = ((slot0 + jf + strat.uni.offset) >> 2) << 2;
#pragma unroll
            for (int h = 0; h < kStrataWindow / 128; ++h) {
                const uint4 b = strata_block(strat.uni.seed, strat.uni.stream, (uint64_t)(win.base >> 2) + h * 32 + lane);
                *reinterpret_cast<uint4 *>(&swin[warp][(h * 32 + lane) * 4]) = b;
            }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const int e = e0 + k;
            if (e >= valid) O[k] = (IdxT)0;
            else if (strat.guide) O[k] = guide_count<IdxT>(W[k], strat.guide);
            else O[k] = strat_count<IdxT>(strat, slot0, W[k], win);
            // the last particle closes the cumulative count at n (the reference's clamp at order[n], App. C)
            if (global_base + start + e == strat.n - 1) O[k] = (IdxT)(strat.guide ? strat.guide : strat.n);
            if (oend_pin && e < valid) {  // a shard's counts are pinned between the agreed closing counts
                const IdxT lo_pin = shard_rank > 0 ? (IdxT)oend_pin[shard_rank - 1] : (IdxT)0, hi_pin = (IdxT)oend_pin[shard_rank];
                O[k] = start + e == n - 1 ? hi_pin : min(max(O[k], lo_pin), hi_pin);
            }
            if (tile_last_O && e == valid - 1) tile_last_O[f * tpf + tile] = O[k];
        }
        IdxT *po = O_out + f * n + start;
        if (sizeof(IdxT) == 4 && ((reinterpret_cast<uintptr_t>(po) & 15) == 0) && e0 + 3 < valid) {
            *reinterpret_cast<int4 *>(po + e0) = make_int4((int)O[0], (int)O[1], (int)O[2], (int)O[3]);
        } else {
#pragma unroll
            for (int k = 0; k < I; ++k)
                if (e0 + k < valid) po[e0 + k] = O[k];
        }
    }
}

// ------------------------------------------------------------------ K3, hot path
// The scan of the README step (and of every stratified resample with library-drawn strata over a power-of-two
// population -- every BASELINE size): no W output, no guide table, Philox stratum uniforms.  Same summation order
// and the SAME count C(W_k) = #{i : u_i <= W_k} as k_scan (the parity tests compare the two through the
// column-uniform path), organised for instruction issue -- the kernel has no function call and no slow path:
//   * ONE stratum window per TILE: the tile's strata are [floor(n*off), floor(n*(off+total))], ~2048 for balanced
//     weights, so each thread draws ~1.1 Philox blocks (4 strata each); a query outside the window (a tile holding
//     a particle with thousands of offspring) draws its block in place;
//   * for n = 2^p the stratum lower bounds j/n are exact, so the count is evaluated literally and branch-free:
//     j = floor(n*W) corrected by at most one so that j/n <= W < (j+1)/n, then
//     C(W) = j + [fl(fl(r_{j+1} * (1/n)) + j/n) <= W]   (u_i of resample.jl:162 for the only stratum that can straddle W);
//     the uniform-weight fallback of safe_softmax (W_k = k/n) and equal weights need no special case;
//   * block-uniform conditions (validity kind, closing particle, alignment) are hoisted out of the particle loop.
constexpr int kHotWindow = 2560;  // stratum words staged per tile (2048 + 25 % slack): 10 KB
#ifndef GENPF_HOT_THREADS
#define GENPF_HOT_THREADS 256
#endif
#ifndef GENPF_HOT_MINB
#define GENPF_HOT_MINB 6  // 40 registers, no spills: 68 -> 62 us at 2^24 (4 -> 60 registers, 7 / 8 -> 32 registers + spills: 63.5 us)
#endif
constexpr int kHotThreads = GENPF_HOT_THREADS;  // 256 threads x 8 particles: per-thread overheads amortised over 8
template <typename IdxT, bool USE_E, int T = kHotThreads>
static __global__ void __launch_bounds__(T, 2048 / T >= 8 ? GENPF_HOT_MINB : 2048 / T)
    k_scan_hot(LwSrc src, int64_t n, int64_t tpf, const Stats *stats, const double *tile_off, IdxT *O_out,
               IdxT *tile_last_O, StratArgs strat, int gate, const double *shard_info, int64_t global_base,
               const double *chunk_info, const double *ew, const double *tile_scale,
               const long long *oend_pin = nullptr, int shard_rank = 0) {
    // oend_pin (multi-GPU particle sharding): the closing counts of all shards as every rank derived them from the
    // exchanged totals (xchg_stats_combine); this shard's counts are pinned to [oend_pin[rank-1], oend_pin[rank]]
    pdl_enter();
    constexpr int I = kTile / T, NW = T / 32;
    __shared__ double sm[NW + 1];
    __shared__ alignas(16) uint32_t swin[kHotWindow];
    const int64_t f = blockIdx.y, tile = blockIdx.x;
    const int kind = stats[f].invalid_kind;
    if (kind == 1 || kind == 4) return;
    if (gate && !stats[f].do_resample) return;
    const int64_t start = tile * kTile;
    const int valid = (int)min((int64_t)kTile, n - start);
    const int e0 = I * (int)threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[I];
    {
        const double *p = (USE_E ? ew : src.p) + f * n + start;
        if (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && e0 + I - 1 < valid) {
#pragma unroll
            for (int k = 0; k < I; k += 2) {
                const double2 a = __ldg(reinterpret_cast<const double2 *>(p + e0 + k));
                v[k] = a.x;
                v[k + 1] = a.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < I; ++k) v[k] = e0 + k < valid ? __ldg(p + e0 + k) : (USE_E ? 0.0 : -INFINITY);
        }
    }
    const double step = strat.step;  // 1/n, exact (n is a power of two)
    double W[I];
    if (kind != 0) {  // uniform fallback of safe_softmax (utils.jl:123-133): block-uniform, rare
        double run = 0.0;
#pragma unroll
        for (int k = 0; k < I; ++k) {
            run += e0 + k < valid ? step : 0.0;
            W[k] = run;
        }
    } else if (USE_E) {
        double escale = tile_scale[f * tpf + tile];
        if (chunk_info) escale *= chunk_info[2 * (f * ((tpf + kChunkTiles - 1) >> kChunkLog2) + (tile >> kChunkLog2)) + 1];
        if (shard_info) escale *= shard_info[1];
        double run = 0.0;
#pragma unroll
        for (int k = 0; k < I; ++k) {
            run += v[k] * escale;
            W[k] = run;
        }
    } else {
        const double M = stats[f].M, inv_S = 1.0 / stats[f].S;
        double run = 0.0;
#pragma unroll
        for (int k = 0; k < I; ++k) {
            run += exp_nonpos(src.fix(v[k]) - M) * inv_S;
            W[k] = run;
        }
    }
    double inc = W[I - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    double lane_excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) lane_excl = 0.0;
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 16 warp totals (+ the tile total in sm[NW])
        double t = lane < NW ? sm[lane] : 0.0;
        double s = t;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        const double ex = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane < NW) sm[lane] = lane == 0 ? 0.0 : ex;
        if (lane == NW - 1) sm[NW] = s;
    }
    __syncthreads();
    double off = tile_off[f * tpf + tile];
    if (chunk_info) {
        const int64_t nchunks = (tpf + kChunkTiles - 1) >> kChunkLog2;
        const double *ci = chunk_info + 2 * (f * nchunks + (tile >> kChunkLog2));
        off = ci[0] + ci[1] * off;
    }
    if (shard_info) off = shard_info[0] + shard_info[1] * off;
    const double base_w = off + (sm[warp] + lane_excl);
    // ---- the tile's stratum window: 0-based strata [jw, jw + 4*ncalls) of this filter
    const double nd = (double)strat.n;
    const IdxT nI = (IdxT)strat.n;
    const int64_t goff = f * strat.n + strat.uni.offset;  // global Philox slot of this filter's stratum 0
    IdxT jw;
    int ncalls;
    {
        const double x0 = off * nd, x1 = (off + sm[NW]) * nd;
        const IdxT jb = x0 <= 0.0 ? (IdxT)0 : (x0 >= nd ? nI : (IdxT)x0);
        const IdxT je = x1 <= 0.0 ? (IdxT)0 : (x1 >= nd ? nI : (IdxT)x1);
        const int64_t wb = ((goff + (int64_t)jb) >> 2) << 2;  // window base slot (a Philox block boundary)
        jw = (IdxT)(wb - goff);
        const int64_t want = (int64_t)je - (int64_t)jw + 3;  // stratum je is queried, +-1 for the floor correction
        ncalls = (int)min((int64_t)(kHotWindow / 4), (want + 3) >> 2);
        for (int c = threadIdx.x; c < ncalls; c += T) {
            const uint4 b = philox_at(strat.uni.seed, strat.uni.stream, (uint64_t)(wb >> 2) + c);
            *reinterpret_cast<uint4 *>(&swin[4 * c]) = b;
        }
    }
    __syncthreads();
    const IdxT wlen = (IdxT)(4 * ncalls);
    IdxT O[I];
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const double Wk = base_w + W[k];
        // n is a power of two on this path, so n*W is an exact scaling: j = floor(n*W) and lo = j/n are exact and
        // j/n <= W < (j+1)/n holds with no correction step (W >= 0; W marginally above 1 lands on j >= n below)
        const double x = Wk * nd;
        const IdxT j = (IdxT)x;
        const double lo = (double)j * step;
        // stratum j+1 (1-based) is the only one that can straddle W: strata <= j lie below, strata >= j+2 above
        const IdxT rel = j - jw;
        uint32_t word;
        if (rel >= 0 && rel < wlen) {
            word = swin[(int)rel];
        } else {  // outside the staged window (a tile spanning thousands of strata): draw the block in place
            const int64_t gs = goff + (int64_t)(j < nI ? j : nI - 1);
            const uint4 b = philox_at(strat.uni.seed, strat.uni.stream, (uint64_t)gs >> 2);
            const int q = (int)(gs & 3);
            word = q == 0 ? b.x : (q == 1 ? b.y : (q == 2 ? b.z : b.w));
        }
        const double r = fma((double)word, 0x1.0p-32, 0x1.0p-33);          // (word + 0.5) * 2^-32, exact
        const double u = __dadd_rn(__dmul_rn(r, step), lo);                // resample.jl:162, two roundings
        O[k] = j >= nI ? nI : j + (IdxT)(u <= Wk ? 1 : 0);
    }
    if (valid < kTile) {
#pragma unroll
        for (int k = 0; k < I; ++k)
            if (e0 + k >= valid) O[k] = (IdxT)0;
    }
    // the last particle closes the cumulative count at n (the reference's clamp at order[n], App. C); a shard's
    // counts are pinned between the closing counts every rank agreed on
    bool closes = (global_base + start + valid == strat.n);
    IdxT close_at = nI;
    if (oend_pin) {
        const IdxT lo_pin = shard_rank > 0 ? (IdxT)oend_pin[shard_rank - 1] : (IdxT)0;
        close_at = (IdxT)oend_pin[shard_rank];
#pragma unroll
        for (int k = 0; k < I; ++k) O[k] = min(max(O[k], lo_pin), close_at);
        closes = (start + valid == n);
    }
#pragma unroll
    for (int k = 0; k < I; ++k) {
        if (e0 + k == valid - 1) {
            if (closes) O[k] = close_at;
            tile_last_O[f * tpf + tile] = O[k];
        }
    }
    IdxT *po = O_out + f * n + start;
    if (sizeof(IdxT) == 4 && ((reinterpret_cast<uintptr_t>(po) & 15) == 0) && e0 + I - 1 < valid) {
#pragma unroll
        for (int k = 0; k < I; k += 4)
            *reinterpret_cast<int4 *>(po + e0 + k) = make_int4((int)O[k], (int)O[k + 1], (int)O[k + 2], (int)O[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < I; ++k)
            if (e0 + k < valid) po[e0 + k] = O[k];
    }
}

#endif  // GENPF_PLUGIN_BUILD
// ------------------------------------------------------------------ searches
// smallest k in [0, n) with a[k] > t, clamped to n-1 (the reference would throw BoundsError, App. C)
template <typename T, typename Q>
__device__ __forceinline__ int64_t upper_bound_clamped(const T *a, int64_t n, Q t) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] > t) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// K4/K6 expand: parent_i = min{k : O_k > i} for output slot i (0-based), output centric so a particle
// owning millions of offspring costs nothing extra.  Per block of 2048 outputs:
//   1. two coarse searches over the per-tile closing counts (tile_last_O) bracket the source tiles,
//   2. those tiles' O values are staged in shared memory (<= 3 tiles; else per-thread global search),
//   3. every source with offspring in the block drops its index at its first output slot (scatter of
//      run heads), and a block-wide max-scan fills the runs -- no per-output binary search.
template <typename IdxT>
struct ExpandSmem {
    static constexpr int kStageCap = (sizeof(IdxT) == 4 ? 3 : 2) * kTile;  // stays under 48 KB static smem
    alignas(16) int32_t sParent[kTile];  // accessed as int4
    int64_t brk[2];
    int32_t warp_max[32];
    IdxT sO[kStageCap + 1];
};

// first tile b in [0, tpf) with tile_last[b] > target (clamped to tpf-1)
template <typename IdxT>
__device__ __forceinline__ int64_t coarse_search(const IdxT *tile_last, int64_t tpf, int64_t target, int64_t guess) {
    // gallop outwards from the guess: resampled populations keep parent ~ output index
    int64_t lo = 0, hi = tpf - 1;
    if (guess > hi) guess = hi;
    if ((int64_t)tile_last[guess] > target) {
        hi = guess;
        int64_t step = 1;
        while (hi - step >= 0 && (int64_t)tile_last[hi - step] > target) {
            hi -= step;
            step <<= 1;
        }
        lo = max((int64_t)0, hi - step + 1);
    } else {
        lo = guess + 1;
        if (lo > hi) return hi;
        int64_t step = 1;
        while (lo + step <= hi && (int64_t)tile_last[lo + step - 1] <= target) {
            lo += step;
            step <<= 1;
        }
        hi = min(hi, lo + step - 1);
    }
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if ((int64_t)tile_last[mid] > target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// On return rel[k] + (returned s0) = source index (local to the filter) of this thread's striped output
// slot k (block of T threads, kTile/T slots each).
template <typename IdxT, int T = kThreads>
__device__ __forceinline__ int64_t block_expand(const IdxT *Of, const IdxT *tile_last, int64_t n_src, int64_t tpf_src,
                                                int64_t i0, int64_t valid, ExpandSmem<IdxT> &sm,
                                                int32_t (&rel)[kTile / T], IdxT O_base = 0, int64_t guess = -1) {
    constexpr int I = kTile / T;
    if (threadIdx.x < 2) {
        const int64_t target = threadIdx.x == 0 ? i0 : i0 + valid - 1;
        sm.brk[threadIdx.x] = coarse_search<IdxT>(tile_last, tpf_src, target, guess >= 0 ? guess : i0 / kTile);
    }
    __syncthreads();
    const int64_t b_lo = sm.brk[0], b_hi = sm.brk[1];
    const int64_t s0 = b_lo * kTile, s1 = min((b_hi + 1) * (int64_t)kTile, n_src);
    const int64_t len = s1 - s0;
    if (len <= ExpandSmem<IdxT>::kStageCap) {
        if (threadIdx.x == 0) sm.sO[0] = b_lo > 0 ? tile_last[b_lo - 1] : O_base;
        for (int j = threadIdx.x; j < (int)len; j += T) sm.sO[1 + j] = Of[s0 + j];
#pragma unroll
        for (int k = 0; k < I; ++k) sm.sParent[k * T + threadIdx.x] = 0;
        __syncthreads();
        const IdxT ibeg = (IdxT)i0, iend = (IdxT)(i0 + valid);
        for (int j = threadIdx.x; j < (int)len; j += T) {
            const IdxT prev = sm.sO[j], cur = sm.sO[j + 1];
            const IdxT pos = prev > ibeg ? prev : ibeg, end = cur < iend ? cur : iend;
            if (pos < end) sm.sParent[(int)(pos - ibeg)] = j;
        }
        __syncthreads();
        // block-wide inclusive max-scan over sParent (blocked I per thread, then warp + cross-warp)
        int32_t a[I];
        if (I == 8) {
            const int4 q0 = reinterpret_cast<const int4 *>(sm.sParent)[2 * threadIdx.x];
            const int4 q1 = reinterpret_cast<const int4 *>(sm.sParent)[2 * threadIdx.x + 1];
            a[0] = q0.x; a[1] = q0.y; a[2] = q0.z; a[3] = q0.w;
            a[4 % I] = q1.x; a[5 % I] = q1.y; a[6 % I] = q1.z; a[7 % I] = q1.w;
        } else if (I == 4) {
            const int4 q0 = reinterpret_cast<const int4 *>(sm.sParent)[threadIdx.x];
            a[0] = q0.x; a[1] = q0.y; a[2 % I] = q0.z; a[3 % I] = q0.w;
        } else {
            const int2 q0 = reinterpret_cast<const int2 *>(sm.sParent)[threadIdx.x];
            a[0] = q0.x; a[1] = q0.y;
        }
#pragma unroll
        for (int k = 1; k < I; ++k) a[k] = max(a[k], a[k - 1]);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int32_t run = a[I - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, run, o);
            if (lane >= o) run = max(run, t);
        }
        if (lane == 31) sm.warp_max[warp] = run;
        int32_t excl = __shfl_up_sync(0xffffffffu, run, 1);
        if (lane == 0) excl = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < T / 32; ++w)
            if (w < warp) excl = max(excl, sm.warp_max[w]);
#pragma unroll
        for (int k = 0; k < I; ++k) a[k] = max(a[k], excl);
        if (I == 8) {
            reinterpret_cast<int4 *>(sm.sParent)[2 * threadIdx.x] = make_int4(a[0], a[1], a[2], a[3]);
            reinterpret_cast<int4 *>(sm.sParent)[2 * threadIdx.x + 1] = make_int4(a[4 % I], a[5 % I], a[6 % I], a[7 % I]);
        } else if (I == 4) {
            reinterpret_cast<int4 *>(sm.sParent)[threadIdx.x] = make_int4(a[0], a[1], a[2 % I], a[3 % I]);
        } else {
            reinterpret_cast<int2 *>(sm.sParent)[threadIdx.x] = make_int2(a[0], a[1]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const int e = tile_elem<T>(k);
            rel[k] = e < valid ? sm.sParent[e] : 0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const int e = tile_elem<T>(k);
            rel[k] = e < valid ? (int32_t)upper_bound_clamped<IdxT, int64_t>(Of + s0, len, i0 + e) : 0;
        }
    }
    __syncthreads();
    return s0;
}

#ifndef GENPF_PLUGIN_BUILD
// update_weights! without priorities fused into the kernels that write the ancestors: full state lw .= 0.0
// (resample.jl:193-195), sub-state lw .= logsumexp(lw) - log(n_v) (resample.jl:208-210)
struct LwFill {
    double *out;      // null: the caller reweights separately (priorities, device filters)
    const Stats *st;  // statistics of the log-weights (not of the priorities / sorted keys)
    int substate;
    __device__ __forceinline__ double value(int64_t f, int64_t n_in) const {
        return substate ? st[f].lse - log((double)n_in) : 0.0;
    }
};

// (register budget left to the compiler: 44 registers; capping at 40 / 32 or letting it take 54 are all slower)
template <typename IdxT, typename OutT>
static __global__ void __launch_bounds__(kThreads)
    k_expand(const IdxT *O, const IdxT *tile_last_O, int64_t n_src, int64_t n_out, int64_t tpf_out,
             const int32_t *order, OutT *parents, int64_t out_base, const Stats *stats, int gate, int residual,
             LwFill fill = LwFill{nullptr, nullptr, 0}) {
    __shared__ ExpandSmem<IdxT> sm;
    int64_t f, tile;
    blk_to_tile(tpf_out, f, tile);
    if (stats) {
        const int kind = stats[f].invalid_kind;
        if (kind == 1 || kind == 4) return;
        if (gate && !stats[f].do_resample) return;
    }
    const int64_t tpf_src = (n_src + kTile - 1) / kTile;
    const IdxT *Of = O + f * n_src;
    const int64_t i0 = tile * kTile;
    int64_t valid = min((int64_t)kTile, n_out - i0);
    if (residual) {  // only the deterministic copies: slots [0, C), C = O[n_src-1]
        int64_t C = (int64_t)Of[n_src - 1];
        valid = min(valid, C - i0);
        if (valid <= 0) return;
    }
    int32_t rel[kItems];
    const int64_t s0 = block_expand<IdxT>(Of, tile_last_O + f * tpf_src, n_src, tpf_src, i0, valid, sm, rel);
    OutT out[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        int64_t q = s0 + rel[k];
        if (order && tile_elem(k) < valid) q = order[f * n_src + q];
        out[k] = (OutT)(q + out_base);
    }
    store_tile<OutT>(parents, f * n_out + i0, valid, out);
    if (fill.out) {
        double w[kItems];
        const double val = fill.value(f, n_src);
#pragma unroll
        for (int k = 0; k < kItems; ++k) w[k] = val;
        store_tile<double>(fill.out, f * n_out + i0, valid, w);
    }
}

// W-based stratified selection (debug / cross-check of the O path): parent_i = min{k : W_k >= u_i}
template <typename OutT>
static __global__ void __launch_bounds__(kThreads)
    k_select_stratified_w(const double *W, int64_t n, const int32_t *order, StratArgs strat, OutT *parents,
                          int64_t out_base) {
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    double u = strat_u(strat, 0, i + 1);
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (W[mid] >= u) hi = mid; else lo = mid + 1;
    }
    parents[i] = (OutT)((order ? order[lo] : lo) + out_base);
}

// K5 inverse-CDF draws for arbitrary (unsorted) uniforms: parent_j = min{k : W_k > u_j}, capped at n
// (Distributions' single-draw rule, resize.jl:284).  Slots below first_slot (residual: the
// deterministic copies) are left alone.
// Guide-table inverse CDF: parent_j = min{k : W_k > u_j} clamped to n_src-1 (rand(Categorical(w)),
// resample.jl:59,113; resize.jl:61,117).  b = floor(n*u) picks the bucket, G[b] .. G[b+1] brackets the answer
// (expected bracket length 1), and the answer is always re-checked against W: two or three scattered
// sector reads per draw instead of a ~20-probe multi-level binary search.
constexpr int kLookupItems = 4;
template <typename IdxT, typename OutT>
static __global__ void __launch_bounds__(kThreads)
    k_lookup(const double *W, const IdxT *G, int64_t B, int64_t n_src, int64_t n_out, UniSrc uni,
             const IdxT *first_slot_O, OutT *parents, int64_t out_base, const Stats *stats, int gate,
             LwFill fill = LwFill{nullptr, nullptr, 0}) {
    constexpr int K = kLookupItems;
    const int64_t f = blockIdx.y;
    if (stats) {
        const int kind = stats[f].invalid_kind;
        if (kind == 1 || kind == 4) return;
        if (gate && !stats[f].do_resample) return;
    }
    const double *Wf = W + f * n_src;
    const IdxT *Gf = G + f * B;
    const int64_t first = first_slot_O ? (int64_t)first_slot_O[f * n_src + n_src - 1] : 0;
    const int64_t j0 = (int64_t)blockIdx.x * (kThreads * K);
    const double nd = (double)B;
    double u[K], wk[K];
    int64_t lo[K], hi[K];
    bool live[K], exact[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t j = j0 + k * kThreads + threadIdx.x;
        live[k] = j < n_out && j >= first;
        u[k] = live[k] ? uni(f * n_out + j) : 0.0;
        const double x = u[k] * nd;
        int64_t b = x >= nd ? B - 1 : (x <= 0.0 ? 0 : (int64_t)x);
        // the bracket's lower end can only be wrong when n*u rounded onto the bucket boundary (or was clamped)
        exact[k] = !(x > (double)b && x < nd);
        lo[k] = b;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t b = lo[k];
        lo[k] = (int64_t)__ldg(Gf + b);
        hi[k] = b + 1 < B ? (int64_t)__ldg(Gf + b + 1) : n_src - 1;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = __ldg(Wf + lo[k]);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (!live[k]) continue;
        int64_t kk = lo[k];
        double w = wk[k];
        bool below_known = false;  // W[kk-1] <= u established by the search itself
        if (hi[k] - kk > 8) {
            int64_t a = kk, c = hi[k];
            while (a < c) {
                const int64_t mid = a + ((c - a) >> 1);
                if (__ldg(Wf + mid) > u[k]) c = mid; else a = mid + 1;
            }
            below_known = a > kk;
            kk = a;
            w = __ldg(Wf + kk);
        }
        while (w <= u[k] && kk < n_src - 1) {
            ++kk;
            w = __ldg(Wf + kk);
            below_known = true;
        }
        if (exact[k] && !below_known)
            while (kk > 0 && __ldg(Wf + kk - 1) > u[k]) --kk;
        const int64_t j = j0 + k * kThreads + threadIdx.x;
        parents[f * n_out + j] = (OutT)(kk + out_base);
        if (fill.out) fill.out[f * n_out + j] = fill.value(f, n_src);
    }
}

// ------------------------------------------------------------------ K5, one sector per draw
// The inverse-CDF draws are a random gather over a table far larger than L2 (512 MB of W at 2^26), so a draw costs
// what its scattered 32-byte sectors cost.  The guide table is therefore stored as one 32-BYTE RECORD per bucket,
//   {g = G[b], span = G[b+1] - G[b], W[g], W[g+1], W[g+2]},
// written by the pass that builds G (its reads of W are monotone, hence streaming): a lookup is ONE sector with no
// dependent second load; only a draw with three or more cumulative weights of its bucket below it (rare: with
// B = n/2 buckets a bucket holds two on average, and heavy particles span many buckets) continues in W itself,
// by bisection inside [g+3, g+span].  Same answers as k_lookup: min{k : W_k > u} clamped to n_src-1.
struct alignas(32) GuideRec {
    int32_t g, span;  // span < 0: the next bucket's start is unknown (last bucket of a build tile)
    double w[3];      // +Inf beyond n_src-1
};
static __global__ void __launch_bounds__(kThreads)
    k_guide_records(const int32_t *O, const int32_t *tile_last_O, const double *W, int64_t n_src, int64_t B,
                    int64_t tpf_b, GuideRec *rec, const Stats *stats, int gate) {
    __shared__ ExpandSmem<int32_t> sm;
    const int64_t f = blockIdx.y, tile = blockIdx.x;
    if (stats) {
        const int kind = stats[f].invalid_kind;
        if (kind == 1 || kind == 4) return;
        if (gate && !stats[f].do_resample) return;
    }
    const int64_t tpf_src = (n_src + kTile - 1) / kTile;
    const int64_t i0 = tile * kTile;
    const int valid = (int)min((int64_t)kTile, B - i0);
    int32_t rel[kItems];
    const int64_t s0 = block_expand<int32_t>(O + f * n_src, tile_last_O + f * tpf_src, n_src, tpf_src, i0, valid, sm, rel);
    // span = the next bucket's start - g: the even slot's neighbour is the thread's own odd slot, the odd slot's is
    // the next lane's even slot (unknown, -1, for lane 31: one bucket in 64 bisects up to n_src-1 instead)
    const double *Wf = W + f * n_src;
    GuideRec *out = rec + f * B + i0;
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    int32_t nxt[kItems];
#pragma unroll
    for (int k = 0; k < kItems; k += 2) {
        nxt[k] = rel[k + 1];
        nxt[k + 1] = __shfl_down_sync(0xffffffffu, rel[k], 1);
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int e = tile_elem(k);
        if (e >= valid) continue;
        const int64_t g = s0 + rel[k];
        int32_t span = -1;
        if (e + 1 < valid && ((k & 1) == 0 || lane < 31)) span = nxt[k] - rel[k];
        GuideRec r;
        r.g = (int32_t)g;
        r.span = span;
        r.w[0] = __ldg(Wf + g);
        r.w[1] = g + 1 < n_src ? __ldg(Wf + g + 1) : kInf;
        r.w[2] = g + 2 < n_src ? __ldg(Wf + g + 2) : kInf;
        int4 *q = reinterpret_cast<int4 *>(out + e);
        q[0] = make_int4(r.g, r.span, __double2loint(r.w[0]), __double2hiint(r.w[0]));
        q[1] = make_int4(__double2loint(r.w[1]), __double2hiint(r.w[1]), __double2loint(r.w[2]), __double2hiint(r.w[2]));
    }
}

#ifndef GENPF_LOOKUP_MINB
#define GENPF_LOOKUP_MINB 5  // 48 registers: more gathers in flight (2.27 -> 2.11 ms at 2^26; 6 and 8 blocks spill too much)
#endif
template <typename OutT>
static __global__ void __launch_bounds__(kThreads, GENPF_LOOKUP_MINB)
    k_lookup_rec(const double *W, const GuideRec *R, int64_t B, int64_t n_src, int64_t n_out, UniSrc uni,
                 const int32_t *first_slot_O, OutT *parents, int64_t out_base, const Stats *stats, int gate,
                 LwFill fill = LwFill{nullptr, nullptr, 0}) {
    constexpr int K = kLookupItems;
    const int64_t f = blockIdx.y;
    if (stats) {
        const int kind = stats[f].invalid_kind;
        if (kind == 1 || kind == 4) return;
        if (gate && !stats[f].do_resample) return;
    }
    const double *Wf = W + f * n_src;
    const int4 *Rf = reinterpret_cast<const int4 *>(R + f * B);
    const int64_t first = first_slot_O ? (int64_t)first_slot_O[f * n_src + n_src - 1] : 0;
    const int64_t j0 = (int64_t)blockIdx.x * (kThreads * K);
    if (j0 + kThreads * K <= first) return;  // residual: a block of deterministic copies only
    const double nd = (double)B;
    double u[K];
    int64_t bk[K];
    bool live[K], exact[K];
    int4 qa[K], qb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t j = j0 + k * kThreads + threadIdx.x;
        live[k] = j < n_out && j >= first;
        u[k] = live[k] ? uni(f * n_out + j) : 0.0;
        const double x = u[k] * nd;
        const int64_t b = x >= nd ? B - 1 : (x <= 0.0 ? 0 : (int64_t)x);
        exact[k] = !(x > (double)b && x < nd);  // n*u rounded onto the bucket boundary (or was clamped)
        bk[k] = b;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        qa[k] = __ldg(Rf + 2 * bk[k]);
        qb[k] = __ldg(Rf + 2 * bk[k] + 1);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (!live[k]) continue;
        const int64_t g = qa[k].x;
        const double w0 = __hiloint2double(qa[k].w, qa[k].z), w1 = __hiloint2double(qb[k].y, qb[k].x),
                     w2 = __hiloint2double(qb[k].w, qb[k].z);
        const double uk = u[k];
        int64_t kk;
        bool below_known = true;  // W[kk-1] <= u established by the comparisons themselves
        if (w0 > uk) {
            kk = g;
            below_known = false;
        } else if (w1 > uk) {
            kk = g + 1;
        } else if (w2 > uk) {
            kk = g + 2;
        } else {
            int64_t a = min(g + 3, n_src - 1), c = qa[k].y >= 0 ? min(g + (int64_t)qa[k].y, n_src - 1) : n_src - 1;
            if (c < a) c = a;
            while (a < c) {
                const int64_t mid = a + ((c - a) >> 1);
                if (__ldg(Wf + mid) > uk) c = mid; else a = mid + 1;
            }
            kk = a;
            double w = __ldg(Wf + kk);
            while (w <= uk && kk < n_src - 1) {  // the table only has to be monotone: the answer is checked against W
                ++kk;
                w = __ldg(Wf + kk);
            }
        }
        if (kk > n_src - 1) kk = n_src - 1;
        if (exact[k] && !below_known)
            while (kk > 0 && __ldg(Wf + kk - 1) > uk) --kk;
        const int64_t j = j0 + k * kThreads + threadIdx.x;
        parents[f * n_out + j] = (OutT)(kk + out_base);
        if (fill.out) fill.out[f * n_out + j] = fill.value(f, n_src);
    }
}

// ------------------------------------------------------------------ K6 residual
// c_i = floor(n_out * w_i) literally (resample.jl:99, resize.jl:103); r_i = n_out*w_i - floor(n_out*w_i).
struct ResidPartials {
    long long *c;  // per tile sum of copies
    double *r;     // per tile sum of residual weights
};
__device__ __forceinline__ void resid_terms(const Stats &st, double v, bool in_range, double n_out_d, double inv_n,
                                            long long &c, double &r) {
    const bool uniform = (st.invalid_kind == 2 || st.invalid_kind == 3);
    double w = uniform ? (in_range ? inv_n : 0.0) : exp(v - st.M) / st.S;
    double nw = n_out_d * w;
    double fl = floor(nw);
    c = (long long)fl;
    r = nw - fl;
}
static __global__ void __launch_bounds__(kThreads)
    k_resid_partials(LwSrc src, int64_t n, int64_t n_out, int64_t tpf, const Stats *stats, ResidPartials out) {
    __shared__ double sm[kWarps];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const Stats st = stats[f];
    if (st.invalid_kind == 1 || st.invalid_kind == 4) return;
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    double v[kItems];
    load_tile(src, f * n + start, valid, v, -INFINITY);
    long long cs = 0;
    double rs = 0.0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        long long c;
        double r;
        resid_terms(st, v[k], tile_elem(k) < valid, (double)n_out, 1.0 / (double)n, c, r);
        cs += c;
        rs += r;
    }
    rs = block_sum(rs, sm);
    double csd = block_sum((double)cs, sm);  // exact: counts < 2^53
    if (threadIdx.x == 0) {
        out.c[blockIdx.y * gridDim.x + blockIdx.x] = (long long)csd;
        out.r[blockIdx.y * gridDim.x + blockIdx.x] = rs;
    }
}
// one block per filter: total residual mass, exclusive tile offsets for counts and normalised residuals.
// 1024 threads x 8 consecutive tiles per round, running carry across rounds.
static __global__ void __launch_bounds__(1024)
    k_resid_finalize(ResidPartials in, int64_t tpf, double *r_total, long long *c_off, double *r_off) {
    constexpr int T = 1024, C = 8, NW = T / 32;
    __shared__ double smd[NW];
    __shared__ long long sml[NW];
    __shared__ double carry_r;
    __shared__ long long carry_c;
    const int64_t f = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double rs = 0.0;
    for (int64_t b = threadIdx.x; b < tpf; b += T) rs += in.r[f * tpf + b];
    rs = warp_sum(rs);
    if (lane == 0) smd[warp] = rs;
    __syncthreads();
    double R = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) R += smd[w];
    if (threadIdx.x == 0) {
        r_total[f] = R;
        carry_r = 0.0;
        carry_c = 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < tpf; base += (int64_t)T * C) {
        double rv[C], rrun = 0.0;
        long long cv[C], crun = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int64_t b = base + (int64_t)threadIdx.x * C + c;
            rv[c] = rrun;
            cv[c] = crun;
            if (b < tpf) {
                rrun += (R > 0.0) ? in.r[f * tpf + b] / R : 0.0;
                crun += in.c[f * tpf + b];
            }
        }
        double rinc = rrun;
        long long cinc = crun;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double ur = __shfl_up_sync(0xffffffffu, rinc, o);
            long long uc = __shfl_up_sync(0xffffffffu, cinc, o);
            if (lane >= o) {
                rinc += ur;
                cinc += uc;
            }
        }
        double rex = __shfl_up_sync(0xffffffffu, rinc, 1);
        long long cex = __shfl_up_sync(0xffffffffu, cinc, 1);
        if (lane == 0) {
            rex = 0.0;
            cex = 0;
        }
        __syncthreads();
        if (lane == 31) {
            smd[warp] = rinc;
            sml[warp] = cinc;
        }
        __syncthreads();
        double rw = 0.0, rtot = 0.0;
        long long cw = 0, ctot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (w < warp) {
                rw += smd[w];
                cw += sml[w];
            }
            rtot += smd[w];
            ctot += sml[w];
        }
        const double cr = carry_r;
        const long long cc = carry_c;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int64_t b = base + (int64_t)threadIdx.x * C + c;
            if (b < tpf) {
                r_off[f * tpf + b] = cr + ((rw + rex) + rv[c]);
                c_off[f * tpf + b] = cc + cw + cex + cv[c];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry_r = cr + rtot;
            carry_c = cc + ctot;
        }
        __syncthreads();
    }
}
template <typename IdxT>
static __global__ void __launch_bounds__(kThreads)
    k_resid_scan(LwSrc src, int64_t n, int64_t n_out, int64_t tpf, const Stats *stats, const double *r_total,
                 const long long *c_off, const double *r_off, IdxT *O_out, IdxT *tile_last_O, WTables rt,
                 IdxT *GO_out, IdxT *G_tile_last, int64_t B) {
    __shared__ double sm[32];
    __shared__ long long smi[32];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const Stats st = stats[f];
    if (st.invalid_kind == 1 || st.invalid_kind == 4) return;
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    double v[kItems], r[kItems], R[kItems];
    long long c[kItems], C[kItems];
    load_tile(src, f * n + start, valid, v, -INFINITY);
    const double Rt = r_total[f];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        resid_terms(st, v[k], tile_elem(k) < valid, (double)n_out, 1.0 / (double)n, c[k], r[k]);
        r[k] = Rt > 0.0 ? r[k] / Rt : 0.0;  // r_weights / sum(r_weights), resample.jl:110
    }
    tile_scan<long long>(c, C, smi);
    tile_scan<double>(r, R, sm);
    const long long co = c_off[f * tpf + tile];
    const double ro = r_off[f * tpf + tile];
    IdxT O[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        long long cc = co + C[k];
        if (cc > n_out) cc = n_out;  // clamp (App. C)
        O[k] = (IdxT)cc;
        R[k] = ro + R[k];
        if (tile_elem(k) == valid - 1) tile_last_O[f * tpf + tile] = O[k];
    }
    store_tile<IdxT>(O_out, f * n + start, valid, O);
    store_tile<double, kThreads>(rt.W, f * n + start, valid, R);
    // guide-table counts of the residual CDF (k_lookup)
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int e = tile_elem(k);
        O[k] = e < valid ? guide_count<IdxT>(R[k], B) : (IdxT)0;
        if (start + e == n - 1) O[k] = (IdxT)B;
        if (e == valid - 1) G_tile_last[f * tpf + tile] = O[k];
    }
    store_tile<IdxT>(GO_out, f * n + start, valid, O);
}

// ------------------------------------------------------------------ K9 reweight after resample
// update_weights! without priorities is LwFill above (written by k_expand / k_lookup).
// with priorities: d_j = lw[parent_j] - lp[parent_j] (resample.jl:197,212)
template <typename InT>
static __global__ void k_prio_ratio(const double *lw, LwSrc lp, const InT *parents, int64_t in_base, int64_t n_in,
                             int64_t n_out, double *d_out) {
    int64_t f = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = (int64_t)parents[f * n_out + j] - in_base;
        d_out[f * n_out + j] = lw[f * n_in + p] - lp.fix(lp.p[f * n_in + p]);
    }
}
// lw_j = d_j + (log(n_out) - lse(d))  full (resample.jl:201, resize.jl:436)
//      = d_j + (lse(lw_old) - lse(d)) sub-state (resample.jl:215-216)
static __global__ void k_prio_shift(double *d, int64_t n_out, const Stats *st_d, const Stats *st_lw, int substate) {
    int64_t f = blockIdx.y;
    const double shift = substate ? (st_lw[f].lse - st_d[f].lse) : (log((double)n_out) - st_d[f].lse);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x)
        d[f * n_out + j] = d[f * n_out + j] + shift;
}

// get_log_norm_weights (utils.jl:148) / get_norm_weights (utils.jl:156)
static __global__ void k_normalize_out(const double *lw, int64_t n, const Stats *stats, double *log_norm, double *norm) {
    const Stats st = stats[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = lw[i];
        if (log_norm) log_norm[i] = v - st.lse;
        if (norm) norm[i] = exp(v - st.M) / st.S;
    }
}

// ------------------------------------------------------------------ K12 weighted mean / variance
// mean = sum(softmax(lw) .* x); var = sum(softmax(lw) .* (x .- mean).^2)  -- two pass, like statistics.jl:13-17,48-54
struct XSrc {
    const double *d;   // fp64 column, or
    const uint8_t *b;  // Bool column promoted to fp64 (README.md:97)
    __device__ __forceinline__ double operator()(int64_t i) const { return d ? d[i] : (double)b[i]; }
};
static __global__ void __launch_bounds__(kThreads)
    k_weighted_moment(LwSrc lw, XSrc x, int64_t n, int64_t tpf, const Stats *stats, const double *center,
                      double *partial) {
    __shared__ double sm[kWarps];
    int64_t f, tile;
    blk_to_tile(tpf, f, tile);
    const Stats st = stats[f];
    const int64_t start = tile * kTile;
    const int64_t valid = min((int64_t)kTile, n - start);
    double v[kItems];
    load_tile(lw, f * n + start, valid, v, -INFINITY);
    const double c = center ? center[f] : 0.0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        int e = tile_elem(k);
        if (e < valid) {
            double xv = x(f * n + start + e);
            double wv = exp(v[k] - st.M) / st.S;
            double t = center ? (xv - c) * (xv - c) : xv;
            acc += wv * t;
        }
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
static __global__ void __launch_bounds__(kThreads) k_sum_partials(const double *partial, int64_t tpf, double *out) {
    __shared__ double sm[kWarps];
    const int64_t f = blockIdx.x;
    double s = 0.0;
    for (int64_t b = threadIdx.x; b < tpf; b += kThreads) s += partial[f * tpf + b];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[f] = s;
}

// ------------------------------------------------------------------ K13 replicate / dereplicate
// pf_replicate! resize.jl:236-244: repeat(x; inner=k) | repeat(x, k); weights copied unchanged
template <typename OutT>
static __global__ void k_replicate(const double *lw, int64_t n, int64_t k, int interleaved, OutT *parents, int64_t out_base,
                            double *lw_out) {
    const int64_t f = blockIdx.y;  // batches: every filter is replicated on its own (parents local to the filter)
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n * k; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t src = interleaved ? (j % n) : (j / k);
        if (parents) parents[f * n * k + j] = (OutT)(src + out_base);
        if (lw_out) lw_out[f * n * k + j] = lw[f * n + src];
    }
}
// pf_dereplicate! resize.jl:267-297.  One thread per retained particle; the block of k replicas is walked
// sequentially exactly like the reference (softmax, then single inverse-CDF draw `cp <= u`, lse - log k).
template <typename OutT>
static __global__ void k_dereplicate(const double *lw, int64_t n, int64_t k, int interleaved, int sample, UniSrc uni,
                              OutT *parents, int64_t out_base, double *lw_out) {
    const int64_t n_new = n / k;
    const int64_t f = blockIdx.y;  // batches: blocks of replicas never straddle two filters
    lw += f * n;
    lw_out += f * n_new;
    if (parents) parents += f * n_new;
    const int64_t uoff = f * n_new;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_new; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t first = interleaved ? b : b * k;
        const int64_t stride = interleaved ? n_new : 1;
        if (!sample) {
            if (parents) parents[b] = (OutT)(first + out_base);
            lw_out[b] = lw[first];
            continue;
        }
        double m = -INFINITY;
        for (int64_t j = 0; j < k; ++j) m = fmax(m, lw[first + j * stride]);
        double s = 0.0;
        for (int64_t j = 0; j < k; ++j) s += exp(lw[first + j * stride] - m);
        const double u = uni(uoff + b);
        int64_t i = 0;
        double cp = exp(lw[first] - m) / s;
        while (cp <= u && i < k - 1) {
            i += 1;
            cp += exp(lw[first + i * stride] - m) / s;
        }
        if (parents) parents[b] = (OutT)(first + i * stride + out_base);
        lw_out[b] = (m == -INFINITY ? -INFINITY : m + log(s)) - log((double)k);
    }
}

static __global__ void k_materialize(LwSrc src, int64_t n, double *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = src.fix(src.p[i]);
}

static __global__ void k_uniforms(UniSrc uni, int64_t n, double *out, int strata) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = strata ? strata_word(philox_at(uni.seed, uni.stream, (uint64_t)(i + uni.offset) >> 2), i + uni.offset)
                        : uni(i);
}

template <typename A, typename B>
static __global__ void k_convert_idx(const A *in, B *out, int64_t n, int64_t add) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (B)((int64_t)in[i] + add);
}

#endif  // GENPF_PLUGIN_BUILD

}  // namespace genpf
