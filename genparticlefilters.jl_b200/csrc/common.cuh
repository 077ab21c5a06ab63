// common.cuh -- shared device primitives for libgenpf_cuda.so (sm_100a).
//
// Tile geometry: every weight-vector kernel works on tiles of kTile = 2048 particles,
// 256 threads x 8 fp64 each, "striped by 16-byte vector": thread t owns elements
// (j*256 + t)*2 + {0,1}, j = 0..3, so every global access is a fully coalesced
// 128-bit load/store (4 KB per warp-instruction group) with no shared-memory transpose.
// The reduce, scan and propagate kernels all use the SAME partition so the per-tile
// partials one kernel writes are the tile offsets the next one consumes.
#pragma once
// GENPF_PLUGIN_BUILD: this header tree is also compiled at RUN TIME by NVRTC (abi_plugin.cu) together with a user's
// model source, to instantiate the model-templated kernels for a plugin (SURVEY 8b "genpf_model_load_cubin").
// NVRTC has no host headers: the fixed-width types and INFINITY are spelled out, CUDA's math is built in.
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long int64_t;
typedef unsigned long uint64_t;
typedef unsigned long uintptr_t;
#ifndef INFINITY
#define INFINITY (__int_as_float(0x7f800000))
#endif
#else
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#endif
// kernels a plugin build instantiates need external (COMDAT) linkage so NVRTC can name them
#ifdef GENPF_PLUGIN_BUILD
#define GENPF_KERNEL __global__
#else
#define GENPF_KERNEL static __global__
#endif

namespace genpf {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kVecs = kItems / 2;
constexpr int kTile = kThreads * kItems;  // 2048 particles per tile
constexpr int kWarps = kThreads / 32;
// Filters of more than kChunkTiles tiles are finalised chunk-wise (one block per chunk, all chunks in parallel) and a
// combine step gives every chunk its {prefix, scale}; the scans compose the pair with the in-chunk tile offsets.
constexpr int kChunkLog2 = 9;
constexpr int64_t kChunkTiles = 1 << kChunkLog2;  // 512 tiles = 2^20 particles per chunk

// per-filter result of the weight reduction (device resident; utils.jl:117-140,163-164, resample.jl:178-182)
struct Stats {
    double M;    // maximum(v)
    double S;    // sum(exp.(v .- M))
    double S2;   // sum(exp.(2 .* (v .- M)))
    double lse;  // logsumexp(v)
    double ess;  // S^2 / S2 == exp(-logsumexp(2 .* lognorm(v)))
    int32_t invalid_kind;
    int32_t do_resample;  // device-side predicate of the fused README step (ess < ess_frac * n)
};

// ------------------------------------------------------------------ peer-memory exchange (replaces NCCL)
// The three per-step exchanges move 24, 8 and 0 bytes per rank: NCCL's launch + protocol latency (tens of
// microseconds each at 8 ranks) dwarfs the payload.  Every rank instead owns an Xchg block that all peers
// have mapped (CUDA IPC); a rank posts its value into slot [rank] of EVERY peer's block with plain NVLink
// stores, fences, then stores the step's epoch into the matching flag; readers spin on their local flags.
// Epochs only grow, so nothing is ever reset.  A bounded spin (about 20 s) turns a lost peer into an error
// code instead of a hang.
constexpr int kMaxPeers = 8;
struct Xchg {
    double stats[kMaxPeers][3];
    long long oend[kMaxPeers];
    unsigned long long flag_stats[kMaxPeers], flag_oend[kMaxPeers], flag_done[kMaxPeers];
    int error;
};
struct XchgPeers {
    Xchg *x[kMaxPeers];
};
__device__ __forceinline__ void xchg_wait(volatile unsigned long long *flag, unsigned long long epoch, int *error) {
    const long long t0 = clock64();
    while (*flag < epoch) {
        if (clock64() - t0 > 40000000000ll) {  // ~20 s at 2 GHz
            *error = 1;
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// what a kernel needs to take part in an exchange; world == 0: the filter is not sharded (or exchanges through NCCL)
struct XchgLink {
    XchgPeers peers;
    int world, rank;
    unsigned long long epoch;  // monotone per shard group (never reset by re-initialisation)
};

// ------------------------------------------------------------------ programmatic dependent launch
// First statement of every kernel of the step chain (host.hpp::launch_pdl): let the NEXT kernel's blocks be scheduled
// as soon as all of ours are running, then wait until the PREVIOUS kernel has completed and its writes are visible.
// Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox_at(uint64_t seed, uint64_t stream, uint64_t idx) {
    return philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}
// 53-bit uniform in [0,1): (x >> 11) * 2^-53
__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
    // x >> 11 = hi * 2^21 + (lo >> 11): two exact 32-bit conversions and one exact fma
    return fma((double)hi, 0x1.0p-32, (double)(lo >> 11) * 0x1.0p-53);
}
__host__ __device__ __forceinline__ uint64_t make_stream(uint32_t purpose, uint64_t step) {
    return ((uint64_t)purpose << 56) | (step & 0x00FFFFFFFFFFFFFFull);
}
enum : uint32_t { kPurposeResample = 1, kPurposeUpdate = 2, kPurposeMH = 3, kPurposeDerep = 4, kPurposeStrata = 5 };

// stratum uniforms generated by the library: 32-bit, FOUR strata per Philox block --
//   r_i = (word[i & 3] of Philox4x32-10(seed, stream, counter = i >> 2) + 0.5) * 2^-32
// so the four consecutive particles a scan thread owns usually share one Philox call.
__device__ __forceinline__ double strata_word(const uint4 &b, int64_t slot) {
    const int s = (int)(slot & 3);
    const uint32_t w = s == 0 ? b.x : (s == 1 ? b.y : (s == 2 ? b.z : b.w));
    return ((double)w + 0.5) * 0x1.0p-32;
}
// source of uniforms for selection: a column, or Philox(seed, stream, counter = slot)
struct UniSrc {
    const double *col;  // nullable
    uint64_t seed, stream;
    int64_t offset;  // global slot offset of local slot 0 (batches / shards)
    __device__ __forceinline__ double operator()(int64_t i) const {
        if (col) return col[i];
        uint4 o = philox_at(seed, stream, (uint64_t)(i + offset));
        return u53(o.x, o.y);
    }
};

// ------------------------------------------------------------------ fp64 constants in the constant bank
// An fp64 instruction can embed only the high word of an immediate, so every full-precision literal costs two UMOV
// (uniform-register loads) in front of its DFMA once the uniform registers run out -- 7 % of the fused step's
// instruction stream (cuobjdump: 291 UMOV per 4096 instructions).  Coefficients read from __constant__ memory
// at a compile-time offset fold into the instruction as a c[bank][offset] operand and cost nothing.
enum : int {
    kcExpLog2e = 0, kcExpLn2Hi, kcExpLn2Lo, kcExpP11, kcExpP10, kcExpP9, kcExpP8, kcExpP7, kcExpP6, kcExpP5, kcExpP4,
    kcExpP3, kcExpP2,
    kcLogA1, kcLogA2, kcLogA3, kcLogB1, kcLogB2, kcLogB3, kcLogB4, kcLn2Hi, kcLn2Lo,
    kcTwoPi, kcSinS6, kcSinS5, kcSinS4, kcSinS3, kcSinS2, kcSinS1, kcCosC6, kcCosC5, kcCosC4, kcCosC3, kcCosC2, kcCosC1,
    kcLog2Pi, kcCount
};
static __constant__ double kC[kcCount] = {
    1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    0x1.af631d0059becp-26, 0x1.28b4057f44145p-22, 0x1.71ddf5749d126p-19, 0x1.a01991ac8730ap-16, 0x1.a01a01b14378fp-13,
    0x1.6c16c187fbe02p-10, 0x1.111111110f225p-7, 0x1.555555554f0cfp-5, 0x1.555555555555ap-3, 0x1.0000000000011p-1,
    1.531383769920937332e-01, 2.222219843214978396e-01, 3.999999999940941908e-01,
    1.479819860511658591e-01, 1.818357216161805012e-01, 2.857142874366239149e-01, 6.666666666666735130e-01,
    6.93147180369123816490e-01, 1.90821492927058770002e-10,
    6.283185307179586476925, 1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02,
    1.8378770664093453,
};

// ------------------------------------------------------------------ exp for normalised weights
// exp(x) for x <= 0 (x = v - max(v); -Inf allowed).  Shifter-based range reduction k = rint(x log2 e),
// two-term Cody-Waite remainder, degree-11 near-minimax polynomial (truncation 0.15 ulp; coefficients fitted
// at Chebyshev nodes in 50-digit arithmetic), exponent add.  Results below 2^-1021 flush to 0
// (absolute error < 4.5e-308).  ~21 instructions against ~45 for the full-range libdevice exp.
__device__ __forceinline__ double exp_nonpos(double x) {
    // no clamp: for x < -708 (or -Inf) the polynomial runs on garbage and the final select discards it; a NaN
    // argument comes out as NaN (callers flag NaN weights from the NaN total)
    const double SH = 6755399441055744.0;  // 2^52 + 2^51
    const double z = fma(x, kC[kcExpLog2e], SH);
    const int k = __double2loint(z);
    const double kf = z - SH;
    double r = fma(kf, kC[kcExpLn2Hi], x);
    r = fma(kf, kC[kcExpLn2Lo], r);
    double p = kC[kcExpP11];
    p = fma(p, r, kC[kcExpP10]);
    p = fma(p, r, kC[kcExpP9]);
    p = fma(p, r, kC[kcExpP8]);
    p = fma(p, r, kC[kcExpP7]);
    p = fma(p, r, kC[kcExpP6]);
    p = fma(p, r, kC[kcExpP5]);
    p = fma(p, r, kC[kcExpP4]);
    p = fma(p, r, kC[kcExpP3]);
    p = fma(p, r, kC[kcExpP2]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return x < -708.0 ? 0.0 : res;
}

// ------------------------------------------------------------------ warp / block reductions
// All block-level helpers are templated on the block size T (threads); a tile is always kTile = 2048
// particles, so a thread owns kTile/T of them.  T = 256 (8 each) is the default; the fused step kernel
// runs T = 512 (4 each) for occupancy.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Exact fp64 maximum in the order-preserving integer domain.  fmax(double) is a 5-6 instruction sequence and
// a 64-bit shuffle is two SHFLs; REDUX reduces a 32-bit integer across the warp in ONE instruction, so the
// warp maximum is taken as: REDUX.MAX of the high words, then REDUX.MAX of the low words among the lanes that
// hold the winning high word.  NaN never wins (callers flag NaN separately; fmax semantics).
__device__ __forceinline__ long long f64_key(double v) {  // monotone double -> int64, NaN -> minimum
    long long b = __double_as_longlong(v);
    b ^= (b >> 63) & 0x7FFFFFFFFFFFFFFFll;
    return v != v ? (long long)0x8000000000000000ull : b;
}
__device__ __forceinline__ double f64_from_key(long long k) {
    k ^= (k >> 63) & 0x7FFFFFFFFFFFFFFFll;
    return __longlong_as_double(k);
}
__device__ __forceinline__ long long warp_max_key(long long k) {
    const int hi = (int)(k >> 32);
    const int hmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo = hi == hmax ? (unsigned)k : 0u;
    const unsigned lmax = __reduce_max_sync(0xffffffffu, lo);
    return ((long long)hmax << 32) | (long long)lmax;
}

// all threads get the result; smem must hold T/32 values; deterministic order
template <int T = kThreads>
__device__ __forceinline__ double block_sum(double v, double *smem) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) r += smem[w];
    return r;
}
template <int T = kThreads>
__device__ __forceinline__ double block_max(double v, double *smem) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = smem[0];
#pragma unroll
    for (int w = 1; w < T / 32; ++w) r = fmax(r, smem[w]);
    return r;
}
template <int T = kThreads>
__device__ __forceinline__ int block_or(int v, int *smem) {
    v = __reduce_or_sync(0xffffffffu, (unsigned)v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) r |= smem[w];
    return r;
}

// ------------------------------------------------------------------ tile load / store (striped by 16 B vector)
// value source: a column, optionally scaled (priority_fn = w -> alpha*w, resample.jl:51-52)
struct LwSrc {
    const double *p;
    double scale;  // 1.0 => identity
    __device__ __forceinline__ double fix(double v) const { return scale == 1.0 ? v : v * scale; }
};

// element index (within the tile) of register slot k of this thread
template <int T = kThreads>
__device__ __forceinline__ int tile_elem(int k) { return ((k >> 1) * T + threadIdx.x) * 2 + (k & 1); }

template <int T = kThreads>
__device__ __forceinline__ void load_tile(const LwSrc &src, int64_t base, int64_t valid, double (&v)[kTile / T],
                                          double fill) {
    const double *p = src.p + base;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
    const int nv = (int)valid;  // <= kTile: 32-bit tile-local arithmetic
#pragma unroll
    for (int j = 0; j < kTile / T / 2; ++j) {
        const int e = (j * T + (int)threadIdx.x) * 2;
        if (vec_ok && e + 1 < nv) {
            double2 d = __ldg(reinterpret_cast<const double2 *>(p + e));
            v[2 * j] = src.fix(d.x);
            v[2 * j + 1] = src.fix(d.y);
        } else {
            v[2 * j] = e < nv ? src.fix(__ldg(p + e)) : fill;
            v[2 * j + 1] = e + 1 < nv ? src.fix(__ldg(p + e + 1)) : fill;
        }
    }
}
template <typename X, int T = kThreads>
__device__ __forceinline__ void store_tile(X *out, int64_t base, int64_t valid, const X (&v)[kTile / T]) {
    X *p = out + base;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & (2 * sizeof(X) - 1)) == 0);
    const int nv = (int)valid;
#pragma unroll
    for (int j = 0; j < kTile / T / 2; ++j) {
        const int e = (j * T + (int)threadIdx.x) * 2;
        if (vec_ok && e + 1 < nv) {
            struct alignas(2 * sizeof(X)) V2 { X a, b; };
            V2 d{v[2 * j], v[2 * j + 1]};
            *reinterpret_cast<V2 *>(p + e) = d;
        } else {
            if (e < nv) p[e] = v[2 * j];
            if (e + 1 < nv) p[e + 1] = v[2 * j + 1];
        }
    }
}

// In-tile inclusive prefix sum over the striped layout.  On return incl[k] is the inclusive sum of all
// tile elements up to and including this thread's slot k; returns the tile total to all threads.
// smem: 32 values of X ((kTile/T/2 rows) x (T/32 warps) = 32 cells for every T).  Deterministic association.
template <typename X, int T = kThreads>
__device__ __forceinline__ X tile_scan(const X (&x)[kTile / T], X (&incl)[kTile / T], X *smem) {
    constexpr int V = kTile / T / 2, NW = T / 32;
    static_assert(V * NW == 32, "tile_scan needs 32 (row, warp) cells");
    X lane_excl[V], inc[V];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        X s = x[2 * j] + x[2 * j + 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            X t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        inc[j] = s;
        X e = __shfl_up_sync(0xffffffffu, s, 1);  // exact exclusive prefix (no subtraction: fp64)
        lane_excl[j] = lane == 0 ? X(0) : e;
    }
    __syncthreads();
    if (lane == 31) {
#pragma unroll
        for (int j = 0; j < V; ++j) smem[j * NW + warp] = inc[j];
    }
    __syncthreads();
    X total;
    {
        // exclusive scan over the 32 (row j, warp w) cells, row-major == element order
        X s = smem[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            X t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        total = __shfl_sync(0xffffffffu, s, 31);
        X e = __shfl_up_sync(0xffffffffu, s, 1);
        __syncthreads();
        if (warp == 0) smem[lane] = lane == 0 ? X(0) : e;  // exclusive
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
        X excl = smem[j * NW + warp] + lane_excl[j];
        incl[2 * j] = excl + x[2 * j];
        incl[2 * j + 1] = incl[2 * j] + x[2 * j + 1];
    }
    return total;
}

}  // namespace genpf
