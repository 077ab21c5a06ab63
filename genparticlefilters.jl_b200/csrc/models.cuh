// models.cuh -- device plugins: transition sample / observation log-density over struct-of-arrays
// particle state.  A plugin is a struct with
//   NF fp64 fields + NB u8 fields per time slice, NP parameters, NAUX per-step host-supplied scalars,
//   initial(P, slice0)                         -- the constant slice before t = 1
//   transition(P, aux, t, prev, next, U, Z)    -- Gen `generate`/`update` sampling the new latents
//   obs_logpdf(P, slice, obs)                  -- log-density of the constrained observation address
// The generic kernels in filter.cuh build pf_initialize / pf_update! / mh-rejuvenation from these.
// Arithmetic that decides values (moving, y) uses explicit __dmul_rn/__dadd_rn so nvcc cannot contract
// to FMA: with identical noise the new state columns are bit-identical to a CPU restatement's.
#pragma once
#include "common.cuh"

namespace genpf {

constexpr int kMaxParams = 8;
constexpr int kMaxAux = 4;
constexpr int kMaxF = 2;  // fp64 fields per slice
constexpr int kMaxB = 2;  // u8 fields per slice

struct ModelParams {
    double v[kMaxParams];
    double aux[kMaxAux];
};

template <int NF, int NB>
struct SliceT {
    double f[NF > 0 ? NF : 1];
    uint8_t b[NB > 0 ? NB : 1];
};

#define GENPF_LOG_2PI 1.8378770664093453  // log(2*pi) as computed by Julia/glibc in fp64

// Gen: logpdf(normal, x, mu, sigma) = -(((x-mu)/sigma)^2 + log(2pi))/2 - log(sigma)
// (x-mu)/sigma is evaluated as (x-mu)*(1/sigma): exact for the README's sigma = 0.25, <= 1 ulp otherwise
// (log-weights are compared at 1e-10 relative).
__device__ __forceinline__ double normal_logpdf(double x, double mu, double inv_sigma, double log_sigma) {
    double z = (x - mu) * inv_sigma;
    return -(__dmul_rn(z, z) + GENPF_LOG_2PI) / 2.0 - log_sigma;
}

// accept iff log(U3) < alpha (Gen mh).  Decision-exact fast path: an fp32 log with a conservative error
// bound decides unless it lands within the bound of alpha, in which case the fp64 log is evaluated.
__device__ __forceinline__ bool mh_accept(double U3, double alpha) {
    if (alpha > 0.0) return true;  // log(U3) <= 0 < alpha
    const float lf = __logf((float)U3);
    const float af = (float)alpha;
    if (fabsf(lf - af) > 1e-4f * (1.0f + fabsf(lf))) return lf < af;
    return log(U3) < alpha;
}

// README.md:43-54 object_motion.  params: v[0]=p_stay .75, v[1]=p_start .25, v[2]=sigma_proc .01,
// v[3]=sigma_obs .25, v[4]=log(sigma_obs), v[5]=1/sigma_obs (host filled).  aux[0] = vel_t = sin(t) from the caller.
struct ObjectMotion {
    static constexpr int NF = 1, NB = 1, NP = 4, NAUX = 1;
    using Slice = SliceT<NF, NB>;
    static __device__ __forceinline__ void initial(const ModelParams &, Slice &s) {
        s.f[0] = 0.0;  // y = 0, moving = false (README.md:44)
        s.b[0] = 0;
    }
    static __device__ __forceinline__ void transition(const ModelParams &p, int64_t, const Slice &prev, Slice &nxt,
                                                      double U, double Z) {
        uint8_t m = U < (prev.b[0] ? p.v[0] : p.v[1]);              // bernoulli: rand() < p
        double mu = __dadd_rn(prev.f[0], m ? p.aux[0] : 0.0);       // y + vel_y
        nxt.f[0] = __dadd_rn(mu, __dmul_rn(p.v[2], Z));             // normal: mu + sigma*randn()
        nxt.b[0] = m;
    }
    static __device__ __forceinline__ double obs_logpdf(const ModelParams &p, const Slice &s, double obs) {
        return normal_logpdf(obs, s.f[0], p.v[5], p.v[4]);
    }
};

// 1-D linear-Gaussian tracker: x_0 ~ N(m0, s0) marginalised into the first transition,
// x_t ~ N(a x_{t-1}, q), y_t ~ N(x_t, r).  params: v[0]=a, v[1]=q, v[2]=r, v[3]=m0, v[4]=s0,
// v[5]=log(r), v[6]=sqrt(a^2 s0^2 + q^2), v[7]=1/r (host filled).
struct LinGauss1D {
    static constexpr int NF = 1, NB = 0, NP = 5, NAUX = 0;
    using Slice = SliceT<NF, NB>;
    static __device__ __forceinline__ void initial(const ModelParams &p, Slice &s) {
        s.f[0] = p.v[3];
        s.b[0] = 0;
    }
    static __device__ __forceinline__ void transition(const ModelParams &p, int64_t t, const Slice &prev, Slice &nxt,
                                                      double, double Z) {
        double sig = (t == 1) ? p.v[6] : p.v[1];
        nxt.f[0] = __dadd_rn(__dmul_rn(p.v[0], prev.f[0]), __dmul_rn(sig, Z));
        nxt.b[0] = 0;
    }
    static __device__ __forceinline__ double obs_logpdf(const ModelParams &p, const Slice &s, double obs) {
        return normal_logpdf(obs, s.f[0], p.v[7], p.v[5]);
    }
};

// ------------------------------------------------------------------ noise policies
// Draw order per particle: init/update = [U1 (bernoulli), Z1 (normal)]; mh = [U2, Z2, U3 (accept)].
// Lean: ONE Philox4x32-10 call per particle per purpose: U = (w0+.5)2^-32, Z = fp32 Box-Muller(w1, w2),
//       U3 = (w3+.5)2^-32.  Counter = global particle slot, stream = (purpose, step).
struct NoiseLean {
    uint64_t seed, stream;
    int64_t offset;
    __device__ __forceinline__ void get(int64_t i, double &U, double &Z, double &U3) const {
        uint4 o = philox_at(seed, stream, (uint64_t)(i + offset));
        U = ((double)o.x + 0.5) * 0x1.0p-32;
        float ua = ((float)(o.y >> 8) + 0.5f) * 0x1.0p-24f;
        float ub = ((float)(o.z >> 8) + 0.5f) * 0x1.0p-24f;
        float rr = sqrtf(-2.0f * __logf(ua));
        Z = (double)(rr * __cosf(6.28318530717958647692f * (ub - 0.5f)));  // angle in [-pi, pi): MUFU range
        U3 = ((double)o.w + 0.5) * 0x1.0p-32;
    }
};
// 53-bit uniforms + fp64 Box-Muller (two Philox calls)
struct NoisePhilox53 {
    uint64_t seed, stream;
    int64_t offset;
    __device__ __forceinline__ void get(int64_t i, double &U, double &Z, double &U3) const {
        uint4 a = philox_at(seed, stream, (uint64_t)(i + offset));
        uint4 b = philox_at(seed, stream ^ (1ull << 55), (uint64_t)(i + offset));
        U = u53(a.x, a.y);
        double ua = 1.0 - u53(a.z, a.w);  // (0,1]
        double ub = u53(b.x, b.y);
        Z = sqrt(-2.0 * log(ua)) * cospi(2.0 * ub);
        U3 = 1.0 - u53(b.z, b.w);  // (0,1]
    }
};
// parity mode: noise supplied as columns (exported from the reference's RNG)
struct NoiseCols {
    const double *U, *Z, *U3;
    __device__ __forceinline__ void get(int64_t i, double &u, double &z, double &u3) const {
        u = U ? U[i] : 0.0;
        z = Z ? Z[i] : 0.0;
        u3 = U3 ? U3[i] : 1.0;
    }
};

}  // namespace genpf
