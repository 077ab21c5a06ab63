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
    return -(__dmul_rn(z, z) + kC[kcLog2Pi]) / 2.0 - log_sigma;
}

// accept iff log(U3) < alpha (Gen mh).  Decision-exact fast path: an fp32 log with a conservative error
// bound decides unless it lands within the bound of alpha, in which case the fp64 log is evaluated.
static __device__ __noinline__ bool mh_accept_slow(double U3, double alpha) { return log(U3) < alpha; }
__device__ __forceinline__ bool mh_accept(double U3, double alpha) {
    if (alpha > 0.0) return true;  // log(U3) <= 0 < alpha
    const float lf = __logf((float)U3);
    const float af = (float)alpha;
    if (fabsf(lf - af) > 1e-4f * (1.0f + fabsf(lf))) return lf < af;
    return mh_accept_slow(U3, alpha);
}

// ------------------------------------------------------------------ optional members of a plugin (detected, never required)
//   transition_logpdf(P, t, prev, x)            log p(x_t | x_{t-1})                         -- Gen `assess`/`update` weight
//   propose(P, t, prev, obs, nxt, U, Z)         x_t ~ q(. | x_{t-1}, obs_t)                  -- Gen `propose` (initialize.jl:55,122; update.jl:85)
//   proposal_logpdf(P, t, prev, obs, x)         log q(x | x_{t-1}, obs_t)                    -- its score
//   translate(P, t, cur, obs, nxt, U, Z)        (new slice, log-weight increment)            -- translator(trace) (update.jl:35-44)
template <class M, class = void>
struct has_proposal { static constexpr bool value = false; };
template <class M>
struct has_proposal<M, decltype((void)&M::proposal_logpdf, (void)&M::propose, (void)&M::transition_logpdf)> {
    static constexpr bool value = true;
};
template <class M, class = void>
struct has_translate { static constexpr bool value = false; };
template <class M>
struct has_translate<M, decltype((void)&M::translate)> { static constexpr bool value = true; };

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
    // generate(model, args, merge(stratum, observations)) (initialize.jl:101-104): latent `fld` (0 = y, 1 = moving)
    // is constrained to `val`, the other one is sampled; returns the log-density of the constraint
    static __device__ __forceinline__ double constrain(const ModelParams &p, int64_t, const Slice &prev, Slice &nxt,
                                                       double U, double Z, int fld, double val) {
        const double pm = prev.b[0] ? p.v[0] : p.v[1];
        if (fld == 1) {
            const uint8_t m = val != 0.0;
            const double mu = __dadd_rn(prev.f[0], m ? p.aux[0] : 0.0);
            nxt.f[0] = __dadd_rn(mu, __dmul_rn(p.v[2], Z));
            nxt.b[0] = m;
            return log(m ? pm : 1.0 - pm);  // logpdf(bernoulli, m, pm)
        }
        const uint8_t m = U < pm;
        const double mu = __dadd_rn(prev.f[0], m ? p.aux[0] : 0.0);
        nxt.f[0] = val;
        nxt.b[0] = m;
        return normal_logpdf(val, mu, 1.0 / p.v[2], log(p.v[2]));
    }
    // custom proposal (initialize.jl:46-62, update.jl:79-96): the motion flag is proposed from a fair coin instead of
    // the sticky prior, y from its prior conditional; importance weight = p(m') / 0.5
    static __device__ __forceinline__ double transition_logpdf(const ModelParams &p, int64_t, const Slice &prev, const Slice &x) {
        const double pm = prev.b[0] ? p.v[0] : p.v[1];
        const double mu = __dadd_rn(prev.f[0], x.b[0] ? p.aux[0] : 0.0);
        return log(x.b[0] ? pm : 1.0 - pm) + normal_logpdf(x.f[0], mu, 1.0 / p.v[2], log(p.v[2]));
    }
    static __device__ __forceinline__ void propose(const ModelParams &p, int64_t, const Slice &prev, double, Slice &nxt,
                                                   double U, double Z) {
        const uint8_t m = U < 0.5;
        const double mu = __dadd_rn(prev.f[0], m ? p.aux[0] : 0.0);
        nxt.f[0] = __dadd_rn(mu, __dmul_rn(p.v[2], Z));
        nxt.b[0] = m;
    }
    static __device__ __forceinline__ double proposal_logpdf(const ModelParams &p, int64_t, const Slice &prev, double,
                                                             const Slice &x) {
        const double mu = __dadd_rn(prev.f[0], x.b[0] ? p.aux[0] : 0.0);
        return log(0.5) + normal_logpdf(x.f[0], mu, 1.0 / p.v[2], log(p.v[2]));
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
    static __device__ __forceinline__ double constrain(const ModelParams &p, int64_t t, const Slice &prev, Slice &nxt,
                                                       double, double, int, double val) {
        const double sig = (t == 1) ? p.v[6] : p.v[1];
        nxt.f[0] = val;
        nxt.b[0] = 0;
        return normal_logpdf(val, __dmul_rn(p.v[0], prev.f[0]), 1.0 / sig, log(sig));
    }
    // custom proposal: the locally optimal one, q(x_t | x_{t-1}, y_t) = N(mu*, s*^2) with 1/s*^2 = 1/q^2 + 1/r^2 and
    // mu* = s*^2 (a x_{t-1} / q^2 + y_t / r^2); the importance weight is then p(y_t | x_{t-1}) for every draw
    static __device__ __forceinline__ void opt_moments(const ModelParams &p, int64_t t, const Slice &prev, double obs,
                                                       double &mu, double &sd) {
        const double sig = (t == 1) ? p.v[6] : p.v[1];
        const double iq = 1.0 / (sig * sig), ir = 1.0 / (p.v[2] * p.v[2]);
        const double var = 1.0 / (iq + ir);
        mu = var * (__dmul_rn(p.v[0], prev.f[0]) * iq + obs * ir);
        sd = sqrt(var);
    }
    static __device__ __forceinline__ double transition_logpdf(const ModelParams &p, int64_t t, const Slice &prev, const Slice &x) {
        const double sig = (t == 1) ? p.v[6] : p.v[1];
        return normal_logpdf(x.f[0], __dmul_rn(p.v[0], prev.f[0]), 1.0 / sig, log(sig));
    }
    static __device__ __forceinline__ void propose(const ModelParams &p, int64_t t, const Slice &prev, double obs, Slice &nxt,
                                                   double, double Z) {
        double mu, sd;
        opt_moments(p, t, prev, obs, mu, sd);
        nxt.f[0] = __dadd_rn(mu, __dmul_rn(sd, Z));
        nxt.b[0] = 0;
    }
    static __device__ __forceinline__ double proposal_logpdf(const ModelParams &p, int64_t t, const Slice &prev, double obs,
                                                             const Slice &x) {
        double mu, sd;
        opt_moments(p, t, prev, obs, mu, sd);
        return normal_logpdf(x.f[0], mu, 1.0 / sd, log(sd));
    }
};

// ------------------------------------------------------------------ noise policies
// Draw order per particle: init/update = [U1 (bernoulli), Z1 (normal)]; mh = [U2, Z2, U3 (accept)].
// A policy serves one README iteration "step s": the update to time s and the mh move applied right before
// it (on slice s-1).  Interface:
//   up(i, U, Z)              noise of pf_update!/pf_initialize to time s
//   mh(i, it, U, Z, U3)      noise of mh iteration `it` on slice s-1
//   both(i, ...)             the two at once (fused step, iteration 0)
__device__ __forceinline__ float fast_bm_radius(float ua) {
    const float x = -2.0f * __logf(ua);  // >= 0 (ua can round up to 1.0f)
    return x * rsqrtf(fmaxf(x, 1e-30f));
}

// Lean (default): ONE Philox4x32-10 call per particle per step s serves BOTH moves (128 bits):
//   w0[31:8] U_mh (24 b)   w1[31:8] U_up (24 b)   w2 U_acc (32 b)   w3[31:8] Box-Muller angle (24 b)
//   {w0[7:0], w1[7:0], w3[7:0]} Box-Muller radius uniform (24 b);  Z_mh = r cos(theta), Z_up = r sin(theta)
// (the two outputs of one Box-Muller transform are independent normals).  fp32 transcendental units.
// mh iterations it >= 1 draw their own call: U = w0, radius w1, angle w2, U_acc = w3.
struct NoiseLean {
    static constexpr bool kIndexed = false;
    uint64_t seed;
    uint64_t step;  // s
    int64_t offset;
    __device__ __forceinline__ void both(int64_t i, double &U_mh, double &Z_mh, double &U_acc, double &U_up,
                                         double &Z_up) const {
        const uint4 o = philox_at(seed, make_stream(kPurposeUpdate, step), (uint64_t)(i + offset));
        U_mh = ((double)(o.x >> 8) + 0.5) * 0x1.0p-24;
        U_up = ((double)(o.y >> 8) + 0.5) * 0x1.0p-24;
        U_acc = ((double)o.z + 0.5) * 0x1.0p-32;
        const uint32_t rb = ((o.x & 0xFFu) << 16) | ((o.y & 0xFFu) << 8) | (o.w & 0xFFu);
        const float ua = ((float)rb + 0.5f) * 0x1.0p-24f;
        const float th = (((float)(o.w >> 8) + 0.5f) * 0x1.0p-24f - 0.5f) * 6.28318530717958647692f;  // [-pi, pi)
        const float rr = fast_bm_radius(ua);
        float sn, cs;
        __sincosf(th, &sn, &cs);
        Z_mh = (double)(rr * cs);
        Z_up = (double)(rr * sn);
    }
    __device__ __forceinline__ void up(int64_t i, double &U, double &Z) const {
        double a, b, c;
        both(i, a, b, c, U, Z);
    }
    __device__ __forceinline__ void mh(int64_t i, int it, double &U, double &Z, double &U3) const {
        if (it == 0) {
            double a, b;
            both(i, U, Z, U3, a, b);
            return;
        }
        const uint4 o = philox_at(seed, make_stream(kPurposeMH, (step << 8) | (uint64_t)(it & 0xFF)), (uint64_t)(i + offset));
        U = ((double)o.x + 0.5) * 0x1.0p-32;
        const float ua = ((float)(o.y >> 8) + 0.5f) * 0x1.0p-24f;
        const float th = (((float)(o.z >> 8) + 0.5f) * 0x1.0p-24f - 0.5f) * 6.28318530717958647692f;
        Z = (double)(fast_bm_radius(ua) * __cosf(th));
        U3 = ((double)o.w + 0.5) * 0x1.0p-32;
    }
};
// ------------------------------------------------------------------ fp64 Box-Muller building blocks
// The production noise of SURVEY 8(c): 53-bit uniforms (x >> 11) * 2^-53 and an fp64 Box-Muller transform.  The
// step kernels are instruction-issue bound, so the three transcendental pieces are written for exactly the
// arguments they get (no special cases, no denormals), ~1 ulp each (checked against long-double references):
//   log_unit   : log(u) for u = k * 2^-53, k in [1, 2^53]  (atanh-series form, reciprocal by MUFU.RCP64H + Newton)
//   sqrt_pos   : sqrt(x) for x in [1e-300, 1e3]            (MUFU.RSQ64H + two Goldschmidt steps + one correction)
//   sincos_2pi : sin / cos of 2 pi a for a in [0, 1)        (exact quadrant reduction, degree-13/14 kernels)
__device__ __forceinline__ double log_unit(double u) {
    int hi = __double2hiint(u);
    const int lo = __double2loint(u);
    hi += 0x3FF00000 - 0x3FE6A09E;  // mantissa into [sqrt(1/2), sqrt(2))
    const int k = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) + 0x3FE6A09E;
    const double f = __hiloint2double(hi, lo) - 1.0;  // exact
    const double d = 2.0 + f;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    double sq = f * r;
    sq = fma(fma(-sq, d, f), r, sq);  // s = f / (2 + f), correctly rounded up to the last bit
    const double z = sq * sq, w = z * z;
    const double t1 = w * fma(w, fma(w, kC[kcLogA1], kC[kcLogA2]), kC[kcLogA3]);
    const double t2 = z * fma(w, fma(w, fma(w, kC[kcLogB1], kC[kcLogB2]), kC[kcLogB3]), kC[kcLogB4]);
    const double R = t1 + t2;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return dk * kC[kcLn2Hi] - ((hfsq - fma(sq, hfsq + R, dk * kC[kcLn2Lo])) - f);
}
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
}
__device__ __forceinline__ void sincos_2pi(double a, double &sn, double &cs) {
    const double SH = 6755399441055744.0;  // 2^52 + 2^51
    const double zq = fma(a, 4.0, SH);
    const int q = __double2loint(zq);             // rint(4a) in 0..4
    const double t = fma(zq - SH, -0.25, a);      // exact, |t| <= 1/8
    const double phi = t * kC[kcTwoPi];
    const double z = phi * phi;
    double ps = fma(z, kC[kcSinS6], kC[kcSinS5]);
    ps = fma(z, ps, kC[kcSinS4]);
    ps = fma(z, ps, kC[kcSinS3]);
    ps = fma(z, ps, kC[kcSinS2]);
    ps = fma(z, ps, kC[kcSinS1]);
    const double sp = fma(phi * z, ps, phi);
    double pc = fma(z, kC[kcCosC6], kC[kcCosC5]);
    pc = fma(z, pc, kC[kcCosC4]);
    pc = fma(z, pc, kC[kcCosC3]);
    pc = fma(z, pc, kC[kcCosC2]);
    pc = fma(z, pc, kC[kcCosC1]);
    const double cp = fma(z * z, pc, fma(z, -0.5, 1.0));
    const bool swap = q & 1;
    double s0 = swap ? cp : sp, c0 = swap ? sp : cp;
    // q: 0 (s, c)  1 (c, -s)  2 (-s, -c)  3 (-c, s)  4 == 0
    const int ss = (q & 2) << 30, sc = ((q + 1) & 2) << 30;
    sn = __hiloint2double(__double2hiint(s0) ^ ss, __double2loint(s0));
    cs = __hiloint2double(__double2hiint(c0) ^ sc, __double2loint(c0));
}
// two independent standard normals from ua in (0, 1], ub in [0, 1)
__device__ __forceinline__ void normal_pair(double ua, double ub, double &z0, double &z1) {
    const double r = sqrt_pos(fmax(-2.0 * log_unit(ua), 1e-300));
    double sn, cs;
    sincos_2pi(ub, sn, cs);
    z0 = r * cs;
    z1 = r * sn;
}

// High-fidelity (SURVEY 8c production noise): 53-bit uniforms + fp64 Box-Muller.  One README iteration draws
// THREE Philox4x32-10 blocks per particle: a = {U_mh, U_up}, b = {Box-Muller radius, angle}, c = {U_acc, -};
// the two outputs of the one transform are the two independent normals Z_mh = r cos, Z_up = r sin.  The unfused
// kernels (k_mh iteration 0, k_propagate) derive their draws from the same three blocks, so fused == separate.
// mh iterations it >= 1 draw two blocks of their own stream.
struct NoisePhilox53 {
    static constexpr bool kIndexed = false;
    uint64_t seed;
    uint64_t step;
    int64_t offset;
    __device__ __forceinline__ void both(int64_t i, double &U_mh, double &Z_mh, double &U_acc, double &U_up,
                                         double &Z_up) const {
        const uint64_t st = make_stream(kPurposeUpdate, step), c = (uint64_t)(i + offset);
        const uint4 a = philox_at(seed, st, c);
        const uint4 b = philox_at(seed, st ^ (1ull << 55), c);
        const uint4 d = philox_at(seed, st ^ (1ull << 54), c);
        U_mh = u53(a.x, a.y);
        U_up = u53(a.z, a.w);
        U_acc = 1.0 - u53(d.x, d.y);  // (0,1]
        normal_pair(1.0 - u53(b.x, b.y), u53(b.z, b.w), Z_mh, Z_up);
    }
    __device__ __forceinline__ void up(int64_t i, double &U, double &Z) const {
        const uint64_t st = make_stream(kPurposeUpdate, step), c = (uint64_t)(i + offset);
        const uint4 a = philox_at(seed, st, c);
        const uint4 b = philox_at(seed, st ^ (1ull << 55), c);
        U = u53(a.z, a.w);
        double zc;
        normal_pair(1.0 - u53(b.x, b.y), u53(b.z, b.w), zc, Z);
    }
    __device__ __forceinline__ void mh(int64_t i, int it, double &U, double &Z, double &U3) const {
        if (it == 0) {
            double a, b;
            both(i, U, Z, U3, a, b);
            return;
        }
        const uint64_t st = make_stream(kPurposeMH, (step << 8) | (uint64_t)(it & 0xFF)), c = (uint64_t)(i + offset);
        const uint4 a = philox_at(seed, st, c);
        const uint4 b = philox_at(seed, st ^ (1ull << 55), c);
        U = u53(a.x, a.y);
        U3 = 1.0 - u53(a.z, a.w);
        double zs;
        normal_pair(1.0 - u53(b.x, b.y), u53(b.z, b.w), Z, zs);
    }
};
// parity mode: noise supplied as columns (exported from the reference's RNG), read by particle slot.
//   mh move      : U, Z, U3        (draw order [U2, Z2, U3])
//   update / init: Uup, Zup        (draw order [U1, Z1]); when both are null the update reads U, Z
// kIndexed tells the fused kernels that the slot must be in range (Philox policies accept any counter).
struct NoiseCols {
    static constexpr bool kIndexed = true;
    const double *U, *Z, *U3;
    const double *Uup, *Zup;
    __device__ __forceinline__ void up(int64_t i, double &u, double &z) const {
        const double *pu = (Uup || Zup) ? Uup : U, *pz = (Uup || Zup) ? Zup : Z;
        u = pu ? pu[i] : 0.0;
        z = pz ? pz[i] : 0.0;
    }
    __device__ __forceinline__ void mh(int64_t i, int, double &u, double &z, double &u3) const {
        u = U ? U[i] : 0.0;
        z = Z ? Z[i] : 0.0;
        u3 = U3 ? U3[i] : 1.0;
    }
    __device__ __forceinline__ void both(int64_t i, double &U_mh, double &Z_mh, double &U_acc, double &U_up,
                                         double &Z_up) const {
        mh(i, 0, U_mh, Z_mh, U_acc);
        up(i, U_up, Z_up);
    }
};

}  // namespace genpf
