// Test plugin: the 1-D linear-Gaussian tracker of SURVEY B.2 written as a USER plugin (the same arithmetic as the
// built-in LinGauss1D, csrc/models.cuh) plus a trace translator.  Compiled at run time by genpf_model_compile.
// params: v[0]=a, v[1]=q, v[2]=r, v[3]=m0, v[4]=s0, v[5]=log(r), v[6]=sqrt(a^2 s0^2 + q^2), v[7]=1/r
struct MyLinGauss {
    static constexpr int NF = 1, NB = 0, NP = 8, NAUX = 0;
    using Slice = genpf::SliceT<NF, NB>;
    static __device__ __forceinline__ void initial(const genpf::ModelParams &p, Slice &s) {
        s.f[0] = p.v[3];
        s.b[0] = 0;
    }
    static __device__ __forceinline__ void transition(const genpf::ModelParams &p, int64_t t, const Slice &prev, Slice &nxt,
                                                      double, double Z) {
        double sig = (t == 1) ? p.v[6] : p.v[1];
        nxt.f[0] = __dadd_rn(__dmul_rn(p.v[0], prev.f[0]), __dmul_rn(sig, Z));
        nxt.b[0] = 0;
    }
    static __device__ __forceinline__ double obs_logpdf(const genpf::ModelParams &p, const Slice &s, double obs) {
        return genpf::normal_logpdf(obs, s.f[0], p.v[7], p.v[5]);
    }
    static __device__ __forceinline__ double constrain(const genpf::ModelParams &p, int64_t t, const Slice &prev, Slice &nxt,
                                                       double, double, int, double val) {
        const double sig = (t == 1) ? p.v[6] : p.v[1];
        nxt.f[0] = val;
        nxt.b[0] = 0;
        return genpf::normal_logpdf(val, __dmul_rn(p.v[0], prev.f[0]), 1.0 / sig, log(sig));
    }
    // translator (update.jl:35-44): deterministic map x -> 2x + 1, scored against the observation
    static __device__ __forceinline__ double translate(const genpf::ModelParams &, int64_t, const Slice &cur, double obs,
                                                       Slice &nxt, double, double) {
        nxt.f[0] = 2.0 * cur.f[0] + 1.0;
        nxt.b[0] = 0;
        return -0.5 * (obs - nxt.f[0]) * (obs - nxt.f[0]);
    }
};
