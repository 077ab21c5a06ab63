/*
 * genpf_oracle.h -- CPU restatement of the GenParticleFilters.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it, and only as the checker / the CPU arm.
 *
 * PARITY PINNING: the reference (Julia) cannot run in this image and ships no
 * seeded golden vectors, so for ANCESTOR INDICES this
 * oracle is "parity unpinned" against real reference output.  What IS pinned:
 * every known-answer invariant the reference's own tests state
 * (test/resample.jl, test/resize.jl, test/utils.jl, test/statistics.jl), see
 * tests/test_oracle_kat.py and tests/golden/.
 *
 * All indices crossing this interface are 1-based int64 (Julia `Vector{Int}`,
 * reference src/view.jl:21).  All weights are fp64.
 */
#ifndef GENPF_ORACLE_H
#define GENPF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MULTINOMIAL = 0, ORC_RESIDUAL = 1, ORC_STRATIFIED = 2 };
enum {
    ORC_FLAG_SORT = 1,      /* sort_particles=true (resample.jl:143-145,156-157) */
    ORC_FLAG_SUBSTATE = 2,  /* ParticleFilterSubState semantics (resample.jl:184-187,205-218) */
    ORC_FLAG_EXACT_CUMSUM = 4 /* long-double cumulative sum instead of the literal fp64 one */
};
/* invalid kinds = the four branches of safe_softmax, utils.jl:119-137 */
enum { ORC_VALID = 0, ORC_INV_NAN_INPUT = 1, ORC_INV_ALL_NEGINF = 2, ORC_INV_ZERO_TOTAL = 3, ORC_INV_NAN_TOTAL = 4 };

/* ---- counter-based RNG shared with the CUDA library (Philox4x32-10) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_uniform53(uint64_t seed, uint64_t stream, uint64_t idx);
void orc_fill_uniform53(uint64_t seed, uint64_t stream, int64_t n, double *out);
double orc_uniform_strata(uint64_t seed, uint64_t stream, uint64_t idx);
void orc_fill_uniform_strata(uint64_t seed, uint64_t stream, int64_t n, double *out);

/* ---- Julia Base / Gen one-liners ---- */
double orc_sum_pairwise(const double *a, int64_t n);      /* Base.sum: pairwise, 1024 block */
double orc_maximum(const double *a, int64_t n);
double orc_logsumexp(const double *v, int64_t n);          /* Gen.logsumexp */
void orc_lognorm(const double *v, int64_t n, double *out); /* utils.jl:100 */
void orc_softmax(const double *v, int64_t n, double *out); /* utils.jl:103-107 */
int32_t orc_safe_softmax(const double *v, int64_t n, double *out); /* utils.jl:117-140, returns invalid kind */
double orc_ess(const double *lw, int64_t n);               /* utils.jl:163-164 + Gen */
double orc_lml_estimate(double log_ml_est, const double *lw, int64_t n); /* utils.jl:174-178 */

/* ---- ancestor selection (resample.jl) ---- */
void orc_sortperm_desc(const double *keys, int64_t n, int64_t *order1);  /* sortperm(v, rev=true), 1-based */
void orc_select_multinomial(const double *w, int64_t n, const double *u, int64_t n_out, int64_t *parents1);
void orc_select_stratified(const double *w, const int64_t *order1_or_null, int64_t n, const double *r,
                           uint32_t flags, int64_t *parents1);
/* search form of the same thing, used to cross-check the literal loop */
void orc_select_stratified_search(const double *w, const int64_t *order1_or_null, int64_t n, const double *r,
                                  uint32_t flags, int64_t *parents1);
void orc_select_residual(const double *w, int64_t n, const double *u, int64_t n_out, int64_t *parents1,
                         int64_t *n_deterministic);
void orc_cumweights(const double *w, const int64_t *order1_or_null, int64_t n, uint32_t flags, double *W);

/* full pf_resample! / pf_*_resize! on the plain-bits members (resample.jl:48-218, resize.jl:46-124,424-438) */
int32_t orc_resample(int32_t method, const double *lw, const double *lp_or_null, int64_t n_in, int64_t n_out,
                     const double *uniforms, uint32_t flags, int64_t *parents1, double *lw_out,
                     double *lml_increment, int32_t *invalid_kind);

/* ---- statistics (statistics.jl:13-17,48-54) ---- */
void orc_mean_var(const double *lw, const double *x, int64_t n, double *mean, double *var);

/* ---- resizing (resize.jl:236-244,267-297,309-334) ---- */
void orc_replicate(const double *lw, int64_t n, int64_t k, int32_t interleaved, int64_t *parents1, double *lw_out);
void orc_dereplicate(const double *lw, int64_t n, int64_t k, int32_t interleaved, int32_t sample,
                     const double *u, int64_t *parents1, double *lw_out);
int64_t orc_coalesce(const double *lw, const int64_t *keys, int64_t n, int64_t *parents1, double *lw_out);

/* ---- pf_optimal_resize! (resize.jl:149-196), find_inv_w_threshold (resize.jl:199-216) ---- */
double orc_find_inv_w_threshold(const double *w, int64_t n, int64_t n_particles);
int32_t orc_optimal_resize(const double *lw, int64_t n, int64_t n_out, double u_rand, int64_t *parents1,
                           double *lw_out, int64_t *n_keep, double *inv_w, int32_t *invalid_kind,
                           int32_t *invalid_kind_strat, int64_t *n_selected);

/* ---- device-plugin models, noise supplied as columns ---- */
typedef struct {
    double p_stay, p_start, sigma_proc, sigma_obs;
} orc_om_params;
void orc_om_transition(const orc_om_params *p, int64_t n, const double *y_prev, const uint8_t *m_prev, double vel,
                       const double *U, const double *Z, double *y_out, uint8_t *m_out);
void orc_om_obs_logpdf_add(const orc_om_params *p, int64_t n, const double *y, double obs, double *lw, int32_t assign);
void orc_om_mh(const orc_om_params *p, int64_t n, const double *y_pp, const uint8_t *m_pp, double *y_cur,
               uint8_t *m_cur, double vel, double obs, const double *U2, const double *Z2, const double *U3,
               uint8_t *accept);

typedef struct {
    double a, q, r, m0, s0;
} orc_lg_params;
void orc_lg_transition(const orc_lg_params *p, int64_t n, const double *x_prev, const double *Z, double *x_out);
void orc_lg_obs_logpdf_add(const orc_lg_params *p, int64_t n, const double *x, double obs, double *lw, int32_t assign);
void orc_lg_mh(const orc_lg_params *p, int64_t n, const double *x_pp, double *x_cur, double obs, const double *Z2,
               const double *U3, uint8_t *accept);

double orc_normal_logpdf(double x, double mu, double sigma);

/* ---- CPU baseline: one README-loop iteration of object_motion with Philox noise (OpenMP when built with it) */
typedef struct {
    int64_t n;
    double *y[2];
    uint8_t *m[2];
    double *y_new[2];
    uint8_t *m_new[2];
    double *lw, *w, *r;
    int64_t *parents;
    double log_ml_est;
    uint64_t seed;
    int32_t cur; /* slot of the newest slice */
} orc_om_filter;
orc_om_filter *orc_om_filter_create(int64_t n, uint64_t seed);
void orc_om_filter_destroy(orc_om_filter *f);
void orc_om_filter_init(orc_om_filter *f, const orc_om_params *p, double vel1, double obs1);
/* ESS -> stratified resample (sort=false) -> MH(t-1) -> update(t); returns ESS before resampling */
double orc_om_filter_step(orc_om_filter *f, const orc_om_params *p, int64_t t, double vel_prev, double obs_prev,
                          double vel_t, double obs_t);
int32_t orc_num_threads(void);
void orc_set_num_threads(int32_t n);

#ifdef __cplusplus
}
#endif
#endif
