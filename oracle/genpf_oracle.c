/*
 * genpf_oracle.c -- literal CPU restatement of the GenParticleFilters.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see genpf_oracle.h).  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).  Loops
 * are kept sequential exactly where the reference is sequential (cumulative
 * sums, residual fill) because ancestor parity depends on that association
 * order.  "parity unpinned" for ancestor indices: see header.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off; no fast-math: FMA
 * contraction would change `rand()*step + lower`, resample.jl:162).
 */
#define _GNU_SOURCE
#include "genpf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ RNG */
/* Philox4x32-10 (Salmon et al. 2011); the CUDA library implements the same
 * function so that seed-generated uniforms are identical on both sides. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void philox_at(uint64_t seed, uint64_t stream, uint64_t idx, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

/* 53-bit uniform in [0,1): (x >> 11) * 2^-53 with x = out[1]:out[0] */
double orc_uniform53(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint32_t o[4];
    philox_at(seed, stream, idx, o);
    uint64_t x = ((uint64_t)o[1] << 32) | o[0];
    return (double)(x >> 11) * 0x1.0p-53;
}

/* stratum uniforms as the CUDA library generates them: 32-bit, four strata per Philox block */
double orc_uniform_strata(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint32_t o[4];
    philox_at(seed, stream, idx >> 2, o);
    return ((double)o[idx & 3] + 0.5) * 0x1.0p-32;
}
void orc_fill_uniform_strata(uint64_t seed, uint64_t stream, int64_t n, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = orc_uniform_strata(seed, stream, (uint64_t)i);
}

void orc_fill_uniform53(uint64_t seed, uint64_t stream, int64_t n, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = orc_uniform53(seed, stream, (uint64_t)i);
}

/* ------------------------------------------------------------------ Base / Gen one-liners */
/* Julia Base.mapreduce_impl(identity, +, A, ifirst, ilast, 1024): sequential
 * below the block size, otherwise split at ifirst + (ilast-ifirst)>>1. */
static double sum_pw(const double *a, int64_t ifirst, int64_t ilast) {
    if (ifirst == ilast) return a[ifirst];
    if (ilast - ifirst < 1024) {
        double v = a[ifirst] + a[ifirst + 1];
        for (int64_t i = ifirst + 2; i <= ilast; ++i) v += a[i];
        return v;
    }
    int64_t imid = ifirst + ((ilast - ifirst) >> 1);
    return sum_pw(a, ifirst, imid) + sum_pw(a, imid + 1, ilast);
}
double orc_sum_pairwise(const double *a, int64_t n) { return n <= 0 ? 0.0 : sum_pw(a, 0, n - 1); }

double orc_maximum(const double *a, int64_t n) {
    /* Base.maximum propagates NaN */
    double m = -INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        if (isnan(a[i])) return NAN;
        if (a[i] > m) m = a[i];
    }
    return m;
}

/* Gen.logsumexp(arr): m = maximum(arr); m == -Inf ? -Inf : m + log(sum(exp.(arr .- m))) */
double orc_logsumexp(const double *v, int64_t n) {
    double m = orc_maximum(v, n);
    if (m == -INFINITY) return -INFINITY;
    double *e = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) e[i] = exp(v[i] - m);
    double s = orc_sum_pairwise(e, n);
    free(e);
    return m + log(s);
}

/* utils.jl:100  lognorm(vs) = vs .- logsumexp(vs) */
void orc_lognorm(const double *v, int64_t n, double *out) {
    double l = orc_logsumexp(v, n);
    for (int64_t i = 0; i < n; ++i) out[i] = v[i] - l;
}

/* utils.jl:103-107  ws = exp.(vs .- maximum(vs)); ws ./ sum(ws) */
void orc_softmax(const double *v, int64_t n, double *out) {
    if (n <= 0) return;
    double m = orc_maximum(v, n);
    for (int64_t i = 0; i < n; ++i) out[i] = exp(v[i] - m);
    double s = orc_sum_pairwise(out, n);
    for (int64_t i = 0; i < n; ++i) out[i] = out[i] / s;
}

/* utils.jl:117-140 */
int32_t orc_safe_softmax(const double *v, int64_t n, double *out) {
    int any_nan = 0, all_neginf = 1;
    for (int64_t i = 0; i < n; ++i) {
        if (isnan(v[i])) any_nan = 1;
        if (!(v[i] == -INFINITY)) all_neginf = 0;
    }
    if (any_nan) { /* :119-122 */
        for (int64_t i = 0; i < n; ++i) out[i] = NAN;
        return ORC_INV_NAN_INPUT;
    }
    if (all_neginf) { /* :123-126  ones ./ length */
        for (int64_t i = 0; i < n; ++i) out[i] = 1.0 / (double)n;
        return ORC_INV_ALL_NEGINF;
    }
    double m = orc_maximum(v, n);
    for (int64_t i = 0; i < n; ++i) out[i] = exp(v[i] - m); /* :128 */
    double total = orc_sum_pairwise(out, n);               /* :129 */
    if (total == 0.0) { /* :130-133 */
        for (int64_t i = 0; i < n; ++i) out[i] = 1.0 / (double)n;
        return ORC_INV_ZERO_TOTAL;
    }
    if (isnan(total)) { /* :134-137 */
        for (int64_t i = 0; i < n; ++i) out[i] = NAN;
        return ORC_INV_NAN_TOTAL;
    }
    double s = orc_sum_pairwise(out, n); /* :139 recomputes sum(ws) */
    for (int64_t i = 0; i < n; ++i) out[i] = out[i] / s;
    return ORC_VALID;
}

/* utils.jl:163-164 -> Gen.effective_sample_size(lnw) = exp(-logsumexp(2 .* lnw)) */
double orc_ess(const double *lw, int64_t n) {
    double *t = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    orc_lognorm(lw, n, t);
    for (int64_t i = 0; i < n; ++i) t[i] = 2.0 * t[i];
    double r = exp(-orc_logsumexp(t, n));
    free(t);
    return r;
}

/* utils.jl:174-178 / Gen.log_ml_estimate */
double orc_lml_estimate(double log_ml_est, const double *lw, int64_t n) {
    return log_ml_est + orc_logsumexp(lw, n) - log((double)n);
}

/* ------------------------------------------------------------------ sortperm(lp, rev=true) */
typedef struct {
    uint64_t t;
    int64_t idx;
} sort_item;
static inline uint64_t total_order_bits(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    /* monotone map fp64 -> u64 with -0.0 < +0.0 (Julia isless) */
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static int cmp_desc_stable(const void *pa, const void *pb) {
    const sort_item *a = (const sort_item *)pa, *b = (const sort_item *)pb;
    if (a->t != b->t) return a->t > b->t ? -1 : 1; /* larger key first */
    return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0); /* ties: ascending original index */
}
/* resample.jl:156-157; Julia's sortperm is stable, also under rev=true */
void orc_sortperm_desc(const double *keys, int64_t n, int64_t *order1) {
    sort_item *it = (sort_item *)malloc(sizeof(sort_item) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        it[i].t = total_order_bits(keys[i]);
        it[i].idx = i;
    }
    qsort(it, (size_t)n, sizeof(sort_item), cmp_desc_stable);
    for (int64_t i = 0; i < n; ++i) order1[i] = it[i].idx + 1;
    free(it);
}

/* ------------------------------------------------------------------ ancestor selection */
/* W_k = sequential left-to-right fp64 sum of w[order[1..k]] (resample.jl:163-166) */
void orc_cumweights(const double *w, const int64_t *order1, int64_t n, uint32_t flags, double *W) {
    if (flags & ORC_FLAG_EXACT_CUMSUM) {
        long double acc = 0.0L;
        for (int64_t k = 0; k < n; ++k) {
            acc += (long double)w[order1 ? order1[k] - 1 : k];
            W[k] = (double)acc;
        }
    } else {
        double acc = 0.0;
        for (int64_t k = 0; k < n; ++k) {
            acc += w[order1 ? order1[k] - 1 : k];
            W[k] = acc;
        }
    }
}

/* Multinomial oracle: inverse CDF with the tie rule of
 * Distributions' single-draw path used at resize.jl:284:
 *   i=1; cp=p[1]; while cp <= u && i < n: i+=1; cp+=p[i]   => smallest i with W_i > u, capped at n.
 * Output order = order of the supplied uniforms (resample.jl:59). */
void orc_select_multinomial(const double *w, int64_t n, const double *u, int64_t n_out, int64_t *parents1) {
    double *W = (double *)malloc(sizeof(double) * (size_t)n);
    orc_cumweights(w, NULL, n, 0, W);
    for (int64_t j = 0; j < n_out; ++j) {
        int64_t lo = 0, hi = n - 1; /* first k in [0,n-1] with W[k] > u, else n-1 */
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (W[mid] > u[j]) hi = mid; else lo = mid + 1;
        }
        parents1[j] = lo + 1;
    }
    free(W);
}

static inline double stratum_lower(int64_t i1, int64_t n) {
    /* element i of Julia's 0.0:1/n:1.0-1/n; exact for power-of-two n */
    return (double)(i1 - 1) / (double)n;
}

/* resample.jl:159-170, literal lazy-draw merge loop; r is dense, indexed by stratum.
 * Clamps: order[0] -> order[1]; running off the end stops at n. */
void orc_select_stratified(const double *w, const int64_t *order1, int64_t n, const double *r, uint32_t flags,
                           int64_t *parents1) {
    int64_t i_old = 0;
    const double weight_step = 1.0 / (double)n;
    double accum = 0.0;
    long double accum_x = 0.0L;
    const int exact = (flags & ORC_FLAG_EXACT_CUMSUM) != 0;
    for (int64_t i_new = 1; i_new <= n; ++i_new) {
        double lower = stratum_lower(i_new, n);
        if (lower + weight_step > accum) {
            double u = r[i_new - 1] * weight_step; /* two roundings, no FMA (resample.jl:162) */
            u = u + lower;
            while (accum < u && i_old < n) {
                double wi = w[order1 ? order1[i_old] - 1 : i_old];
                if (exact) { accum_x += (long double)wi; accum = (double)accum_x; }
                else accum += wi;
                i_old += 1;
            }
        }
        int64_t k = i_old < 1 ? 1 : i_old;
        parents1[i_new - 1] = order1 ? order1[k - 1] : k;
    }
}

/* parent_i = order[min{k : W_k >= u_i}] -- outcome-identical to the loop above */
void orc_select_stratified_search(const double *w, const int64_t *order1, int64_t n, const double *r, uint32_t flags,
                                  int64_t *parents1) {
    double *W = (double *)malloc(sizeof(double) * (size_t)n);
    orc_cumweights(w, order1, n, flags, W);
    const double weight_step = 1.0 / (double)n;
    for (int64_t i = 1; i <= n; ++i) {
        double u = r[i - 1] * weight_step;
        u = u + stratum_lower(i, n);
        int64_t lo = 0, hi = n - 1;
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (W[mid] >= u) hi = mid; else lo = mid + 1;
        }
        parents1[i - 1] = order1 ? order1[lo] : lo + 1;
    }
    free(W);
}

/* resample.jl:96-115 / resize.jl:100-119.  Draw for output slot j uses u[j]. */
void orc_select_residual(const double *w, int64_t n, const double *u, int64_t n_out, int64_t *parents1,
                         int64_t *n_deterministic) {
    int64_t n_resampled = 0;
    for (int64_t i = 0; i < n; ++i) {
        double c = floor((double)n_out * w[i]); /* floor(Int, n_particles * w) :99 */
        int64_t n_copies = (int64_t)c;
        if (n_copies <= 0) continue;
        if (n_resampled + n_copies > n_out) n_copies = n_out - n_resampled; /* clamp (App. C) */
        for (int64_t j = 0; j < n_copies; ++j) parents1[n_resampled + j] = i + 1;
        n_resampled += n_copies;
    }
    if (n_deterministic) *n_deterministic = n_resampled;
    if (n_resampled < n_out) { /* :108-115 */
        double *rw = (double *)malloc(sizeof(double) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            double nw = (double)n_out * w[i];
            rw[i] = nw - floor(nw);
        }
        double s = orc_sum_pairwise(rw, n);
        for (int64_t i = 0; i < n; ++i) rw[i] = rw[i] / s;
        orc_select_multinomial(rw, n, u + n_resampled, n_out - n_resampled, parents1 + n_resampled);
        free(rw);
    }
}

/* pf_{multinomial,residual,stratified}_resample! and the two resize variants,
 * restricted to the plain-bits members of the state (SURVEY 8b cut line).
 *   prologue  resample.jl:51-57 / 88-94 / 147-153 ; resize.jl:48-56 / 89-96
 *   epilogue  update_weights! resample.jl:190-218 ; resize.jl:424-438
 * lml_increment = logsumexp(lw) - log(n_in)  (update_lml_est!, resample.jl:178-182; 0 for substates :184-187)
 * NaN weights: the reference crashes later (App. C); we return the kind and leave outputs untouched. */
int32_t orc_resample(int32_t method, const double *lw, const double *lp, int64_t n_in, int64_t n_out,
                     const double *uniforms, uint32_t flags, int64_t *parents1, double *lw_out,
                     double *lml_increment, int32_t *invalid_kind) {
    if (n_in <= 0 || n_out <= 0) return -1;
    if (method == ORC_STRATIFIED && n_out != n_in) return -1; /* no stratified resize, resize.jl:16-27 */
    if ((flags & ORC_FLAG_SUBSTATE) && n_out != n_in) return -1;
    const double *prio = lp ? lp : lw;
    double *w = (double *)malloc(sizeof(double) * (size_t)n_in);
    int32_t kind = orc_safe_softmax(prio, n_in, w);
    if (invalid_kind) *invalid_kind = kind;
    double lse_lw = orc_logsumexp(lw, n_in);
    if (lml_increment) *lml_increment = (flags & ORC_FLAG_SUBSTATE) ? 0.0 : lse_lw - log((double)n_in);
    if (kind == ORC_INV_NAN_INPUT || kind == ORC_INV_NAN_TOTAL) {
        free(w);
        return 0;
    }
    if (method == ORC_MULTINOMIAL) {
        orc_select_multinomial(w, n_in, uniforms, n_out, parents1);
    } else if (method == ORC_RESIDUAL) {
        orc_select_residual(w, n_in, uniforms, n_out, parents1, NULL);
    } else if (method == ORC_STRATIFIED) {
        int64_t *order = NULL;
        if (flags & ORC_FLAG_SORT) {
            order = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_in);
            orc_sortperm_desc(prio, n_in, order);
        }
        orc_select_stratified(w, order, n_in, uniforms, flags, parents1);
        free(order);
    } else {
        free(w);
        return -4;
    }
    free(w);
    /* update_weights! */
    if (!lp) {
        double v = (flags & ORC_FLAG_SUBSTATE) ? lse_lw - log((double)n_in) : 0.0; /* :193-195 / :208-210 */
        for (int64_t j = 0; j < n_out; ++j) lw_out[j] = v;
    } else {
        double *d = (double *)malloc(sizeof(double) * (size_t)n_out);
        for (int64_t j = 0; j < n_out; ++j) d[j] = lw[parents1[j] - 1] - lp[parents1[j] - 1]; /* :197 / :212 */
        double lse_d = orc_logsumexp(d, n_out);
        double shift = (flags & ORC_FLAG_SUBSTATE) ? (lse_lw - lse_d)              /* :215-216 */
                                                   : (log((double)n_out) - lse_d); /* :201, resize.jl:436 */
        for (int64_t j = 0; j < n_out; ++j) lw_out[j] = d[j] + shift;
        free(d);
    }
    return 0;
}

/* ------------------------------------------------------------------ statistics.jl:13-17,48-54 */
void orc_mean_var(const double *lw, const double *x, int64_t n, double *mean, double *var) {
    double *w = (double *)malloc(sizeof(double) * (size_t)n);
    double *t = (double *)malloc(sizeof(double) * (size_t)n);
    orc_softmax(lw, n, w); /* get_norm_weights: plain softmax, not safe_softmax */
    for (int64_t i = 0; i < n; ++i) t[i] = w[i] * x[i];
    double mu = orc_sum_pairwise(t, n);
    for (int64_t i = 0; i < n; ++i) {
        double d = x[i] - mu;
        t[i] = w[i] * (d * d);
    }
    double v = orc_sum_pairwise(t, n);
    if (mean) *mean = mu;
    if (var) *var = v;
    free(w);
    free(t);
}

/* ------------------------------------------------------------------ resize.jl */
/* pf_replicate! resize.jl:236-244: repeat(x; inner=k) | repeat(x, k); weights copied unchanged */
void orc_replicate(const double *lw, int64_t n, int64_t k, int32_t interleaved, int64_t *parents1, double *lw_out) {
    for (int64_t j = 0; j < n * k; ++j) {
        int64_t src = interleaved ? (j % n) : (j / k);
        parents1[j] = src + 1;
        lw_out[j] = lw[src];
    }
}

/* pf_dereplicate! resize.jl:267-297 */
void orc_dereplicate(const double *lw, int64_t n, int64_t k, int32_t interleaved, int32_t sample, const double *u,
                     int64_t *parents1, double *lw_out) {
    int64_t n_new = n / k;
    if (!sample) { /* :272-276 keepfirst: idxs = 1:k:n | 1:n_new */
        for (int64_t b = 0; b < n_new; ++b) {
            int64_t src = interleaved ? b : b * k;
            parents1[b] = src + 1;
            lw_out[b] = lw[src];
        }
        return;
    }
    double *blk = (double *)malloc(sizeof(double) * (size_t)k);
    double *wb = (double *)malloc(sizeof(double) * (size_t)k);
    for (int64_t b = 0; b < n_new; ++b) { /* :278-291 */
        for (int64_t j = 0; j < k; ++j) blk[j] = lw[interleaved ? b + j * n_new : b * k + j];
        orc_softmax(blk, k, wb);
        /* rand(Categorical(weights)): i=1; cp=p[1]; while cp <= u && i < k: i++; cp += p[i] */
        int64_t i = 0;
        double cp = wb[0];
        while (cp <= u[b] && i < k - 1) {
            i += 1;
            cp += wb[i];
        }
        parents1[b] = (interleaved ? b + i * n_new : b * k + i) + 1;
        lw_out[b] = orc_logsumexp(blk, k) - log((double)k);
    }
    free(blk);
    free(wb);
}

/* pf_coalesce! resize.jl:309-334 with integer keys standing in for by(trace).
 * Un-shifted exp(w) accumulation as in the reference (:318).  Output order:
 * first-occurrence order (the reference's Dict order is unspecified; compare as sets). */
int64_t orc_coalesce(const double *lw, const int64_t *keys, int64_t n, int64_t *parents1, double *lw_out) {
    if (n <= 0) return 0;
    double *acc = (double *)calloc((size_t)n, sizeof(double));
    int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    sort_item *it = (sort_item *)malloc(sizeof(sort_item) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        it[i].t = ~((uint64_t)keys[i] ^ 0x8000000000000000ull); /* ascending keys under cmp_desc_stable */
        it[i].idx = i;
    }
    qsort(it, (size_t)n, sizeof(sort_item), cmp_desc_stable);
    for (int64_t s = 0; s < n;) { /* groups of equal key; first index = smallest idx in the group */
        int64_t e = s;
        while (e < n && it[e].t == it[s].t) ++e;
        for (int64_t j = s; j < e; ++j) first[it[j].idx] = it[s].idx;
        s = e;
    }
    for (int64_t i = 0; i < n; ++i) acc[first[i]] += exp(lw[i]); /* particle order, like the reference loop */
    int64_t n_new = 0;
    for (int64_t i = 0; i < n; ++i)
        if (first[i] == i) n_new++;
    double log_n_ratio = log((double)n_new) - log((double)n);
    int64_t o = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (first[i] != i) continue;
        parents1[o] = i + 1;
        lw_out[o] = log(acc[i]) + log_n_ratio; /* :327 */
        o++;
    }
    free(acc);
    free(first);
    free(it);
    return n_new;
}

/* ------------------------------------------------------------------ models */
/* Gen: logpdf(normal, x, mu, sigma) = -(((x-mu)/sigma)^2 + log(2pi))/2 - log(sigma) */
double orc_normal_logpdf(double x, double mu, double sigma) {
    double z = (x - mu) / sigma;
    return -(z * z + log(2.0 * M_PI)) / 2.0 - log(sigma);
}
/* ---- pf_optimal_resize! (resize.jl:149-196) + find_inv_w_threshold (resize.jl:199-216), literal ---- */
static int cmp_double_asc(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
static double julia_eps(double x) { /* eps(x): distance to the next larger float of |x| */
    x = fabs(x);
    return nextafter(x, INFINITY) - x;
}
double orc_find_inv_w_threshold(const double *w, int64_t n, int64_t n_particles) {
    double *s = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    memcpy(s, w, sizeof(double) * (size_t)n);
    qsort(s, (size_t)n, sizeof(double), cmp_double_asc); /* :200 */
    int64_t A = n;                                       /* :202 */
    double B = 0.0;                                      /* :203 */
    for (int64_t r = 0; r < n; ++r) {
        double kappa = s[r];
        A -= 1;
        B += kappa;
        double n_check = B / kappa + (double)A;                 /* :208 */
        if (n_check <= (double)n_particles + julia_eps(n_check)) { /* :209 */
            double c = ((double)n_particles - (double)A) / B;   /* :211 */
            free(s);
            return c;
        }
    }
    free(s);
    return (double)n_particles; /* :214 */
}
/* returns 0, or -3 when the loop selects a number of particles different from n_resample (the reference's
 * @assert at resize.jl:181 would throw); *n_selected reports what the loop produced */
int32_t orc_optimal_resize(const double *lw, int64_t n, int64_t n_out, double u_rand, int64_t *parents1,
                           double *lw_out, int64_t *n_keep_out, double *inv_w_out, int32_t *invalid_kind,
                           int32_t *invalid_kind_strat, int64_t *n_selected) {
    double *w = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int32_t kind = orc_safe_softmax(lw, n, w); /* :152 */
    if (invalid_kind) *invalid_kind = kind;
    double c = orc_find_inv_w_threshold(w, n, n_out); /* :155 */
    if (inv_w_out) *inv_w_out = c;
    int64_t *strat = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    double *slw = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *sw = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int64_t n_keep = 0, n_strat = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (c * w[i] >= 1.0) parents1[n_keep++] = i + 1; /* :156,159,177 */
        else { strat[n_strat] = i + 1; slw[n_strat++] = lw[i]; }
    }
    if (n_keep_out) *n_keep_out = n_keep;
    int64_t n_res = n_out - n_keep; /* :162 */
    int32_t kind2 = n_strat > 0 ? orc_safe_softmax(slw, n_strat, sw) : 0; /* :166-167 */
    if (invalid_kind_strat) *invalid_kind_strat = kind2;
    double step = 1.0 / (double)n_res; /* :170 */
    double u = u_rand * step;          /* :171 */
    int64_t k = 0;
    for (int64_t i = 0; i < n_strat; ++i) { /* :172-178 */
        u = u - sw[i];
        if (u < 0) {
            if (n_keep + k < n_out) parents1[n_keep + k] = strat[i];
            ++k;
            u += step;
        }
    }
    if (n_selected) *n_selected = k;
    double log_n_ratio = log((double)n_out) - log((double)n); /* :187 */
    double log_tot = orc_logsumexp(lw, n);                    /* :188 */
    double res_lw = log_tot - log(c);                         /* :190 */
    for (int64_t j = 0; j < n_keep && j < n_out; ++j) lw_out[j] = lw[parents1[j] - 1] + log_n_ratio; /* :192 */
    for (int64_t j = n_keep; j < n_out; ++j) lw_out[j] = res_lw + log_n_ratio;                        /* :193 */
    free(w); free(strat); free(slw); free(sw);
    return k == n_res ? 0 : -3;
}



/* README.md:47-49: moving ~ bernoulli(moving ? .75 : .25) [rand() < p]; y ~ normal(y + vel, 0.01) [mu + sigma*randn()] */
void orc_om_transition(const orc_om_params *p, int64_t n, const double *y_prev, const uint8_t *m_prev, double vel,
                       const double *U, const double *Z, double *y_out, uint8_t *m_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t mp = m_prev ? m_prev[i] : 0;
        double yp = y_prev ? y_prev[i] : 0.0;
        uint8_t m = U[i] < (mp ? p->p_stay : p->p_start);
        double mu = yp + (m ? vel : 0.0);
        double sz = p->sigma_proc * Z[i];
        y_out[i] = mu + sz;
        m_out[i] = m;
    }
}

/* README.md:50 observation y_obs ~ normal(y, 0.25); initialize.jl:39-41 assigns, update.jl:21 accumulates */
void orc_om_obs_logpdf_add(const orc_om_params *p, int64_t n, const double *y, double obs, double *lw, int32_t assign) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double l = orc_normal_logpdf(obs, y[i], p->sigma_obs);
        lw[i] = assign ? l : lw[i] + l;
    }
}

/* Gen mh(trace, select(tau=>:moving, tau=>:y)) while tau is the last step of the trace
 * (rejuvenate.jl:40-53, README.md:72-73): regenerate from the prior given slice tau-1,
 * weight = obs logpdf ratio, accept iff log(rand()) < weight. */
void orc_om_mh(const orc_om_params *p, int64_t n, const double *y_pp, const uint8_t *m_pp, double *y_cur,
               uint8_t *m_cur, double vel, double obs, const double *U2, const double *Z2, const double *U3,
               uint8_t *accept) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t mp = m_pp ? m_pp[i] : 0;
        double yp = y_pp ? y_pp[i] : 0.0;
        uint8_t m = U2[i] < (mp ? p->p_stay : p->p_start);
        double mu = yp + (m ? vel : 0.0);
        double sz = p->sigma_proc * Z2[i];
        double y = mu + sz;
        double alpha = orc_normal_logpdf(obs, y, p->sigma_obs) - orc_normal_logpdf(obs, y_cur[i], p->sigma_obs);
        uint8_t acc = log(U3[i]) < alpha;
        if (acc) {
            y_cur[i] = y;
            m_cur[i] = m;
        }
        if (accept) accept[i] = acc;
    }
}

/* 1-D linear-Gaussian tracker: x_t ~ N(a x_{t-1}, q), y_t ~ N(x_t, r) */
void orc_lg_transition(const orc_lg_params *p, int64_t n, const double *x_prev, const double *Z, double *x_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double mu = p->a * x_prev[i];
        double sz = p->q * Z[i];
        x_out[i] = mu + sz;
    }
}
void orc_lg_obs_logpdf_add(const orc_lg_params *p, int64_t n, const double *x, double obs, double *lw, int32_t assign) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double l = orc_normal_logpdf(obs, x[i], p->r);
        lw[i] = assign ? l : lw[i] + l;
    }
}
void orc_lg_mh(const orc_lg_params *p, int64_t n, const double *x_pp, double *x_cur, double obs, const double *Z2,
               const double *U3, uint8_t *accept) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double mu = p->a * x_pp[i];
        double sz = p->q * Z2[i];
        double x = mu + sz;
        double alpha = orc_normal_logpdf(obs, x, p->r) - orc_normal_logpdf(obs, x_cur[i], p->r);
        uint8_t acc = log(U3[i]) < alpha;
        if (acc) x_cur[i] = x;
        if (accept) accept[i] = acc;
    }
}

/* ------------------------------------------------------------------ CPU baseline filter (bench only) */
/* Production-noise convention shared with the CUDA library ("lean" Philox noise, models.cuh NoiseLean):
 * ONE Philox call per particle per README iteration serves the mh move and the update:
 * w0[31:8] U_mh, w1[31:8] U_up, w2 U_acc, w3[31:8] Box-Muller angle, {w0,w1,w3}[7:0] radius uniform;
 * Z_mh = r cos(theta), Z_up = r sin(theta). */
static inline void lean_noise(uint64_t seed, uint64_t stream, uint64_t idx, double *U_mh, double *Z_mh, double *U_acc,
                              double *U_up, double *Z_up) {
    uint32_t o[4];
    philox_at(seed, stream, idx, o);
    *U_mh = ((double)(o[0] >> 8) + 0.5) * 0x1.0p-24;
    *U_up = ((double)(o[1] >> 8) + 0.5) * 0x1.0p-24;
    *U_acc = ((double)o[2] + 0.5) * 0x1.0p-32;
    uint32_t rb = ((o[0] & 0xFFu) << 16) | ((o[1] & 0xFFu) << 8) | (o[3] & 0xFFu);
    float ua = ((float)rb + 0.5f) * 0x1.0p-24f;
    float th = (((float)(o[3] >> 8) + 0.5f) * 0x1.0p-24f - 0.5f) * 6.28318530717958647692f;
    float rr = sqrtf(-2.0f * logf(ua));
    *Z_mh = (double)(rr * cosf(th));
    *Z_up = (double)(rr * sinf(th));
}
#define ORC_STREAM(purpose, step) (((uint64_t)(purpose) << 56) | ((uint64_t)(step) & 0x00FFFFFFFFFFFFFFull))

orc_om_filter *orc_om_filter_create(int64_t n, uint64_t seed) {
    orc_om_filter *f = (orc_om_filter *)calloc(1, sizeof(orc_om_filter));
    f->n = n;
    f->seed = seed;
    for (int s = 0; s < 2; ++s) {
        f->y[s] = (double *)calloc((size_t)n, sizeof(double));
        f->m[s] = (uint8_t *)calloc((size_t)n, 1);
        f->y_new[s] = (double *)calloc((size_t)n, sizeof(double));
        f->m_new[s] = (uint8_t *)calloc((size_t)n, 1);
    }
    f->lw = (double *)calloc((size_t)n, sizeof(double));
    f->w = (double *)calloc((size_t)n, sizeof(double));
    f->r = (double *)calloc((size_t)n, sizeof(double));
    f->parents = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    return f;
}
void orc_om_filter_destroy(orc_om_filter *f) {
    if (!f) return;
    for (int s = 0; s < 2; ++s) {
        free(f->y[s]); free(f->m[s]); free(f->y_new[s]); free(f->m_new[s]);
    }
    free(f->lw); free(f->w); free(f->r); free(f->parents);
    free(f);
}

/* pf_initialize, initialize.jl:31-44 on object_motion(1): slice 0 = (false, 0.0), README.md:44 */
void orc_om_filter_init(orc_om_filter *f, const orc_om_params *p, double vel1, double obs1) {
    const int64_t n = f->n;
    f->cur = 1;
    f->log_ml_est = 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double U, Z, a_, b_, c_;
        lean_noise(f->seed, ORC_STREAM(2, 1), (uint64_t)i, &a_, &b_, &c_, &U, &Z);
        f->y[0][i] = 0.0;
        f->m[0][i] = 0;
        uint8_t m = U < p->p_start;
        double sz = p->sigma_proc * Z;
        f->y[1][i] = (0.0 + (m ? vel1 : 0.0)) + sz;
        f->m[1][i] = m;
        f->lw[i] = orc_normal_logpdf(obs1, f->y[1][i], p->sigma_obs);
        f->parents[i] = i + 1;
    }
}

/* One README loop iteration (README.md:66-77) with the resample forced:
 * ESS (utils.jl:163) -> pf_resample!(:stratified, sort_particles=false) -> pf_rejuvenate!(mh) -> pf_update! */
double orc_om_filter_step(orc_om_filter *f, const orc_om_params *p, int64_t t, double vel_prev, double obs_prev,
                          double vel_t, double obs_t) {
    const int64_t n = f->n;
    const int cur = f->cur, prv = cur ^ 1;
    /* ESS + normalised weights (safe_softmax) */
    double m = -INFINITY;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (int64_t i = 0; i < n; ++i)
        if (f->lw[i] > m) m = f->lw[i];
    double s = 0.0, s2 = 0.0;
#pragma omp parallel for reduction(+ : s, s2) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double e = exp(f->lw[i] - m);
        f->w[i] = e;
        s += e;
        s2 += e * e;
    }
    double ess = s * s / s2;
    f->log_ml_est += (m + log(s)) - log((double)n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        f->w[i] = f->w[i] / s;
        f->r[i] = orc_uniform_strata(f->seed, ORC_STREAM(1, t), (uint64_t)i);
    }
    /* stratified merge loop (sequential, resample.jl:159-170) */
    orc_select_stratified(f->w, NULL, n, f->r, 0, f->parents);
    /* gather both window slices, lw = 0, then MH(t-1) and update(t) per particle */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int64_t a = f->parents[i] - 1;
        double y_pp = f->y[prv][a], y_c = f->y[cur][a];
        uint8_t m_pp = f->m[prv][a], m_c = f->m[cur][a];
        double U2, Z2, U3, U1, Z1;
        lean_noise(f->seed, ORC_STREAM(2, t), (uint64_t)i, &U2, &Z2, &U3, &U1, &Z1);
        uint8_t mq = U2 < (m_pp ? p->p_stay : p->p_start);
        double sz = p->sigma_proc * Z2;
        double yq = (y_pp + (mq ? vel_prev : 0.0)) + sz;
        double alpha = orc_normal_logpdf(obs_prev, yq, p->sigma_obs) - orc_normal_logpdf(obs_prev, y_c, p->sigma_obs);
        if (log(U3) < alpha) { y_c = yq; m_c = mq; }
        uint8_t mn = U1 < (m_c ? p->p_stay : p->p_start);
        double sz1 = p->sigma_proc * Z1;
        double yn = (y_c + (mn ? vel_t : 0.0)) + sz1;
        /* new window: (t-1 -> slot cur stays), t -> slot prv */
        f->y_new[cur][i] = y_c;  f->m_new[cur][i] = m_c;
        f->y_new[prv][i] = yn;   f->m_new[prv][i] = mn;
        f->lw[i] = 0.0 + orc_normal_logpdf(obs_t, yn, p->sigma_obs);
    }
    for (int s2i = 0; s2i < 2; ++s2i) { /* update_refs!: swap (utils.jl:10-15) */
        double *ty = f->y[s2i]; f->y[s2i] = f->y_new[s2i]; f->y_new[s2i] = ty;
        uint8_t *tm = f->m[s2i]; f->m[s2i] = f->m_new[s2i]; f->m_new[s2i] = tm;
    }
    f->cur = prv;
    return ess;
}

int32_t orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py's reference arm sets the thread count explicitly */
void orc_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
