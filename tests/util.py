"""Shared helpers for the GPU parity tests."""
import numpy as np

EPS = np.finfo(np.float64).eps


def tie_tolerance(n):
    """Documented fp64 cumulative-sum tie window: |W_k - u_i| <= 8*eps*sqrt(n)."""
    return 8.0 * EPS * np.sqrt(float(n))


def check_parents(p_gpu, p_ref, W_ref, u, order=None, max_frac=1e-4, p_exact=None):
    """Ancestors must be bit-exact except at documented cumulative-sum ties.

    A mismatch at output i is accepted only if every cumulative-weight boundary separating the two
    choices lies within tie_tolerance(n) of the threshold u_i.  The literal oracle sums sequentially in fp64, so
    its OWN cumulative weights drift by a random walk of ~eps*sqrt(k) (8.2*eps*sqrt(n) observed at 2^24); when
    `p_exact` (the oracle with a long-double cumulative sum, SURVEY 8c "exact mode") is given, a mismatch up to
    4x the window is also accepted provided the GPU agrees with the exact-mode ancestor there.
    Returns (n_mismatch, max_gap)."""
    n = W_ref.size
    mism = np.flatnonzero(p_gpu != p_ref)
    if mism.size == 0:
        return 0, 0.0
    assert mism.size <= max(8, int(max_frac * p_ref.size)), f"{mism.size} ancestor mismatches of {p_ref.size}"
    if p_exact is not None:
        # the parallel fp64 scan agrees with the long-double cumulative sum almost everywhere (SURVEY 8c "exact mode")
        n_ex = int(np.sum(p_gpu != p_exact))
        assert n_ex <= max(4, int(2e-6 * p_ref.size)), f"{n_ex} ancestors differ from the exact-mode oracle"
    if order is not None:
        inv = np.empty(n, dtype=np.int64)
        inv[order] = np.arange(n)
        a, b = inv[p_gpu[mism]], inv[p_ref[mism]]
    else:
        a, b = p_gpu[mism], p_ref[mism]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    tol = tie_tolerance(n)
    max_gap = 0.0
    for i, l, h in zip(mism, lo, hi):
        gap = np.max(np.abs(W_ref[l:h] - u[i]))
        max_gap = max(max_gap, gap)
        ok = gap <= tol or (p_exact is not None and p_gpu[i] == p_exact[i] and gap <= 4 * tol)
        assert ok, f"ancestor mismatch at output {i} is not a cumulative-sum tie (gap {gap:.3e} > {tol:.3e})"
    return mism.size, max_gap


def strat_u(r, n):
    """u_i = r_i*(1/n) + (i-1)/n, two roundings (resample.jl:160-162)."""
    step = 1.0 / n
    return r * step + np.arange(n) / n


def weights(rng, n, kind):
    if kind == "A":  # N(0,1)
        return rng.normal(0.0, 1.0, n)
    if kind == "B":  # N(0,5^2): heavy degeneracy
        return rng.normal(0.0, 5.0, n)
    if kind == "C":  # all equal
        return np.full(n, -3.5)
    raise ValueError(kind)


def oracle_readme_step(orc, st, t, obs_prev, obs_t, r, U2, Z2, U3, U1, Z1, p_gpu=None, mh=True):
    """One README iteration (README.md:66-77, resample forced, stratified with sort_particles=false) on the CPU
    oracle with every draw supplied: resample.jl:143-175 -> rejuvenate.jl:40-53 (mh on slice t-1) -> update.jl:12-25.
    `st` = dict(y_pp, m_pp, y, m, lw) holding slices t-2 (None for t == 2), t-1 and the log-weights; returns the
    new dict plus (p_ref, n_tie) -- ancestors are tie-checked against `p_gpu` when given, and the GPU's choice is
    then followed so the populations stay aligned (ties are documented)."""
    import math
    n = st["lw"].size
    p_ref, lw0, inc, kind = orc.resample("stratified", st["lw"], r)
    assert kind == 0 and np.all(lw0 == 0.0)
    n_tie = 0
    p = p_ref
    if p_gpu is not None:
        if not np.array_equal(p_gpu, p_ref):
            W_ref = orc.cumweights(orc.softmax(st["lw"]))
            p_exact = orc.resample("stratified", st["lw"], r, exact=True)[0]
            n_tie, _ = check_parents(p_gpu, p_ref, W_ref, strat_u(r, n), p_exact=p_exact)
        p = p_gpu
    y_pp = None if st["y_pp"] is None else st["y_pp"][p]
    m_pp = None if st["m_pp"] is None else st["m_pp"][p]
    y, m = st["y"][p], st["m"][p]
    acc = np.zeros(n, dtype=bool)
    if mh:
        y, m, acc = orc.om_mh(y_pp, m_pp, y, m, math.sin(t - 1.0), obs_prev, U2, Z2, U3)
    y_new, m_new = orc.om_transition(y, m, math.sin(float(t)), U1, Z1)
    lw = orc.om_obs_logpdf(y_new, obs_t, lw0)
    return dict(y_pp=y, m_pp=m, y=y_new, m=m_new, lw=lw), p_ref, n_tie, inc, acc
