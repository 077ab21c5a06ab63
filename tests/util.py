"""Shared helpers for the GPU parity tests."""
import numpy as np

EPS = np.finfo(np.float64).eps


def tie_tolerance(n):
    """Documented fp64 cumulative-sum tie window: |W_k - u_i| <= 8*eps*sqrt(n)."""
    return 8.0 * EPS * np.sqrt(float(n))


def check_parents(p_gpu, p_ref, W_ref, u, order=None, max_frac=1e-4):
    """Ancestors must be bit-exact except at documented cumulative-sum ties.

    A mismatch at output i is accepted only if every cumulative-weight boundary separating the two
    choices lies within tie_tolerance(n) of the threshold u_i.  Returns (n_mismatch, max_gap)."""
    n = W_ref.size
    mism = np.flatnonzero(p_gpu != p_ref)
    if mism.size == 0:
        return 0, 0.0
    assert mism.size <= max(8, int(max_frac * p_ref.size)), f"{mism.size} ancestor mismatches of {p_ref.size}"
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
        assert gap <= tol, f"ancestor mismatch at output {i} is not a cumulative-sum tie (gap {gap:.3e} > {tol:.3e})"
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
