"""GPU parity tests of the host-array C ABI (genpf_logsumexp/ess/normalize/resample/...) against the
CPU oracle and the golden fixtures.  Bit-exact for indices except documented fp64 cumulative-sum ties
(tests/util.py); 1e-10 relative for fp64 scalars and weights (BASELINE.json north_star)."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import GOLDEN_CASES
from util import check_parents, strat_u, tie_tolerance, weights

pytestmark = pytest.mark.gpu
RTOL = 1e-10  # north_star: log-weights, ESS, mean/var within 1e-10 relative in fp64


def raw_resample(g, method, lw, u=None, lp=None, n_out=None, flags=0, seed=0):
    L = g._lib
    lib = g.load()
    lw = np.ascontiguousarray(lw, dtype=np.float64)
    n_in = lw.size
    n_out = n_in if n_out is None else n_out
    parents = np.full(n_out, -7, dtype=np.int64)
    lw_out = np.full(n_out, np.nan)
    inc, kind = C.c_double(np.nan), C.c_int32(-1)
    st = lib.genpf_resample(L.METHODS[method], L.ptr(lw), L.ptr(lp), n_in, n_out, L.ptr(u), seed, flags,
                            L.ptr(parents), L.ptr(lw_out), C.byref(inc), C.byref(kind))
    return st, parents, lw_out, inc.value, kind.value


def gpu_cumweights(g, lw):
    W = np.empty(lw.size)
    g._lib.check(g.load().genpf_debug_cumweights(g._lib.ptr(lw), lw.size, 0, g._lib.ptr(W)))
    return W


SIZES = [1, 2, 3, 100, 2047, 2048, 2049, 4096, 100_003, (1 << 20) + 3]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("kind", ["A", "B", "C"])
def test_lse_ess_normalize(g, orc, n, kind):
    lw = weights(np.random.default_rng(n), n, kind)
    assert g.logsumexp_host(lw) == pytest.approx(orc.logsumexp(lw), rel=RTOL, abs=1e-12)

    class S:
        log_weights = lw
    assert g.effective_sample_size(S) == pytest.approx(orc.ess(lw), rel=RTOL)
    np.testing.assert_allclose(g.get_log_norm_weights(S), orc.lognorm(lw), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(g.get_norm_weights(S), orc.softmax(lw), rtol=RTOL, atol=1e-300)
    assert abs(g.get_norm_weights(S).sum() - 1) < 1e-12  # test/utils.jl:5-8


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_scalars(g, golden, case):
    lw = golden[f"{case}/lw"]
    assert g.logsumexp_host(lw) == pytest.approx(float(golden[f"{case}/lse"]), rel=RTOL)
    lib, L = g.load(), g._lib
    m, v = C.c_double(), C.c_double()
    x = golden[f"{case}/x"]
    L.check(lib.genpf_weighted_mean_var(L.ptr(lw), L.ptr(x), lw.size, 0, C.byref(m), C.byref(v)))
    gm, gv = golden[f"{case}/mean_var"]
    assert m.value == pytest.approx(gm, rel=RTOL, abs=1e-14) and v.value == pytest.approx(gv, rel=RTOL)


def test_invalid_kinds(g):
    lib, L = g.load(), g._lib
    for arr, expect in [([0.0, np.nan, 1.0], 1), ([-np.inf] * 5, 2), ([0.0, np.inf], 4), ([0.0, -np.inf], 0)]:
        lw = np.array(arr)
        kind = C.c_int32(-1)
        L.check(lib.genpf_normalize(L.ptr(lw), lw.size, 0, None, None, None, None, C.byref(kind)))
        assert kind.value == expect
    assert math.isnan(g.logsumexp_host(np.array([0.0, np.nan])))
    assert g.logsumexp_host(np.full(7, -np.inf)) == -np.inf
    st = lib.genpf_logsumexp(None, 0, 0, C.byref(C.c_double()))
    assert st == L.ERR_INVALID_ARG  # empty input is an error status, not a crash (App. C)


def test_philox_uniforms_bit_exact(g, orc):
    lib, L = g.load(), g._lib
    for seed, stream, n in [(0, 0, 1000), (12345, 7, 4099), (2**63 + 5, (1 << 56) | 3, 2048)]:
        out = np.empty(n)
        L.check(lib.genpf_uniforms(seed, stream, n, 0, L.ptr(out)))
        np.testing.assert_array_equal(out, orc.uniforms(seed, stream, n))
        L.check(lib.genpf_uniforms(seed, stream, n, L.UNIFORMS_STRATA, L.ptr(out)))
        np.testing.assert_array_equal(out, orc.uniforms_strata(seed, stream, n))
        assert out.min() > 0.0 and out.max() < 1.0


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_resample(g, golden, case):
    """Golden ancestors (independent restatement) are reproduced bit-for-bit at these sizes."""
    lw, u = golden[f"{case}/lw"], golden[f"{case}/u"]
    n = lw.size
    st, p, lw_out, inc, kind = raw_resample(g, "stratified", lw, u)
    assert st == 0 and kind == 0
    np.testing.assert_array_equal(p, golden[f"{case}/strat/parents"])
    assert np.all(lw_out == 0.0)
    assert inc == pytest.approx(float(golden[f"{case}/lse"]) - math.log(n), rel=RTOL)
    st, p, *_ = raw_resample(g, "stratified", lw, u, flags=g._lib.SORT_PARTICLES)
    np.testing.assert_array_equal(p, golden[f"{case}/strat_sorted/parents"])
    lp = lw / 2
    st, p, lw_out, _, _ = raw_resample(g, "stratified", lw, u, lp=lp)
    np.testing.assert_array_equal(p, golden[f"{case}/strat_prio/parents"])
    np.testing.assert_allclose(lw_out, golden[f"{case}/strat_prio/lw_out"], rtol=RTOL, atol=1e-12)
    st, p, lw_out, inc, _ = raw_resample(g, "stratified", lw, u, lp=lp, flags=g._lib.SUBSTATE)
    np.testing.assert_allclose(lw_out, golden[f"{case}/strat_prio/lw_out_sub"], rtol=RTOL, atol=1e-12)
    assert inc == 0.0
    st, p, lw_out, _, _ = raw_resample(g, "stratified", lw, u, flags=g._lib.SUBSTATE)
    np.testing.assert_allclose(lw_out, golden[f"{case}/lw_out_sub"], rtol=RTOL, atol=1e-12)
    for n_out in (n, n // 2, n + n // 2):
        us = golden[f"{case}/u_{n_out}"]
        st, p, lw_out, inc, _ = raw_resample(g, "multinomial", lw, us, n_out=n_out)
        assert st == 0
        np.testing.assert_array_equal(p, golden[f"{case}/multi_{n_out}/parents"])
        assert np.all(lw_out == 0.0)
        st, p, *_ = raw_resample(g, "residual", lw, us, n_out=n_out)
        np.testing.assert_array_equal(p, golden[f"{case}/resid_{n_out}/parents"])
    # 1-based indices for Julia
    st, p1, *_ = raw_resample(g, "stratified", lw, u, flags=g._lib.INDEX_BASE1)
    np.testing.assert_array_equal(p1, golden[f"{case}/strat/parents"] + 1)


@pytest.mark.parametrize("n", [1 << 16, 100_003, 1 << 20, 1 << 22])
@pytest.mark.parametrize("kind", ["A", "B", "C"])
@pytest.mark.parametrize("sort", [False, True])
def test_stratified_vs_oracle(g, orc, n, kind, sort):
    rng = np.random.default_rng(n + 17)
    lw = weights(rng, n, kind)
    r = rng.random(n)
    flags = g._lib.SORT_PARTICLES if sort else 0
    st, p, lw_out, inc, k = raw_resample(g, "stratified", lw, r, flags=flags)
    assert st == 0 and k == 0
    p_ref, lw_ref, inc_ref, _ = orc.resample("stratified", lw, r, sort=sort)
    order = orc.sortperm_desc(lw) if sort else None
    W_ref = orc.cumweights(orc.softmax(lw), order)
    nm, gap = check_parents(p, p_ref, W_ref, strat_u(r, n), order)
    assert inc == pytest.approx(inc_ref, rel=RTOL)
    assert np.all(lw_out == 0.0)
    if not sort:
        assert np.all(np.diff(p) >= 0)  # monotone ancestors


@pytest.mark.parametrize("n", [5000, 1 << 18, (1 << 20) + 1])
@pytest.mark.parametrize("kind", ["A", "B"])
def test_selection_exact_given_gpu_cumweights(g, orc, n, kind):
    """Feeding the GPU's own cumulative weights to the reference search rule must give EXACTLY the GPU's
    ancestors: isolates the selection kernels from summation-order effects."""
    rng = np.random.default_rng(n + 3)
    lw = weights(rng, n, kind)
    r = rng.random(n)
    W = gpu_cumweights(g, lw)
    assert np.all(np.diff(W) >= -4 * np.finfo(float).eps)
    assert abs(W[-1] - 1) < 1e-12
    W_ref = orc.cumweights(orc.softmax(lw))
    assert np.max(np.abs(W - W_ref)) < tie_tolerance(n)
    st, p, *_ = raw_resample(g, "stratified", lw, r)
    expect = np.minimum(np.searchsorted(W, strat_u(r, n), side="left"), n - 1)  # min{k: W_k >= u}
    np.testing.assert_array_equal(p, expect)
    st, p, *_ = raw_resample(g, "multinomial", lw, r)
    expect = np.minimum(np.searchsorted(W, r, side="right"), n - 1)  # min{k: W_k > u}
    np.testing.assert_array_equal(p, expect)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 4097, 100_003, 1 << 18])
@pytest.mark.parametrize("kind", ["A", "B", "C", "spike"])
def test_inverse_cdf_adversarial_uniforms(g, n, kind):
    """The guide-table lookup (k_lookup) must return min{k : W_k > u} for uniforms sitting exactly on bucket
    boundaries, on cumulative weights themselves, at 0 and just below 1, and across long runs of zero-weight
    particles (bracket longer than the linear-scan limit)."""
    rng = np.random.default_rng(7 * n + len(kind))
    if kind == "spike":  # a few carriers, long runs of -inf and of negligible weights between them
        lw = np.full(n, -np.inf)
        lw[rng.integers(0, n, max(1, n // 1000))] = 0.0
        tiny = rng.integers(0, n, max(1, n // 3))
        lw[tiny] = np.where(np.isinf(lw[tiny]), -60.0, lw[tiny])
        lw[rng.integers(0, n)] = 3.0
    else:
        lw = weights(rng, n, kind)
    W = gpu_cumweights(g, lw)
    B = max(1, n >> 1)
    one_m = np.nextafter(1.0, 0.0)
    pieces = [rng.random(257), np.array([0.0, one_m, 0.5]), np.arange(min(B, 4096)) / B,
              rng.integers(0, B, 512) / B, np.nextafter(rng.integers(1, B + 1, 512) / B, 0.0),
              W[rng.integers(0, n, 1024)], np.nextafter(W[rng.integers(0, n, 1024)], 0.0),
              np.nextafter(W[rng.integers(0, n, 1024)], 1.0)]
    u = np.clip(np.concatenate(pieces), 0.0, one_m)
    for n_out in {u.size, max(1, u.size // 3)}:
        uu = np.ascontiguousarray(u[:n_out])
        st, p, *_ = raw_resample(g, "multinomial", lw, uu, n_out=n_out)
        assert st == 0
        # the parallel scan's W may wobble by an ulp inside a run of zero-weight particles, so min{k : W_k > u}
        # is taken over the running maximum; a different answer is accepted only as a cumulative-sum tie
        expect = np.minimum(np.searchsorted(np.maximum.accumulate(W), uu, side="right"), n - 1)
        ok = (W[p] > uu) | (p == n - 1)
        below = np.where(p > 0, W[np.maximum(p - 1, 0)] <= uu, True)
        assert np.all(ok & below), "returned index is not a crossing of the cumulative weights"
        check_parents(p, expect, W, uu, max_frac=0.5)  # a third of the uniforms sit exactly on W values


@pytest.mark.parametrize("n,n_out", [(100_003, 100_003), (1 << 18, 1 << 18), (1 << 18, 150_000), (70_000, 1 << 18)])
@pytest.mark.parametrize("kind", ["A", "B", "C"])
def test_multinomial_residual_vs_oracle(g, orc, n, n_out, kind):
    rng = np.random.default_rng(n + n_out)
    lw = weights(rng, n, kind)
    u = rng.random(n_out)
    w = orc.softmax(lw)
    W_ref = orc.cumweights(w)
    st, p, lw_out, inc, _ = raw_resample(g, "multinomial", lw, u, n_out=n_out)
    assert st == 0
    p_ref, _, inc_ref, _ = orc.resample("multinomial", lw, u, n_out=n_out)
    check_parents(p, p_ref, W_ref, u)
    assert inc == pytest.approx(inc_ref, rel=RTOL) and np.all(lw_out == 0) and lw_out.size == n_out
    st, p, lw_out, inc, _ = raw_resample(g, "residual", lw, u, n_out=n_out)
    assert st == 0
    p_ref, nd = orc.select_residual(w, u, n_out)
    # deterministic copies: exact integer arithmetic on floor(n_out*w)
    np.testing.assert_array_equal(p[:nd], p_ref[:nd])
    if nd < n_out:
        nw = n_out * w
        rw = nw - np.floor(nw)
        R_ref = orc.cumweights(rw / orc.load().orc_sum_pairwise(rw.ctypes.data, rw.size))
        check_parents(p[nd:], p_ref[nd:], R_ref, u[nd:])
    copies = np.bincount(p, minlength=n)
    assert np.all(copies >= np.floor(n_out * w))  # test/resample.jl:46-52


@pytest.mark.parametrize("method", ["multinomial", "residual", "stratified"])
def test_priorities_and_lml(g, orc, method):
    rng = np.random.default_rng(99)
    n = 50_000
    lw = rng.normal(0, 2, n)
    u = rng.random(n)
    lp = lw / 2  # priority_fn = w -> w/2 (test/resample.jl:14-23)
    for flags, sub in [(0, False), (g._lib.SUBSTATE, True)]:
        st, p, lw_out, inc, _ = raw_resample(g, method, lw, u, lp=lp, flags=flags)
        p_ref, lw_ref, inc_ref, _ = orc.resample(method, lw, u, lp=lp, substate=sub)
        W_ref = orc.cumweights(orc.softmax(lp))
        if method == "stratified":
            nm, _ = check_parents(p, p_ref, W_ref, strat_u(u, n))
        elif method == "multinomial":
            nm, _ = check_parents(p, p_ref, W_ref, u)
        else:
            nm = int(np.sum(p != p_ref))
            assert nm <= 8
        if nm == 0:
            np.testing.assert_allclose(lw_out, lw_ref, rtol=RTOL, atol=1e-11)
        # lml preserved (test/resample.jl:23,70,119,160)
        lml0 = orc.logsumexp(lw) - math.log(n)
        lml1 = inc + orc.logsumexp(lw_out) - math.log(n)
        assert lml1 == pytest.approx(lml0, abs=1e-9)


@pytest.mark.parametrize("method", ["multinomial", "residual", "stratified"])
def test_kat_invalid_weights_policy(g, method):
    """test/resample.jl:25-31,72-78,121-127."""
    n = 100
    state = g.ParticleFilterState(list(range(n)), np.full(n, -np.inf))
    with pytest.raises(g.GenPFErrorException, match="Invalid weights."):
        g.pf_resample(state, method, check=True)
    assert np.all(np.isneginf(state.log_weights)) and state.log_ml_est == 0.0  # nothing was mutated
    with pytest.warns(UserWarning, match="All input values are -Inf"):
        g.pf_resample(state, method, check="warn", sort_particles=False)
    assert np.all(state.log_weights == 0.0)
    assert state.log_ml_est == -np.inf
    with pytest.raises(g.GenPFErrorException, match="not recognized"):
        g.pf_resample(state, "systematic")
    lib, L = g.load(), g._lib
    lw = np.zeros(4)
    out_p, out_w = np.zeros(4, dtype=np.int64), np.zeros(4)
    assert lib.genpf_resample(9, L.ptr(lw), None, 4, 4, None, 0, 0, L.ptr(out_p), L.ptr(out_w), None, None) \
        == L.ERR_UNKNOWN_METHOD
    assert lib.genpf_resample(L.STRATIFIED, L.ptr(lw), None, 4, 2, None, 0, 0, L.ptr(out_p), L.ptr(out_w), None,
                              None) == L.ERR_INVALID_ARG
    # NaN weights: kind reported, outputs untouched
    lw = np.array([0.0, np.nan, 0.0, 0.0])
    st, p, lw_out, inc, kind = raw_resample(g, method, lw, np.full(4, 0.5))
    assert st == 0 and kind == 1 and np.all(p == -7) and np.all(np.isnan(lw_out))


@pytest.mark.parametrize("n", [100, 128, 4096, 1 << 16])
def test_kat_equal_weights_identity(g, n):
    """test/resample.jl:35-40,82-87."""
    u = np.random.default_rng(n).random(n)
    lw = np.full(n, 0.7)
    for method, flags in [("residual", 0), ("stratified", 0), ("stratified", g._lib.SORT_PARTICLES)]:
        st, p, lw_out, *_ = raw_resample(g, method, lw, u, flags=flags)
        np.testing.assert_array_equal(p, np.arange(n))


def test_seeded_uniforms_match_supplied(g, orc):
    """uniforms == NULL draws Philox(seed, stream 0, slot): same ancestors as passing them explicitly."""
    rng = np.random.default_rng(5)
    n = 30_000
    lw = rng.normal(0, 1, n)
    for method in ("multinomial", "residual", "stratified"):
        # stratified draws 32-bit stratum uniforms (4 per Philox block), the others 53-bit per slot
        u = orc.uniforms_strata(42, 0, n) if method == "stratified" else orc.uniforms(42, 0, n)
        a = raw_resample(g, method, lw, None, seed=42)[1]
        b = raw_resample(g, method, lw, u)[1]
        np.testing.assert_array_equal(a, b)


def test_views_and_segmented(g, orc):
    """test/resample.jl:130-162: per-block local parents, block and global lml preserved."""
    rng = np.random.default_rng(8)
    n, nb = 100, 5
    for method in ("multinomial", "residual", "stratified"):
        for prio in (None, lambda w: w / 2):
            traces = [("tr", i) for i in range(n)]
            state = g.ParticleFilterState(traces, rng.normal(0, 1, n))
            total0 = g.log_ml_estimate(state)
            for b in range(nb):
                sub = state[b * 20:(b + 1) * 20]
                old = sub.traces
                lml_b = g.log_ml_estimate(sub)
                g.pf_resample(sub, method, priority_fn=prio, sort_particles=False)
                assert all(p < 20 for p in sub.parents)
                assert sub.traces == [old[p] for p in sub.parents]
                assert g.log_ml_estimate(sub) == pytest.approx(lml_b, abs=1e-10)
            assert g.log_ml_estimate(state) == pytest.approx(total0, abs=1e-10)
    # segmented C entry point == per-segment oracle
    lib, L = g.load(), g._lib
    n = 10_000
    lw = rng.normal(0, 1.5, n)
    u = rng.random(n)
    offs = np.array([0, 1000, 1001, 5000, 10_000], dtype=np.int64)
    parents = np.empty(n, dtype=np.int64)
    lw_out = np.empty(n)
    kinds = np.zeros(4, dtype=np.int32)
    L.check(lib.genpf_resample_segmented(L.STRATIFIED, L.ptr(lw), None, n, offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                         4, L.ptr(u), 0, 0, L.ptr(parents), L.ptr(lw_out),
                                         kinds.ctypes.data_as(C.POINTER(C.c_int32))))
    for s in range(4):
        a, b = offs[s], offs[s + 1]
        p_ref, lw_ref, _, _ = orc.resample("stratified", lw[a:b], u[a:b], substate=True)
        np.testing.assert_array_equal(parents[a:b], p_ref)
        np.testing.assert_allclose(lw_out[a:b], lw_ref, rtol=RTOL)


def test_full_state_api_like_reference(g):
    """test/resample.jl:5-23 shape: new_traces == old_traces[parents], lml preserved, with/without priorities."""
    rng = np.random.default_rng(21)
    n = 100
    for method in ("multinomial", "residual", "stratified"):
        for prio in (None, lambda w: w / 2):
            traces = [object() for _ in range(n)]
            state = g.ParticleFilterState(traces, rng.normal(0, 1, n))
            lml0 = g.get_lml_est(state)
            old = list(state.traces)
            g.pf_resample(state, method, priority_fn=prio)
            assert all(state.traces[j] is old[state.parents[j]] for j in range(n))
            assert g.get_lml_est(state) == pytest.approx(lml0, abs=1e-10)
    # resize (test/resize.jl:3-84)
    for method in ("multinomial", "residual"):
        for n_new in (50, 150):
            state = g.ParticleFilterState([object() for _ in range(n)], rng.normal(0, 1, n))
            lml0, old = g.get_lml_est(state), list(state.traces)
            g.pf_resize(state, n_new, method)
            assert len(state.traces) == len(state.log_weights) == len(state.parents) == n_new
            assert all(state.traces[j] is old[state.parents[j]] for j in range(n_new))
            assert g.get_lml_est(state) == pytest.approx(lml0, abs=1e-10)


def test_mean_var(g, orc):
    """statistics.jl:13-17,48-54 and test/statistics.jl:10-18."""
    rng = np.random.default_rng(31)
    lib, L = g.load(), g._lib
    for n in (50, 2048, 100_001, 1 << 20):
        lw, x = rng.normal(0, 2, n), rng.normal(3, 2, n)
        m, v = C.c_double(), C.c_double()
        L.check(lib.genpf_weighted_mean_var(L.ptr(lw), L.ptr(x), n, 0, C.byref(m), C.byref(v)))
        mr, vr = orc.mean_var(lw, x)
        assert m.value == pytest.approx(mr, rel=RTOL, abs=1e-13) and v.value == pytest.approx(vr, rel=RTOL)

    class Tr(dict):
        pass
    state = g.ParticleFilterState([Tr(slope=5.0) for _ in range(100)], rng.normal(0, 1, 100))
    assert g.mean(state, "slope") == pytest.approx(5.0, abs=1e-12)
    assert g.var(state, "slope") == pytest.approx(0.0, abs=1e-6)


def test_replicate_dereplicate_coalesce(g, orc):
    """test/resize.jl:116-254."""
    rng = np.random.default_rng(41)
    n, k = 20, 5
    for layout in ("contiguous", "interleaved"):
        traces = [object() for _ in range(n)]
        state = g.ParticleFilterState(traces, rng.normal(0, 1, n))
        lw0, lml0 = state.log_weights.copy(), g.get_lml_est(state)
        g.pf_replicate(state, k, layout=layout)
        expect = np.repeat(np.arange(n), k) if layout == "contiguous" else np.tile(np.arange(n), k)
        np.testing.assert_array_equal(state.parents, expect)
        assert all(state.traces[j] is traces[expect[j]] for j in range(n * k))
        assert g.get_lml_est(state) == pytest.approx(lml0, abs=1e-10)
        g.pf_dereplicate(state, k, layout=layout)
        np.testing.assert_array_equal(state.log_weights, lw0)
        assert all(a is b for a, b in zip(state.traces, traces))
    # :sample against the oracle with the same uniforms
    n = 10_000
    lw, u = rng.normal(0, 1, n), rng.random(n // 5)
    for interleaved in (False, True):
        state = g.ParticleFilterState(list(range(n)), lw)
        g.pf_dereplicate(state, 5, layout="interleaved" if interleaved else "contiguous", method="sample", uniforms=u)
        q, out = orc.dereplicate(lw, 5, interleaved, True, u)
        np.testing.assert_array_equal(state.parents, q)
        np.testing.assert_allclose(state.log_weights, out, rtol=RTOL)
    # coalesce
    n = 5000
    vals = rng.integers(-2, 3, n)
    lw = rng.normal(0, 1, n)
    state = g.ParticleFilterState([("slope", int(v)) for v in vals], lw)
    lml0 = g.get_lml_est(state)
    g.pf_coalesce(state, by=lambda tr: tr[1])
    p_ref, lw_ref = orc.coalesce(lw, vals)
    np.testing.assert_array_equal(state.parents, p_ref)
    np.testing.assert_allclose(state.log_weights, lw_ref, rtol=RTOL)
    assert len(state.traces) == len(np.unique(vals)) <= 5
    assert g.get_lml_est(state) == pytest.approx(lml0, abs=1e-6)


def test_coalesce_large_groups_deterministic(g, orc):
    """Groups longer than the 4096-particle chunk are summed chunk-wise in a fixed order: matches the oracle to
    1e-10 and is bit-identical run to run (groups within one chunk are added in the reference's own order; only
    the device exp() differs from libm's in the last bit)."""
    rng = np.random.default_rng(77)
    for n, n_keys in ((60_000, 3), (60_000, 40)):
        vals = rng.integers(0, n_keys, n)
        lw = rng.normal(0, 1, n)
        outs = []
        for _ in range(2):
            state = g.ParticleFilterState([("k", int(v)) for v in vals], lw)
            g.pf_coalesce(state, by=lambda tr: tr[1])
            outs.append((state.parents.copy(), state.log_weights.copy()))
        p_ref, lw_ref = orc.coalesce(lw, vals)
        np.testing.assert_array_equal(outs[0][0], p_ref)
        np.testing.assert_allclose(outs[0][1], lw_ref, rtol=RTOL)
        np.testing.assert_array_equal(outs[0][1], outs[1][1])
        pm = g.proportionmap(g.ParticleFilterState([{"k": int(v)} for v in vals], lw), "k")
        w = orc.softmax(lw)
        for v, p in pm.items():
            assert p == pytest.approx(w[vals == v].sum(), rel=RTOL)


def test_large_properties(g):
    """BASELINE-size properties that need no oracle: 2^24 particles, stratified."""
    n = 1 << 24
    rng = np.random.default_rng(77)
    lw = rng.normal(0, 1, n)
    st, p, lw_out, inc, kind = raw_resample(g, "stratified", lw, None, seed=3)
    assert st == 0 and kind == 0
    assert p[0] >= 0 and p[-1] < n and np.all(np.diff(p) >= 0)
    w = np.exp(lw - lw.max())
    w /= w.sum()
    copies = np.bincount(p, minlength=n)
    # stratified: copies_i in {floor(n w_i) - 1 .. ceil(n w_i) + 1}
    assert np.all(np.abs(copies - n * w) < 2.0)
    assert inc == pytest.approx(np.log(np.exp(lw - lw.max()).sum()) + lw.max() - math.log(n), rel=1e-9)
    # idempotence: resampling equal weights is the identity
    st, p2, *_ = raw_resample(g, "stratified", lw_out, None, seed=4)
    np.testing.assert_array_equal(p2, np.arange(n))


def test_chunked_large_filter(g, orc):
    """n > 2^24 particles: the finalize runs per 2^24-particle chunk + combine (same maths as a multi-GPU shard)."""
    n = (1 << 25) + 12345
    rng = np.random.default_rng(123)
    lw = rng.normal(0, 1.5, n)
    assert g.logsumexp_host(lw) == pytest.approx(orc.logsumexp(lw), rel=RTOL)

    class S:
        log_weights = lw
    assert g.effective_sample_size(S) == pytest.approx(orc.ess(lw), rel=RTOL)
    r = rng.random(n)
    st, p, lw_out, inc, kind = raw_resample(g, "stratified", lw, r)
    assert st == 0 and kind == 0
    p_ref, _, inc_ref, _ = orc.resample("stratified", lw, r)
    W_ref = orc.cumweights(orc.softmax(lw))
    nm, gap = check_parents(p, p_ref, W_ref, strat_u(r, n))
    assert inc == pytest.approx(inc_ref, rel=RTOL)
    W = gpu_cumweights(g, lw)
    assert abs(W[-1] - 1) < 1e-11 and np.max(np.abs(W - W_ref)) < tie_tolerance(n)
    u = rng.random(1 << 20)
    st, p, *_ = raw_resample(g, "multinomial", lw, u, n_out=u.size)
    np.testing.assert_array_equal(p, np.minimum(np.searchsorted(W, u, side="right"), n - 1))
    # all -Inf over several chunks: uniform fallback is still the identity for stratified
    st, p, lw_out, inc, kind = raw_resample(g, "stratified", np.full(n, -np.inf), r)
    assert kind == 2 and inc == -np.inf
    assert np.mean(p == np.arange(n)) > 0.999999


@pytest.mark.parametrize("n", [1000, 2048, 70_001, 1 << 20])
def test_sort_particles_ties_and_signed_zero(g, orc, n):
    """sortperm(lp, rev=true) semantics of the hand-written radix sort: stable (ties keep ascending index),
    Julia isless order (0.0 before -0.0 under rev), negative and positive keys, -Inf weights."""
    rng = np.random.default_rng(n)
    vals = np.array([0.0, -0.0, 1.5, -2.25, -2.25, 3.0, -np.inf, 1e-300, -1e-300])
    lw = vals[rng.integers(0, len(vals), n)]
    lw[rng.integers(0, n)] = 7.0  # a unique maximum
    r = rng.random(n)
    st, p, lw_out, inc, kind = raw_resample(g, "stratified", lw, r, flags=g._lib.SORT_PARTICLES)
    assert st == 0 and kind == 0
    # thousands of identical addends make the SEQUENTIAL fp64 cumulative sum drift linearly (~n*eps/2), so the
    # reference here is the exact-cumsum oracle (long double), which the blocked GPU scan must match
    p_ref, _, inc_ref, _ = orc.resample("stratified", lw, r, sort=True, exact=True)
    order = orc.sortperm_desc(lw)
    W_ref = orc.cumweights(orc.softmax(lw), order, exact=True)
    check_parents(p, p_ref, W_ref, strat_u(r, n), order)
    assert inc == pytest.approx(inc_ref, rel=RTOL)
    # the permutation itself: ancestors of distinct weight classes appear in descending weight order
    assert np.all(np.diff(lw[p]) <= 0)


@pytest.mark.parametrize("n", [1000, (1 << 18) + 5, 1 << 21])
@pytest.mark.parametrize("kind", ["random", "equal", "small_runs", "long_runs", "one_run", "int_like"])
def test_sortperm_exact(g, orc, n, kind):
    """The radix sort's permutation == sortperm(keys, rev=true) (stable) on every path of the hybrid sort: plain
    keys, all-equal keys (every digit skipped), keys agreeing on the top 40 bits in short runs (repaired in
    place), in long runs and in a single run (device-gated full sort)."""
    rng = np.random.default_rng(n + len(kind))
    if kind == "random":
        keys = rng.normal(0, 3, n)
    elif kind == "equal":
        keys = np.full(n, -1.25)
    elif kind == "small_runs":  # ~8 keys per cluster, differing only in the last mantissa bits
        keys = rng.integers(1, max(2, n // 8), n).astype(float) * (1.0 + 1e-13 * rng.integers(-50, 50, n))
    elif kind == "long_runs":   # clusters of ~1000 keys within 4e-9 relative of each other
        keys = rng.integers(1, max(2, n // 1000), n).astype(float) * (1.0 + 1e-13 * rng.integers(-50, 50, n))
    elif kind == "one_run":
        keys = 2.0 + 1e-12 * rng.normal(size=n)
    else:                       # many exact duplicates, signed zeros, infinities
        keys = rng.integers(-3, 4, n).astype(float)
        keys[rng.integers(0, n, n // 50)] = -np.inf
        keys[rng.integers(0, n, n // 50)] = -0.0
    out = np.empty(n, dtype=np.int64)
    g._lib.check(g.load().genpf_debug_sortperm(g._lib.ptr(keys), n, 0, g._lib.ptr(out)))
    np.testing.assert_array_equal(out, orc.sortperm_desc(keys))


def raw_optimal_resize(g, lw, n_out, u=None, flags=0, seed=0):
    L, lib = g._lib, g.load()
    lw = np.ascontiguousarray(lw, dtype=np.float64)
    parents = np.full(n_out, -7, dtype=np.int64)
    lw_out = np.full(n_out, np.nan)
    n_keep, inv_w, kinds = C.c_int64(-1), C.c_double(np.nan), (C.c_int32 * 2)(-1, -1)
    up = None if u is None else C.byref(C.c_double(u))
    st = lib.genpf_optimal_resize(L.ptr(lw), lw.size, n_out, up, seed, flags, L.ptr(parents), L.ptr(lw_out),
                                  C.byref(n_keep), C.byref(inv_w), kinds)
    return st, parents, lw_out, n_keep.value, inv_w.value, (kinds[0], kinds[1])


@pytest.mark.parametrize("n", [100, 1000, 4097, 100_003, 1 << 20])
@pytest.mark.parametrize("kind", ["A", "B"])
def test_optimal_resize_vs_oracle(g, orc, n, kind):
    """pf_optimal_resize! (resize.jl:149-216) against the literal CPU restatement with the same rand()."""
    rng = np.random.default_rng(n + 11)
    lw = weights(rng, n, kind)
    for n_out in sorted({max(1, n // 7), n // 2, n - 1, n}):
        u = float(rng.random())
        ref = orc.optimal_resize(lw, n_out, u)
        st, p, lw_out, n_keep, inv_w, kinds = raw_optimal_resize(g, lw, n_out, u)
        assert st == 0 and kinds == (0, 0) if n_out > ref["n_keep"] else st == 0
        assert inv_w == pytest.approx(ref["inv_w"], rel=1e-9)
        w, _ = orc.safe_softmax(lw)
        borderline = np.any(np.abs(ref["inv_w"] * w - 1.0) < 1e-9)  # keep test c*w >= 1 decided by rounding
        if borderline:
            continue
        assert n_keep == ref["n_keep"]
        np.testing.assert_array_equal(p[:n_keep], ref["parents0"][:n_keep])
        np.testing.assert_allclose(lw_out, ref["lw_out"], rtol=RTOL)
        n_res = n_out - n_keep
        if n_res == 0:
            continue
        assert ref["status"] == 0
        # systematic draws over the remainder: exact except at cumulative-sum ties
        keep = np.zeros(n, dtype=bool)
        keep[p[:n_keep]] = True
        strat = np.flatnonzero(~keep)
        C_ref = orc.cumweights(orc.safe_softmax(lw[strat])[0])
        step = 1.0 / n_res
        thr = u * step + np.arange(n_res) * step
        q_gpu, q_ref = np.searchsorted(strat, p[n_keep:]), np.searchsorted(strat, ref["parents0"][n_keep:])
        assert np.all(strat[q_gpu] == p[n_keep:]) and np.all(np.diff(p[n_keep:]) > 0)
        check_parents(q_gpu, q_ref, C_ref, thr)


def test_optimal_resize_kats_and_errors(g):
    """test/resize.jl:86-113 through the host mirror (pf_resize!(state, n, :optimal))."""
    rng = np.random.default_rng(3)
    n = 100
    for n_particles in (25, 50):
        traces = [object() for _ in range(n)]
        state = g.ParticleFilterState(traces, rng.normal(-20, 1.5, n))
        lw0, lml0 = state.log_weights.copy(), g.get_lml_est(state)
        g.pf_resize(state, n_particles, "optimal", uniform=0.25)
        assert len(state.traces) == n_particles
        assert all(state.traces[j] is traces[state.parents[j]] for j in range(n_particles))
        assert g.get_lml_est(state) == pytest.approx(lml0, rel=1e-3)
        assert len(set(state.parents.tolist())) == n_particles  # unique parents (the point of the algorithm)
    state = g.ParticleFilterState(list(range(n)), np.full(n, -np.inf))
    with pytest.raises(g.GenPFErrorException):
        g.pf_optimal_resize(state, 50, check=True)
    g.pf_optimal_resize(state, 50, check=False, uniform=0.3)
    assert len(state.traces) == 50 and np.all(np.isneginf(state.log_weights))
    state = g.ParticleFilterState(list(range(n)), np.zeros(n))
    with pytest.raises(AssertionError):
        g.pf_optimal_resize(state, n + 1)  # @assert n_particles <= n_old, resize.jl:183
    # library-drawn uniform: deterministic in the seed
    lw = rng.normal(0, 1, 5000)
    a = raw_optimal_resize(g, lw, 1000, None, seed=9)
    b = raw_optimal_resize(g, lw, 1000, None, seed=9)
    c = raw_optimal_resize(g, lw, 1000, None, seed=10)
    assert a[0] == 0 and np.array_equal(a[1], b[1]) and not np.array_equal(a[1], c[1])
    u9 = np.empty(1)
    g._lib.check(g.load().genpf_uniforms(9, 0, 1, 0, g._lib.ptr(u9)))
    d = raw_optimal_resize(g, lw, 1000, float(u9[0]))
    np.testing.assert_array_equal(a[1], d[1])


def test_proportionmap(g, orc):
    """StatsBase.proportionmap(state, addr) (statistics.jl:91-130): sum of normalised weights per distinct value."""
    rng = np.random.default_rng(23)
    n = 20_000
    vals = rng.integers(-3, 4, n)
    lw = rng.normal(0, 2, n)
    state = g.ParticleFilterState([{"slope": int(v)} for v in vals], lw)
    pm = g.proportionmap(state, "slope")
    w = orc.softmax(lw)
    assert set(pm) == set(int(v) for v in np.unique(vals))
    for v, p in pm.items():
        assert p == pytest.approx(w[vals == v].sum(), rel=RTOL)
    assert sum(pm.values()) == pytest.approx(1.0, rel=1e-12)
    pm2 = g.proportionmap(state, "slope", f=lambda v: v > 0)
    assert pm2[True] == pytest.approx(w[vals > 0].sum(), rel=RTOL)
    # degenerate: one value, and all -Inf weights (uniform fallback of get_norm_weights' safe path)
    state = g.ParticleFilterState([{"slope": 1}] * 50, np.zeros(50))
    assert g.proportionmap(state, "slope") == {1: pytest.approx(1.0)}


def test_sample_unweighted_traces(g, orc):
    """Gen.sample_unweighted_traces (sub-state method utils.jl:189-194): draws in proportion to the normalised
    weights, state untouched; with supplied uniforms = the inverse-CDF rule of the multinomial path."""
    rng = np.random.default_rng(5)
    n = 10_000
    lw = rng.normal(0, 2, n)
    state = g.ParticleFilterState(list(range(n)), lw)
    u = rng.random(2500)
    out = g.sample_unweighted_traces(state, 2500, uniforms=u)
    p_ref, *_ = orc.resample("multinomial", lw, u, n_out=2500)
    assert out == [int(p) for p in p_ref]
    np.testing.assert_array_equal(state.log_weights, lw)
    sub = state[100:600]
    out = g.sample_unweighted_traces(sub, 500, uniforms=u[:500])
    p_ref, *_ = orc.resample("multinomial", lw[100:600], u[:500])
    assert out == [100 + int(p) for p in p_ref]
