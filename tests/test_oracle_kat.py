"""CPU: pins the C oracle (oracle/genpf_oracle.c) against
 (1) the golden fixtures produced by the independent pure-Python restatement (tests/golden/make_golden.py),
 (2) every known-answer invariant the reference's own tests state (SURVEY.md 8c "golden vectors / KATs").
Reference test files are cited per test (paths relative to the reference root)."""
import math

import numpy as np
import pytest

from conftest import GOLDEN_CASES


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_scalars_and_weights(orc, golden, case):
    lw = golden[f"{case}/lw"]
    assert orc.logsumexp(lw) == golden[f"{case}/lse"]
    assert orc.ess(lw) == golden[f"{case}/ess"]
    w, kind = orc.safe_softmax(lw)
    assert kind == 0
    np.testing.assert_array_equal(w, golden[f"{case}/w"])
    m, v = orc.mean_var(lw, golden[f"{case}/x"])
    np.testing.assert_array_equal([m, v], golden[f"{case}/mean_var"])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_philox(orc, golden, case):
    seed = {"n100_s1": 11, "n1000_s5": 12, "n2048_s2": 13, "n3000_s1": 14}[case]
    u = golden[f"{case}/u"]
    np.testing.assert_array_equal(orc.uniforms(seed, 0, u.size), u)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_stratified(orc, golden, case):
    lw, u = golden[f"{case}/lw"], golden[f"{case}/u"]
    p, lw_out, inc, kind = orc.resample("stratified", lw, u)
    np.testing.assert_array_equal(p, golden[f"{case}/strat/parents"])
    assert kind == 0 and np.all(lw_out == 0.0)
    assert inc == golden[f"{case}/lse"] - math.log(lw.size)
    np.testing.assert_array_equal(orc.sortperm_desc(lw), golden[f"{case}/order"])
    p, *_ = orc.resample("stratified", lw, u, sort=True)
    np.testing.assert_array_equal(p, golden[f"{case}/strat_sorted/parents"])
    p, lw_out, _, _ = orc.resample("stratified", lw, u, lp=lw / 2)
    np.testing.assert_array_equal(p, golden[f"{case}/strat_prio/parents"])
    np.testing.assert_array_equal(lw_out, golden[f"{case}/strat_prio/lw_out"])
    p, lw_out, inc, _ = orc.resample("stratified", lw, u, lp=lw / 2, substate=True)
    np.testing.assert_array_equal(lw_out, golden[f"{case}/strat_prio/lw_out_sub"])
    assert inc == 0.0
    _, lw_out, _, _ = orc.resample("stratified", lw, u, substate=True)
    np.testing.assert_array_equal(lw_out, golden[f"{case}/lw_out_sub"])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_multinomial_residual_and_resize(orc, golden, case):
    lw = golden[f"{case}/lw"]
    n = lw.size
    for n_out in (n, n // 2, n + n // 2):
        us = golden[f"{case}/u_{n_out}"]
        p, lw_out, inc, _ = orc.resample("multinomial", lw, us, n_out=n_out)
        np.testing.assert_array_equal(p, golden[f"{case}/multi_{n_out}/parents"])
        assert np.all(lw_out == 0.0) and lw_out.size == n_out
        assert inc == golden[f"{case}/lse"] - math.log(n)  # pre-resize n (resize.jl:56)
        p, *_ = orc.resample("residual", lw, us, n_out=n_out)
        np.testing.assert_array_equal(p, golden[f"{case}/resid_{n_out}/parents"])


def test_lazy_loop_equals_search_form(orc):
    """the lazy-draw loop of resample.jl:160-170 is outcome-identical to min{k: W_k >= u_i}."""
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 100, 1000, 4096, 10007):
        for sigma in (0.0, 1.0, 6.0):
            lw = rng.normal(0, sigma, n)
            w = orc.softmax(lw)
            r = rng.random(n)
            a = orc.select_stratified(w, r)
            b = orc.select_stratified(w, r, search=True)
            np.testing.assert_array_equal(a, b)
            order = orc.sortperm_desc(lw)
            np.testing.assert_array_equal(orc.select_stratified(w, r, order=order),
                                          orc.select_stratified(w, r, order=order, search=True))


# ---- test/utils.jl:5-10
def test_kat_normalisation_and_ess(orc):
    rng = np.random.default_rng(1)
    lw = rng.normal(0, 3, 1000)
    assert abs(np.exp(orc.lognorm(lw)).sum() - 1) < 1e-12
    w = orc.softmax(lw)
    assert abs(w.sum() - 1) < 1e-12
    assert orc.ess(lw) == pytest.approx(w.sum() ** 2 / (w ** 2).sum(), rel=1e-12)


# ---- test/resample.jl:35-40,82-87: equal weights => residual and stratified leave the traces unchanged
@pytest.mark.parametrize("n", [100, 128, 1024, 4096])
def test_kat_equal_weights_identity(orc, n):
    rng = np.random.default_rng(2)
    lw = np.full(n, -1.25)
    u = rng.random(n)
    for method in ("residual", "stratified"):
        p, lw_out, _, kind = orc.resample(method, lw, u)
        np.testing.assert_array_equal(p, np.arange(n))
        assert kind == 0 and np.all(lw_out == 0)
    p, *_ = orc.resample("stratified", lw, u, sort=True)
    np.testing.assert_array_equal(p, np.arange(n))  # stable sort keeps index order


# ---- n*fl(1/n) < 1 for n = 49 => residual makes zero deterministic copies
def test_kat_residual_literal_floor(orc):
    n = 49
    w = orc.softmax(np.zeros(n))
    _, nd = orc.select_residual(w, np.random.default_rng(3).random(n))
    assert nd == 0 and n * (1.0 / n) < 1.0


# ---- test/resample.jl:25-31,72-78,121-127: all -Inf => invalid, uniform fallback, final lw == 0
@pytest.mark.parametrize("method", ["multinomial", "residual", "stratified"])
def test_kat_all_neginf(orc, method):
    n = 100
    lw = np.full(n, -np.inf)
    p, lw_out, inc, kind = orc.resample(method, lw, np.random.default_rng(4).random(n))
    assert kind == 2
    assert np.all(lw_out == 0.0)
    assert inc == -np.inf  # log_ml_est becomes -Inf (App. C)
    assert p.min() >= 0 and p.max() < n


def test_kat_invalid_kinds(orc):
    assert orc.safe_softmax(np.array([0.0, np.nan]))[1] == 1
    assert orc.safe_softmax(np.array([-np.inf, -np.inf]))[1] == 2
    assert orc.safe_softmax(np.array([0.0, np.inf]))[1] == 4
    w, k = orc.safe_softmax(np.array([0.0, -np.inf]))
    assert k == 0 and w[0] == 1.0 and w[1] == 0.0


# ---- test/resample.jl:12,23,54,70,102,119: lml estimate preserved by every resample (with and w/o priorities)
@pytest.mark.parametrize("method", ["multinomial", "residual", "stratified"])
@pytest.mark.parametrize("prio", [False, True])
def test_kat_lml_preserved(orc, method, prio):
    rng = np.random.default_rng(5)
    n = 500
    lw = rng.normal(0, 2, n)
    lml0 = orc.logsumexp(lw) - math.log(n)
    p, lw_out, inc, _ = orc.resample(method, lw, rng.random(n), lp=(lw / 2 if prio else None))
    lml1 = inc + orc.logsumexp(lw_out) - math.log(n)
    assert lml1 == pytest.approx(lml0, abs=1e-10)


# ---- test/resample.jl:46-52,61-68,93-100: copies >= floor(n w)
def test_kat_min_copies(orc):
    rng = np.random.default_rng(6)
    n = 1000
    lw = rng.normal(0, 2, n)
    w = orc.softmax(lw)
    p, *_ = orc.resample("residual", lw, rng.random(n))
    copies = np.bincount(p, minlength=n)
    assert np.all(copies >= np.floor(n * w))
    p, *_ = orc.resample("stratified", lw, rng.random(n), sort=True)
    copies = np.bincount(p, minlength=n)
    i = np.argmax(w)
    assert copies[i] >= math.floor(n * w[i])


# ---- test/resample.jl:130-162: per-view resampling keeps every block's and the whole state's lml
def test_kat_views_compose(orc):
    rng = np.random.default_rng(7)
    n, nb = 100, 5
    lw = rng.normal(0, 1, n)
    total0 = orc.logsumexp(lw)
    new = lw.copy()
    for b in range(nb):
        sl = slice(b * 20, (b + 1) * 20)
        p, lw_out, inc, _ = orc.resample("stratified", lw[sl], rng.random(20), substate=True)
        assert inc == 0.0 and p.max() < 20
        assert orc.logsumexp(lw_out) == pytest.approx(orc.logsumexp(lw[sl]), abs=1e-12)
        new[sl] = lw_out
    assert orc.logsumexp(new) == pytest.approx(total0, abs=1e-12)


# ---- test/statistics.jl:10-18: degenerate distributions
def test_kat_degenerate_mean_var(orc):
    lw = np.random.default_rng(8).normal(0, 1, 50)
    m, v = orc.mean_var(lw, np.full(50, 5.0))
    assert m == pytest.approx(5.0, abs=1e-12) and v == pytest.approx(0.0, abs=1e-6)


# ---- test/resize.jl:116-182: replicate layouts, lml invariant, dereplicate keepfirst round-trips
@pytest.mark.parametrize("interleaved", [False, True])
def test_kat_replicate_roundtrip(orc, interleaved):
    rng = np.random.default_rng(9)
    n, k = 20, 5
    lw = rng.normal(0, 1, n)
    p, lw2 = orc.replicate(lw, k, interleaved)
    expect = np.tile(np.arange(n), k) if interleaved else np.repeat(np.arange(n), k)
    np.testing.assert_array_equal(p, expect)
    np.testing.assert_array_equal(lw2, lw[expect])
    assert orc.logsumexp(lw2) - math.log(n * k) == pytest.approx(orc.logsumexp(lw) - math.log(n), abs=1e-12)
    q, lw3 = orc.dereplicate(lw2, k, interleaved)
    np.testing.assert_array_equal(lw3, lw)
    np.testing.assert_array_equal(p[q], np.arange(n))


# ---- test/resize.jl:184-225: dereplicate :sample weight = logsumexp(block) - log k
def test_kat_dereplicate_sample(orc):
    rng = np.random.default_rng(10)
    n, k = 100, 5
    lw = rng.normal(0, 1, n)
    q, out = orc.dereplicate(lw, k, False, True, rng.random(n // k))
    for b in range(n // k):
        assert b * k <= q[b] < (b + 1) * k
        assert out[b] == pytest.approx(orc.logsumexp(lw[b * k:(b + 1) * k]) - math.log(k), abs=1e-12)


# ---- test/resize.jl:227-254: coalesce
def test_kat_coalesce(orc):
    rng = np.random.default_rng(11)
    n = 100
    keys = rng.integers(-2, 3, n)
    lw = rng.normal(0, 1, n)
    p, out = orc.coalesce(lw, keys)
    assert len(p) == len(np.unique(keys)) <= 5
    np.testing.assert_array_equal(np.sort(keys[p]), np.unique(keys))
    for j, i in enumerate(p):
        assert i == np.flatnonzero(keys == keys[i])[0]
    assert orc.logsumexp(out) - math.log(len(p)) == pytest.approx(orc.logsumexp(lw) - math.log(n), abs=1e-6)


# ---- README.md:43-54 model pieces (Gen logpdf closed forms, test/update.jl:8-10)
def test_kat_object_motion_pieces(orc):
    n = 8
    U = np.array([0.1, 0.3, 0.2, 0.8, 0.7, 0.74, 0.76, 0.0])
    Z = np.linspace(-1, 1, n)
    y, m = orc.om_transition(None, None, math.sin(1.0), U, Z)
    np.testing.assert_array_equal(m, (U < 0.25).astype(np.uint8))
    np.testing.assert_array_equal(y, (0.0 + np.where(m, math.sin(1.0), 0.0)) + 0.01 * Z)
    lw = orc.om_obs_logpdf(y, 0.3)
    ref = -0.5 * ((0.3 - y) / 0.25) ** 2 - 0.5 * math.log(2 * math.pi) - math.log(0.25)
    np.testing.assert_allclose(lw, ref, rtol=1e-14)
    y2, m2 = orc.om_transition(y, m, math.sin(2.0), U, Z)
    np.testing.assert_array_equal(m2, (U < np.where(m, 0.75, 0.25)).astype(np.uint8))
    # MH: alpha >= 0 always accepts; log(U3) < alpha decides otherwise (Gen mh)
    yq, mq, acc = orc.om_mh(y, m, y2, m2, math.sin(2.0), 5.0, U, Z, np.full(n, 1.0 - 1e-12))
    prop, _ = orc.om_transition(y, m, math.sin(2.0), U, Z)
    alpha = orc.om_obs_logpdf(prop, 5.0) - orc.om_obs_logpdf(y2, 5.0)
    np.testing.assert_array_equal(acc, math.log(1.0 - 1e-12) < alpha)


@pytest.mark.parametrize("n_particles", [25, 50])
def test_optimal_resize_kat(orc, n_particles):
    """test/resize.jl:86-105: length, parents, kept weights shifted by log(N/n), lml preserved to rtol 1e-3."""
    rng = np.random.default_rng(5)
    n = 100
    lw = rng.normal(-20.0, 1.5, n)  # line-model-like log-weights: lml estimate well away from zero
    w, _ = orc.safe_softmax(lw)
    thresh = orc.find_inv_w_threshold(w, n_particles)
    keep = np.flatnonzero(thresh * w >= 1)
    r = orc.optimal_resize(lw, n_particles, 0.37)
    assert r["status"] == 0 and r["n_selected"] == n_particles - keep.size
    assert r["n_keep"] == keep.size and r["inv_w"] == thresh
    np.testing.assert_array_equal(r["parents0"][:keep.size], keep)
    np.testing.assert_allclose(r["lw_out"][:keep.size], lw[keep] + np.log(n_particles) - np.log(n), rtol=1e-12)
    drawn = r["parents0"][keep.size:]
    assert np.all(np.diff(drawn) > 0) and not np.intersect1d(drawn, keep).size  # unique, in index order
    old = orc.logsumexp(lw) - np.log(n)
    new = orc.logsumexp(r["lw_out"]) - np.log(n_particles)
    assert new == pytest.approx(old, rel=1e-3)


def test_optimal_resize_threshold_properties(orc):
    """find_inv_w_threshold (resize.jl:199-216): c*B + A = N at the returned threshold; identity at N = n."""
    rng = np.random.default_rng(6)
    w, _ = orc.safe_softmax(rng.normal(0, 2, 1000))
    for N in (1, 10, 500, 999):
        c = orc.find_inv_w_threshold(w, N)
        assert np.sum(np.minimum(c * w, 1.0)) == pytest.approx(N, rel=0.02)  # expected offspring sum to ~N
    r = orc.optimal_resize(np.log(w), 1000, 0.5)
    assert r["n_keep"] == 1000 and r["status"] == 0
    np.testing.assert_array_equal(r["parents0"], np.arange(1000))


def test_optimal_resize_invalid_weights(orc):
    """test/resize.jl:107-113: all -Inf weights fall back to uniform, every output weight is -Inf."""
    r = orc.optimal_resize(np.full(100, -np.inf), 50, 0.3)
    assert r["kind"] == 2 and r["kind_strat"] == 2 and r["n_keep"] == 0 and r["status"] == 0
    assert np.all(np.isneginf(r["lw_out"]))
    np.testing.assert_array_equal(r["parents0"], np.arange(0, 100, 2))


def test_optimal_resize_golden(orc):
    """The C oracle's pf_optimal_resize! reproduces the independent pure-Python restatement
    (tests/golden/make_golden.py::optimal_resize) bit for bit: threshold, keep set, systematic draws, weights."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2_optimal.npz"))
    g1 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
    keys = sorted(k for k in z.files if k.endswith("/u"))
    assert len(keys) == 12
    for key in keys:
        case, tag, _ = key.split("/")
        N = int(tag.split("_")[1])
        lw = g1[f"{case}/lw"]
        r = orc.optimal_resize(lw, N, float(z[key]))
        assert r["status"] == 0 and r["n_keep"] == int(z[f"{case}/{tag}/n_keep"])
        assert r["inv_w"] == float(z[f"{case}/{tag}/inv_w"])
        np.testing.assert_array_equal(r["parents0"], z[f"{case}/{tag}/parents"])
        np.testing.assert_array_equal(r["lw_out"], z[f"{case}/{tag}/lw_out"])


def test_resize_golden(orc):
    """pf_replicate! / pf_dereplicate! (keepfirst and sample) / pf_coalesce!: the C oracle reproduces the independent
    pure-Python restatement (tests/golden/make_golden.py) bit for bit."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, g1 = np.load(os.path.join(here, "golden_v3_resize.npz")), np.load(os.path.join(here, "golden_v1.npz"))
    for case in ("n100_s1", "n3000_s1"):
        lw = g1[f"{case}/lw"]
        for tag, inter in (("contiguous", False), ("interleaved", True)):
            p, w = orc.replicate(lw, 3, inter)
            np.testing.assert_array_equal(p, z[f"{case}/replicate3_{tag}/parents"])
            np.testing.assert_array_equal(w, z[f"{case}/replicate3_{tag}/lw_out"])
            for mtag, smp in (("keepfirst", False), ("sample", True)):
                p, w = orc.dereplicate(lw, 5, inter, smp, z[f"{case}/u_derep"])
                np.testing.assert_array_equal(p, z[f"{case}/dereplicate5_{tag}_{mtag}/parents"])
                np.testing.assert_array_equal(w, z[f"{case}/dereplicate5_{tag}_{mtag}/lw_out"])
        p, w = orc.coalesce(lw, z[f"{case}/coalesce/keys"])
        np.testing.assert_array_equal(p, z[f"{case}/coalesce/parents"])
        np.testing.assert_array_equal(w, z[f"{case}/coalesce/lw_out"])
