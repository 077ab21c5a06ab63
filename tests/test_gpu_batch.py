"""GPU: batches of independent filters (BASELINE config 5: 4096 filters x 4096 particles) against the CPU oracle,
filter by filter -- per-filter resample decisions inside one launch (the reference's views resample independently,
test/resample.jl:130-162; README.md:68-74), pf_replicate! + residual / multinomial resize per filter
(resize.jl:87-124,236-244), and sort_particles=true for a whole batch in one launch."""
import math

import numpy as np
import pytest

from util import oracle_readme_step

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _noisy_init(g, pf, obs, U, Z):
    L = g._lib
    L.check(g.load().genpf_initialize_with_noise(pf._h, L.ptr(pf._obs(obs)), L.ptr(pf.model.aux(1)), L.ptr(U), L.ptr(Z)))
    pf.t = 1


def _mixed_weights(rng, nf, n):
    """Filters 0, 2, ... degenerate (ESS << n/2), the others nearly flat (ESS ~ n)."""
    lw = np.empty((nf, n))
    for f in range(nf):
        lw[f] = rng.normal(0, 4.0 if f % 2 == 0 else 0.1, n)
    return lw.ravel()


@pytest.mark.parametrize("method", ["stratified", "multinomial", "residual"])
def test_batch_per_filter_decisions_vs_oracle(g, orc, method):
    nf, n = 5, 3000
    N = nf * n
    rng = np.random.default_rng(31)
    model = g.DeviceModel("object_motion")
    pf = g.DevicePFState(model, n, n_filters=nf, seed=2)
    obs1, obs2 = rng.normal(0, 0.3, nf), rng.normal(0.4, 0.3, nf)
    U, Z = rng.random(N), rng.normal(size=N)
    _noisy_init(g, pf, obs1, U, Z)
    lw = _mixed_weights(rng, nf, n)
    pf.log_weights = lw
    lml0 = np.atleast_1d(g.log_ml_estimate(pf))
    r, U2, Z2, U3, U1, Z1 = rng.random(N), rng.random(N), rng.normal(size=N), rng.random(N), rng.random(N), rng.normal(size=N)
    g.pf_step_with_noise(pf, 2, obs1, obs2, method=method, ess_thresh=0.5, uniforms=r, U2=U2, Z2=Z2, U3=U3, U1=U1, Z1=Z1)
    p, y2, y1, m2, lw2 = pf.parents, pf.field("y", 2), pf.field("y", 1), pf.field("moving", 2), pf.log_weights
    lml2 = np.atleast_1d(g.log_ml_estimate(pf))
    decided = []
    for f in range(nf):
        sl = slice(f * n, (f + 1) * n)
        ys, ms = orc.om_transition(None, None, math.sin(1.0), U[sl], Z[sl])
        ess = orc.ess(lw[sl])
        resample = ess < 0.5 * n
        decided.append(resample)
        if resample:
            if method == "stratified":
                st = dict(y_pp=None, m_pp=None, y=ys, m=ms, lw=lw[sl])
                st, p_ref, n_tie, inc, _ = oracle_readme_step(orc, st, 2, obs1[f], obs2[f], r[sl], U2[sl], Z2[sl], U3[sl],
                                                              U1[sl], Z1[sl], p_gpu=p[sl])
                assert n_tie == 0
                y_new, m_new, y_mid, lw_ref = st["y"], st["m"], st["y_pp"], st["lw"]
            else:
                p_ref, lw0, inc, _ = orc.resample(method, lw[sl], r[sl])
                np.testing.assert_array_equal(p[sl], p_ref)
                y_mid, m_mid, _ = orc.om_mh(None, None, ys[p_ref], ms[p_ref], math.sin(1.0), obs1[f], U2[sl], Z2[sl], U3[sl])
                y_new, m_new = orc.om_transition(y_mid, m_mid, math.sin(2.0), U1[sl], Z1[sl])
                lw_ref = orc.om_obs_logpdf(y_new, obs2[f], lw0)
            lml_ref = lml0[f] - (orc.logsumexp(lw[sl]) - math.log(n)) + inc + orc.logsumexp(lw_ref) - math.log(n)
        else:  # only pf_update!: identity ancestors, no rejuvenation, weights accumulate (update.jl:21)
            np.testing.assert_array_equal(p[sl], np.arange(n))
            y_mid = ys
            y_new, m_new = orc.om_transition(ys, ms, math.sin(2.0), U1[sl], Z1[sl])
            lw_ref = orc.om_obs_logpdf(y_new, obs2[f], lw[sl])
            lml_ref = lml0[f] - (orc.logsumexp(lw[sl]) - math.log(n)) + orc.logsumexp(lw_ref) - math.log(n)
        np.testing.assert_array_equal(y2[sl], y_new)
        np.testing.assert_array_equal(m2[sl], m_new)
        np.testing.assert_array_equal(y1[sl], y_mid)
        np.testing.assert_allclose(lw2[sl], lw_ref, rtol=RTOL, atol=1e-12)
        assert lml2[f] == pytest.approx(lml_ref, rel=RTOL, abs=1e-9)
    assert any(decided) and not all(decided)


@pytest.mark.parametrize("method", ["stratified", "residual"])
def test_batch_step_mixed_decisions_library_noise(g, method):
    """genpf_step on a batch whose filters disagree (round 1 refused this): passing filters keep their population."""
    nf, n = 6, 4096
    rng = np.random.default_rng(5)
    model = g.DeviceModel("object_motion")
    obs1, obs2 = rng.normal(0, 0.3, nf), rng.normal(0.4, 0.3, nf)
    pf = g.pf_initialize(model, (1,), obs1, n, n_filters=nf, seed=9)
    lw = _mixed_weights(rng, nf, n)
    pf.log_weights = lw
    y1 = pf.field("y", 1)
    ess = g.pf_step(pf, 2, obs1, obs2, method=method, ess_thresh=0.5)
    p, lw2, y2, y1b = pf.parents, pf.log_weights, pf.field("y", 2), pf.field("y", 1)
    n_pass = 0
    for f in range(nf):
        sl = slice(f * n, (f + 1) * n)
        inc = -(((obs2[f] - y2[sl]) / 0.25) ** 2 + math.log(2 * math.pi)) / 2 - math.log(0.25)
        if ess[f] >= 0.5 * n:
            n_pass += 1
            np.testing.assert_array_equal(p[sl], np.arange(n))
            np.testing.assert_array_equal(y1b[sl], y1[sl])
            np.testing.assert_allclose(lw2[sl], lw[sl] + inc, rtol=1e-9, atol=1e-9)
        else:
            assert not np.array_equal(p[sl], np.arange(n)) and p[sl].min() >= 0 and p[sl].max() < n
            if method == "stratified":
                assert np.all(np.diff(p[sl]) >= 0)
            np.testing.assert_allclose(lw2[sl], inc, rtol=1e-9, atol=1e-9)
    assert 0 < n_pass < nf


def test_batch_replicate_then_resize_vs_oracle(g, orc):
    """config 5: pf_replicate!(k=2) then residual (or multinomial) resize back, every filter of the batch at once."""
    nf, n, k = 4, 1024, 2
    rng = np.random.default_rng(77)
    model = g.DeviceModel("object_motion")
    for layout in ("contiguous", "interleaved"):
        for method in ("residual", "multinomial"):
            obs1 = rng.normal(0, 0.3, nf)
            pf = g.pf_initialize(model, (1,), obs1, n, n_filters=nf, seed=4)
            g.pf_update(pf, (2,), None, obs1 + 0.1)
            lw0, y0, lml0 = pf.log_weights, pf.field("y", 2), np.atleast_1d(g.log_ml_estimate(pf))
            g.pf_replicate(pf, k, layout=layout)
            assert len(pf) == n * k
            p, lw1, y1 = pf.parents, pf.log_weights, pf.field("y", 2)
            for f in range(nf):
                p_ref, lw_ref = orc.replicate(lw0[f * n:(f + 1) * n], k, interleaved=(layout == "interleaved"))
                sl = slice(f * n * k, (f + 1) * n * k)
                np.testing.assert_array_equal(p[sl], p_ref)
                np.testing.assert_array_equal(lw1[sl], lw_ref)
                np.testing.assert_array_equal(y1[sl], y0[f * n:(f + 1) * n][p_ref])
            np.testing.assert_allclose(np.atleast_1d(g.log_ml_estimate(pf)), lml0, atol=1e-10)  # test/resize.jl:131
            u = rng.random(nf * n)
            g.pf_resize(pf, n, method, uniforms=u)
            assert len(pf) == n
            p2, lw2, y2 = pf.parents, pf.log_weights, pf.field("y", 2)
            for f in range(nf):
                sl_big, sl = slice(f * n * k, (f + 1) * n * k), slice(f * n, (f + 1) * n)
                p_ref, lw_ref, _, _ = orc.resample(method, lw1[sl_big], u[sl], n_out=n)
                np.testing.assert_array_equal(p2[sl], p_ref)
                np.testing.assert_array_equal(lw2[sl], lw_ref)
                np.testing.assert_array_equal(y2[sl], y1[sl_big][p_ref])
            np.testing.assert_allclose(np.atleast_1d(g.log_ml_estimate(pf)), lml0, atol=1e-9)
            g.pf_step(pf, 3, obs1 + 0.1, obs1 + 0.2, method="stratified", ess_thresh=1.0)  # keeps working
            assert np.isfinite(pf.log_weights).all()
    # dereplicate on a batch: keepfirst round trip (test/resize.jl:147-182)
    pf = g.pf_initialize(model, (1,), np.zeros(nf), n, n_filters=nf, seed=8)
    lw0, y0 = pf.log_weights, pf.field("y", 1)
    g.pf_replicate(pf, 3)
    g.pf_dereplicate(pf, 3)
    np.testing.assert_array_equal(pf.log_weights, lw0)
    np.testing.assert_array_equal(pf.field("y", 1), y0)


@pytest.mark.parametrize("n", [1000, 4096])
def test_batch_sorted_stratified_vs_oracle(g, orc, n):
    """sort_particles=true (the reference default, resample.jl:143-145) for a whole batch: the segmented in-block
    sort must give sortperm(lp, rev=true) per filter, ties in ascending index."""
    nf = 7
    rng = np.random.default_rng(n)
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), np.zeros(nf), n, n_filters=nf, seed=3)
    lw = rng.normal(0, 2.0, nf * n)
    lw[:n] = np.round(lw[:n])          # many exact ties in filter 0
    lw[n:2 * n] = -1.25                # all equal in filter 1
    pf.log_weights = lw
    y0 = pf.field("y", 1)
    r = rng.random(nf * n)
    g.pf_resample(pf, "stratified", sort_particles=True, uniforms=r)
    p, y1 = pf.parents, pf.field("y", 1)
    for f in range(nf):
        sl = slice(f * n, (f + 1) * n)
        p_ref, lw_ref, _, _ = orc.resample("stratified", lw[sl], r[sl], sort=True)
        np.testing.assert_array_equal(p[sl], p_ref)
        np.testing.assert_array_equal(y1[sl], y0[sl][p_ref])
    assert np.all(pf.log_weights == 0.0)


@pytest.mark.parametrize("noise", ["lean", "philox53"])
def test_batch_shards_reproduce_the_whole_batch(g, noise):
    """Batch sharding (north star: "batches of independent filters shard with no communication at all", BASELINE
    config 5): two handles holding filters [0, 3) and [3, 8) of a batch give, filter for filter and bit for bit, what
    ONE handle holding all 8 gives -- every Philox counter is a global batch slot (genpf_filter_set_first_filter) --
    through per-filter resample decisions, mh, updates and a replicate + residual-resize cycle."""
    nf, n, T = 8, 4096, 12
    rng = np.random.default_rng(77)
    obs = np.cumsum(rng.normal(0, 0.3, (T + 1, nf)), axis=0)
    model = g.DeviceModel("object_motion")

    def run(f0, f1):
        st = g.pf_initialize(model, (1,), obs[0, f0:f1], n, n_filters=f1 - f0, seed=5, noise=noise, first_filter=f0)
        for t in range(2, T + 1):
            g.pf_step(st, t, obs[t - 2, f0:f1], obs[t - 1, f0:f1], ess_thresh=0.5, return_ess=False)
            if t == 6:
                g.pf_replicate(st, 2)
                g.pf_resize(st, n, "residual")
        return (st.field("y", T).reshape(f1 - f0, -1), st.field("moving", T).reshape(f1 - f0, -1),
                st.log_weights.reshape(f1 - f0, -1), np.atleast_1d(g.log_ml_estimate(st)))

    whole = run(0, nf)
    for f0, f1 in ((0, 3), (3, 8)):
        part = run(f0, f1)
        for a, b in zip(whole, part):
            np.testing.assert_array_equal(a[f0:f1], b)
    assert np.isfinite(whole[2]).all()
