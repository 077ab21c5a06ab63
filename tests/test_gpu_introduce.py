"""GPU: pf_introduce! on device filters (resize.jl:351-421) -- new chains generated under the whole observation
history, against the CPU oracle with supplied noise; existing particles keep their place with log_ml_est folded in."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_chains(orc, obs, U, Z):
    """obs[tau-1] = y_tau; U, Z: (T, m).  Returns (y_T, y_{T-1}, moving_T, moving_{T-1}, lw)."""
    T = len(obs)
    y, mv = orc.om_transition(None, None, math.sin(1.0), U[0], Z[0])
    w = orc.om_obs_logpdf(y, obs[0])
    yp, mp = np.zeros_like(y), np.zeros_like(mv)
    for tau in range(2, T + 1):
        yp, mp = y, mv
        y, mv = orc.om_transition(yp, mp, math.sin(float(tau)), U[tau - 1], Z[tau - 1])
        w = orc.om_obs_logpdf(y, obs[tau - 1], lw=w)
    return y, yp, mv, mp, w


@pytest.mark.parametrize("nf", [1, 3])
def test_introduce_vs_oracle(g, orc, nf):
    L, lib = g._lib, g.load()
    n, m, T = 5000, 777, 4
    rng = np.random.default_rng(12)
    obs = rng.normal(0, 0.5, (T, nf))
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), obs[0], n, n_filters=nf, seed=4)
    for t in range(2, T + 1):
        g.pf_update(pf, (t,), None, obs[t - 1])
        if t == 3:
            g.pf_resample(pf, "stratified", sort_particles=False)  # log_ml_est != 0 from here on
    lw0, y0, yp0, m0 = pf.log_weights, pf.field("y", T), pf.field("y", T - 1), pf.field("moving", T)
    lml0 = np.atleast_1d(pf.log_ml_est)  # the accumulated field (resample.jl:178-182), not the total estimate
    total0 = np.atleast_1d(g.log_ml_estimate(pf))
    assert np.all(lml0 != 0.0)
    U, Z = rng.random((T, nf, m)), rng.normal(size=(T, nf, m))
    aux = np.concatenate([model.aux(tau) for tau in range(1, T + 1)])
    L.check(lib.genpf_introduce(pf._h, m, L.ptr(np.ascontiguousarray(obs)), L.ptr(aux), 0, L.ptr(np.ascontiguousarray(U)),
                                L.ptr(np.ascontiguousarray(Z))))
    N = n + m
    assert len(pf) == N
    lw1, y1, yp1, m1 = pf.log_weights, pf.field("y", T), pf.field("y", T - 1), pf.field("moving", T)
    # log_ml_estimate(state) = log_ml_est + logsumexp(lw) - log(n) (utils.jl): the folded estimate must now come from
    # the weights alone
    for f in range(nf):
        old, new = slice(f * N, f * N + n), slice(f * N + n, (f + 1) * N)
        o0 = slice(f * n, (f + 1) * n)
        np.testing.assert_array_equal(y1[old], y0[o0])
        np.testing.assert_array_equal(yp1[old], yp0[o0])
        np.testing.assert_array_equal(m1[old], m0[o0])
        np.testing.assert_array_equal(lw1[old], lw0[o0] + lml0[f])  # log_weights .+= log_ml_est (resize.jl:363)
        y, yp, mv, mp, w = _oracle_chains(orc, obs[:, f], U[:, f], Z[:, f])
        np.testing.assert_array_equal(y1[new], y)
        np.testing.assert_array_equal(yp1[new], yp)
        np.testing.assert_array_equal(m1[new], mv)
        np.testing.assert_allclose(lw1[new], w, rtol=1e-10, atol=1e-12)
        assert np.atleast_1d(pf.log_ml_est)[f] == 0.0  # log_ml_est = 0 (resize.jl:364)
        lml_now = np.atleast_1d(g.log_ml_estimate(pf))[f]
        assert lml_now == pytest.approx(orc.logsumexp(lw1[f * N:(f + 1) * N]) - math.log(N), rel=1e-10)
        # the old half alone still carries the estimate it had before: lse(lw + lml) - log n == total0
        assert orc.logsumexp(lw1[old]) - math.log(n) == pytest.approx(total0[f], rel=1e-10)
    # the filter goes on: a README iteration over the enlarged population
    ess = g.pf_step(pf, T + 1, obs[T - 1], obs[T - 1] + 0.1, ess_thresh=1.0)
    assert np.isfinite(ess).all() and len(pf) == N and np.isfinite(pf.log_weights).all()


@pytest.mark.parametrize("noise", ["lean", "philox53"])
def test_introduce_library_noise_is_a_prior_chain(g, orc, noise):
    """Without resampling the old particles are prior chains too: both halves estimate the same marginal likelihood."""
    n, T = 200_000, 3
    rng = np.random.default_rng(2)
    obs = rng.normal(0, 0.4, T)
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), float(obs[0]), n, seed=8, noise=noise)
    for t in range(2, T + 1):
        g.pf_update(pf, (t,), None, float(obs[t - 1]))
    g.pf_introduce(pf, None, None, list(obs), n)
    lw = pf.log_weights
    assert len(pf) == 2 * n and np.isfinite(lw).all()
    assert orc.logsumexp(lw[n:]) == pytest.approx(orc.logsumexp(lw[:n]), abs=0.05)
    q = np.linspace(0.05, 0.95, 19)  # sanity only (quantile s.e. ~ 0.005 here): exact parity is the noise-column test above
    y = pf.field("y", T)
    assert np.abs(np.quantile(y[n:], q) - np.quantile(y[:n], q)).max() < 0.03  # same marginal of y_T
    assert abs(pf.field("moving", T)[n:].mean() - pf.field("moving", T)[:n].mean()) < 0.01


def test_introduce_more_particles_than_one_launch_wave(g, orc):
    """The kernel's grid is capped (grid-stride loop): 1.5 M new chains against the oracle at both ends of the range."""
    L, lib = g._lib, g.load()
    n, m, T = 4096, 1_500_000, 2
    rng = np.random.default_rng(3)
    obs = rng.normal(0, 0.5, (T, 1))
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), obs[0], n, seed=4)
    g.pf_update(pf, (2,), None, obs[1])
    U, Z = rng.random((T, 1, m)), rng.normal(size=(T, 1, m))
    aux = np.concatenate([model.aux(tau) for tau in range(1, T + 1)])
    L.check(lib.genpf_introduce(pf._h, m, L.ptr(np.ascontiguousarray(obs)), L.ptr(aux), 0, L.ptr(np.ascontiguousarray(U)),
                                L.ptr(np.ascontiguousarray(Z))))
    y, yp, mv, mp, w = _oracle_chains(orc, obs[:, 0], U[:, 0], Z[:, 0])
    np.testing.assert_array_equal(pf.field("y", T)[n:], y)
    np.testing.assert_array_equal(pf.field("moving", T)[n:], mv)
    np.testing.assert_allclose(pf.log_weights[n:], w, rtol=1e-10, atol=1e-12)


def test_introduce_with_proposal_and_errors(g, orc):
    A, Q, R, M0, S0 = 0.9, 1.0, 0.8, 0.1, 1.3
    model = g.DeviceModel("lingauss1d", (A, Q, R, M0, S0))
    obs = [0.4, -0.7, 2.5]
    pf = g.pf_initialize(model, (1,), obs[0], 4096, seed=2)
    for t in (2, 3):
        g.pf_update(pf, (t,), None, obs[t - 1])
    g.pf_introduce(pf, None, None, obs, 1000, proposal=True)
    lw = pf.log_weights
    assert len(pf) == 5096 and np.isfinite(lw).all()
    # the locally optimal proposal's weight is prod_t p(y_t | x_{t-1}): far less spread than the prior chains' weights
    assert np.std(lw[4096:]) < np.std(lw[:4096])
    with pytest.raises(ValueError, match="needs 3 observations"):
        g.pf_introduce(pf, None, None, obs[:2], 10)
    with pytest.raises(g.GenPFError, match="n_particles must be >= 1"):
        g._lib.check(g.load().genpf_introduce(pf._h, 0, None, None, 0, None, None))
