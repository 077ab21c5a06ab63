"""GPU: models registered at run time (genpf_model_compile) and the optional plugin members -- custom proposals
(initialize.jl:46-62, update.jl:79-96, rejuvenate.jl:134-148) and trace translators (update.jl:35-44)."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-10
A, Q, R, M0, S0 = 0.9, 1.0, 0.8, 0.1, 1.3
P8 = (A, Q, R, M0, S0, math.log(R), math.sqrt(A * A * S0 * S0 + Q * Q), 1.0 / R)


@pytest.fixture(scope="module")
def plugin(g):
    src = open(os.path.join(ROOT, "tests", "plugins", "my_lingauss.cu")).read()
    return g.DeviceModel.from_source("my_lingauss", src, "MyLinGauss", fields={"x": 0}, params=P8)


@pytest.mark.parametrize("noise", ["lean", "philox53"])
def test_plugin_from_source_reproduces_builtin_bit_for_bit(g, plugin, noise):
    """register lingauss1d from source at run time and reproduce the built-in model exactly --
    same kernels (k_propagate, k_mh, k_step_fused), same Philox streams, through the README loop."""
    n = 50_000
    rng = np.random.default_rng(1)
    obs = rng.normal(0, 1, 8)
    ref = g.pf_initialize(g.DeviceModel("lingauss1d", (A, Q, R, M0, S0)), (1,), obs[0], n, seed=11, noise=noise)
    dyn = g.pf_initialize(plugin, (1,), obs[0], n, seed=11, noise=noise)

    def same():
        np.testing.assert_array_equal(dyn.log_weights, ref.log_weights)
        np.testing.assert_array_equal(dyn.field("x", dyn.t), ref.field("x", ref.t))
        assert g.effective_sample_size(dyn) == g.effective_sample_size(ref)

    same()
    for t in range(2, 6):  # fused step (k_scan_hot + k_step_fused from the plugin image)
        for s in (ref, dyn):
            g.pf_step(s, t, obs[t - 2], obs[t - 1], method="stratified", ess_thresh=1.0)
        np.testing.assert_array_equal(dyn.parents, ref.parents)
        same()
    for s in (ref, dyn):  # separate kernels: residual resample, 2 mh sweeps, update; then mh_iters = 0 fused
        g.pf_resample(s, "residual", sort_particles=False)
        g.pf_rejuvenate(s, g.mh, (5, obs[4]), 2)
        g.pf_update(s, (6,), None, obs[5])
        g.pf_step(s, 7, obs[5], obs[6], method="stratified", ess_thresh=1.0, mh_iters=0)
    same()
    np.testing.assert_array_equal(dyn.accepts, ref.accepts)
    assert g.log_ml_estimate(dyn) == g.log_ml_estimate(ref)
    for s in (ref, dyn):  # pf_introduce!: k_introduce from the plugin image
        g.pf_introduce(s, None, None, list(obs[:7]), 1234)
    assert len(dyn) == n + 1234
    same()
    assert g.mean(dyn, (7, "x")) == g.mean(ref, (7, "x"))


def test_plugin_translator_update(g, orc, plugin):
    """pf_update!(state, translator) (update.jl:35-44): (new slice, increment) from the plugin's translate."""
    n = 10_000
    pf = g.pf_initialize(plugin, (1,), 0.3, n, seed=5)
    x1, lw1 = pf.field("x", 1), pf.log_weights
    g.pf_update(pf, (2,), None, 1.7, translator=True)
    x2 = 2.0 * x1 + 1.0
    np.testing.assert_array_equal(pf.field("x", 2), x2)
    np.testing.assert_allclose(pf.log_weights, lw1 - 0.5 * (1.7 - x2) * (1.7 - x2), rtol=1e-12)
    assert g.effective_sample_size(pf) == pytest.approx(orc.ess(pf.log_weights), rel=RTOL)
    with pytest.raises(g.GenPFError, match="no translator"):
        g.pf_update(g.pf_initialize(g.DeviceModel("lingauss1d"), (1,), 0.0, 64), (2,), None, 0.0, translator=True)
    with pytest.raises(g.GenPFError, match="no custom proposal"):
        g.pf_update(pf, (3,), None, 0.0, proposal=True)


def _lg_opt(x_prev, y, sig):
    var = 1.0 / (1.0 / sig ** 2 + 1.0 / R ** 2)
    return var * (A * x_prev / sig ** 2 + y / R ** 2), math.sqrt(var)


def test_custom_proposal_lingauss_vs_closed_form(g, orc):
    """Custom proposals with supplied noise: x ~ q, lw += log p(x|prev) + log p(y|x) - log q(x); for the locally
    optimal proposal of the linear-Gaussian model that increment equals log p(y_t | x_{t-1}) for every particle."""
    L, lib = g._lib, g.load()
    n = 20_000
    rng = np.random.default_rng(3)
    model = g.DeviceModel("lingauss1d", (A, Q, R, M0, S0))
    pf = g.DevicePFState(model, n, seed=1)
    sig1 = math.sqrt(A * A * S0 * S0 + Q * Q)
    Z = rng.normal(size=n)
    U = np.zeros(n)
    L.check(lib.genpf_initialize_proposal(pf._h, L.ptr(pf._obs(0.4)), None, L.ptr(U), L.ptr(Z)))
    pf.t = 1
    mu, sd = _lg_opt(np.full(n, M0), 0.4, sig1)
    x1 = mu + sd * Z
    np.testing.assert_allclose(pf.field("x", 1), x1, rtol=1e-13, atol=1e-15)  # the device contracts mu + sd*Z into an fma
    lw1 = orc.normal_logpdf(x1, A * M0, sig1) + orc.normal_logpdf(0.4, x1, R) - orc.normal_logpdf(x1, mu, sd)
    np.testing.assert_allclose(pf.log_weights, lw1, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(pf.log_weights, orc.normal_logpdf(0.4, A * M0, math.sqrt(sig1 ** 2 + R ** 2)), rtol=1e-9)
    Z = rng.normal(size=n)
    x1 = pf.field("x", 1)
    L.check(lib.genpf_update_proposal(pf._h, 2, L.ptr(pf._obs(-0.7)), None, L.ptr(U), L.ptr(Z)))
    pf.t = 2
    mu, sd = _lg_opt(x1, -0.7, Q)
    x2 = mu + sd * Z
    np.testing.assert_allclose(pf.field("x", 2), x2, rtol=1e-13, atol=1e-15)
    inc = orc.normal_logpdf(-0.7, A * x1, math.sqrt(Q * Q + R * R))
    np.testing.assert_allclose(pf.log_weights, lw1 + inc, rtol=1e-9, atol=1e-9)
    # move_reweight(trace, proposal, ...) (rejuvenate.jl:134-148): both importance weights equal p(y | x_{t-1}), so
    # the relative weight vanishes for the optimal proposal while the slice is re-proposed
    Z2 = rng.normal(size=n)
    lw_before = pf.log_weights
    L.check(lib.genpf_rejuvenate_reweight_proposal(pf._h, 2, L.ptr(pf._obs(-0.7)), None, 1, L.ptr(U), L.ptr(Z2)))
    np.testing.assert_allclose(pf.field("x", 2), mu + sd * Z2, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(pf.log_weights, lw_before, rtol=1e-9, atol=1e-8)
    # library noise through the reference-shaped calls: the optimal proposal beats the bootstrap filter's ESS
    obs = [0.4, -0.7, 2.5, 2.9]
    boot = g.pf_initialize(model, (1,), obs[0], n, seed=2)
    opt = g.pf_initialize(model, (1,), obs[0], n, seed=2, proposal=True)
    for t in range(2, 5):
        g.pf_update(boot, (t,), None, obs[t - 1])
        g.pf_update(opt, (t,), None, obs[t - 1], proposal=True)
    assert g.effective_sample_size(opt) > g.effective_sample_size(boot)
    assert g.log_ml_estimate(opt) == pytest.approx(g.log_ml_estimate(boot), abs=0.1)
    g.pf_move_reweight(opt, g.move_reweight, (4, obs[3]), 1, proposal=True)
    assert np.isfinite(opt.log_weights).all()


def test_custom_proposal_object_motion(g, orc):
    """object_motion's proposal flips a fair coin for `moving` instead of the sticky prior: weight = p(m') / 0.5."""
    L, lib = g._lib, g.load()
    n = 20_000
    rng = np.random.default_rng(8)
    model = g.DeviceModel("object_motion")
    pf = g.DevicePFState(model, n, seed=1)
    U, Z = rng.random(n), rng.normal(size=n)
    L.check(lib.genpf_initialize_proposal(pf._h, L.ptr(pf._obs(0.2)), L.ptr(model.aux(1)), L.ptr(U), L.ptr(Z)))
    pf.t = 1
    m1 = (U < 0.5).astype(np.uint8)
    y1 = (0.0 + np.where(m1 == 1, math.sin(1.0), 0.0)) + 0.01 * Z
    np.testing.assert_array_equal(pf.field("moving", 1), m1)
    np.testing.assert_array_equal(pf.field("y", 1), y1)
    lw1 = np.log(np.where(m1 == 1, 0.25, 0.75)) - math.log(0.5) + orc.om_obs_logpdf(y1, 0.2)
    np.testing.assert_allclose(pf.log_weights, lw1, rtol=1e-9, atol=1e-9)
    U, Z = rng.random(n), rng.normal(size=n)
    L.check(lib.genpf_update_proposal(pf._h, 2, L.ptr(pf._obs(0.9)), L.ptr(model.aux(2)), L.ptr(U), L.ptr(Z)))
    pf.t = 2
    m2 = (U < 0.5).astype(np.uint8)
    y2 = (y1 + np.where(m2 == 1, math.sin(2.0), 0.0)) + 0.01 * Z
    pm = np.where(m1 == 1, 0.75, 0.25)
    lw2 = lw1 + np.log(np.where(m2 == 1, pm, 1 - pm)) - math.log(0.5) + orc.om_obs_logpdf(y2, 0.9)
    np.testing.assert_array_equal(pf.field("y", 2), y2)
    np.testing.assert_allclose(pf.log_weights, lw2, rtol=1e-9, atol=1e-9)
    # same target: the proposal-based filter and the bootstrap filter agree on the evidence
    a = g.pf_initialize(model, (1,), 0.2, 200_000, seed=3)
    b = g.pf_initialize(model, (1,), 0.2, 200_000, seed=4, proposal=True)
    assert g.log_ml_estimate(b) == pytest.approx(g.log_ml_estimate(a), abs=0.03)
