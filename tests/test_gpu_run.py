"""GPU: genpf_run_steps / pf_run -- T README iterations (README.md:66-77) enqueued by one asynchronous call, eager
or replayed as a CUDA graph, against the same iterations issued one genpf_step at a time (each filter decides
ess < ess_thresh * n on the device in both) and, for the forced-resample case, against the CPU oracle's loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _snapshot(g, pf, t):
    return (pf.field("y", t), pf.field("y", t - 1), pf.field("moving", t), pf.log_weights,
            np.atleast_1d(g.log_ml_estimate(pf)), np.atleast_1d(g.effective_sample_size(pf)))


@pytest.mark.parametrize("nf,n", [(1, 100), (1, 70_001), (6, 4096)])
@pytest.mark.parametrize("ess_thresh", [0.5, 1.0])
def test_run_steps_equals_step_loop(g, nf, n, ess_thresh):
    T = 12
    rng = np.random.default_rng(5)
    obs = rng.normal(0, 0.5, (T + 1, nf))
    model = g.DeviceModel("object_motion")
    snaps = []
    for mode in ("loop", "eager", "graph"):
        pf = g.pf_initialize(model, (1,), obs[0], n, n_filters=nf, seed=9) if nf > 1 else \
            g.pf_initialize(model, (1,), float(obs[0, 0]), n, seed=9)
        if mode == "loop":
            for r in range(1, T + 1):
                g.pf_step(pf, 1 + r, obs[r - 1], obs[r], ess_thresh=ess_thresh, return_ess=False)
        else:
            g.pf_run(pf, 2, obs, ess_thresh=ess_thresh, graph=(mode == "graph"))
        assert pf.t == T + 1
        snaps.append(_snapshot(g, pf, T + 1))
    for other in snaps[1:]:
        for a, b in zip(snaps[0], other):
            np.testing.assert_array_equal(a, b)
    assert np.isfinite(snaps[0][3]).all() and np.isfinite(snaps[0][4]).all()


def test_run_steps_twice_and_continue(g):
    """A second graph run replaces the first one's graph; a plain step afterwards continues from its state."""
    n, T = 5000, 5
    rng = np.random.default_rng(6)
    obs = rng.normal(0, 0.5, 2 * T + 2)
    model = g.DeviceModel("object_motion")
    a = g.pf_initialize(model, (1,), float(obs[0]), n, seed=3)
    b = g.pf_initialize(model, (1,), float(obs[0]), n, seed=3)
    g.pf_run(a, 2, obs[:T + 1], graph=True)
    g.pf_run(a, T + 2, obs[T:2 * T + 1], graph=True)
    g.pf_step(a, 2 * T + 2, obs[2 * T], obs[2 * T + 1], return_ess=False)
    for r in range(1, 2 * T + 2):
        g.pf_step(b, 1 + r, obs[r - 1], obs[r], return_ess=False)
    for x, y in zip(_snapshot(g, a, 2 * T + 2), _snapshot(g, b, 2 * T + 2)):
        np.testing.assert_array_equal(x, y)


def test_run_steps_argument_errors(g):
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), 0.1, 256, seed=1)
    with pytest.raises(g.GenPFError, match="t_cur"):
        g.pf_run(pf, 5, np.zeros(3))
    with pytest.raises(g.GenPFError, match="fused stratified"):
        g.pf_run(pf, 2, np.zeros(3), method="residual")
