"""CPU-only checks of the host mirror's pure-Python pieces (no kernel is launched): pf_introduce! (resize.jl:351-421),
choiceproduct (utils.jl:56-95), the stratified host pf_initialize layout (utils.jl:29-55, initialize.jl:93-108)."""
import math

import numpy as np
import pytest


class LineTrace:
    def __init__(self, model, args, choices, score):
        self.model, self.args, self.choices, self.score = model, args, dict(choices), score

    def __getitem__(self, addr):
        return self.choices[addr]

    def update(self, new_args, argdiffs, constraints):
        """Gen.update for the toy model: extend to new_args[0] points; constrained y's are scored, an `outlier_i`
        flag may be constrained too (prior 0.1), nothing is ever discarded."""
        (n_new,) = new_args
        ch, incr = dict(self.choices), 0.0
        slope = ch["slope"]
        for i in range(self.args[0] + 1, n_new + 1):
            flag = constraints.get(("outlier", i), False)
            incr += math.log(0.1 if flag else 0.9) if ("outlier", i) in constraints else 0.0
            ch[("outlier", i)] = flag
            y = constraints.get(("y", i), slope * i)
            sd = 10.0 if flag else 1.0
            if ("y", i) in constraints:
                incr += -0.5 * (((y - slope * i) / sd) ** 2 + math.log(2 * math.pi)) - math.log(sd)
            ch[("y", i)] = y
        return LineTrace(self.model, new_args, ch, self.score + incr), incr, None, {}


class LineModel:
    """slope ~ uniform_discrete(-2, 2); y_i ~ normal(slope * i, 1): the reference tests' line_model in miniature."""

    def generate(self, args, constraints):
        (n,) = args
        rng = np.random.default_rng(abs(hash(tuple(sorted(constraints.items(), key=repr)))) % (2 ** 32))
        w = 0.0
        if "slope" in constraints:
            slope = constraints["slope"]
            w += -math.log(5.0)
        else:
            slope = int(rng.integers(-2, 3))
        ch = {"slope": slope}
        for i in range(1, n + 1):
            key = ("y", i)
            if key in constraints:
                y = constraints[key]
                w += -0.5 * ((y - slope * i) ** 2 + math.log(2 * math.pi))
            else:
                y = slope * i + rng.normal()
            ch[key] = y
        return LineTrace(self, args, ch, w), w


@pytest.fixture(scope="module")
def api():
    import genpf_b200 as g
    return g


def test_choiceproduct(api):
    strata = list(api.choiceproduct(("a", [1, 2]), ("b", [3])))
    assert strata == [{"a": 1, "b": 3}, {"a": 2, "b": 3}]
    assert list(api.choiceproduct({"slope": [-1, 0, 1]})) == [{"slope": -1}, {"slope": 0}, {"slope": 1}]
    assert len(list(api.choiceproduct(("a", [1, 2, 3]), ("b", [0, 1])))) == 6


@pytest.mark.parametrize("layout", ["contiguous", "interleaved"])
def test_host_stratified_initialize(api, layout):
    """test/initialize.jl:39-64: with no observations every weight is log p(slope) + log K = 0; strata layout."""
    model = LineModel()
    strata = list(api.choiceproduct(("slope", [-2, -1, 0, 1, 2])))
    state = api.pf_initialize(model, (0,), {}, 100, strata=strata, layout=layout)
    assert np.allclose(state.log_weights, 0.0)
    state = api.pf_initialize(model, (1,), {("y", 1): 0.0}, 100, strata=strata, layout=layout)
    for k, slope in enumerate(range(-2, 3)):
        idx = range(20 * k, 20 * (k + 1)) if layout == "contiguous" else range(k, 100, 5)
        assert all(state.traces[i]["slope"] == slope and state.traces[i][("y", 1)] == 0.0 for i in idx)
    # left-over particles (103 = 5 * 20 + 3) take strata drawn with replacement
    state = api.pf_initialize(model, (0,), {}, 103, strata=strata, layout=layout)
    assert len(state.traces) == 103 and all(tr["slope"] in range(-2, 3) for tr in state.traces[100:])


def test_pf_introduce(api):
    """resize.jl:351-378: log_ml_est folded into the old weights, new traces appended with their generate weights."""
    model = LineModel()
    obs = {("y", 1): 0.5}
    state = api.pf_initialize(model, (1,), obs, 10)
    lw0 = state.log_weights.copy()
    state.log_ml_est = -1.25
    api.pf_introduce(state, None, None, obs, 5)
    assert len(state.traces) == 15 and len(state.log_weights) == 15 and state.log_ml_est == 0.0
    np.testing.assert_allclose(state.log_weights[:10], lw0 - 1.25)
    assert all(tr[("y", 1)] == 0.5 for tr in state.traces[10:])
    # custom proposal: weight = model weight - proposal score (resize.jl:410-413)
    api.pf_introduce(state, model, (1,), obs, 3, proposal=lambda: ({"slope": 1}, math.log(0.5)))
    assert len(state.traces) == 18 and all(tr["slope"] == 1 for tr in state.traces[15:])
    expect = -math.log(5.0) - 0.5 * ((0.5 - 1.0) ** 2 + math.log(2 * math.pi)) - math.log(0.5)
    np.testing.assert_allclose(state.log_weights[15:], expect)
    assert api.get_traces(state)[0] is state.traces[0] and api.get_log_weights(state).shape == (18,)


@pytest.mark.parametrize("layout", ["interleaved", "contiguous"])
def test_host_stratified_update(api, layout):
    """update.jl:193-210: every particle is updated under merge(stratum, observations) and gains log(n_strata)."""
    model = LineModel()
    state = api.pf_initialize(model, (1,), {("y", 1): 0.0}, 100, strata=[{"slope": 1}])
    lw0 = state.log_weights.copy()
    strata = [{("outlier", 2): False}, {("outlier", 2): True}]
    api.pf_update(state, (2,), None, {("y", 2): 2.5}, strata=strata, layout=layout)
    for k, flag in enumerate((False, True)):
        idx = list(range(50 * k, 50 * (k + 1))) if layout == "contiguous" else list(range(k, 100, 2))
        assert all(state.traces[i][("outlier", 2)] is flag for i in idx)
        sd = 10.0 if flag else 1.0
        incr = math.log(0.1 if flag else 0.9) - 0.5 * (((2.5 - 2.0) / sd) ** 2 + math.log(2 * math.pi)) - math.log(sd)
        np.testing.assert_allclose(state.log_weights[idx], lw0[idx] + incr + math.log(2))
