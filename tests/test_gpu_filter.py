"""GPU parity tests of the device-resident path (plugin models): pf_initialize / pf_update! /
pf_resample! / pf_rejuvenate!(mh) / mean / var / resizing, against the CPU oracle with the SAME noise
columns (parity mode) and against closed forms (README posterior flip, Kalman filter)."""
import ctypes as C
import math

import numpy as np
import pytest

from util import check_parents, strat_u

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def noisy_init(g, pf, obs, U, Z):
    L = g._lib
    L.check(g.load().genpf_initialize_with_noise(pf._h, L.ptr(pf._obs(obs)), L.ptr(pf.model.aux(1)), L.ptr(U), L.ptr(Z)))
    pf.t = 1


def noisy_update(g, pf, t, obs, U, Z):
    L = g._lib
    L.check(g.load().genpf_update_with_noise(pf._h, t, L.ptr(pf._obs(obs)), L.ptr(pf.model.aux(t)), L.ptr(U), L.ptr(Z)))
    pf.t = t


def noisy_mh(g, pf, tau, obs, U2, Z2, U3):
    L = g._lib
    acc = np.zeros(pf.n_filters, dtype=np.int64)
    L.check(g.load().genpf_rejuvenate_mh_with_noise(pf._h, tau, L.ptr(pf._obs(obs)), L.ptr(pf.model.aux(tau)),
                                                    L.ptr(U2), L.ptr(Z2), L.ptr(U3), L.ptr(acc)))
    return acc


@pytest.mark.parametrize("n", [100, 2048, 5000, (1 << 17) + 7])
def test_object_motion_parity_with_noise(g, orc, n):
    """Full README-style sequence with supplied noise: state columns bit-identical, log-weights 1e-10."""
    rng = np.random.default_rng(n)
    model = g.DeviceModel("object_motion")
    pf = g.DevicePFState(model, n, seed=1)
    T = 6
    obs = np.concatenate([rng.normal(0, 0.3, 3), np.sin(np.arange(4, T + 1)).cumsum() + rng.normal(0, 0.3, T - 3)])
    vel = [math.sin(float(t)) for t in range(0, T + 1)]
    # t = 1
    U, Z = rng.random(n), rng.normal(size=n)
    noisy_init(g, pf, obs[0], U, Z)
    y1, m1 = orc.om_transition(None, None, vel[1], U, Z)
    lw = orc.om_obs_logpdf(y1, obs[0])
    np.testing.assert_array_equal(pf.field("y", 1), y1)
    np.testing.assert_array_equal(pf.field("moving", 1), m1)
    np.testing.assert_allclose(pf.log_weights, lw, rtol=RTOL)
    ys, ms = {0: np.zeros(n), 1: y1}, {0: np.zeros(n, dtype=np.uint8), 1: m1}
    lml = 0.0
    for t in range(2, T + 1):
        # ESS
        assert g.effective_sample_size(pf) == pytest.approx(orc.ess(lw), rel=RTOL)
        assert g.log_ml_estimate(pf) == pytest.approx(lml + orc.logsumexp(lw) - math.log(n), rel=RTOL, abs=1e-9)
        # resample (alternate the three methods), then MH on tau = t-1, then update to t
        method = ["stratified", "residual", "multinomial"][t % 3]
        r = rng.random(n)
        g.pf_resample(pf, method, sort_particles=False, uniforms=r)
        p_ref, lw_new, inc, _ = orc.resample(method, lw, r)
        p = pf.parents
        W_ref = orc.cumweights(orc.softmax(lw))
        if method == "stratified":
            nm, _ = check_parents(p, p_ref, W_ref, strat_u(r, n))
        elif method == "multinomial":
            nm, _ = check_parents(p, p_ref, W_ref, r)
        else:
            nm = int(np.sum(p != p_ref))
        assert nm == 0  # no ties expected at these sizes
        lml += inc
        for tau in (t - 2, t - 1):
            ys[tau], ms[tau] = ys[tau][p_ref], ms[tau][p_ref]
        lw = lw_new
        np.testing.assert_array_equal(pf.field("y", t - 1), ys[t - 1])
        np.testing.assert_array_equal(pf.log_weights, lw)
        U2, Z2, U3 = rng.random(n), rng.normal(size=n), rng.random(n)
        acc = noisy_mh(g, pf, t - 1, obs[t - 2], U2, Z2, U3)
        yq, mq, a_ref = orc.om_mh(ys[t - 2] if t > 2 else None, ms[t - 2] if t > 2 else None, ys[t - 1], ms[t - 1],
                                  vel[t - 1], obs[t - 2], U2, Z2, U3)
        ys[t - 1], ms[t - 1] = yq, mq
        np.testing.assert_array_equal(pf.accepts, a_ref)  # trace changes iff accepted (test/rejuvenate.jl:30-50)
        assert acc[0] == a_ref.sum()
        np.testing.assert_array_equal(pf.field("y", t - 1), yq)
        np.testing.assert_array_equal(pf.field("moving", t - 1), mq)
        np.testing.assert_array_equal(pf.log_weights, lw)  # move-accept leaves weights alone
        U, Z = rng.random(n), rng.normal(size=n)
        noisy_update(g, pf, t, obs[t - 1], U, Z)
        ys[t], ms[t] = orc.om_transition(ys[t - 1], ms[t - 1], vel[t], U, Z)
        lw = orc.om_obs_logpdf(ys[t], obs[t - 1], lw)
        np.testing.assert_array_equal(pf.field("y", t), ys[t])
        np.testing.assert_array_equal(pf.field("moving", t), ms[t])
        np.testing.assert_allclose(pf.log_weights, lw, rtol=RTOL, atol=1e-12)
    m_gpu, v_gpu = g.mean(pf, (T, "y")), g.var(pf, (T, "y"))
    m_ref, v_ref = orc.mean_var(lw, ys[T])
    assert m_gpu == pytest.approx(m_ref, rel=RTOL) and v_gpu == pytest.approx(v_ref, rel=RTOL)
    m_gpu = g.mean(pf, (T, "moving"))  # Bool promoted to fp64 (README.md:97)
    assert m_gpu == pytest.approx(orc.mean_var(lw, ms[T].astype(float))[0], rel=RTOL, abs=1e-15)


def test_lingauss_parity_with_noise(g, orc):
    rng = np.random.default_rng(5)
    n = 10_000
    params = (0.9, 1.0, 1.0, 0.0, 1.0)
    model = g.DeviceModel("lingauss1d", params)
    pf = g.DevicePFState(model, n)
    sig1 = math.sqrt(0.9 ** 2 * 1.0 + 1.0)
    Z = rng.normal(size=n)
    noisy_init(g, pf, 0.3, np.zeros(n), Z)
    x1 = orc.lg_transition(np.zeros(n), Z, (0.9, sig1, 1.0, 0.0, 1.0))
    lw = orc.lg_obs_logpdf(x1, 0.3, params)
    np.testing.assert_array_equal(pf.field("x", 1), x1)
    np.testing.assert_allclose(pf.log_weights, lw, rtol=RTOL)
    Z = rng.normal(size=n)
    noisy_update(g, pf, 2, -0.2, np.zeros(n), Z)
    x2 = orc.lg_transition(x1, Z, params)
    lw = orc.lg_obs_logpdf(x2, -0.2, params, lw)
    np.testing.assert_array_equal(pf.field("x", 2), x2)
    np.testing.assert_allclose(pf.log_weights, lw, rtol=RTOL)
    Z2, U3 = rng.normal(size=n), rng.random(n)
    noisy_mh(g, pf, 2, -0.2, np.zeros(n), Z2, U3)
    xq, acc = orc.lg_mh(x1, x2, -0.2, Z2, U3, params)
    np.testing.assert_array_equal(pf.field("x", 2), xq)
    np.testing.assert_array_equal(pf.accepts, acc)


def readme_observations(T=10, seed=3):
    """README.md:87-89: moving = t > 5, one fixed draw."""
    rng = np.random.default_rng(seed)
    y, obs = 0.0, []
    for t in range(1, T + 1):
        y = y + (math.sin(t) if t > 5 else 0.0) + 0.01 * rng.normal()
        obs.append(y + 0.25 * rng.normal())
    return np.array(obs)


def run_readme_filter(g, n, seed, method="residual", keep_history=True, noise="lean"):
    """README.md:60-79."""
    obs = readme_observations()
    model = g.DeviceModel("object_motion")
    state = g.pf_initialize(model, (1,), obs[0], n, seed=seed, keep_history=keep_history, noise=noise)
    for t in range(2, len(obs) + 1):
        if g.effective_sample_size(state) < 0.5 * n:
            g.pf_resample(state, method, sort_particles=False)
            g.pf_rejuvenate(state, g.mh, (t - 1, obs[t - 2]))
        g.pf_update(state, (t,), None, obs[t - 1])
    return state


@pytest.mark.parametrize("noise", ["lean", "philox53"])
def test_readme_posterior_flip(g, noise):
    """README.md:97-104: P(moving) flips from low at t=5 to high at t=6 (their draw: 0.07 / 0.95)."""
    for n in (100, 100_000):
        flips = 0
        for seed in range(5):
            state = run_readme_filter(g, n, seed, noise=noise)
            m5, m6 = g.mean(state, (5, "moving")), g.mean(state, (6, "moving"))
            v5 = g.var(state, (5, "moving"))
            assert 0.0 <= m5 <= 1.0 and 0.0 <= m6 <= 1.0
            assert v5 == pytest.approx(m5 * (1 - m5), abs=1e-9)  # Bernoulli variance
            flips += (m5 < 0.5 < m6)
        assert flips >= 4, f"posterior flip seen in only {flips}/5 runs at n={n}"


def kalman(obs, a, q, r, m0, s0):
    m, P, lz = m0, s0 ** 2, 0.0
    for y in obs:
        mp, Pp = a * m, a * a * P + q * q
        S = Pp + r * r
        K = Pp / S
        lz += -0.5 * ((y - mp) ** 2 / S + math.log(2 * math.pi * S))
        m, P = mp + K * (y - mp), (1 - K) * Pp
    return m, P, lz


@pytest.mark.parametrize("method", ["stratified", "multinomial", "residual"])
def test_lingauss_kalman(g, method):
    """bootstrap PF on the linear-Gaussian tracker against the Kalman filter."""
    a, q, r, m0, s0 = 0.9, 1.0, 1.0, 0.0, 1.0
    rng = np.random.default_rng(2)
    T, n = 30, 1 << 18
    x, obs = rng.normal(m0, s0), []
    for _ in range(T):
        x = a * x + q * rng.normal()
        obs.append(x + r * rng.normal())
    model = g.DeviceModel("lingauss1d", (a, q, r, m0, s0))
    state = g.pf_initialize(model, (1,), obs[0], n, seed=9)
    for t in range(2, T + 1):
        g.pf_resample(state, method, sort_particles=False)
        g.pf_update(state, (t,), None, obs[t - 1])
    m, P, lz = kalman(obs, a, q, r, m0, s0)
    ess = g.effective_sample_size(state)
    se = math.sqrt(P / ess)
    assert abs(g.mean(state, (T, "x")) - m) < 6 * se
    assert g.var(state, (T, "x")) == pytest.approx(P, rel=0.05)
    assert g.log_ml_estimate(state) == pytest.approx(lz, abs=0.05)


def test_production_noise_distribution(g):
    """Philox noise checked distributionally: KS on Z, Bernoulli rate on U."""
    from scipy import stats
    n = 1 << 20
    for noise in ("lean", "philox53"):
        model = g.DeviceModel("lingauss1d", (0.0, 1.0, 1.0, 0.0, 0.0))  # x_1 = Z exactly
        state = g.pf_initialize(model, (1,), 0.0, n, seed=123, noise=noise)
        z = state.field("x", 1)
        assert stats.kstest(z, "norm").pvalue > 1e-3
        assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
        model = g.DeviceModel("object_motion")
        state = g.pf_initialize(model, (1,), 0.0, n, seed=7, noise=noise)
        frac = state.field("moving", 1).mean()
        assert abs(frac - 0.25) < 4 * math.sqrt(0.25 * 0.75 / n)
        # different seeds / steps give different streams
        s2 = g.pf_initialize(model, (1,), 0.0, n, seed=8, noise=noise)
        assert np.mean(s2.field("y", 1) == state.field("y", 1)) < 0.01


def test_production_noise_tails_and_lattice(g):
    """a KS test at 2^20 sees neither a truncated tail nor a coarse lattice.  At 2^26 draws the
    philox53 normals (53-bit uniforms, fp64 Box-Muller) must populate the tails like N(0,1) -- |Z| > 4, > 5, the
    maximum -- match the first four moments and take (almost) all-distinct values; the lean policy (24-bit radius
    uniform, fp32 transcendentals) is checked against its DOCUMENTED limits: |Z| <= sqrt(-2 log(2^-25)) = 5.89 and a
    2^24-point radius lattice, i.e. it is the throughput policy, not the production one."""
    n, reps = 1 << 24, 4
    model = g.DeviceModel("lingauss1d", (0.0, 1.0, 1.0, 0.0, 0.0))  # x_1 = Z exactly
    for noise in ("philox53", "lean"):
        z = np.concatenate([g.pf_initialize(model, (1,), 0.0, n, seed=1000 + k, noise=noise).field("x", 1)
                            for k in range(reps)])
        N = z.size
        a = np.abs(z)
        assert abs(z.mean()) < 5 / math.sqrt(N) and abs(z.var() - 1) < 5 * math.sqrt(2 / N)
        assert abs(np.mean(z ** 3)) < 5 * math.sqrt(15 / N) and abs(np.mean(z ** 4) - 3) < 5 * math.sqrt(96 / N)
        for thr, p in ((3.0, 2.6997960632601866e-03), (4.0, 6.334248366623973e-05), (5.0, 5.733031437583869e-07)):
            cnt, exp = int(np.sum(a > thr)), N * p
            if noise == "lean" and thr == 5.0:
                continue  # the 24-bit radius grid is too coarse out there (documented)
            assert abs(cnt - exp) < 5 * math.sqrt(exp) + 1, (noise, thr, cnt, exp)
        distinct = np.unique(z).size / N
        if noise == "philox53":
            assert 5.2 < a.max() < 6.8            # E[max |Z|] of 2^26 draws = 5.6; 53-bit uniforms reach 8.57
            assert distinct > 0.999               # 2^26 draws from a 2^53 x 2^53 grid: collisions are birthday-rare
        else:
            assert a.max() <= 5.8871 + 1e-3       # sqrt(-2 log((0 + 0.5) 2^-24))
            assert 0.5 < distinct < 0.9           # fp32 values: 2^26 draws on a 2^23-per-binade lattice (measured 0.67)


def test_multinomial_counts_chi_square(g):
    """The reference draws multinomial ancestors with rand(Categorical(w), n) (resample.jl:59; an alias sampler, not
    reproducible from uniforms), the library by inverse CDF of its Philox uniforms: the two must agree in
    DISTRIBUTION.  Chi-square of the ancestor counts of 2^20 draws against n*w over 256 particles, device and
    host-array paths, + independence of consecutive draws (lag-1 pairs)."""
    from scipy import stats
    n_src, n_out = 256, 1 << 20
    rng = np.random.default_rng(4)
    lw = rng.normal(0, 1.5, n_src)
    w = np.exp(lw - lw.max())
    w /= w.sum()
    for seed in (1, 2):
        state = g.ParticleFilterState(list(range(n_src)), lw.copy())
        g.pf_resize(state, n_out, "multinomial", seed=seed)
        p = np.asarray(state.parents)
        counts = np.bincount(p, minlength=n_src)
        chi2 = np.sum((counts - n_out * w) ** 2 / (n_out * w))
        assert stats.chi2.sf(chi2, n_src - 1) > 1e-4, chi2
        # lag-1 independence on a coarse 8 x 8 partition of the CDF
        edges = np.searchsorted(np.cumsum(w), np.linspace(0, 1, 9)[1:-1])
        cls = np.searchsorted(edges, p, side="right")
        pair = np.bincount(cls[:-1] * 8 + cls[1:], minlength=64).reshape(8, 8)
        pc = np.bincount(cls, minlength=8) / cls.size
        exp = np.outer(pc, pc) * (cls.size - 1)
        assert stats.chi2.sf(np.sum((pair - exp) ** 2 / exp), 49) > 1e-4


def test_step_equals_separate_calls(g):
    """genpf_step (one C call per README iteration) == ESS / resample / mh / update issued separately."""
    obs = readme_observations()
    n = 50_000
    model = g.DeviceModel("object_motion")
    for method in ("stratified", "residual", "multinomial"):
        a = g.pf_initialize(model, (1,), obs[0], n, seed=11)
        b = g.pf_initialize(model, (1,), obs[0], n, seed=11)
        for t in range(2, len(obs) + 1):
            ess_a = g.pf_step(a, t, obs[t - 2], obs[t - 1], method=method, ess_thresh=0.5)
            ess_b = g.effective_sample_size(b)
            assert ess_a[0] == ess_b
            if ess_b < 0.5 * n:
                g.pf_resample(b, method, sort_particles=False)
                g.pf_rejuvenate(b, g.mh, (t - 1, obs[t - 2]))
            g.pf_update(b, (t,), None, obs[t - 1])
            np.testing.assert_array_equal(a.log_weights, b.log_weights)
            np.testing.assert_array_equal(a.field("y", t), b.field("y", t))
            np.testing.assert_array_equal(a.field("moving", t - 1), b.field("moving", t - 1))
        assert g.log_ml_estimate(a) == g.log_ml_estimate(b)


@pytest.mark.parametrize("nf,n", [(7, 3000), (3, 3001), (5, 2049), (2, 1)])
def test_batched_filters_match_single(g, orc, nf, n):
    """n_filters independent filters (views, view.jl:16-48; config 5) == each filter run alone.
    Odd n exercises the unaligned (scalar) load/store paths of every tile kernel."""
    rng = np.random.default_rng(13)
    model = g.DeviceModel("object_motion")
    batch = g.DevicePFState(model, n, n_filters=nf)
    obs1, obs2 = rng.normal(0, 0.3, nf), rng.normal(0, 0.3, nf)
    U, Z = rng.random(nf * n), rng.normal(size=nf * n)
    U2, Z2 = rng.random(nf * n), rng.normal(size=nf * n)
    r = rng.random(nf * n)
    noisy_init(g, batch, obs1, U, Z)
    ess = np.atleast_1d(g.effective_sample_size(batch))
    g.pf_resample(batch, "stratified", sort_particles=False, uniforms=r)
    noisy_update(g, batch, 2, obs2, U2, Z2)
    lml = np.atleast_1d(g.log_ml_estimate(batch))
    mean_b = np.atleast_1d(g.mean(batch, (2, "y")))
    for f in range(nf):
        sl = slice(f * n, (f + 1) * n)
        one = g.DevicePFState(model, n)
        noisy_init(g, one, obs1[f], U[sl], Z[sl])
        assert g.effective_sample_size(one) == ess[f]
        g.pf_resample(one, "stratified", sort_particles=False, uniforms=r[sl])
        noisy_update(g, one, 2, obs2[f], U2[sl], Z2[sl])
        np.testing.assert_array_equal(batch.parents[sl], one.parents)  # local parents
        np.testing.assert_array_equal(batch.log_weights[sl], one.log_weights)
        np.testing.assert_array_equal(batch.field("y", 2)[sl], one.field("y", 2))
        assert g.log_ml_estimate(one) == lml[f]
        assert g.mean(one, (2, "y")) == mean_b[f]
    # the fused step on the batch (Philox noise keyed by the global slot f*n + i) == separate calls on the batch
    a = g.pf_initialize(model, (1,), obs1, n, n_filters=nf, seed=3)
    b = g.pf_initialize(model, (1,), obs1, n, n_filters=nf, seed=3)
    g.pf_step(a, 2, obs1, obs2, method="stratified", ess_thresh=1.0)
    g.pf_resample(b, "stratified", sort_particles=False)
    g.pf_rejuvenate(b, g.mh, (1, obs1))
    g.pf_update(b, (2,), None, obs2)
    np.testing.assert_array_equal(a.parents, b.parents)
    np.testing.assert_array_equal(a.log_weights, b.log_weights)
    np.testing.assert_array_equal(a.field("y", 2), b.field("y", 2))
    np.testing.assert_array_equal(a.field("moving", 1), b.field("moving", 1))


def test_device_resizing(g, orc):
    """pf_replicate! / pf_dereplicate! / pf_coalesce! / pf_resize! on device state (test/resize.jl)."""
    rng = np.random.default_rng(17)
    n, k = 1000, 4
    model = g.DeviceModel("object_motion")
    for layout in ("contiguous", "interleaved"):
        pf = g.pf_initialize(model, (1,), 0.1, n, seed=5)
        g.pf_update(pf, (2,), None, 0.2)
        lw0, y0, lml0 = pf.log_weights, pf.field("y", 2), g.log_ml_estimate(pf)
        g.pf_replicate(pf, k, layout=layout)
        assert len(pf) == n * k
        expect = np.repeat(np.arange(n), k) if layout == "contiguous" else np.tile(np.arange(n), k)
        np.testing.assert_array_equal(pf.parents, expect)
        np.testing.assert_array_equal(pf.field("y", 2), y0[expect])
        np.testing.assert_array_equal(pf.log_weights, lw0[expect])
        assert g.log_ml_estimate(pf) == pytest.approx(lml0, abs=1e-10)  # test/resize.jl:131,144
        g.pf_update(pf, (3,), None, 0.3)  # the replicated population keeps working
        g.pf_dereplicate(pf, k, layout=layout)
        assert len(pf) == n
        np.testing.assert_array_equal(pf.field("y", 2), y0)
    # coalesce: replicas are identical particles and collapse back (test/resize.jl:227-254)
    pf = g.pf_initialize(model, (1,), 0.1, n, seed=6)
    lml0 = g.log_ml_estimate(pf)
    y0 = pf.field("y", 1)
    g.pf_replicate(pf, 3)
    g.pf_coalesce(pf)
    assert len(pf) == len(np.unique(y0))
    assert g.log_ml_estimate(pf) == pytest.approx(lml0, abs=1e-6)
    # resize through residual / multinomial resampling (test/resize.jl:3-84)
    for method in ("multinomial", "residual"):
        for n_new in (n // 2, n + n // 2):
            pf = g.pf_initialize(model, (1,), 0.1, n, seed=7)
            lw0, y0, lml0 = pf.log_weights, pf.field("y", 1), g.log_ml_estimate(pf)
            u = rng.random(n_new)
            g.pf_resize(pf, n_new, method, uniforms=u)
            p_ref, lw_ref, _, _ = orc.resample(method, lw0, u, n_out=n_new)
            assert len(pf) == n_new
            np.testing.assert_array_equal(pf.parents, p_ref)
            np.testing.assert_array_equal(pf.field("y", 1), y0[p_ref])
            assert g.log_ml_estimate(pf) == pytest.approx(lml0, abs=1e-10)
            g.pf_update(pf, (2,), None, 0.0)


def test_device_optimal_resize(g, orc):
    """pf_resize!(state, n, :optimal) on device state == the host-array call on the same weights + a gather."""
    model = g.DeviceModel("object_motion")
    n = 5000
    for n_new in (n // 5, n // 2, n):
        pf = g.pf_initialize(model, (1,), 0.1, n, seed=21)
        g.pf_update(pf, (2,), None, 0.3)
        lw0, y0, m0 = pf.log_weights, pf.field("y", 2), pf.field("moving", 2)
        ref = orc.optimal_resize(lw0, n_new, 0.61)
        g.pf_resize(pf, n_new, "optimal", uniform=0.61)
        assert len(pf) == n_new
        p = pf.parents
        assert len(set(p.tolist())) == n_new
        np.testing.assert_array_equal(p[:ref["n_keep"]], ref["parents0"][:ref["n_keep"]])
        assert np.mean(p == ref["parents0"]) > 0.999
        np.testing.assert_array_equal(pf.field("y", 2), y0[p])
        np.testing.assert_array_equal(pf.field("moving", 2), m0[p])
        np.testing.assert_allclose(pf.log_weights, ref["lw_out"], rtol=1e-10)
        g.pf_update(pf, (3,), None, 0.2)  # the resized population keeps working
        assert np.isfinite(g.effective_sample_size(pf))


@pytest.mark.parametrize("layout", ["contiguous", "interleaved"])
def test_device_stratified_initialize(g, orc, layout):
    """Stratified pf_initialize (initialize.jl:93-108 + stratified_map!, utils.jl:29-55; test/initialize.jl:39-64)."""
    L = g._lib
    lib = g.load()
    n, K = 1000, 2
    rng = np.random.default_rng(12)
    model = g.DeviceModel("object_motion")
    pf = g.DevicePFState(model, n, seed=3)
    U, Z = rng.random(n), rng.normal(size=n)
    vals = np.array([0.0, 1.0])
    lay = L.LAYOUT_CONTIGUOUS if layout == "contiguous" else L.LAYOUT_INTERLEAVED
    L.check(lib.genpf_initialize_stratified(pf._h, L.ptr(pf._obs(0.3)), L.ptr(model.aux(1)), model.fields["moving"],
                                            L.ptr(vals), K, lay, L.ptr(U), L.ptr(Z)))
    pf.t = 1
    k = np.arange(n) // (n // K) if layout == "contiguous" else np.arange(n) % K
    m = vals[k].astype(np.uint8)
    vel = math.sin(1.0)
    y = (0.0 + np.where(m == 1, vel, 0.0)) + 0.01 * Z
    np.testing.assert_array_equal(pf.field("moving", 1), m)  # every particle sits in its stratum
    np.testing.assert_array_equal(pf.field("y", 1), y)
    lw = np.log(np.where(m == 1, 0.25, 0.75)) + orc.om_obs_logpdf(y, 0.3) + math.log(K)  # initialize.jl:104
    np.testing.assert_allclose(pf.log_weights, lw, rtol=RTOL)
    assert g.effective_sample_size(pf) == pytest.approx(orc.ess(lw), rel=RTOL)
    # same target as the plain initialisation: the log marginal likelihood estimates agree
    big = 200_000
    a = g.pf_initialize(model, (1,), 0.3, big, seed=5)
    b = g.pf_initialize(model, (1,), 0.3, big, seed=6, strata=("moving", [False, True]), layout=layout)
    assert g.log_ml_estimate(b) == pytest.approx(g.log_ml_estimate(a), abs=0.02)
    assert g.mean(b, (1, "moving")) == pytest.approx(g.mean(a, (1, "moving")), abs=0.01)
    g.pf_update(b, (2,), None, 0.4)  # the stratified population keeps working
    # left-over particles (n not a multiple of K) take strata drawn with replacement; a continuous latent as stratum
    grid = np.linspace(-0.5, 0.5, 5)
    c = g.pf_initialize(model, (1,), 0.3, 1003, seed=7, strata=("y", grid), layout=layout)
    yy = c.field("y", 1)
    kk = np.arange(1000) // 200 if layout == "contiguous" else np.arange(1000) % 5
    np.testing.assert_array_equal(yy[:1000], grid[kk])
    assert np.all(np.isin(yy[1000:], grid)) and np.isfinite(c.log_weights).all()
    # stratified update (update.jl:193-210): the constraint applies to the new slice, weights gain
    # log p(moving_2 | moving_1) + obs log-density + log K
    n2 = 1000
    pf2 = g.DevicePFState(model, n2, seed=9)
    U, Z = rng.random(n2), rng.normal(size=n2)
    noisy_init(g, pf2, 0.1, U, Z)
    y1, m1 = orc.om_transition(None, None, math.sin(1.0), U, Z)
    lw1 = orc.om_obs_logpdf(y1, 0.1)
    U, Z = rng.random(n2), rng.normal(size=n2)
    L.check(lib.genpf_update_stratified(pf2._h, 2, L.ptr(pf2._obs(0.5)), L.ptr(model.aux(2)), model.fields["moving"],
                                        L.ptr(vals), K, lay, L.ptr(U), L.ptr(Z)))
    pf2.t = 2
    k2 = np.arange(n2) // (n2 // K) if layout == "contiguous" else np.arange(n2) % K
    m2 = vals[k2].astype(np.uint8)
    y2 = (y1 + np.where(m2 == 1, math.sin(2.0), 0.0)) + 0.01 * Z
    p_move = np.where(m1 == 1, 0.75, 0.25)
    lw2 = lw1 + np.log(np.where(m2 == 1, p_move, 1.0 - p_move)) + orc.om_obs_logpdf(y2, 0.5) + math.log(K)
    np.testing.assert_array_equal(pf2.field("moving", 2), m2)
    np.testing.assert_array_equal(pf2.field("y", 2), y2)
    np.testing.assert_allclose(pf2.log_weights, lw2, rtol=RTOL)
    lg = g.pf_initialize(g.DeviceModel("lingauss1d", (0.9, 1.0, 1.0, 0.0, 1.0)), (1,), 0.2, 600, seed=8,
                         strata=("x", [-1.0, 0.0, 1.0]), layout=layout)
    x = lg.field("x", 1)
    assert set(np.unique(x)) == {-1.0, 0.0, 1.0}
    sig = math.sqrt(0.81 + 1.0)
    expect = orc.normal_logpdf(x, 0.0, sig) + orc.normal_logpdf(0.2, x, 1.0) + math.log(3)
    np.testing.assert_allclose(lg.log_weights, expect, rtol=1e-9)


@pytest.mark.parametrize("noise", ["lean", "philox53"])
def test_checkpoint_resume(g, tmp_path, noise):
    """save -> load continues bit-identically (fields, log-weights, ancestors, log_ml_est) to the uninterrupted run."""
    model = g.DeviceModel("object_motion")
    obs = [0.1, 0.2, 0.0, 0.4, 1.1, 1.6, 1.4]

    def advance(pf, ts):
        for t in ts:
            if g.effective_sample_size(pf) < 0.9 * len(pf):
                g.pf_resample(pf, "residual" if t % 2 else "stratified", sort_particles=False)
                g.pf_rejuvenate(pf, g.mh, (t - 1, obs[t - 2]))
            g.pf_update(pf, (t,), None, obs[t - 1])

    a = g.pf_initialize(model, (1,), obs[0], 3000, seed=31, noise=noise)
    advance(a, range(2, 5))
    path = str(tmp_path / "ckpt.npz")
    a.save(path)
    b = g.DevicePFState.load(path)
    assert len(b) == len(a) and b.t == a.t
    np.testing.assert_array_equal(b.log_weights, a.log_weights)
    assert g.log_ml_estimate(b) == g.log_ml_estimate(a)
    advance(a, range(5, 8))
    advance(b, range(5, 8))
    for name in ("y", "moving"):
        np.testing.assert_array_equal(b.field(name, 7), a.field(name, 7))
        np.testing.assert_array_equal(b.field(name, 6), a.field(name, 6))
    np.testing.assert_array_equal(b.log_weights, a.log_weights)
    np.testing.assert_array_equal(b.parents, a.parents)
    assert g.log_ml_estimate(b) == g.log_ml_estimate(a)


def test_device_move_reweight(g, orc):
    """pf_move_reweight!(state, move_reweight, (select(tau),)) (rejuvenate.jl:74-90,125-132) with supplied noise:
    slice tau regenerated from the conditional prior for every particle, log_weights += the regenerate weight."""
    L = g._lib
    n = 5000
    rng = np.random.default_rng(8)
    model = g.DeviceModel("object_motion")
    pf = g.DevicePFState(model, n, seed=2)
    vel = [math.sin(float(t)) for t in range(4)]
    U, Z = rng.random(n), rng.normal(size=n)
    noisy_init(g, pf, 0.1, U, Z)
    y1, m1 = orc.om_transition(None, None, vel[1], U, Z)
    U, Z = rng.random(n), rng.normal(size=n)
    noisy_update(g, pf, 2, 0.4, U, Z)
    y2, m2 = orc.om_transition(y1, m1, vel[2], U, Z)
    lw = orc.om_obs_logpdf(y2, 0.4, orc.om_obs_logpdf(y1, 0.1))
    U2, Z2 = rng.random(n), rng.normal(size=n)
    L.check(g.load().genpf_rejuvenate_reweight_with_noise(pf._h, 2, L.ptr(pf._obs(0.4)), L.ptr(pf.model.aux(2)),
                                                          L.ptr(U2), L.ptr(Z2)))
    yq, mq = orc.om_transition(y1, m1, vel[2], U2, Z2)
    lw_ref = lw + (orc.om_obs_logpdf(yq, 0.4) - orc.om_obs_logpdf(y2, 0.4))
    np.testing.assert_array_equal(pf.field("y", 2), yq)
    np.testing.assert_array_equal(pf.field("moving", 2), mq)
    np.testing.assert_allclose(pf.log_weights, lw_ref, rtol=RTOL, atol=1e-12)
    assert g.effective_sample_size(pf) == pytest.approx(orc.ess(lw_ref), rel=RTOL)
    # library noise through the reference-shaped call; weights stay a proper importance sample of the same target
    lml0 = g.log_ml_estimate(pf)
    g.pf_rejuvenate(pf, g.move_reweight, (2, 0.4), 2, method="reweight")
    assert np.isfinite(pf.log_weights).all() and not np.array_equal(pf.field("y", 2), yq)
    assert g.log_ml_estimate(pf) == pytest.approx(lml0, abs=0.5)


def test_device_proportionmap(g):
    """proportionmap(state, t => :moving) on device state equals mean(state, t => :moving) for a Bool field."""
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), 0.1, 30_000, seed=4)
    for t in range(2, 8):
        g.pf_update(pf, (t,), None, 0.4 * t)
    pm = g.proportionmap(pf, (7, "moving"))
    assert set(pm) <= {True, False} and sum(pm.values()) == pytest.approx(1.0, rel=1e-12)
    assert pm.get(True, 0.0) == pytest.approx(g.mean(pf, (7, "moving")), rel=1e-10)
    lw, mv = pf.log_weights, pf.field("moving", 7)
    w = np.exp(lw - lw.max())
    w /= w.sum()
    assert pm.get(False, 0.0) == pytest.approx(w[mv == 0].sum(), rel=1e-10)
    ys = g.proportionmap(pf, (7, "y"), max_values=8)  # continuous field: every particle its own value
    assert len(ys) == 8


def test_history_lineage(g):
    """mean(state, tau=>addr) for a slice that left the window is resolved through the ancestry log."""
    obs = readme_observations()
    n = 4000
    model = g.DeviceModel("object_motion")
    state = g.pf_initialize(model, (1,), obs[0], n, seed=2, keep_history=True)
    cols, par = {}, []
    for t in range(2, 8):
        g.pf_resample(state, "stratified", sort_particles=False)
        p = state.parents
        for tau in cols:
            cols[tau] = cols[tau][p]
        g.pf_rejuvenate(state, g.mh, (t - 1, obs[t - 2]))
        cols[t - 1] = state.field("moving", t - 1)
        g.pf_update(state, (t,), None, obs[t - 1])
    for tau in (1, 3, 5):
        np.testing.assert_array_equal(state.field("moving", tau), cols[tau])
        w = np.exp(state.log_weights - state.log_weights.max())
        w /= w.sum()
        assert g.mean(state, (tau, "moving")) == pytest.approx(float(np.sum(w * cols[tau])), rel=1e-9, abs=1e-12)
    plain = g.pf_initialize(model, (1,), obs[0], n, seed=2)
    g.pf_update(plain, (2,), None, obs[1])
    g.pf_update(plain, (3,), None, obs[2])
    with pytest.raises(g.GenPFError):
        plain.field("moving", 1)  # not resident without GENPF_KEEP_HISTORY


def test_device_priorities_and_check(g, orc):
    rng = np.random.default_rng(23)
    n = 20_000
    model = g.DeviceModel("object_motion")
    pf = g.pf_initialize(model, (1,), 0.4, n, seed=3)
    lw0, lml0 = pf.log_weights, g.log_ml_estimate(pf)
    u = rng.random(n)
    g.pf_resample(pf, "stratified", priority_fn=0.5, sort_particles=False, uniforms=u)  # w -> w/2
    p_ref, lw_ref, _, _ = orc.resample("stratified", lw0, u, lp=lw0 / 2)
    np.testing.assert_array_equal(pf.parents, p_ref)
    np.testing.assert_allclose(pf.log_weights, lw_ref, rtol=RTOL, atol=1e-11)
    assert g.log_ml_estimate(pf) == pytest.approx(lml0, abs=1e-9)
    pf.log_weights = np.full(n, -np.inf)
    with pytest.raises(g.GenPFErrorException, match="Invalid weights."):
        g.pf_resample(pf, "stratified", check=True)
    with pytest.warns(UserWarning, match="All input values are -Inf"):
        g.pf_resample(pf, "stratified", sort_particles=False)
    assert np.all(pf.log_weights == 0.0)
    with pytest.raises(g.GenPFErrorException, match="not recognized"):
        g.pf_resample(pf, "systematic")


def test_baseline_size_step_properties(g):
    """BASELINE size (2^24 particles): size-independent properties of the fused README iteration, and
    equality with the separately issued calls."""
    obs = readme_observations()
    n = 1 << 24
    model = g.DeviceModel("object_motion")
    a = g.pf_initialize(model, (1,), obs[0], n, seed=5)
    b = g.pf_initialize(model, (1,), obs[0], n, seed=5)
    for t in range(2, 5):
        ess = g.pf_step(a, t, obs[t - 2], obs[t - 1], method="stratified", ess_thresh=1.0)
        assert 1.0 <= ess[0] <= n
        p = a.parents
        assert p[0] >= 0 and p[-1] < n and np.all(np.diff(p) >= 0)  # stratified ancestors are sorted
        counts = np.bincount(p, minlength=n)
        assert counts.sum() == n
        lw = a.log_weights
        assert np.all(np.isfinite(lw))
        g.pf_resample(b, "stratified", sort_particles=False)
        g.pf_rejuvenate(b, g.mh, (t - 1, obs[t - 2]))
        g.pf_update(b, (t,), None, obs[t - 1])
        np.testing.assert_array_equal(p, b.parents)
        np.testing.assert_array_equal(lw, b.log_weights)
        np.testing.assert_array_equal(a.field("y", t), b.field("y", t))
    assert g.log_ml_estimate(a) == g.log_ml_estimate(b)
    # moving flag is a Bool column: mean in [0,1], var = m(1-m)
    m, v = g.mean(a, (4, "moving")), g.var(a, (4, "moving"))
    assert 0.0 <= m <= 1.0 and v == pytest.approx(m * (1 - m), abs=1e-9)
