#!/usr/bin/env python
"""Runs the non-headline BASELINE.json configs on one GPU and prints one JSON line each.
  config 1: README object_motion T=10, 100 particles, residual + MH when ESS < n/2  (latency + posterior flip)
  config 3: 1-D linear-Gaussian tracker, 2^24 particles, T=1000, stratified every step, Kalman check
  config 5 (single-GPU share): 512 independent object_motion filters x 4096 particles, MH every step
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def kalman(obs, a, q, r, m0, s0):
    m, P, lz = m0, s0 ** 2, 0.0
    for y in obs:
        mp, Pp = a * m, a * a * P + q * q
        S = Pp + r * r
        K = Pp / S
        lz += -0.5 * ((y - mp) ** 2 / S + math.log(2 * math.pi * S))
        m, P = mp + K * (y - mp), (1 - K) * Pp
    return m, P, lz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T3", type=int, default=1000)
    ap.add_argument("--n3", type=int, default=1 << 24)
    ap.add_argument("--filters5", type=int, default=512)
    args = ap.parse_args()
    import torch

    import genpf_b200 as g

    # ---- config 1
    rng = np.random.default_rng(3)
    y, obs = 0.0, []
    for t in range(1, 11):
        y = y + (math.sin(t) if t > 5 else 0.0) + 0.01 * rng.normal()
        obs.append(y + 0.25 * rng.normal())
    model = g.DeviceModel("object_motion")
    lat, flips = [], 0
    for seed in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        state = g.pf_initialize(model, (1,), obs[0], 100, seed=seed, keep_history=True)
        for t in range(2, 11):
            if g.effective_sample_size(state) < 50:
                g.pf_resample(state, "residual")
                g.pf_rejuvenate(state, g.mh, (t - 1, obs[t - 2]))
            g.pf_update(state, (t,), None, obs[t - 1])
        state.sync()
        lat.append((time.perf_counter() - t0) / 9)
        m5, m6 = g.mean(state, (5, "moving")), g.mean(state, (6, "moving"))
        flips += m5 < 0.5 < m6
    # the same filter through genpf_run_steps: the 9 iterations enqueued by ONE asynchronous call, eager and as a CUDA
    # graph (stratified resample decided per step on the device: the fused step has no residual form)
    run_us = {}
    for mode in ("eager", "graph"):
        ts = []
        for seed in range(21):
            st = g.pf_initialize(model, (1,), obs[0], 100, seed=seed)
            st.sync()
            t0 = time.perf_counter()
            g.pf_run(st, 2, np.array(obs), ess_thresh=0.5, graph=(mode == "graph"))
            st.sync()
            ts.append((time.perf_counter() - t0) / 9)
        run_us[mode] = 1e6 * sorted(ts[1:])[10]
    print(json.dumps({"config": 1, "workload": "README object_motion T=10 n=100 residual+MH (README.md:60-79)",
                      "wall_us_per_step_median": 1e6 * sorted(lat)[len(lat) // 2], "posterior_flip_runs": f"{flips}/20",
                      "last_run": {"mean_moving_5": m5, "mean_moving_6": m6},
                      "run_steps_us_per_step": run_us,
                      "run_steps_note": "genpf_run_steps, stratified + mh + update, ess < n/2 decided on the device; "
                                        "wall time of the whole call incl. graph capture + instantiate, / 9 steps"}), flush=True)

    # ---- config 3
    a, q, r, m0, s0 = 0.9, 1.0, 1.0, 0.0, 1.0
    rng = np.random.default_rng(2)
    x, obs3 = rng.normal(m0, s0), []
    for _ in range(args.T3):
        x = a * x + q * rng.normal()
        obs3.append(x + r * rng.normal())
    model3 = g.DeviceModel("lingauss1d", (a, q, r, m0, s0))
    m, P, lz = kalman(obs3, a, q, r, m0, s0)
    for noise in ("philox53", "lean"):
        state = g.pf_initialize(model3, (1,), obs3[0], args.n3, seed=5, noise=noise)
        state.sync()
        t0 = time.perf_counter()
        for t in range(2, args.T3 + 1):
            g.pf_step(state, t, obs3[t - 2], obs3[t - 1], method="stratified", ess_thresh=1.0, mh_iters=0, return_ess=False)
        state.sync()
        dt = time.perf_counter() - t0
        pm, pv, lml = g.mean(state, (args.T3, "x")), g.var(state, (args.T3, "x")), g.log_ml_estimate(state)
        ups = args.n3 * (args.T3 - 1) / dt
        print(json.dumps({"config": 3, "workload": f"lingauss1d n={args.n3} T={args.T3} stratified every step", "noise": noise,
                          "particle_updates_per_s": ups, "ms_per_step": 1e3 * dt / (args.T3 - 1),
                          "frac_of_68B_roofline": ups * 68 / 6463.3e9,
                          "pf_mean": pm, "kalman_mean": m, "pf_var": pv, "kalman_var": P, "pf_lml": lml, "kalman_logZ": lz}),
              flush=True)
        del state

    # ---- config 5 (one GPU's share of the 4096-filter batch)
    nf, n = args.filters5, 4096
    T5 = 50
    rng = np.random.default_rng(4)
    obs5 = np.cumsum(rng.normal(0, 0.3, (T5, nf)), axis=0)
    for noise in ("philox53", "lean"):
        state = g.pf_initialize(model, (1,), obs5[0], n, n_filters=nf, seed=9, noise=noise)
        state.sync()
        t0 = time.perf_counter()
        for t in range(2, T5 + 1):
            g.pf_step(state, t, obs5[t - 2], obs5[t - 1], method="stratified", ess_thresh=1.0, mh_iters=1, return_ess=False)
        state.sync()
        dt = time.perf_counter() - t0
        ess = g.effective_sample_size(state)
        # pf_replicate! x2 then residual resize back to n, per filter (resize.jl:87-124,236-244), timed on their own
        state.sync()
        t1 = time.perf_counter()
        g.pf_replicate(state, 2)
        g.pf_resize(state, n, "residual")
        state.sync()
        dt_rs = time.perf_counter() - t1
        print(json.dumps({"config": 5, "workload": f"{nf} independent object_motion filters x {n} particles, MH every step",
                          "noise": noise, "particle_updates_per_s": nf * n * (T5 - 1) / dt, "ms_per_step": 1e3 * dt / (T5 - 1),
                          "ess_min": float(ess.min()), "ess_max": float(ess.max()),
                          "replicate_x2_plus_residual_resize_ms": 1e3 * dt_rs}), flush=True)
        del state


if __name__ == "__main__":
    main()
