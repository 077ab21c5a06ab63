#!/usr/bin/env python
"""BASELINE.json configs[4]: 4096 independent object_motion filters x 4096 particles, MH rejuvenation + pf_replicate
resizing, batch-sharded over the GPUs this is launched on (torchrun; 4096 / world filters per GPU, NO communication
in the data path -- SURVEY 8e "batches of independent filters shard with no communication at all").
Every step: per-filter ESS decision on the device (ess < n/2), stratified resample + mh + update (genpf_step);
every 10th step pf_replicate!(x2) + residual resize back to n per filter (resize.jl:87-124,236-244).
Rank 0 prints one JSON line: aggregate particle-updates/s = filters * particles * steps / max-over-ranks device time."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes as C

    import torch

    import genpf_b200 as g
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    g._lib.check(g.load().genpf_set_device(lr))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nf_total, n = int(os.environ.get("CONFIG5_FILTERS", "4096")), 4096
    nf = nf_total // world
    noise = os.environ.get("CONFIG5_NOISE", "philox53")
    T, W = 40, 12  # the warm-up contains one replicate + resize cycle (first-touch allocations)
    rng = np.random.default_rng(4)
    obs_all = np.cumsum(rng.normal(0, 0.3, (T + W + 2, nf_total)), axis=0)
    obs = np.ascontiguousarray(obs_all[:, rank * nf:(rank + 1) * nf])  # this rank's filters
    model = g.DeviceModel("object_motion")
    st = g.pf_initialize(model, (1,), obs[0], n, n_filters=nf, seed=9, noise=noise, first_filter=rank * nf)
    sp = C.c_void_p()
    g._lib.check(g.load().genpf_filter_stream(st._h, C.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value)

    use_run = os.environ.get("CONFIG5_RUN_STEPS", "1") == "1"

    def advance(t0, k):
        """steps t0 .. t0+k-1; the stretch up to the next multiple of 10 goes out as ONE genpf_run_steps call (the
        per-step Python/ctypes overhead otherwise exceeds the kernels' 55 us per step at this size)"""
        t = t0
        while t < t0 + k:
            stop = min(t0 + k, (t // 10 + 1) * 10 + 1)  # run through the next multiple of 10 (inclusive)
            if use_run:
                g.pf_run(st, t, obs[t - 2:stop - 1], ess_thresh=0.5, mh_iters=1)
            else:
                for u in range(t, stop):
                    g.pf_step(st, u, obs[u - 2], obs[u - 1], method="stratified", ess_thresh=0.5, mh_iters=1, return_ess=False)
            t = stop
            if (t - 1) % 10 == 0:
                g.pf_replicate(st, 2)
                g.pf_resize(st, n, "residual")
        return t

    t = advance(2, W)
    st.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t = advance(t, T)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    ess = g.effective_sample_size(st)
    stats = torch.tensor([float(np.min(ess)), -float(np.max(ess)), -float(np.isfinite(st.log_weights).all())],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.MIN)
    if rank == 0:
        msv = float(ms.item())
        print(json.dumps({"config": 5, "workload": f"{nf_total} independent object_motion filters x {n} particles: per-filter ESS "
                          "decision + stratified resample + mh + update per step, replicate x2 + residual resize every 10th step",
                          "n_gpus": world, "filters_per_gpu": nf, "driver": "genpf_run_steps" if use_run else "genpf_step per step", "noise": noise, "steps": T, "ms_per_step": msv / T,
                          "particle_updates_per_s": nf_total * n * T / (msv * 1e-3), "communication": "none (batch sharding)",
                          "ess_min": float(stats[0].item()), "ess_max": -float(stats[1].item()),
                          "all_weights_finite": bool(-stats[2].item() > 0.5)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
