#!/usr/bin/env python
"""BASELINE.json configs[3]: object_motion, 2^27 particles in ONE filter sharded over the GPUs this is
launched on (strong scaling: 2^27 / world per GPU).  Run with torchrun (world 1 runs the unsharded fused path).
Prints one JSON line on rank 0."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import genpf_b200 as g
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    g._lib.check(g.load().genpf_set_device(lr))
    n_total = 1 << int(os.environ.get("CONFIG4_LOG2", "27"))
    T, W = 30, 5
    rng = np.random.default_rng(3)
    y, obs = 0.0, []
    for t in range(1, T + W + 3):
        y = y + (math.sin(t) if t > 5 else 0.0) + 0.01 * rng.normal()
        obs.append(y + 0.25 * rng.normal())
    model = g.DeviceModel("object_motion")
    if world > 1:
        import torch.distributed as dist
        from genpf_b200.sharded import ShardedFilter
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        sf = ShardedFilter(model, n_total // world, seed=7)
        sf.initialize(obs[0])
        step = lambda t: sf.step(t, obs[t - 2], obs[t - 1])
        stream = sf.stream
    else:
        import ctypes as C
        st = g.pf_initialize(model, (1,), obs[0], n_total, seed=7)
        step = lambda t: g.pf_step(st, t, obs[t - 2], obs[t - 1], method="stratified", ess_thresh=1.0, return_ess=False)
        sp = C.c_void_p()
        g._lib.check(g.load().genpf_filter_stream(st._h, C.byref(sp)))
        stream = torch.cuda.ExternalStream(sp.value)
    # CONFIG4_TILT = c > 0: IMBALANCED shards -- before every step rank r's log-weights get + c*r (outside the timed
    # region), so shard r carries a mass ~ e^{c r}: offspring migrate towards the high ranks over NVLink
    tilt = float(os.environ.get("CONFIG4_TILT", "0"))
    BYTES_PER_OFFSPRING = 38  # y_t 8 + y_{t-1} 8 + moving 1+1 + lw 8 + e 8 + parent 4, stored into the owner's HBM
    t = 2
    for _ in range(W):
        step(t)
        t += 1
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms_sum, fracs = 0.0, []
    if tilt > 0.0 and world > 1:
        T = 8
        for _ in range(T):
            lw = sf.state.log_weights
            sf.state.log_weights = lw + tilt * rank
            sf.state.sync()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(t)
            e1.record(stream)
            torch.cuda.synchronize()
            ms_sum += e0.elapsed_time(e1)
            fracs.append(sf.exchange_summary()[1])
            t += 1
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(T):
            step(t)
            t += 1
        e1.record(stream)
        torch.cuda.synchronize()
        ms_sum = e0.elapsed_time(e1)
    ms = torch.tensor([ms_sum], dtype=torch.float64, device="cuda")
    extra = {}
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ranges, frac = sf.exchange_summary()
        if fracs:
            frac = float(np.mean(fracs))
        ess, lml, kind = sf.stats()
        per_gpu = frac * (n_total // world) * BYTES_PER_OFFSPRING
        extra = {"cross_shard_offspring_fraction": frac, "nvlink_bytes_per_step_per_gpu": per_gpu,
                 "nvlink_GBps_per_gpu_over_the_step": per_gpu / (float(ms.item()) / T * 1e-3) / 1e9,
                 "nvlink_note": "bytes of offspring stored into a peer's HBM (mean over GPUs) / whole step time; the stores "
                                "overlap the step kernel, peak 770 GB/s per direction is not approached",
                 "tilt": tilt, "ess": ess}
    if rank == 0:
        msv = float(ms.item())
        print(json.dumps({"config": 4, "workload": f"object_motion 2^{int(math.log2(n_total))} particles, one filter, "
                          f"{world} GPU(s), stratified + MH + update per step", "n_gpus": world, "steps": T,
                          "ms_per_step": msv / T, "particle_updates_per_s": n_total * T / (msv * 1e-3), **extra}), flush=True)
    if world > 1:
        dist.barrier()
        sf.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
