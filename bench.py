#!/usr/bin/env python
"""bench.py -- headline benchmark: particle-updates/s of the object_motion filter at 2^24 particles per GPU.

One "step" = one README loop iteration (README.md:66-77) with the resample forced:
    ESS -> pf_resample!(:stratified, sort_particles=false) -> pf_rejuvenate!(mh) -> pf_update!
over 2^24 particles (BASELINE.json configs).  One JSON line on stdout:
    value        whole-job particle-updates/s, state resident in HBM, CUDA events on the filter's stream
    e2e          the same through the public API (genpf_b200.pf_step) with host observation buffers in and
                 the ESS read back every step
    roofline     dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
    cpu_baseline the CPU oracle (OpenMP port of the reference algorithms) on a bounded sample, same box
`--impl reference` times that CPU port alone (the Julia reference cannot run in this image: no julia).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec"
N_PARTICLES = 1 << 24
# SURVEY.md 8(d) algorithmic bytes per particle-update, per kernel of the step (sum = 117)
KERNEL_ALGO_BYTES = {
    "k_propagate": 34,  # R y,m,lw 17 + W y,m,lw 17
    "k_scan": 8,        # R lw
    "k_expand": 4,      # W parents (int32)
    "k_gather": 44,     # R window 18 + W window 18 + W lw 8
    "k_mh": 27,         # R window 18 + W slice 9
    # fused kernels: what is left of the 117 B once gather -> MH -> update stay in registers:
    # R window 18 + W window 18 + W lw 8 + W parents 4 (the scan's 8 B read is k_scan's)
    "k_step_fused": 48,
    "k_step_push": 48,   # the same arithmetic, stores routed to the owning GPU (multi-GPU)
    "k_reduce": 8,
}
STEP_ALGO_BYTES = 117
HEADLINE_NOISE = "philox53"  # 53-bit uniforms + fp64 Box-Muller is the production policy


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_of(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
    return None


def observations(T, seed=3):
    """README.md:87-89 generator extended to T steps (moving for t > 5)."""
    rng = np.random.default_rng(seed)
    y, obs = 0.0, []
    for t in range(1, T + 1):
        y = y + (math.sin(t) if t > 5 else 0.0) + 0.01 * rng.normal()
        obs.append(y + 0.25 * rng.normal())
    return np.array(obs)


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, device, period=0.004):
        import threading
        self.rows, self.stop_flag, self.err = [], False, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(device)
            self.nv, self.h = nv, h
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return

        def loop():
            while not self.stop_flag:
                try:
                    sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                    try:
                        rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        rs = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    pw = self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    self.rows.append((sm, rs, pw))
                except Exception as e:  # noqa: BLE001
                    self.err = repr(e)
                    return
                time.sleep(period)

        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.err and not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err]}
        self.t.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm,
                "reasons": [nm for b, nm in names.items() if bits & b], "samples": len(self.rows),
                "power_w_max": max(r[2] for r in self.rows)}


def cpu_arm(steps, warmup, n_sample, omp=True, threads=None):
    """The CPU port of the reference path (oracle/) on the host cores: same step, bounded sample.
    torchrun exports OMP_NUM_THREADS=1 to its workers, so the thread count is set explicitly."""
    from oracle import oracle as orc
    T = warmup + steps + 1
    obs = observations(T)
    f = orc.OMFilter(n_sample, seed=0, omp=omp, threads=threads or os.cpu_count())
    f.init(math.sin(1.0), obs[0])
    t = 2
    for _ in range(warmup):
        f.step(t, math.sin(t - 1.0), obs[t - 2], math.sin(float(t)), obs[t - 1])
        t += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        f.step(t, math.sin(t - 1.0), obs[t - 2], math.sin(float(t)), obs[t - 1])
        t += 1
    dt = time.perf_counter() - t0
    return n_sample * steps / dt, dt / steps, f.threads()


WORKLOAD = ("object_motion 2^24 particles/GPU: ESS + stratified resample(sort_particles=false) + MH rejuvenation "
            "+ update per step (README.md:66-77, resample forced)")


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia/Gen cannot be installed here (a Julia package; no julia
    binary in the image, no network), so the C/OpenMP port of the reference algorithms (oracle/) is timed on all
    host cores on the SAME workload (2^24 particles per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.particles
    value, sec_per_step, cores = cpu_arm(args.steps, args.warmup, n_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "particles_per_gpu": n_sample,
                   "note": "Julia/Gen cannot run here (no julia in the image); this is the C/OpenMP port of the "
                           "reference algorithms (oracle/), faster than the real reference (no Gen trace overhead); "
                           "at --gpus N it still times ONE 2^24-particle filter on the host cores"},
        "cpu_baseline": {"value": value, "unit": "particle-updates/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} particles per step (the full 2^24 workload), {args.steps} steps, "
                                   f"{cores} OpenMP threads"},
        "e2e": {"value": value, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse_profile(buf):
    prof = {}
    for row in buf.value.decode().strip().splitlines():
        name, cnt, tot = row.split("\t")
        name = name.strip("()").split("<")[0]
        c, tt = prof.get(name, (0, 0.0))
        prof[name] = (c + int(cnt), tt + float(tot))
    return prof


def resample_microbench(g, torch, peak, log2n=26, reps=5):
    """BASELINE metric, second half: "resample HBM GB/s vs peak" (SURVEY 8d config 2).  Host-array semantics with the
    inputs resident in HBM (GENPF_DEVICE_PTRS), n = 2^26 fp64 log-weights ~ N(0,1) (512 MB > L2, and L2 flushed
    between repetitions); time = summed CUDA-event durations of the call's kernels, median of `reps`;
    GB/s = algorithmic bytes (ESS 8 B, resample 24 B per particle) / time."""
    import ctypes as C
    L, lib = g._lib, g.load()
    n = 1 << log2n
    gen = torch.Generator(device="cuda").manual_seed(0)
    lw = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    parents = torch.empty(n, dtype=torch.int64, device="cuda")
    lw_out = torch.empty(n, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    inc, kind, ess = C.c_double(), C.c_int32(), C.c_double()

    def timed(fn):
        ts = []
        for _ in range(reps + 1):
            flush.zero_()
            torch.cuda.synchronize()
            L.check(lib.genpf_profile_begin())
            fn()
            buf = C.create_string_buffer(1 << 14)
            L.check(lib.genpf_profile_end(buf, len(buf)))
            prof = parse_profile(buf)
            ts.append((sum(v[1] for v in prof.values()), {k: v[1] for k, v in prof.items()}))
        ts = sorted(ts[1:], key=lambda t: t[0])  # first repetition = warm-up (allocations)
        return ts[len(ts) // 2]

    rows = {}

    def add(name, algo, fn):
        ms, parts = timed(fn)
        rows[name] = {"ms": ms, "GBps": algo * n / ms / 1e6, "frac": algo * n / ms / 1e6 / peak,
                      "algo_bytes_per_particle": algo, "kernels_ms": parts}

    add("ess+logsumexp", 8, lambda: L.check(lib.genpf_ess(lw.data_ptr(), n, L.DEVICE_PTRS, C.byref(ess))))
    for name, method, flags in (("stratified(sort_particles=false)", L.STRATIFIED, 0),
                                ("stratified(sort_particles=true)", L.STRATIFIED, L.SORT_PARTICLES),
                                ("multinomial", L.MULTINOMIAL, 0), ("residual", L.RESIDUAL, 0)):
        add(name, 24, lambda m=method, f=flags: L.check(lib.genpf_resample(
            m, lw.data_ptr(), None, n, n, None, 1, f | L.DEVICE_PTRS, parents.data_ptr(), lw_out.data_ptr(),
            C.byref(inc), C.byref(kind))))
    del lw, parents, lw_out, flush
    torch.cuda.empty_cache()
    return {"n": n, "dist": "N(0,1)", "l2": "flushed between repetitions; 512 MB input > L2", "reps": reps,
            "peak_GBps": peak, "rows": rows}


def host_array_e2e(g, torch, n, reps=5):
    """The drop-in call for arbitrary Gen models (north star path 1): genpf_resample with HOST buffers -- log_weights
    in, Int64 ancestors + new log_weights out -- copies inside the timed region (pinned host memory)."""
    import ctypes as C
    L, lib = g._lib, g.load()
    lw = torch.randn(n, dtype=torch.float64).pin_memory()
    parents = torch.empty(n, dtype=torch.int64).pin_memory()
    lw_out = torch.empty(n, dtype=torch.float64).pin_memory()
    inc, kind = C.c_double(), C.c_int32()
    out = {}
    for name, method, flags in (("stratified(sort_particles=false)", L.STRATIFIED, 0), ("residual", L.RESIDUAL, 0)):
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            L.check(lib.genpf_resample(method, lw.data_ptr(), None, n, n, None, 1, flags, parents.data_ptr(),
                                       lw_out.data_ptr(), C.byref(inc), C.byref(kind)))
            ts.append(time.perf_counter() - t0)
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        out[name] = {"ms": t * 1e3, "particles_per_s": n / t, "h2d_bytes": 8 * n, "d2h_bytes": 16 * n,
                     "pcie_GBps": 24 * n / t / 1e9}
    return {"n": n, "call": "genpf_resample(host pointers, pinned)", "rows": out}


def main():
    # library banners (e.g. "NCCL version ...") go to stderr: stdout carries only the JSON line
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="genpf")
    ap.add_argument("--particles", type=int, default=N_PARTICLES)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-micro", action="store_true", help="skip the resample microbench + host-array e2e objects")
    ap.add_argument("--noise", default=HEADLINE_NOISE, choices=["lean", "philox53"],
                    help="noise policy of the headline `value` (the other one is reported beside it at 1 GPU)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C

    import torch

    import genpf_b200 as g
    L = g._lib
    lib = g.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    L.check(lib.genpf_set_device(local_rank))
    # GENPF_BENCH_SHARD1=1 (diagnostic, under torchrun --nproc-per-node 1): the sharded step with a world of one, to
    # separate what the push kernel itself costs from what the peers cost
    sharded = world > 1 or os.environ.get("GENPF_BENCH_SHARD1") == "1"
    if sharded:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.particles
    K, W = args.steps, args.warmup
    T = W + 3 * K + 6
    obs = observations(T)
    model = g.DeviceModel("object_motion")
    auxs = [np.array([math.sin(float(t))]) for t in range(T + 2)]
    obs_arr = [np.array([o]) for o in obs]
    method = L.STRATIFIED

    def measure(noise, sampler_on):
        """W warm-up + K timed steps (device resident) + K profiled steps + K end-to-end steps for one noise policy."""
        shard_info = None
        if not sharded:
            state = g.pf_initialize(model, (1,), obs[0], n, seed=1234, noise=noise)
            sp = C.c_void_p()
            L.check(lib.genpf_filter_stream(state._h, C.byref(sp)))
            stream = torch.cuda.ExternalStream(sp.value)

            def raw_step(t):  # asynchronous: nothing is copied back
                L.check(lib.genpf_step(state._h, t, L.ptr(obs_arr[t - 2]), L.ptr(auxs[t - 1]), L.ptr(obs_arr[t - 1]),
                                       L.ptr(auxs[t]), method, 1.0, 1, None))
                state.t = t

            def e2e_step(t, pin_prev, pin_t):
                return g.pf_step(state, t, pin_prev, pin_t, method="stratified", ess_thresh=1.0, mh_iters=1,
                                 return_ess=True)
        else:
            # ONE filter of world * n particles, slots sharded contiguously over the ranks:
            # shard totals + closing barrier exchanged through peer memory by the step's own kernels (two
            # synchronisation points, no NCCL) + NVLink P2P push of offspring
            from genpf_b200.sharded import ShardedFilter
            sf = ShardedFilter(model, n, seed=1234, noise=noise)
            sf.initialize(obs[0])
            stream = sf.stream

            def raw_step(t):
                sf.step(t, obs[t - 2], obs[t - 1])

            def e2e_step(t, pin_prev, pin_t):
                sf.step(t, pin_prev[0], pin_t[0])
                return np.array([sf.stats()[0]])

        t = 2
        for _ in range(W):
            raw_step(t)
            t += 1
        # ---- timed region 1: device-resident throughput
        sampler = ClockSampler(local_rank) if (rank == 0 and sampler_on) else None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.genpf_launch_count()
        e0.record(stream)
        for _ in range(K):
            raw_step(t)
            t += 1
        e1.record(stream)
        barrier()
        launches = lib.genpf_launch_count() - launches0
        ms = e0.elapsed_time(e1)
        if sharded:
            ranges, frac = sf.exchange_summary()
            shard_info = {"cross_shard_offspring_fraction": frac,
                          "nvlink_bytes_per_step_per_gpu": frac * n * 38.0,  # parents 4 + two slices 18 + lw 8 + e 8
                          "exchange_per_step": "24 B of shard totals + one closing barrier per rank as NVLink P2P stores "
                                               "+ epoch flags polled by the step's own kernels (no NCCL inside the "
                                               "step; closing counts derived on every rank)"}
        # ---- timed region 2: per-kernel CUDA events (same K steps again) for the roofline of the dominant kernel
        L.check(lib.genpf_profile_begin())
        for _ in range(K):
            raw_step(t)
            t += 1
        buf = C.create_string_buffer(1 << 16)
        L.check(lib.genpf_profile_end(buf, len(buf)))
        prof = parse_profile(buf)
        # ---- timed region 3: end to end through the public API: host obs in, ESS back, every step
        pin_prev, pin_t = np.empty(1), np.empty(1)
        barrier()
        w0 = time.perf_counter()
        for _ in range(K):
            pin_prev[0], pin_t[0] = obs[t - 2], obs[t - 1]
            ess = e2e_step(t, pin_prev, pin_t)
            t += 1
        barrier()
        e2e_s = time.perf_counter() - w0
        assert np.isfinite(ess).all()
        clocks = sampler.stop() if sampler else None
        times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        if sharded:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            sf.close()
        ms_max, e2e_ms_max = times.tolist()
        return dict(ms=ms_max, e2e_ms=e2e_ms_max, launches=int(launches), prof=prof, clocks=clocks,
                    shard_info=shard_info)

    head = measure(args.noise, True)
    other_noise = "lean" if args.noise == "philox53" else "philox53"
    other = measure(other_noise, False) if world == 1 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak_gbs()
    total_updates = float(n) * world * K

    def summarise(m):
        v = total_updates / (m["ms"] * 1e-3)
        step_kernel_ms = sum(x[1] for x in m["prof"].values()) / K
        return {"value": v, "ms_per_step": m["ms"] / K, "e2e_value": total_updates / (m["e2e_ms"] * 1e-3),
                "roofline_frac": STEP_ALGO_BYTES * n / (step_kernel_ms * 1e-3) / 1e9 / peak,
                "kernels_ms_per_step": {k: x[1] / K for k, x in sorted(m["prof"].items())}}

    hs = summarise(head)
    prof = head["prof"]
    # SURVEY 8(d)'s 117 B/particle-update describes the whole step, so the roofline figure is formed over the
    # step's launches together (k_scan + k_finalize_fast + k_step_fused, or their sharded counterparts): fusion
    # removed the gather->MH->update round trips, so apportioning the 117 B to the fused kernel alone would
    # credit it with bytes it never moves.  The dominant kernel's own numbers are reported next to it.
    dom_name, (dom_cnt, dom_ms) = max(prof.items(), key=lambda kv: kv[1][1])
    per_launch_ms = dom_ms / dom_cnt
    prof_total = sum(v[1] for v in prof.values())
    step_kernel_ms = prof_total / K
    algo_b = STEP_ALGO_BYTES * n
    achieved = algo_b / (step_kernel_ms * 1e-3) / 1e9
    def tr(k):
        """DRAM bytes per launch of kernel k from the committed ncu capture of THIS build (profiles/traffic.json, written
        by tools/ncu_traffic.py): the policy-specific entry first; the lean kernels move the same buffers as the
        philox53 ones (same loads and stores, only the noise arithmetic differs), so they fall back to that capture."""
        for key in (f"{k}@{args.noise}", k, f"{k}@philox53"):
            v = traffic_of(key)
            if v is not None:
                return v
        return None

    dom_traffic = tr(dom_name)
    step_traffic = sum(tr(k) for k in prof) if all(tr(k) is not None for k in prof) else None
    dominant = {
        "name": dom_name, "ms_per_launch": per_launch_ms, "share_of_step": dom_ms / prof_total,
        "dram_bytes_per_launch": dom_traffic,
        "dram_gbs": (dom_traffic / (per_launch_ms * 1e-3) / 1e9) if dom_traffic else None,
        "dram_frac_of_peak": (dom_traffic / (per_launch_ms * 1e-3) / 1e9 / peak) if dom_traffic else None,
        "own_algo_bytes_per_update": KERNEL_ALGO_BYTES.get(dom_name),
        "bound": "instruction issue (ncu summary under profiles/)",
    }
    noise_desc = {"lean": "lean: ONE Philox4x32-10 block/particle/step (24-bit Bernoulli uniforms, 32-bit accept uniform, "
                          "fp32 Box-Muller pair)",
                  "philox53": "philox53 (SURVEY 8c production noise): 53-bit uniforms (x>>11)*2^-53 + fp64 Box-Muller, "
                              "three Philox4x32-10 blocks/particle/step"}
    line = {
        "metric": METRIC, "value": hs["value"], "unit": "particle-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": hs["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "particles_per_gpu": n, "l2": "inputs larger than L2 (>=1.2 GB touched per step)",
                   "parallelism": "1 GPU" if world == 1 else f"one filter of {world}x2^24 particles sharded over "
                                  f"{world} GPUs: shard totals exchanged through peer memory (NVLink stores + epoch flags), offspring pushed to the owner GPU over NVLink P2P inside the step kernel",
                   "sharding": head["shard_info"],
                   "noise": noise_desc[args.noise], "algo_bytes_per_update": STEP_ALGO_BYTES,
                   "resample_note": "stratified with sort_particles=false (the reference README uses :residual and "
                                    "the stratified default sorts; both are in resample_microbench)"},
        "clocks": head["clocks"],
        "e2e": {"value": hs["e2e_value"], "unit": "particle-updates/s",
                "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 48,
                "note": "genpf_b200.pf_step / ShardedFilter.step: obs/aux scalars in (kernel arguments), "
                        "Stats(ESS) read back per step"},
        "gpu_launches": head["launches"],
        "roofline": {"bound": "hbm", "kernel": dom_name,
                     "scope": "the step's launches together (" + " + ".join(sorted(prof)) + "); SURVEY 8(d): "
                              "117 B/particle-update is defined on the whole step",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": step_traffic, "peak_source": peak_src,
                     "algo_bytes_per_launch": algo_b, "ms_per_launch": step_kernel_ms,
                     "wall_frac": (STEP_ALGO_BYTES * n / (hs["ms_per_step"] * 1e-3) / 1e9) / peak,
                     "dominant_kernel": dominant,
                     "kernels_ms_per_step": hs["kernels_ms_per_step"]},
    }
    if other is not None:
        line["noise_policies"] = {args.noise: dict(hs, headline=True), other_noise: dict(summarise(other), headline=False)}
    if world == 1 and not args.no_micro:
        line["resample_microbench"] = resample_microbench(g, torch, peak)
        line["e2e_host_arrays"] = host_array_e2e(g, torch, n)
    if world == 1 and not args.no_cpu:
        n_sample = n
        v1, s1, cores = cpu_arm(2, 1, n_sample)
        steps_cpu = max(2, min(40, int(12.0 / max(s1, 1e-3))))
        v, s, cores = cpu_arm(steps_cpu, 1, n_sample)
        line["cpu_baseline"] = {"value": v, "unit": "particle-updates/s", "cores": cores, "kind": "port",
                                "sample": f"{n_sample} particles x {steps_cpu} steps of the same step "
                                          "(C/OpenMP port of the reference algorithms; Julia absent from the image)"}
    emit(line)
    if sharded:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
