#!/usr/bin/env python
"""Resampling microbenchmark (BASELINE.json configs[1], SURVEY.md 8d "config 2").

log_weights[n], n in 2^16..2^28, distributions A = N(0,1), B = N(0,5^2), C = all equal; methods multinomial /
stratified(sort=false) / stratified(sort=true) / residual, plus ESS+logsumexp alone.  Inputs are resident in
HBM (GENPF_DEVICE_PTRS); time = sum of the library's kernel durations (CUDA events on the launching stream,
genpf_profile_begin/end), median over repetitions; GB/s = algorithmic bytes (ESS 8 B, resample 24 B
per particle) / time, against MEASURED_PEAKS.json.  Writes one JSON object per line.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=16)
    ap.add_argument("--max-log2", type=int, default=26)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default=None)
    ap.add_argument("--dists", default="ABC", help="subset of A (N(0,1)), B (N(0,25)), C (equal)")
    args = ap.parse_args()
    import torch

    import genpf_b200 as g
    L, lib = g._lib, g.load()
    L.check(lib.genpf_set_device(torch.cuda.current_device()))
    peak = 6463.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)
    out_f = open(args.out, "w") if args.out else None

    def kernel_ms(fn):
        times = []
        for _ in range(args.reps):
            flush.zero_()
            torch.cuda.synchronize()
            L.check(lib.genpf_profile_begin())
            fn()
            buf = C.create_string_buffer(1 << 14)
            L.check(lib.genpf_profile_end(buf, len(buf)))
            rows = [r.split("\t") for r in buf.value.decode().strip().splitlines()]
            times.append((sum(float(r[2]) for r in rows), {r[0].strip("()").split("<")[0]: float(r[2]) for r in rows}))
        times.sort(key=lambda t: t[0])
        return times[len(times) // 2]

    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = 1 << lg
        gen = torch.Generator(device="cuda").manual_seed(0)
        base = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
        for dist_name, mk in (("A:N(0,1)", lambda: base), ("B:N(0,25)", lambda: base * 5.0),
                              ("C:equal", lambda: torch.zeros_like(base))):
            if dist_name[0] not in args.dists:
                continue
            lw = mk()
            parents = torch.empty(n, dtype=torch.int64, device="cuda")
            lw_out = torch.empty(n, dtype=torch.float64, device="cuda")
            inc, kind, ess = C.c_double(), C.c_int32(), C.c_double()

            def ess_fn():
                L.check(lib.genpf_ess(lw.data_ptr(), n, L.DEVICE_PTRS, C.byref(ess)))

            ms, parts = kernel_ms(ess_fn)
            rec = {"op": "ess+logsumexp", "n": n, "dist": dist_name, "ms": ms, "algo_bytes": 8 * n,
                   "GBps": 8 * n / ms / 1e6, "frac_of_measured_peak": 8 * n / ms / 1e6 / peak, "kernels_ms": parts}
            print(json.dumps(rec), flush=True)
            if out_f:
                out_f.write(json.dumps(rec) + "\n")
            for method, flags, label in ((L.STRATIFIED, 0, "stratified(sort=false)"),
                                         (L.STRATIFIED, L.SORT_PARTICLES, "stratified(sort=true)"),
                                         (L.MULTINOMIAL, 0, "multinomial"), (L.RESIDUAL, 0, "residual")):
                def rs_fn():
                    L.check(lib.genpf_resample(method, lw.data_ptr(), None, n, n, None, 1, flags | L.DEVICE_PTRS,
                                               parents.data_ptr(), lw_out.data_ptr(), C.byref(inc), C.byref(kind)))
                ms, parts = kernel_ms(rs_fn)
                rec = {"op": "resample:" + label, "n": n, "dist": dist_name, "ms": ms, "algo_bytes": 24 * n,
                       "GBps": 24 * n / ms / 1e6, "frac_of_measured_peak": 24 * n / ms / 1e6 / peak,
                       "kernels_ms": parts}
                print(json.dumps(rec), flush=True)
                if out_f:
                    out_f.write(json.dumps(rec) + "\n")
    if out_f:
        out_f.close()


if __name__ == "__main__":
    main()
