#!/usr/bin/env python
"""Packs the output of tools/export_golden.jl (run where Julia + Gen + GenParticleFilters exist) into
tests/golden/reference_julia.npz.  Once that file is committed, tests/test_reference_golden.py pins the CPU oracle
and the CUDA path against REAL reference output (ancestor indices converted to 0-based).

    python tools/import_golden.py out_dir
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = sys.argv[1]
    man = json.load(open(os.path.join(src, "manifest.json")))
    out = {}
    for a in man["arrays"]:
        dt = {"f64": "<f8", "i64": "<i8"}[a["dtype"]]
        x = np.fromfile(os.path.join(src, a["name"].replace("/", "__") + ".bin"), dtype=dt)
        assert x.size == a["length"], a
        if a["name"].endswith("/parents"):
            x = x - 1  # Julia is 1-based
        out[a["name"]] = x
    dst = os.path.join(ROOT, "tests", "golden", "reference_julia.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {len(out)} arrays to {dst} (reference {man.get('reference')}, julia {man.get('julia')})")


if __name__ == "__main__":
    main()
