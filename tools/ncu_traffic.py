#!/usr/bin/env python
"""profiles/traffic.json from `ncu --page raw --csv` exports: dram__bytes_read.sum + dram__bytes_write.sum per launch,
averaged over the captured launches of each kernel.  usage: tools/ncu_traffic.py CAPTURE_NAME raw.csv [raw2.csv ...]
Kernels templated on the noise policy are keyed "name@policy" (bench.py looks "k_step_fused@philox53" up first)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    capture, files = sys.argv[1], sys.argv[2:]
    acc = collections.defaultdict(list)
    for f in files:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            full = r[col["Kernel Name"]]
            name = full.replace("void ", "").replace("genpf::", "").split("<")[0].split("(")[0]
            key = name
            if "NoisePhilox53" in full:
                key += "@philox53"
            elif "NoiseLean" in full:
                key += "@lean"
            if name == "k_scan_hot":
                key += "@host" if "<int, 0" in full or "<int, false" in full else ""
            b = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[col[m]].replace(",", "")) * UNIT[units[col[m]]]
            n = int(r[col["launch__grid_size"]].replace(",", "")) if "launch__grid_size" in col else 0
            acc[(key, n)].append(b)
    out = {}
    for (key, n), v in sorted(acc.items()):
        k = key if key not in out else f"{key}#grid{n}"
        out[k] = {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v), "grid": n, "capture": capture}
    p = os.path.join(ROOT, "profiles", "traffic.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps({k: round(v["dram_bytes_per_launch"] / 1e6, 1) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
