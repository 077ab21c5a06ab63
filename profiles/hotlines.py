"""Aggregate `ncu --page source --csv --print-source cuda,sass` per CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python hotlines.py src.csv [kernel-substr ...]"""
import collections
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
filters = sys.argv[2:] or [""]
cur_file = cur_fn = hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1].replace("genpf::", "").split("(")[0][:60]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        k = (cur_fn, cur_file, int(r[0]))
        agg[k][0] += num(r[hdr.index("Instructions Executed")])
        agg[k][1] += num(r[hdr.index("# Samples")])
        agg[k][2] = r[1].strip()[:100]
for flt in filters:
    fns = sorted({k[0] for k in agg if flt in k[0]})
    for fn in fns:
        items = [(k, v) for k, v in agg.items() if k[0] == fn]
        tot = sum(v[0] for _, v in items)
        tots = sum(v[1] for _, v in items)
        print(f"===== {fn}: {tot} warp-instructions, {tots} samples")
        for k, v in sorted(items, key=lambda kv: -kv[1][0])[:60]:
            print(f"{100 * v[0] / max(tot, 1):5.1f}% inst {100 * v[1] / max(tots, 1):5.1f}% smp  {k[1]}:{k[2]:4d}  {v[2]}")
