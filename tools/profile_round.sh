#!/bin/bash
# usage (on the GPU box): tools/profile_round.sh TAG  -> gpurun_out/TAG_{step,micro}_raw.csv, TAG_step_source.csv
# ncu reports stay in /tmp on the box (tens of MB each); only the CSV pages travel back.
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"k_scan_hot|k_step_fused|k_finalize|k_chunk" -s 16 -c 4 \
    -o /tmp/${TAG}_step -f python bench.py --steps 3 --warmup 3 --no-cpu --no-micro > $OUT/${TAG}_ncu_step.log 2>&1
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > $OUT/${TAG}_step_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_step.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_step_source.csv 2>/dev/null
python profiles/hotlines.py /tmp/${TAG}_step_source.csv k_step_fused k_scan_hot k_finalize k_chunk > $OUT/${TAG}_step_hotlines.txt 2>&1
ncu --set full --clock-control none -k regex:"k_reduce|k_expand|k_lookup|k_guide|k_scan|k_radix_onesweep|k_resid_scan|k_radix_prepare|k_sort_finish" -c 44 \
    -o /tmp/${TAG}_micro -f python tools/microbench.py --min-log2 26 --max-log2 26 --reps 1 --dists A > $OUT/${TAG}_ncu_micro.log 2>&1
ncu -i /tmp/${TAG}_micro.ncu-rep --page raw --csv > $OUT/${TAG}_micro_raw.csv 2>/dev/null
