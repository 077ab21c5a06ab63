#!/bin/bash
# usage (on an N-GPU box): tools/run_multigpu.sh TAG N [quick]  -> gpurun_out/TAG_*: shard tests on all GPUs, bench.py at N
# (and N/2), BASELINE config 4 (2^27 particles in ONE filter, balanced and imbalanced shards) and config 5 (4096 filters,
# batch-sharded).  "quick" trims to what an 8-GPU call should spend box time on.
TAG=${1:-mg}
N=${2:-8}
QUICK=${3:-}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ -n "$QUICK" ]; then
  timeout 300 python -m pytest tests/test_gpu_shard.py -x -q -k "all_gpus or world2_push" 2>&1 | tail -5 > $OUT/${TAG}_tests.log
else
  timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -5 > $OUT/${TAG}_tests.log
fi
timeout 300 $TR --nproc-per-node $N --master-port 29700 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
if [ "$N" -ge 4 ] && [ -z "$QUICK" ]; then
  H=$((N / 2))
  timeout 300 $TR --nproc-per-node $H --master-port 29701 bench.py --gpus $H --steps 50 --warmup 10 --no-cpu > $OUT/${TAG}_bench_n$H.json 2> $OUT/${TAG}_bench_n$H.err
fi
: > $OUT/${TAG}_configs.jsonl
WORLDS="1 $N"
[ -n "$QUICK" ] && WORLDS="$N"
for W in $WORLDS; do
  timeout 300 $TR --nproc-per-node $W --master-port 29702 tools/run_config4.py 2>> $OUT/${TAG}_configs.err | grep '^{' >> $OUT/${TAG}_configs.jsonl
done
TILTS="0.35 1.0"
[ -n "$QUICK" ] && TILTS="0.35"
for TILT in $TILTS; do
  CONFIG4_TILT=$TILT timeout 300 $TR --nproc-per-node $N --master-port 29703 tools/run_config4.py 2>> $OUT/${TAG}_configs.err | grep '^{' >> $OUT/${TAG}_configs.jsonl
done
NOISES="philox53 lean"
[ -n "$QUICK" ] && NOISES="philox53"
for NOISE in $NOISES; do
  CONFIG5_NOISE=$NOISE timeout 300 $TR --nproc-per-node $N --master-port 29705 tools/run_config5_mgpu.py 2>> $OUT/${TAG}_configs.err | grep '^{' >> $OUT/${TAG}_configs.jsonl
done
