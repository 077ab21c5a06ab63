#!/bin/bash
# usage (on an N-GPU box): tools/run_multigpu.sh TAG N  -> gpurun_out/TAG_*: shard tests on all GPUs, bench.py at N (and N/2),
# BASELINE config 4 (2^27 particles in ONE filter, balanced and imbalanced shards) and config 5 (4096 filters, batch-sharded)
TAG=${1:-mg}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -5 > $OUT/${TAG}_tests.log
timeout 300 $TR --nproc-per-node $N --master-port 29700 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
if [ "$N" -ge 4 ]; then
  H=$((N / 2))
  timeout 300 $TR --nproc-per-node $H --master-port 29701 bench.py --gpus $H --steps 50 --warmup 10 --no-cpu > $OUT/${TAG}_bench_n$H.json 2> $OUT/${TAG}_bench_n$H.err
fi
: > $OUT/${TAG}_configs.jsonl
for W in 1 $N; do
  timeout 300 $TR --nproc-per-node $W --master-port 29702 tools/run_config4.py >> $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_configs.err
done
CONFIG4_TILT=0.35 timeout 300 $TR --nproc-per-node $N --master-port 29703 tools/run_config4.py >> $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_configs.err
CONFIG4_TILT=1.0 timeout 300 $TR --nproc-per-node $N --master-port 29704 tools/run_config4.py >> $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_configs.err
for NOISE in philox53 lean; do
  CONFIG5_NOISE=$NOISE timeout 300 $TR --nproc-per-node $N --master-port 29705 tools/run_config5_mgpu.py >> $OUT/${TAG}_configs.jsonl 2>> $OUT/${TAG}_configs.err
done
