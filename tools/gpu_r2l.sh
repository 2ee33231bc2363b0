#!/bin/bash
set -u
TAG=${1:-r2l}
OUT=gpurun_out
mkdir -p $OUT
for steps in 20 200; do
BENCH_GATHER=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps $steps --warmup 5 --skip-e2e \
   > $OUT/${TAG}_bench_n8_s$steps.json 2> $OUT/${TAG}_bench_n8_s$steps.err
echo "bench n8 steps=$steps rc=$?"; tail -2 $OUT/${TAG}_bench_n8_s$steps.err | cut -c1-300
done
timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"
