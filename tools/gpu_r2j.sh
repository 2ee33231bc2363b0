#!/bin/bash
# 8 GPUs: where does the scaling stand (C2 weak, C5 strong)
set -u
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
for steps in 20 200; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps $steps --warmup 5 --skip-e2e \
   > $OUT/${TAG}_bench_n8_s$steps.json 2> $OUT/${TAG}_bench_n8_s$steps.err
echo "bench n8 steps=$steps rc=$?"; tail -2 $OUT/${TAG}_bench_n8_s$steps.err | cut -c1-300
done
python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"
