#!/bin/bash
set -u
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e > $OUT/${TAG}_scale_n2.json 2> $OUT/${TAG}_scale_n2.err
echo "n2 rc=$?"; tail -2 $OUT/${TAG}_scale_n2.err | cut -c1-200
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-channelizer > $OUT/${TAG}_scale_n1.json 2> $OUT/${TAG}_scale_n1.err
echo "n1 rc=$?"
