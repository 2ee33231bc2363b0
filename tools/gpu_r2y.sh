#!/bin/bash
# 8 GPUs: the full default line at N = 8 and N = 2 after the e2e change
set -u
TAG=${1:-r2y}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29521"
for n in 8 2; do
SECONDS=0
timeout 500 $TR --nproc-per-node $n bench.py --gpus $n --steps 20 --warmup 5 > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
echo "n$n rc=$? wall ${SECONDS}s"; tail -1 $OUT/${TAG}_scale_n$n.err | cut -c1-200
done
SECONDS=0
timeout 300 $TR --nproc-per-node 8 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $OUT/${TAG}_ref_n8.json 2> $OUT/${TAG}_ref_n8.err
echo "ref n8 rc=$? wall ${SECONDS}s"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu > $OUT/${TAG}_scale_n1.json 2> $OUT/${TAG}_scale_n1.err
echo "n1 rc=$?"
