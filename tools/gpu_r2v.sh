#!/bin/bash
# 8 GPUs: the driver's SCALE sequence (N = 1, 2, 4, 8 at 20 steps, the full default line incl. e2e and the C5 sub-record)
set -u
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29521"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_scale_n1.json 2> $OUT/${TAG}_scale_n1.err
echo "n1 rc=$?"
for n in 2 4 8; do
SECONDS=0
timeout 500 $TR --nproc-per-node $n bench.py --gpus $n --steps 20 --warmup 5 > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
echo "n$n rc=$? wall ${SECONDS}s"; tail -1 $OUT/${TAG}_scale_n$n.err | cut -c1-200
done
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --skip-e2e --skip-channelizer > $OUT/${TAG}_scale_n8_b.json 2> $OUT/${TAG}_scale_n8_b.err
echo "n8 repeat rc=$?"
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 200 --warmup 5 --skip-e2e --skip-channelizer > $OUT/${TAG}_scale_n8_s200.json 2> $OUT/${TAG}_scale_n8_s200.err
echo "n8 s200 rc=$?"
