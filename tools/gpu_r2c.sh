#!/bin/bash
# two GPUs: the gather over NCCL, the sharded bench line with the channelizer sub-record
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_gather.py -q > $OUT/${TAG}_pytest_gather.log 2>&1
echo "pytest gather rc=$?"; tail -5 $OUT/${TAG}_pytest_gather.log
for steps in 20 200; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $steps --warmup 5 \
   > $OUT/${TAG}_bench_n2_s$steps.json 2> $OUT/${TAG}_bench_n2_s$steps.err
echo "bench n2 steps=$steps rc=$?"; tail -3 $OUT/${TAG}_bench_n2_s$steps.err; cat $OUT/${TAG}_bench_n2_s$steps.json
done
python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"; cat $OUT/${TAG}_bench_n1.json
