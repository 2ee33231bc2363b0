#!/bin/bash
# two GPUs: the peer-memory gather (tests + bench), NCCL mode beside it
set -u
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gather.py -q > $OUT/${TAG}_pytest_gather.log 2>&1
echo "pytest gather rc=$?"; tail -8 $OUT/${TAG}_pytest_gather.log
for mode in peer nccl; do
BENCH_GATHER=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e \
   > $OUT/${TAG}_bench_n2_$mode.json 2> $OUT/${TAG}_bench_n2_$mode.err
echo "bench n2 $mode rc=$?"; tail -3 $OUT/${TAG}_bench_n2_$mode.err | cut -c1-300
done
python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"
