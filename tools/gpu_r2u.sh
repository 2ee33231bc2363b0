#!/bin/bash
# 4 GPUs: the default bench line at N = 4 and N = 2 (copy-engine gather, balanced, channelizer sub-record), gather tests
set -u
TAG=${1:-r2u}
OUT=gpurun_out
mkdir -p $OUT
for n in 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 20 --warmup 5 \
   > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
echo "bench n$n rc=$?"; tail -2 $OUT/${TAG}_bench_n$n.err | cut -c1-300
done
timeout 400 python -m pytest tests/test_gpu_gather.py -x -q -m gpu > $OUT/${TAG}_pytest_gather.log 2>&1
echo "pytest gather rc=$?"; tail -3 $OUT/${TAG}_pytest_gather.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"
