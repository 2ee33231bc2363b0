#!/bin/bash
# 8 GPUs: copy-engine gather (mode 2) with the shrinking slab schedule; peer gather tests; C5 at the per-rank size on one GPU
set -u
TAG=${1:-r2o}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 python -m pytest tests/test_gpu_gather.py -x -q -m gpu > $OUT/${TAG}_pytest_gather.log 2>&1
echo "pytest gather rc=$?"; tail -3 $OUT/${TAG}_pytest_gather.log
for tag in s20 s20b; do
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --skip-e2e > $OUT/${TAG}_bench_n8_$tag.json 2> $OUT/${TAG}_bench_n8_$tag.err
echo "bench n8 $tag rc=$?"; tail -2 $OUT/${TAG}_bench_n8_$tag.err | cut -c1-300
done
timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 5 --skip-e2e > $OUT/${TAG}_bench_n8_s200.json 2> $OUT/${TAG}_bench_n8_s200.err
echo "bench n8 s200 rc=$?"
timeout 300 python bench.py --workload channelizer --log2-block 26 --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_c5_n1_log26.json 2> $OUT/${TAG}_c5_n1_log26.err
echo "c5 log26 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu --skip-ncu --skip-channelizer > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"
