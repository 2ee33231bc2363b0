#!/bin/bash
# 8 GPUs: balanced partition calibrated under load + peer-memory gather; the pure-kernel floor (diag); NCCL transport beside it
set -u
TAG=${1:-r2n}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
for steps in 20 200; do
BENCH_GATHER=peer timeout 300 $TR bench.py --gpus 8 --steps $steps --warmup 5 --skip-e2e > $OUT/${TAG}_bench_n8_s$steps.json 2> $OUT/${TAG}_bench_n8_s$steps.err
echo "bench n8 steps=$steps rc=$?"; tail -2 $OUT/${TAG}_bench_n8_s$steps.err | cut -c1-300
done
BENCH_GATHER=peer timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --skip-e2e > $OUT/${TAG}_bench_n8_s20b.json 2> $OUT/${TAG}_bench_n8_s20b.err
echo "bench n8 steps=20 (repeat) rc=$?"
BENCH_GATHER=nccl timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --skip-e2e > $OUT/${TAG}_bench_n8_s20_nccl.json 2> $OUT/${TAG}_bench_n8_s20_nccl.err
echo "bench n8 steps=20 nccl rc=$?"
timeout 200 $TR tools/diag_scale.py --steps 100 > $OUT/${TAG}_diag_n8.json 2> $OUT/${TAG}_diag_n8.err
echo "diag rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench n1 rc=$?"
