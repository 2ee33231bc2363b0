#!/bin/bash
# 1 GPU: where does the Filter-API e2e loop wait?  copy and resident producers, several step sizes
set -u
TAG=${1:-r2w}
OUT=gpurun_out
mkdir -p $OUT
python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench
wl = bench.workload("am")
np.asarray(wl["t1"], dtype=np.float32).tofile("/dev/shm/t1.f32")
np.asarray(wl["t2"], dtype=np.float32).tofile("/dev/shm/t2.f32")
PY
COMMON="--fs 19200000 --freq -1234000 --mod am --d1 40 --d2 10 --taps1 /dev/shm/t1.f32 --taps2 /dev/shm/t2.f32 --samples-per-pass 268435456 --passes 6 --warmup-steps 4 --pipeline 1"
for prod in copy resident; do
for step in 67108864 33554432 134217728; do
echo "producer=$prod step=$step"
LD_LIBRARY_PATH=cuda_sdr_b200 timeout 120 build/bin/filter_api_bench $COMMON --step $step --threads 16 --producer $prod
done
done > $OUT/${TAG}_e2e.log 2>&1
cat $OUT/${TAG}_e2e.log | cut -c1-600
