#!/bin/bash
# 1 GPU: parity after the windowKernel staging rewrite and the padded / wide rows kernels; FIR sweep; C5 at N = 1
set -u
TAG=${1:-r2q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_channelizer.py tests/test_gpu_chain.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload firsweep --steps 5 --warmup 3 > $OUT/${TAG}_firsweep.json 2> $OUT/${TAG}_firsweep.err
echo "firsweep rc=$?"
timeout 300 python bench.py --workload channelizer --steps 10 --warmup 3 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_c5_n1.json 2> $OUT/${TAG}_c5_n1.err
echo "c5 rc=$?"; cut -c1-300 $OUT/${TAG}_c5_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|window|direct' -s 8 -c 6 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launches rc=$?"
