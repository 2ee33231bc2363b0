#!/bin/bash
# 1 GPU: pfb256 (last-tap skip, wrap-free loads) parity + C5
set -u
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload channelizer --steps 10 --warmup 3 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_c5_n1_$i.json 2> $OUT/${TAG}_c5_n1_$i.err
echo "c5 rc=$?"; cut -c120-300 $OUT/${TAG}_c5_n1_$i.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|window|direct' -s 8 -c 4 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launches rc=$?"
