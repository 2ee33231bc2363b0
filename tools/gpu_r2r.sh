#!/bin/bash
# 1 GPU: paired-stream audio kernel parity + C5; ncu of the M = 16 / 32 rows cells
set -u
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_channelizer.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload channelizer --steps 10 --warmup 3 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_c5_n1.json 2> $OUT/${TAG}_c5_n1.err
echo "c5 rc=$?"; cut -c1-300 $OUT/${TAG}_c5_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|window|direct' -s 8 -c 4 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'Kernel' -c 3 -f -o $OUT/${TAG}_prof_cells python tools/fir_cells.py 1024x32 2048x64 1024x64 > $OUT/${TAG}_ncu_cells.log 2>&1
echo "ncu cells rc=$?"; grep "T=" $OUT/${TAG}_ncu_cells.log
