#!/bin/bash
set -u
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
C5_SHORT="python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0"
ncu --set full --clock-control none --import-source on -k regex:'pfb256Kernel' -s 4 -c 1 -f -o $OUT/${TAG}_prof_pfb256 $C5_SHORT > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/${TAG}_prof_pfb256.ncu-rep --page raw --csv > $OUT/${TAG}_pfb256_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof_pfb256.ncu-rep --page source --csv > $OUT/${TAG}_pfb256_source.csv 2>/dev/null
ls -la $OUT/${TAG}_*
