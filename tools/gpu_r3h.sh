#!/bin/bash
set -u
OUT=gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-channelizer --skip-ncu > $OUT/r3h_bench.json 2> $OUT/r3h_bench.err
echo "bench rc=$?"; cut -c1-260 $OUT/r3h_bench.json; tail -2 $OUT/r3h_bench.err
