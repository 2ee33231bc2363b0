#!/bin/bash
set -u
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-channelizer > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
