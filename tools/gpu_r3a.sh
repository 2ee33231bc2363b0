#!/bin/bash
set -u
TAG=${1:-r3a}
OUT=gpurun_out
mkdir -p $OUT
B200SDR_NO_WIDE_ROWS=1 timeout 300 python bench.py --workload firsweep --steps 5 --warmup 3 > $OUT/${TAG}_firsweep_nowide.json 2> $OUT/${TAG}_firsweep_nowide.err
echo "firsweep rc=$?"
