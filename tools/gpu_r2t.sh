#!/bin/bash
# 1 GPU: rows unroll parity + sweep; full default bench (e2e through the Filter API with the pooled producer)
set -u
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "fir" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload firsweep --steps 5 --warmup 3 > $OUT/${TAG}_firsweep.json 2> $OUT/${TAG}_firsweep.err
echo "firsweep rc=$?"
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err | cut -c1-300
