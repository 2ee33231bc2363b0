#!/bin/bash
# 1 GPU: FIR / channelizer parity after the windowKernel store change, C5 at N=1, a 2-rank-free sanity run of the balanced bench path
set -u
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_channelizer.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload channelizer --steps 10 --warmup 3 --skip-e2e --skip-cpu --skip-ncu > $OUT/${TAG}_c5_n1.json 2> $OUT/${TAG}_c5_n1.err
echo "c5 rc=$?"; cut -c1-600 $OUT/${TAG}_c5_n1.json
timeout 300 python bench.py --workload firsweep --steps 5 --warmup 3 > $OUT/${TAG}_firsweep.json 2> $OUT/${TAG}_firsweep.err
echo "firsweep rc=$?"
