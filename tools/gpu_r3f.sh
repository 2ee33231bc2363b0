#!/bin/bash
set -u
TAG=${1:-r3f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_graph.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log | cut -c1-300
