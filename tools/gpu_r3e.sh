#!/bin/bash
set -u
TAG=${1:-r3e}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
