#!/bin/bash
# 1 GPU: new sweep parity tests; ncu --set full of the C5 kernels (pfb256 + audio windowKernel) and of three FIR sweep cells
set -u
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "sweep" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
C5_SHORT="python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu --skip-ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pfb256Kernel|windowKernel' -s 8 -c 2 -f -o $OUT/${TAG}_prof_c5 $C5_SHORT > $OUT/${TAG}_ncu_c5.log 2>&1
echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'Kernel' -c 3 -f -o $OUT/${TAG}_prof_cells python tools/fir_cells.py 4096x1 1024x64 32x64 > $OUT/${TAG}_ncu_cells.log 2>&1
echo "ncu cells rc=$?"; tail -4 $OUT/${TAG}_ncu_cells.log
ls -la $OUT/${TAG}_*
