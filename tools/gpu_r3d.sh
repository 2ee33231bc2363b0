#!/bin/bash
# 1 GPU: the per-channel route of C5 on tcgen05 (bench line) and one ncu --set full capture of channelTcKernel
set -u
TAG=${1:-r3d}
OUT=gpurun_out
mkdir -p $OUT
B200SDR_PFB=0 timeout 300 python bench.py --workload channelizer --log2-block 27 --steps 5 --warmup 3 --skip-cpu --skip-e2e --skip-ncu > $OUT/${TAG}_bench_gemm_tc.json 2> $OUT/${TAG}_bench_gemm_tc.err
echo "bench tc rc=$?"; cut -c120-260 $OUT/${TAG}_bench_gemm_tc.json
B200SDR_PFB=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'channelTcKernel' -s 2 -c 1 -f -o $OUT/${TAG}_prof_tc \
  python bench.py --workload channelizer --log2-block 25 --steps 2 --warmup 1 --warmup-seconds 0 --skip-cpu --skip-e2e --skip-ncu > $OUT/${TAG}_ncu_tc.log 2>&1
echo "ncu tc rc=$?"
