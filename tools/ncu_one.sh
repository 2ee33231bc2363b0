#!/bin/bash
# ncu --set full of the chain kernel under the environment given on the command line:
#   gpurun -- 'bash tools/ncu_one.sh tag B200SDR_FUSED=1 B200SDR_CHAIN_CTAS=1'
TAG=$1; shift
CMD="python bench.py --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu"
env "$@" $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
env "$@" ncu --set full --clock-control none --import-source on -k regex:'rowsKernel|directKernel|chainKernel|channelKernel' -s 4 -c 1 \
    -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
