#!/bin/bash
# A handful of chain-kernel configurations, kernel time only:  gpurun -- 'bash tools/sweep_small.sh [am|wbfm]'
WL=${1:-am}
run() {
  env "$@" python bench.py --workload $WL --steps 100 --warmup 10 --warmup-seconds 0.3 --skip-e2e --skip-cpu 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['roofline']['kernel_ms'],4), d['config']['kernel_variant'])"
}
run B200SDR_FUSED=0
run B200SDR_FUSED=1
run B200SDR_FUSED=1 B200SDR_CHAIN_AUDIO_WARPS=2
run B200SDR_FUSED=1 B200SDR_CHAIN_WARPS=6
run B200SDR_FUSED=1 B200SDR_CHAIN_WARPS=4
run B200SDR_FUSED=1 B200SDR_CHAIN_RPT=1 B200SDR_CHAIN_WARPS=12
run B200SDR_FUSED=1 B200SDR_CHAIN_RPT=1 B200SDR_CHAIN_WARPS=8
run B200SDR_FUSED=1 B200SDR_CHAIN_STAGES=1
run B200SDR_FUSED=1 B200SDR_CHAIN_MMA=0 B200SDR_CHAIN_AUDIO_WARPS=1
