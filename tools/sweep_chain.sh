#!/bin/bash
# Tuning sweep of the fused chain kernel's knobs on one B200 (device-resident, kernel time only).
#   gpurun --timeout 900 -- 'bash tools/sweep_chain.sh [am|wbfm]'  ->  gpurun_out/sweep_<workload>.jsonl
WL=${1:-am}
OUT=gpurun_out/sweep_$WL.jsonl
mkdir -p gpurun_out; : > $OUT
run() {
  env "$@" python bench.py --workload $WL --steps 100 --warmup 10 --warmup-seconds 0.3 --skip-e2e --skip-cpu 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'env': '$*', 'kernel_ms': d['roofline']['kernel_ms'], 'frac': d['roofline']['frac'], 'variant': d['config']['kernel_variant']}))" >> $OUT
}
run B200SDR_FUSED=0
for rpt in 4 2; do
  for warps in 4 8; do
    for stages in 1 2 3; do
      run B200SDR_FUSED=1 B200SDR_CHAIN_RPT=$rpt B200SDR_CHAIN_WARPS=$warps B200SDR_CHAIN_STAGES=$stages
    done
  done
done
run B200SDR_FUSED=1 B200SDR_CHAIN_MMA=0
run B200SDR_FUSED=1 B200SDR_CHAIN_CTAS=1
cat $OUT
