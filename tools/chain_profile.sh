#!/bin/bash
# Where the compute warps of the chain kernel spend their cycles (B200SDR_CHAIN_PROFILE=1 instruments and synchronises).
for e in "B200SDR_CHAIN_STAGES=2" "B200SDR_CHAIN_RPT=2 B200SDR_CHAIN_STAGES=3" "B200SDR_CHAIN_CTAS=1 B200SDR_CHAIN_STAGES=3"; do
echo "== $e"; env B200SDR_FUSED=1 B200SDR_CHAIN_PROFILE=1 $e python bench.py --steps 3 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu 2>&1 | grep "chain profile" | tail -1
done
