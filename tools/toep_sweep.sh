#!/bin/bash
# Toeplitz chain kernel: knob sweep (short benches) + one ncu --set full capture.   gpurun --timeout 900 -- 'bash tools/toep_sweep.sh'
OUT=gpurun_out
mkdir -p $OUT
WL=${WL:-am}
B="python bench.py --workload $WL --steps 60 --warmup 5 --skip-e2e --skip-cpu"
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 120 $B > $OUT/sw_$label.json 2> $OUT/sw_$label.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/sw_$label.json") if l.startswith("{")][-1])
    print("$label", round(d["roofline"]["kernel_ms"], 4), "ms", round(d["roofline"]["frac"], 3), d["config"].get("kernel_variant"))
except Exception as e:
    print("$label", "failed", e); print(open("$OUT/sw_$label.err").read()[-800:])
PY
}
run default A=1
run g1 B200SDR_TOEP_G=1
run g1w8 B200SDR_TOEP_G=1 B200SDR_TOEP_WARPS=8
run g1w6 B200SDR_TOEP_G=1 B200SDR_TOEP_WARPS=6
run g1s3 B200SDR_TOEP_G=1 B200SDR_TOEP_STAGES=3
run g1s3w6 B200SDR_TOEP_G=1 B200SDR_TOEP_STAGES=3 B200SDR_TOEP_WARPS=6
run g2s3w3 B200SDR_TOEP_G=2 B200SDR_TOEP_STAGES=3 B200SDR_TOEP_WARPS=3
run g2w8 B200SDR_TOEP_G=2 B200SDR_TOEP_WARPS=8
run g2w6 B200SDR_TOEP_G=2 B200SDR_TOEP_WARPS=6
run g2w5 B200SDR_TOEP_G=2 B200SDR_TOEP_WARPS=5
run g2w3 B200SDR_TOEP_G=2 B200SDR_TOEP_WARPS=3
run g2w2 B200SDR_TOEP_G=2 B200SDR_TOEP_WARPS=2
run aw2 B200SDR_TOEP_AUDIO_WARPS=2
BS="python bench.py --workload $WL --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu"
ncu --set full --clock-control none --import-source on -k regex:'toepKernel' -s 4 -c 1 -f -o $OUT/toep2_prof $BS > $OUT/toep2_ncu.log 2>&1
echo "ncu rc=$?"
