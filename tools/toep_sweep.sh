#!/bin/bash
# Toeplitz chain kernel: knob sweep (short benches) + one ncu --set full capture.   gpurun --timeout 900 -- 'bash tools/toep_sweep.sh'
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 60 --warmup 5 --skip-e2e --skip-cpu"
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
run g1s3 B200SDR_TOEP_G=1 B200SDR_TOEP_STAGES=3
run s3 B200SDR_TOEP_STAGES=3
run w8 B200SDR_TOEP_WARPS=8
run w6 B200SDR_TOEP_WARPS=6
run w3 B200SDR_TOEP_WARPS=3
run i2f B200SDR_TOEP_MAGIC=0
run aw2 B200SDR_TOEP_AUDIO_WARPS=2
BS="python bench.py --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu"
ncu --set full --clock-control none --import-source on -k regex:'toepKernel' -s 4 -c 1 -f -o $OUT/toep1_prof $BS > $OUT/toep1_ncu.log 2>&1
echo "ncu rc=$?"
