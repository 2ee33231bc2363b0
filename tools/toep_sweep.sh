#!/bin/bash
# Toeplitz chain kernel: parity subset, knob sweep (short benches), optional ncu --set full capture (NCU=tag).
#   gpurun --timeout 900 -- 'NCU=toep5 bash tools/toep_sweep.sh'
OUT=gpurun_out
mkdir -p $OUT
run() {  # workload, label, env...
  local wl=$1; local label=$2; shift; shift
  env "$@" timeout 120 python bench.py --workload $wl --steps 60 --warmup 5 --skip-e2e --skip-cpu > $OUT/sw_$label.json 2> $OUT/sw_$label.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/sw_$label.json") if l.startswith("{")][-1])
    print("$label", round(d["roofline"]["kernel_ms"], 4), "ms kernel", round(d["ms_per_step"], 4), "ms/step", round(d["roofline"]["frac"], 3), d["config"].get("kernel_variant"))
except Exception as e:
    print("$label", "failed", e); print(open("$OUT/sw_$label.err").read()[-800:])
PY
}
timeout 400 python -m pytest tests/test_gpu_chain.py -m gpu -x -q > $OUT/sw_pytest.log 2>&1; tail -4 $OUT/sw_pytest.log
run am default A=1
run am aw1 B200SDR_TOEP_AUDIO_WARPS=1
run am aw3 B200SDR_TOEP_AUDIO_WARPS=3
run am g1w6 B200SDR_TOEP_G=1 B200SDR_TOEP_WARPS=6

run am w8 B200SDR_TOEP_WARPS=8
run wbfm fm_default A=1
run wbfm fm_g2 B200SDR_TOEP_G=2
run wbfm fm_aw2 B200SDR_TOEP_AUDIO_WARPS=2
run wbfm fm_g2w4 B200SDR_TOEP_G=2 B200SDR_TOEP_WARPS=4
if [ -n "$NCU" ]; then
BS="python bench.py --workload am --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu"
ncu --set full --clock-control none --import-source on -k regex:'toepKernel' -s 4 -c 1 -f -o $OUT/${NCU}_prof $BS > $OUT/${NCU}_ncu.log 2>&1
echo "ncu rc=$?"
fi
