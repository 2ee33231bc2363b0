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
run am pdl1 A=1
run am pdl0 B200SDR_TOEP_PDL=0
run am pdl1_aw3 B200SDR_TOEP_AUDIO_WARPS=3
run wbfm fm_pdl1 A=1
run wbfm fm_pdl0 B200SDR_TOEP_PDL=0
if [ -n "$NCU" ]; then
BS="python bench.py --workload am --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu"
ncu --set full --clock-control none --import-source on -k regex:'toepKernel' -s 4 -c 1 -f -o $OUT/${NCU}_prof $BS > $OUT/${NCU}_ncu.log 2>&1
echo "ncu rc=$?"
fi
