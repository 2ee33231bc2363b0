#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/pfb_check.sh'
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_channelizer.py -m gpu -x -q > $OUT/pfb_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pfb_pytest.log
B="python bench.py --workload channelizer --steps 4 --warmup 3 --warmup-seconds 0 --skip-cpu"
timeout 200 python bench.py --workload channelizer --steps 10 --warmup 3 --skip-cpu > $OUT/pfb_bench_1.json 2> $OUT/pfb_bench_1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$OUT/pfb_bench_1.json") if l.startswith("{")][-1])
print(round(d["value"]), "Msps", round(d["ms_per_step"], 4), "ms/step", d["config"].get("kernel_variant"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfbKernel|directKernel|windowKernel' -s 8 -c 4 --csv --log-file $OUT/pfb_launches.csv $B > $OUT/pfb_ncu.log 2>&1
grep -v "^==" $OUT/pfb_launches.csv | cut -d, -f5,15 | tail -5
ncu --set full --clock-control none --import-source on -k regex:'pfbKernel' -s 3 -c 1 -f -o $OUT/pfb1_prof $B > $OUT/pfb1_ncu.log 2>&1; echo "ncu rc=$?"
