#!/bin/bash
# Quick GPU visit for the Toeplitz chain kernel: chain parity tests, then a short bench of both routes.
#   gpurun --timeout 900 -- 'bash tools/toep_check.sh'
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -x -q > $OUT/toep_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/toep_pytest.log
B="python bench.py --steps 100 --warmup 5 --skip-e2e --skip-cpu"
for WL in am wbfm; do
  timeout 120 $B --workload $WL > $OUT/toep_bench_$WL.json 2> $OUT/toep_bench_$WL.err; echo "$WL toeplitz rc=$?"
  B200SDR_TOEPLITZ=0 timeout 120 $B --workload $WL > $OUT/toep_bench_${WL}_old.json 2> $OUT/toep_bench_${WL}_old.err; echo "$WL old rc=$?"
  python - <<PY
import json
for tag in ("$WL", "${WL}_old"):
    try:
        d = json.loads([l for l in open(f"$OUT/toep_bench_{tag}.json") if l.startswith("{")][-1])
        print(tag, round(d["ms_per_step"], 4), "ms", round(d["roofline"]["frac"], 3), d["config"].get("kernel_variant"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"$OUT/toep_bench_{tag}.err").read()[-1500:])
PY
done
