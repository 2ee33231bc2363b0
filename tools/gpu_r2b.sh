#!/bin/bash
# Round 2, second GPU visit: host-layer fixes, gather, new bench line.
set -u
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_gather.py tests/test_gpu_reference_dropin.py -q > $OUT/${TAG}_pytest_graph.log 2>&1
echo "pytest graph rc=$?"; tail -7 $OUT/${TAG}_pytest_graph.log
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
python bench.py --workload firsweep --steps 3 --warmup 1 > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err
echo "bench c4 rc=$?"; tail -3 $OUT/${TAG}_bench_c4.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2b_bench_c4.json'))
    print({k: d['roofline'][k] for k in ('frac_min', 'frac_median', 'frac_max', 'cells_at_or_above_0.70')}, d['config']['wall_seconds'])
    for T in [32, 64, 128, 256, 512, 1024, 2048, 4096]:
        print(T, ' '.join('%.2f%s' % (c['frac'], c['bound'][0]) for c in d['cells'] if c['taps'] == T))
except Exception as e:
    print('c4 parse failed', e)
PY
