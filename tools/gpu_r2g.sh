#!/bin/bash
# 2 GPUs: why is a step slower when both run? + tcgen05 probe + window kernel check
set -u
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 build/bin/tc_probe > $OUT/${TAG}_tc_probe.log 2>&1; echo "tc_probe rc=$?"; tail -6 $OUT/${TAG}_tc_probe.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/diag_scale.py --steps 200 \
  > $OUT/${TAG}_diag.json 2> $OUT/${TAG}_diag.err
echo "diag rc=$?"; tail -3 $OUT/${TAG}_diag.err; cat $OUT/${TAG}_diag.json
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_channelizer.py -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
python bench.py --workload channelizer --log2-block 27 --steps 10 --warmup 3 > $OUT/${TAG}_bench_c5_27.json 2> $OUT/${TAG}_bench_c5_27.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_c5_27.json')); print('c5 2^27:', d['value'], d['ms_per_step'])"
python bench.py --workload firsweep --steps 3 --warmup 1 > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err
python - <<PY
import json
try:
    d = json.load(open('$OUT/${TAG}_bench_c4.json'))
    print({k: d['roofline'][k] for k in ('frac_min', 'frac_median', 'frac_max', 'cells_at_or_above_0.70')})
    for T in [32, 64, 128, 256, 512, 1024, 2048, 4096]:
        print(T, ' '.join('%.2f%s' % (c['frac'], c['bound'][0]) for c in d['cells'] if c['taps'] == T))
except Exception as e:
    print('c4 parse failed', e)
PY
