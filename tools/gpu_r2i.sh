#!/bin/bash
set -u
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_channelizer.py -q -x -k "small_mixed or eight_channels or sharding" > $OUT/${TAG}_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -15 $OUT/${TAG}_pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_channelizer.py tests/test_gpu_ops.py -q > $OUT/${TAG}_pytest_chan.log 2>&1; echo "pytest chan rc=$?"; tail -6 $OUT/${TAG}_pytest_chan.log
for tc in 1 0; do
B200SDR_PFB=0 B200SDR_CHANNEL_TC=$tc timeout 300 python bench.py --workload channelizer --log2-block 27 --steps 5 --warmup 3 > $OUT/${TAG}_bench_gemm_$tc.json 2> $OUT/${TAG}_bench_gemm_$tc.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_gemm_$tc.json')); print('gemm route tc=$tc', d['value'], d['ms_per_step'], d['config']['route'][:60])"
done
python bench.py --workload channelizer --steps 10 --warmup 3 > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_c5.json')); print('c5', d['value'], d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|window|channel' -s 8 -c 4 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 > $OUT/${TAG}_ncu_launches.log 2>&1
python - <<PY
import csv
for r in list(csv.reader(open('$OUT/${TAG}_c5_launches.csv')))[-4:]: print(r[4][:44], r[-3], r[-1])
PY
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
