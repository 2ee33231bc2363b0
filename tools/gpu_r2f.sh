#!/bin/bash
set -u
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_channelizer.py tests/test_gpu_ops.py -q > $OUT/${TAG}_pytest_chan.log 2>&1
echo "pytest channelizer+ops rc=$?"; tail -8 $OUT/${TAG}_pytest_chan.log
python bench.py --workload channelizer --steps 10 --warmup 3 > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
echo "bench c5 rc=$?"; tail -2 $OUT/${TAG}_bench_c5.err; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_c5.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|window|direct' -s 8 -c 4 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 > $OUT/${TAG}_ncu_launches.log 2>&1
cut -d, -f5,15 $OUT/${TAG}_c5_launches.csv | tail -5
python bench.py --workload firsweep --steps 3 --warmup 1 > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err
echo "bench c4 rc=$?"; tail -3 $OUT/${TAG}_bench_c4.err; python - <<PY
import json
try:
    d = json.load(open('$OUT/${TAG}_bench_c4.json'))
    print({k: d['roofline'][k] for k in ('frac_min', 'frac_median', 'frac_max', 'cells_at_or_above_0.70')})
    for T in [32, 64, 128, 256, 512, 1024, 2048, 4096]:
        print(T, ' '.join('%.2f%s' % (c['frac'], c['bound'][0]) for c in d['cells'] if c['taps'] == T))
except Exception as e:
    print('c4 parse failed', e)
PY
C3_SHORT="python bench.py --workload wbfm --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu --skip-channelizer --skip-ncu"
ncu --set full --clock-control none --import-source on -k regex:'toepKernel' -s 4 -c 1 -f -o $OUT/${TAG}_prof_c3 $C3_SHORT > $OUT/${TAG}_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
ncu -i $OUT/${TAG}_prof_c3.ncu-rep --page raw --csv > $OUT/${TAG}_c3_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof_c3.ncu-rep --page source --csv > $OUT/${TAG}_c3_source.csv 2>/dev/null
