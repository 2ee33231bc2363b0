#!/bin/bash
set -u
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_channelizer.py -q -x -k "pfb256_am_only" > $OUT/${TAG}_sanitizer.log 2>&1
echo "sanitizer rc=$?"; tail -15 $OUT/${TAG}_sanitizer.log
timeout 1200 python -m pytest tests/test_gpu_channelizer.py -q > $OUT/${TAG}_pytest_chan.log 2>&1
echo "pytest channelizer rc=$?"; tail -15 $OUT/${TAG}_pytest_chan.log
for pfb256 in 1 0; do
B200SDR_PFB256=$pfb256 python bench.py --workload channelizer --steps 10 --warmup 3 > $OUT/${TAG}_bench_c5_$pfb256.json 2> $OUT/${TAG}_bench_c5_$pfb256.err
echo "bench c5 pfb256=$pfb256 rc=$?"; tail -2 $OUT/${TAG}_bench_c5_$pfb256.err; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_c5_$pfb256.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['route'])"
done
B200SDR_PFB256=1 python bench.py --workload channelizer --log2-block 27 --steps 10 --warmup 3 > $OUT/${TAG}_bench_c5_27.json 2> $OUT/${TAG}_bench_c5_27.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_c5_27.json')); print('2^27:', d['value'], d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|window|direct' -s 8 -c 6 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 > $OUT/${TAG}_ncu_launches.log 2>&1
cat $OUT/${TAG}_c5_launches.csv | tail -8
