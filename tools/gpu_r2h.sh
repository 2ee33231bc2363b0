#!/bin/bash
set -u
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_chain.py -q -x > $OUT/${TAG}_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -5 $OUT/${TAG}_pytest_chain.log
for tiled in 1 0; do
B200SDR_TOEP_TILED_AUDIO=$tiled python bench.py --workload wbfm --steps 100 --warmup 5 --skip-e2e --skip-cpu --skip-channelizer --skip-ncu > $OUT/${TAG}_bench_c3_$tiled.json 2> $OUT/${TAG}_bench_c3_$tiled.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_c3_$tiled.json')); print('c3 tiled=$tiled', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:'pfb|window|toep' -s 8 -c 4 --csv --log-file $OUT/${TAG}_c5_launches.csv \
   python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 > $OUT/${TAG}_ncu_launches.log 2>&1
python - <<PY
import csv
for r in list(csv.reader(open('$OUT/${TAG}_c5_launches.csv')))[-12:]: print(r[4][:44], r[-3], r[-1])
PY
