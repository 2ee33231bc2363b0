#!/bin/bash
# One GPU visit: parity tests, the bench line, the ncu launch list and one --set full capture of K1/K2.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
# Everything lands in gpurun_out/<tag>_*; summarise with tools/ncu_summary.py into profiles/.
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "bench reference rc=$?"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
cat $OUT/${TAG}_bench.json
BENCH_SHORT="python bench.py --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu"
KERNELS='rowsKernel|directKernel|chainKernel|channelKernel|toepKernel'
$BENCH_SHORT > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -s 6 -c 10 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH_SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s 6 -c 2 \
    -f -o $OUT/${TAG}_prof $BENCH_SHORT > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"

# the channelizer workload (C5): bench line + ncu --set full of its two kernels (short block for the capture)
python bench.py --workload channelizer --steps 20 --warmup 3 --skip-cpu > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
echo "bench c5 rc=$?"
python bench.py --workload wbfm --steps 200 --warmup 5 --skip-cpu --skip-e2e > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err
echo "bench c3 rc=$?"
C5_SHORT="python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 --skip-cpu"
ncu --set full --clock-control none --import-source on -k regex:'pfbKernel|windowKernel' -s 6 -c 2 \
    -f -o $OUT/${TAG}_prof_c5 $C5_SHORT > $OUT/${TAG}_ncu_full_c5.log 2>&1
echo "ncu c5 rc=$?"
