#!/bin/bash
# The round's final GPU visit (one B200): smoke, all GPU parity tests, the reference arm, the default bench line, C3 / C5 / C4
# lines, the launch list of the bench command and ncu --set full captures of the C2 kernel and the final C5 kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2_final.sh r2_final'
set -u
TAG=${1:-r2_final}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "bench reference rc=$?"
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
timeout 300 python bench.py --workload wbfm --steps 200 --warmup 5 --skip-cpu --skip-e2e > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err
echo "bench c3 rc=$?"
timeout 300 python bench.py --workload channelizer --steps 20 --warmup 3 --skip-cpu --skip-e2e > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
echo "bench c5 rc=$?"
timeout 300 python bench.py --workload firsweep --steps 5 --warmup 3 > $OUT/${TAG}_firsweep.json 2> $OUT/${TAG}_firsweep.err
echo "firsweep rc=$?"
BENCH_SHORT="python bench.py --steps 5 --warmup 3 --warmup-seconds 0 --skip-e2e --skip-cpu --skip-ncu --skip-channelizer"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -s 6 -c 12 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH_SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'toepKernel' -s 6 -c 1 -f -o $OUT/${TAG}_prof_c2 $BENCH_SHORT > $OUT/${TAG}_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
C5_SHORT="python bench.py --workload channelizer --log2-block 27 --steps 3 --warmup 3 --warmup-seconds 0 --skip-cpu --skip-e2e --skip-ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pfb256Kernel|windowKernel' -s 8 -c 2 -f -o $OUT/${TAG}_prof_c5 $C5_SHORT > $OUT/${TAG}_ncu_c5.log 2>&1
echo "ncu c5 rc=$?"
ls -la $OUT/${TAG}_* | head -40
