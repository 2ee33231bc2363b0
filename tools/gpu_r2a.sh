#!/bin/bash
# Round 2, first GPU visit: the new parity/graph tests, then the whole GPU suite, the Filter-API throughput, baseline bench lines.
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_reference_dropin.py -q -x > $OUT/${TAG}_pytest_graph.log 2>&1
echo "pytest graph rc=$?"; tail -5 $OUT/${TAG}_pytest_graph.log
timeout 900 python -m pytest tests/test_gpu_channelizer.py -q > $OUT/${TAG}_pytest_chan.log 2>&1
echo "pytest channelizer rc=$?"; tail -5 $OUT/${TAG}_pytest_chan.log
timeout 900 python -m pytest tests/test_gpu_chain.py -q -k "ring or full_size" > $OUT/${TAG}_pytest_ring.log 2>&1
echo "pytest ring rc=$?"; tail -5 $OUT/${TAG}_pytest_ring.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest all rc=$?"; tail -5 $OUT/${TAG}_pytest.log
# Filter-API throughput: producer writes into the pinned block, one fused node, event-pipelined D2H
python - <<'PY' > $OUT/${TAG}_taps.log 2>&1
import numpy as np, sys
sys.path.insert(0, '.')
from cuda_sdr_b200 import taps
fs=19.2e6
taps.lowpass(101, 0.45*fs/40, fs).astype(np.float32).tofile('/dev/shm/t1.f32')
taps.lowpass(129, 0.45*48e3, fs/40).astype(np.float32).tofile('/dev/shm/t2.f32')
PY
for pipe in 0 1; do for step in 4 64; do
  oracle/_ref/ref_chain_ours_hdr --fs 19.2e6 --freq -1.234e6 --mod am --d1 40 --d2 10 --taps1 /dev/shm/t1.f32 --taps2 /dev/shm/t2.f32 \
    --fused 1 --synth-samples $((1<<28)) --step $((step<<20)) --repeat 6 --warmup-steps 4 --pipeline $pipe 2>&1 | tail -1
done; done | tee $OUT/${TAG}_filter_api.jsonl
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
python bench.py --workload wbfm --steps 50 --warmup 5 --skip-cpu --skip-e2e > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err
echo "bench c3 rc=$?"; cat $OUT/${TAG}_bench_c3.json
python bench.py --workload channelizer --steps 10 --warmup 3 --skip-cpu > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
echo "bench c5 rc=$?"; cat $OUT/${TAG}_bench_c5.json
