#!/bin/bash
# N = 2 on one 2-GPU box: how often to gather.  gpurun --gpus 2 --timeout 600 -- 'bash tools/scale2.sh'
OUT=gpurun_out/scale2
mkdir -p $OUT
for G in 8 32 64; do
  BENCH_GATHER_EVERY=$G timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload am --steps 256 --warmup 5 --skip-cpu --skip-e2e > $OUT/am_g$G.json 2> $OUT/am_g$G.err
  echo "am N=2 gather_every=$G $(python -c "import json; d=json.loads([l for l in open('$OUT/am_g$G.json') if l.startswith('{')][-1]); print(round(d['value']), 'Msps', round(d['ms_per_step'],4), 'ms/step')" 2>&1 | tail -1)"
done
timeout 100 python bench.py --workload am --steps 256 --warmup 5 --skip-cpu --skip-e2e > $OUT/am_n1.json 2>$OUT/am_n1.err
echo "am N=1 $(python -c "import json; d=json.loads([l for l in open('$OUT/am_n1.json') if l.startswith('{')][-1]); print(round(d['value']), 'Msps', round(d['ms_per_step'],4), 'ms/step')")"
