#!/bin/bash
# N = 1 and N = 2 on one 2-GPU box (cheap check of the multi-GPU paths).  gpurun --gpus 2 --timeout 600 -- 'bash tools/scale2.sh'
OUT=gpurun_out/scale2
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_channelizer.py -m gpu -x -q 2>&1 | tail -3
for WL in am channelizer; do
  STEPS=$([ $WL = channelizer ] && echo 20 || echo 200)
  timeout -k 5 200 python bench.py --workload $WL --steps $STEPS --warmup 5 --skip-cpu --skip-e2e > $OUT/${WL}_n1.json 2> $OUT/${WL}_n1.err
  timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload $WL --steps $STEPS --warmup 5 --skip-cpu --skip-e2e > $OUT/${WL}_n2.json 2> $OUT/${WL}_n2.err
  for N in 1 2; do
    echo "$WL N=$N $(python -c "import json; d=json.loads([l for l in open('$OUT/${WL}_n$N.json') if l.startswith('{')][-1]); print(round(d['value']), 'Msps', round(d['ms_per_step'],4), 'ms/step', round(d['roofline']['frac'],3), d['config']['parallelism'][:40])" 2>&1 | tail -1)"
  done
done
tail -3 $OUT/channelizer_n2.err
