#!/bin/bash
# Scaling runs on ONE box: bench.py at N = 1, 2, 4, 8 for the headline chain (weak scaling, time segments) and the C5
# channelizer (strong scaling, channels).  gpurun --gpus 8 --timeout 900 -- 'bash tools/scale_run.sh'
OUT=gpurun_out/scale
mkdir -p $OUT
for N in ${NS:-1 2 4 8}; do
  for WL in am channelizer; do
    STEPS=$([ $WL = am ] && echo 200 || echo 20)
    if [ $N = 1 ]; then
      timeout -k 5 200 python bench.py --workload $WL --steps $STEPS --warmup 5 --skip-cpu > $OUT/${WL}_n$N.json 2> $OUT/${WL}_n$N.err
    else
      timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --workload $WL --steps $STEPS --warmup 5 --skip-cpu > $OUT/${WL}_n$N.json 2> $OUT/${WL}_n$N.err
    fi
    echo "$WL N=$N rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('$OUT/${WL}_n$N.json') if l.startswith('{')][-1]); print(round(d['value']), 'Msps', round(d['ms_per_step'],4), 'ms/step')" 2>&1 | tail -1)"
  done
done
