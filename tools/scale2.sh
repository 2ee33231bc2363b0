#!/bin/bash
# N = 1 and N = 2 on one 2-GPU box (cheap check of the multi-GPU path).  gpurun --gpus 2 --timeout 600 -- 'bash tools/scale2.sh'
OUT=gpurun_out/scale2
mkdir -p $OUT
for WL in am wbfm channelizer; do
  STEPS=$([ $WL = channelizer ] && echo 10 || echo 200)
  timeout -k 5 200 python bench.py --workload $WL --steps $STEPS --warmup 5 --skip-cpu --skip-e2e > $OUT/${WL}_n1.json 2> $OUT/${WL}_n1.err
  timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload $WL --steps $STEPS --warmup 5 --skip-cpu --skip-e2e > $OUT/${WL}_n2.json 2> $OUT/${WL}_n2.err
  for N in 1 2; do
    echo "$WL N=$N $(python -c "import json; d=json.loads([l for l in open('$OUT/${WL}_n$N.json') if l.startswith('{')][-1]); print(round(d['value']), 'Msps', round(d['ms_per_step'],4), 'ms/step', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
  done
done
timeout 300 python -m pytest tests/test_gpu_channelizer.py tests/test_gpu_reference_dropin.py -m gpu -x -q 2>&1 | tail -3
# Filter-API driver with different step sizes (host memcpy into the pinned staging buffer dominates)
python - <<'PY'
import json, os, subprocess, tempfile, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
wl = bench.workload("am")
exe = "oracle/_ref/ref_chain_ours_hdr"
n = 1 << 24
with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
    np.random.default_rng(1).integers(-100, 101, size=2 * n, dtype=np.int8).tofile(tmp + "/in.i8")
    np.asarray(wl["t1"], dtype=np.float32).tofile(tmp + "/t1.f32"); np.asarray(wl["t2"], dtype=np.float32).tofile(tmp + "/t2.f32")
    base = ["--fs", repr(wl["fs"]), "--freq", repr(wl["f"]), "--mod", "am", "--d1", str(wl["d1"]), "--d2", str(wl["d2"]),
            "--taps1", tmp + "/t1.f32", "--taps2", tmp + "/t2.f32", "--in", tmp + "/in.i8", "--repeat", "16", "--fused", "1"]
    for step in (1 << 20, 4 << 20, 16 << 20, 64 << 20):
        r = subprocess.run([exe] + base + ["--step", str(step)], capture_output=True, text=True)
        print("filter api fused, step", step >> 20, "MiB:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
PY
