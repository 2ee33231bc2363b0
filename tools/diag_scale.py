#!/usr/bin/env python3
"""Why is a C2 step slower when several GPUs run at once?  Run under torchrun (one rank per GPU):

  torchrun --nproc-per-node N tools/diag_scale.py [--steps 200]

Per rank, ms per step of the fused chain kernel in these settings (CUDA events around the loop, barrier before each):
  alone_r   : only rank r runs, the others idle                       (the single-GPU figure of that very GPU)
  together  : every rank runs the same loop at once, NO gather at all (platform effects: power, clocks, host)
  gather_K  : together, with b200sdr_gather every K steps             (NCCL kernels next to the persistent kernel)
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    args = ap.parse_args()
    ctx = bench.Ctx(args)
    import torch
    import cuda_sdr_b200 as sdr

    wl = bench.workload("am")
    n = 1 << 28
    chain = sdr.Chain(wl["fs"], wl["f"], wl["t1"], wl["d1"], wl["mod"], fm_gain=wl["gain"], audio_taps=wl["t2"], audio_decim=wl["d2"], device=ctx.local_rank)
    n_audio = chain.counts(n)[2]
    x = sdr.synth.device_int8_iq(n, ctx.dev, seed=1 + ctx.rank)
    out = torch.empty(4, n_audio, dtype=torch.float32, device=ctx.dev)

    def loop(steps, gather=None, every=0, views=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            if gather is None:
                chain.process_device(x, 0, out=out[k & 3])
            else:
                slab, slot = (k // every) % 3, k % every
                if slot == 0:
                    gather.acquire(slab)
                chain.process_device(x, 0, out=views[slab][slot])
                if slot == every - 1 or k == steps - 1:
                    gather.submit(slab, [(slot + 1) * n_audio] * ctx.world)
        if gather is not None:
            gather.finish()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # warm up (clocks)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:
        chain.process_device(x, 0, out=out[0])
    torch.cuda.synchronize()
    res = {}
    for r in range(ctx.world):
        ctx.barrier()
        if ctx.rank == r:
            loop(20)
            res[f"alone_{r}"] = loop(args.steps)
        ctx.barrier()
    for rep in range(2):
        ctx.barrier()
        loop(20)
        ctx.barrier()
        res[f"together_{rep}"] = loop(args.steps)
    for every in (32, 5, 1):
        g = ctx.make_gather([every * n_audio] * ctx.world, 3)
        views = [g.slab(s).view(every, n_audio) for s in range(3)]
        ctx.barrier()
        loop(2 * every + 3, g, every, views)
        g.finish()
        ctx.barrier()
        res[f"gather_{every}"] = loop(args.steps, g, every, views)
        ctx.barrier()
        g.close()
    allres = ctx.gather_objects(res)
    if ctx.rank == 0:
        print(json.dumps({"world": ctx.world, "steps": args.steps, "ms_per_step_by_rank": allres}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
