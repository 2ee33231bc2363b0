#!/usr/bin/env python3
"""Headline benchmark: input Msps through the int8 -> mix -> FIR -> demod -> audio-FIR chain.

  python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the fp64 oracle port on host cores

Workload = BASELINE.json configs[1] (C2, AM broadcast chain): per GPU and per step one block of 2^28
synthetic int8 IQ samples (512 MiB, larger than L2, so no L2 flush is needed between iterations)
-> mix by -1.234 MHz -> 101-tap low-pass, decimate by 40 -> |.| -> 129-tap low-pass, decimate by 10
-> 48 kHz-class audio.  At N > 1 every rank processes its own time segment of one long stream
(no collective on the filter path) and the decimated audio is gathered to rank 0 over NCCL inside the
timed region; per-GPU work is fixed, so scaling is "weak".

One JSON line is printed by rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 19.2e6            # 48 kHz x 40 x 10 ("20 Msps-class")
F_SHIFT = -1.234e6
T1, D1, T2, D2 = 101, 40, 129, 10
LOG2_BLOCK = 28
METRIC = "input Msps through int8->mix->FIR->demod chain"
UNIT = "Msps"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--warmup-seconds", type=float, default=0.5,
                    help="keep running untimed warm-up steps until this much wall time has passed (a step is ~0.2 ms; "
                         "the SM clock needs far longer than 3 steps to leave its idle state)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--log2-block", type=int, default=LOG2_BLOCK, help="log2 of input samples per GPU per step")
    ap.add_argument("--workload", choices=["am", "wbfm", "channelizer"], default="am")
    ap.add_argument("--channels", type=int, default=256, help="channelizer workload: total channels (sharded over the GPUs)")
    ap.add_argument("--shard", choices=["auto", "channels", "time"], default="auto",
                    help="channelizer workload, N > 1: shard by channel or by time segment (auto: time on the filter-bank route)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def workload(name: str):
    from cuda_sdr_b200 import taps
    if name == "am":
        return dict(name="C2 AM broadcast chain", fs=FS, f=F_SHIFT, t1=taps.lowpass(T1, 0.45 * FS / D1, FS), d1=D1, mod=0, gain=1.0,
                    t2=taps.lowpass(T2, 0.45 * 48e3, FS / D1), d2=D2)
    from cuda_sdr_b200 import fm_gain
    return dict(name="C3 WBFM chain", fs=FS, f=2.5e6, t1=taps.lowpass(545, 100e3, FS), d1=80, mod=1, gain=fm_gain(FS / 80, 75e3),
                t2=taps.lowpass(273, 0.45 * 48e3, FS / 80), d2=5)


def config_dict(args, wl, extra=None):
    n = 1 << args.log2_block
    cfg = {
        "workload": f"{wl['name']}: 2^{args.log2_block} synthetic int8 IQ samples per GPU per step at {wl['fs'] / 1e6:.1f} Msps-class rate "
                    f"-> mix -> {len(wl['t1'])}-tap FIR decimate-by-{wl['d1']} -> {'AM' if wl['mod'] == 0 else 'FM'} demod -> "
                    f"{len(wl['t2'])}-tap audio FIR decimate-by-{wl['d2']}",
        "samples_per_gpu_per_step": n,
        "rf_taps": len(wl["t1"]), "rf_decimation": wl["d1"], "audio_taps": len(wl["t2"]), "audio_decimation": wl["d2"],
        "modulation": "am" if wl["mod"] == 0 else "fm",
        "phase_mode": "exact (64-bit fixed-point turns)",
        "l2": f"input block {2 * n >> 20} MiB per step exceeds the 126 MB L2; no flush between iterations",
        "parallelism": "overlapped time segments, one per GPU; NCCL send/recv gather of the audio to rank 0 on a side stream, "
                       "overlapping the next step's kernel",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range seen
        lo, hi = min(power), max(power)
        loaded = [c for c, p in zip(sm, power) if p >= lo + 0.5 * (hi - lo)] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": hi}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (fp64, OpenMP) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_chain_rate(wl, target_seconds: float, max_log2: int = 26):
    """Returns (Msps, cores, sample description, samples, seconds) for the fp64 oracle on host cores."""
    import numpy as np
    from oracle import oracle as orc

    spec = orc.ChainSpec(wl["fs"], wl["f"], wl["t1"], wl["d1"], wl["mod"], wl["gain"], wl["t2"], wl["d2"])
    rng = np.random.default_rng(0x5D120001)
    probe = 1 << 20
    x = rng.integers(-100, 101, size=2 * probe, dtype=np.int8)
    orc.chain(spec, x)  # warm up threads and pages
    t0 = time.perf_counter()
    orc.chain(spec, x)
    rate = probe / (time.perf_counter() - t0)
    log2 = max(20, min(max_log2, int(np.floor(np.log2(max(rate * target_seconds, 1.0))))))
    n = 1 << log2
    x = rng.integers(-100, 101, size=2 * n, dtype=np.int8)
    t0 = time.perf_counter()
    out, _, _ = orc.chain(spec, x)
    dt = time.perf_counter() - t0
    assert out.size == spec.num_outputs(n)
    return n / dt / 1e6, orc.num_threads(), f"2^{log2} samples of the same workload, one pass, fp64 oracle (oracle/oracle.c), OpenMP", n, dt


def reference_api_rates(wl, log2_samples: int = 24):
    """Throughput of the chain driven through the REFERENCE'S Filter API in <= 1 MiB steps, host buffers in and out
    (oracle/ref/ref_chain.cpp, built by oracle/ref/build_ref.sh into oracle/_ref/):
      reference_cuda_pipeline : the reference's own host framework (compiled in place) + a plain restated gsdr  (= B1)
      ours_filter_api_fused   : this repo's libgpusdrpipeline.so, the same Filter contract, ONE fused node, 4 MiB steps
    Returns {} when the binaries are not there (they need /root/reference at build time)."""
    import numpy as np

    out = {}
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    naive, ours = os.path.join(ref_dir, "ref_chain_naive"), os.path.join(ref_dir, "ref_chain_ours_hdr")
    if not os.path.exists(naive) and not os.path.exists(ours):
        return out
    n = 1 << log2_samples
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        rng = np.random.default_rng(0x5D120001)
        rng.integers(-100, 101, size=2 * n, dtype=np.int8).tofile(os.path.join(tmp, "in.i8"))
        np.asarray(wl["t1"], dtype=np.float32).tofile(os.path.join(tmp, "t1.f32"))
        np.asarray(wl["t2"], dtype=np.float32).tofile(os.path.join(tmp, "t2.f32"))
        base = ["--fs", repr(wl["fs"]), "--freq", repr(wl["f"]), "--mod", "am" if wl["mod"] == 0 else "fm", "--d1", str(wl["d1"]),
                "--d2", str(wl["d2"]), "--taps1", os.path.join(tmp, "t1.f32"), "--taps2", os.path.join(tmp, "t2.f32"),
                "--in", os.path.join(tmp, "in.i8")]
        runs = [("reference_cuda_pipeline", naive, ["--repeat", "4"],
                 "reference host framework compiled in place + restated one-thread-per-output gsdr kernels (B1), <= 1 MiB steps, "
                 "pinned host in / host out"),
                ("ours_filter_api_fused", ours, ["--repeat", "16", "--fused", "1", "--step", str(4 << 20)],
                 "this repo's libgpusdrpipeline.so through the same Filter contract: CudaMemcpy -> ONE fused node -> CudaMemcpy, "
                 "4 MiB steps")]
        for key, exe, extra, what in runs:
            if not os.path.exists(exe):
                continue
            try:
                res = subprocess.run([exe] + base + extra, capture_output=True, text=True, timeout=300)
                info = json.loads(res.stdout.strip().splitlines()[-1])
                out[key] = {"value": info["msps"], "unit": UNIT, "samples": info["samples"], "what": what}
            except Exception as e:  # a baseline, never fatal for the bench line
                out[key] = {"value": None, "error": str(e)[:200]}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = workload(args.workload)
    import numpy as np
    from oracle import oracle as orc

    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the
    # reference arm on one core at N > 1
    orc.set_num_threads(len(os.sched_getaffinity(0)))
    spec = orc.ChainSpec(wl["fs"], wl["f"], wl["t1"], wl["d1"], wl["mod"], wl["gain"], wl["t2"], wl["d2"])
    # bounded sample per step so that steps+warmup finish within minutes on any host
    rng = np.random.default_rng(0x5D120001)
    probe = 1 << 20
    x = rng.integers(-100, 101, size=2 * probe, dtype=np.int8)
    orc.chain(spec, x)
    t0 = time.perf_counter()
    orc.chain(spec, x)
    rate = probe / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)
    log2 = max(20, min(26, int(np.floor(np.log2(max(rate * min(budget, 10.0), 1.0))))))
    n = 1 << log2
    x = rng.integers(-100, 101, size=2 * n, dtype=np.int8)
    for _ in range(args.warmup):
        orc.chain(spec, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.chain(spec, x)
    dt = time.perf_counter() - t0
    msps = n * args.steps / dt / 1e6
    sample = f"each step = 2^{log2} samples of the workload (bounded sample), fp64 oracle port on host cores, OpenMP"
    line = {
        "impl": "reference", "metric": METRIC, "value": msps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, wl, {"note": "the reference has no CPU DSP path and its kernels (gsdr) are absent; this arm is the "
                                                  "repo's CPU restatement (oracle port), not the reference's CUDA pipeline"}),
        "cpu_baseline": {"value": msps, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": msps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import cuda_sdr_b200 as sdr
    from cuda_sdr_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; stdout must carry ONE JSON line, so fd 1
        # points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            probe = torch.zeros(1, device=dev)
            dist.all_reduce(probe)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    wl = workload(args.workload)
    n = 1 << args.log2_block
    chain = sdr.Chain(wl["fs"], wl["f"], wl["t1"], wl["d1"], wl["mod"], fm_gain=wl["gain"], audio_taps=wl["t2"], audio_decim=wl["d2"],
                      device=local_rank)
    n_rf, n_demod, n_audio = chain.counts(n)
    x = sdr.synth.device_int8_iq(n, dev, seed=0x5D120001 + rank)  # this rank's time segment of the stream
    demod = torch.empty(n_demod, dtype=torch.float32, device=dev)
    # Audio of GATHER_EVERY consecutive steps is collected in one of two slabs; a full slab is gathered to rank 0 with ONE
    # batched send/recv on a side stream while the next slab fills (the exchange is ~1 % of the input bytes: what it costs
    # is host-side enqueue time per call, hence the batching)
    GATHER_EVERY = int(os.environ.get("BENCH_GATHER_EVERY", "32"))  # measured at N = 2: 8 -> 0.1065, 32 -> 0.1043, 64 -> 0.1042 ms/step
    slabs = [torch.empty(GATHER_EVERY, n_audio, dtype=torch.float32, device=dev) for _ in range(2)]
    first_index = rank * n  # absolute sample index of the segment (mixer phase)
    gathered = [torch.empty(world, GATHER_EVERY, n_audio, dtype=torch.float32, device=dev) for _ in range(2)] if (world > 1 and rank == 0) else None
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    slab_full = [torch.cuda.Event() for _ in range(2)]
    slab_drained = [torch.cuda.Event() for _ in range(2)]
    step_no = [0]

    def gather(slab, count):
        """Decimated audio of every rank -> rank 0 (the only exchange on the path), on the side stream."""
        if world == 1:
            return
        slab_full[slab].record()
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(slab_full[slab])
            if rank == 0:
                gathered[slab][0, :count].copy_(slabs[slab][:count], non_blocking=True)
                ops = [dist.P2POp(dist.irecv, gathered[slab][r, :count], r) for r in range(1, world)]
            else:
                ops = [dist.P2POp(dist.isend, slabs[slab][:count], 0)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            slab_drained[slab].record()

    k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    fused = chain.fused

    def step(i=None, last=False):
        k = step_no[0]
        step_no[0] += 1
        slab, slot = (k // GATHER_EVERY) & 1, k % GATHER_EVERY
        if world > 1 and slot == 0:
            torch.cuda.current_stream().wait_event(slab_drained[slab])  # this slab's previous gather has drained
        if i is not None and not fused:
            k1_events[i][0].record()
        got = chain.process_device(x, first_index, out=slabs[slab][slot], scratch=demod)  # fused: ONE kernel; else K1 + K2
        if i is not None and not fused:
            k1_events[i][1].record()
        assert got.numel() == n_audio
        if slot == GATHER_EVERY - 1 or last:
            gather(slab, slot + 1)
            step_no[0] += GATHER_EVERY - 1 - slot  # a partial slab at the end of a phase: start the next phase on a fresh slab

    def barrier():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm_stream)
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: first the local kernel alone until the clocks are up (time-based, so NO communication in it -- ranks would
    # run different counts), then max(W, 3) complete steps including the gather, in lockstep on every rank
    warm_steps = 0
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < args.warmup_seconds:
        chain.process_device(x, first_index, out=slabs[0][0], scratch=demod)
        warm_steps += 1
        if warm_steps % 16 == 0:
            torch.cuda.synchronize()
    n_warm = max(args.warmup, 3)
    for w in range(n_warm):
        step(last=(w == n_warm - 1))
        warm_steps += 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sdr._native.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        step(i, last=(i == args.steps - 1))
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm_stream)  # the last gathers are part of the timed region
    t_end.record()
    barrier()
    launches = sdr._native.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    # fused route: ONE kernel per step and nothing else on the stream, so the kernel's average duration is the step time
    # (launch gaps included: an upper bound); events between the launches would serialise what programmatic dependent
    # launch overlaps.  Two-kernel route: events around K1 + K2 of every step.
    k1_ms = total_ms / args.steps if fused else statistics.mean(a.elapsed_time(b) for a, b in k1_events)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- end to end: host (pinned) input -> C-ABI host call -> host output, H2D/D2H inside the timed region
    e2e = None
    if not args.skip_e2e:
        xh = torch.empty(2 * n, dtype=torch.int8).pin_memory()
        xh.copy_(x)
        outh = torch.empty(n_audio, dtype=torch.float32).pin_memory()
        e2e_steps = max(2, min(args.steps, 5))
        chain.process_host(xh, first_index, out=outh)  # warm-up: allocates the staging slots
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = chain.process_host(xh, first_index, out=outh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert got.numel() == n_audio
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n * e2e_steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n_audio,
               "steps": e2e_steps, "api": "b200sdr_chain_process_host (pinned host buffers; double-buffered H2D/K1/K2/D2H)"}
        del xh
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if fused:  # one kernel: 2 B/sample of int8 IQ in, 4 B per audio sample out, nothing in between
            alg_bytes = 2.0 * n + 4.0 * n_audio
            kernel_name = ("toepKernel" if chain.variant.startswith("toeplitz<") else "chainKernel") + \
                " (convert+mix+FIR+decimate+demod+audio FIR, persistent, one launch per step)"
        else:      # K1 + K2: the demodulated stream makes one round trip through HBM
            alg_bytes = 2.0 * n + 8.0 * n_demod + 4.0 * n_audio
            kernel_name = "rowsKernel + directKernel (two launches per step)"
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        traffic = None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel variant (profiles/)
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if chain.variant.startswith(t["variant_prefix"]) and t["samples_per_launch"] == n:
                traffic = t["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warm_steps, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl, {"kernel_variant": chain.variant}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": kernel_name, "kernel_ms": k1_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.skip_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
            v, cores, sample, _, _ = cpu_chain_rate(wl, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        if world == 1 and not args.skip_e2e:
            line.update(reference_api_rates(wl))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------
# C5: wideband channelizer, channels sharded over the GPUs (strong scaling: the same input on every GPU)
# ---------------------------------------------------------------------------------------------------
def run_channelizer(args):
    import torch
    import torch.distributed as dist

    import cuda_sdr_b200 as sdr
    from cuda_sdr_b200 import sharding, taps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    fs, T1, D1, T2, D2 = 153.6e6, 4097, 640, 273, 5            # 153.6 Msps = 48 kHz x 640 x 5
    total = args.channels
    log2n = args.log2_block if args.log2_block != LOG2_BLOCK else 29  # 3.5 s of signal per step: long segments amortise the per-launch costs at N = 8
    n = 1 << log2n
    freqs = [(c - total / 2) * 600e3 + 100e3 for c in range(total)]  # 600 kHz raster
    mods = [c & 1 for c in range(total)]                             # alternating AM / FM
    t1 = taps.lowpass(T1, 100e3, fs)
    t2 = taps.lowpass(T2, 0.45 * 48e3, fs / D1)
    gain = sdr.fm_gain(fs / D1, 75e3)
    x = sdr.synth.device_int8_iq(n, dev, seed=0x5D120005)           # identical on every rank: stands in for a broadcast feed
    # Two decompositions, no collective on the filter path either way:
    #   filter-bank route (all channels on one raster: ONE pass yields every channel) -> overlapped TIME segments, every GPU
    #     produces all channels for its share of the audio outputs;
    #   per-channel route -> channel c on rank c mod G, input replicated.
    probe = sdr.Channelizer(fs, freqs, mods, t1, D1, t2, D2, fm_gains=[gain] * total, device=local_rank)
    by_time = probe.variant.startswith("pfb<") and args.shard != "channels"
    if by_time:
        mine = list(range(total))
        ch = probe
        _, n_audio_total = ch.counts(n)
        window = sharding.chain_window(T1, D1, T2, any(mods), True)
        seg = sharding.time_segment(n_audio_total, world, rank, D1 * D2, window)
        x_mine = x[2 * seg.first_input: 2 * (seg.first_input + seg.input_count)]
        n_audio = seg.output_count
        counts_audio = [sharding.time_segment(n_audio_total, world, r, D1 * D2, window).output_count for r in range(world)]
        shapes = [(total, counts_audio[r]) for r in range(world)]
    else:
        mine = sharding.channels_of_rank(total, world, rank)
        ch = sdr.Channelizer(fs, [freqs[c] for c in mine], [mods[c] for c in mine], t1, D1, t2, D2, fm_gains=[gain] * len(mine), device=local_rank)
        del probe
        x_mine = x
        _, n_audio = ch.counts(n)
        shapes = [(len(sharding.channels_of_rank(total, world, r)), n_audio) for r in range(world)]
    n_demod = (n_audio - 1) * D2 + T2
    scratch = torch.empty(len(mine), n_demod, dtype=torch.float32, device=dev)
    outs = [torch.empty(len(mine), n_audio, dtype=torch.float32, device=dev) for _ in range(2)]
    gathered = [[torch.empty(*shapes[r], dtype=torch.float32, device=dev) for r in range(world)] for _ in range(2)] \
        if (world > 1 and rank == 0) else None
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    done_evt = [torch.cuda.Event() for _ in range(2)]
    drained = [torch.cuda.Event() for _ in range(2)]
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    step_no = [0]

    def step(i=None):
        buf = step_no[0] & 1
        step_no[0] += 1
        if world > 1:
            torch.cuda.current_stream().wait_event(drained[buf])
        if i is not None:
            k_events[i][0].record()
        ch.run(x_mine, n_audio, out=outs[buf], scratch=scratch)
        if i is not None:
            k_events[i][1].record()
        if world > 1:  # gather this step's audio of every rank's channels to rank 0 on the side stream
            done_evt[buf].record()
            with torch.cuda.stream(comm):
                comm.wait_event(done_evt[buf])
                if rank == 0:
                    gathered[buf][0].copy_(outs[buf], non_blocking=True)
                    ops = [dist.P2POp(dist.irecv, gathered[buf][r], r) for r in range(1, world)]
                else:
                    ops = [dist.P2POp(dist.isend, outs[buf], 0)]
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                drained[buf].record()

    def barrier():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)
            dist.barrier()
        torch.cuda.synchronize()

    t_warm = time.perf_counter()
    warm = 0
    while time.perf_counter() - t_warm < args.warmup_seconds:
        ch.run(x_mine, n_audio, out=outs[0], scratch=scratch)
        warm += 1
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step()
        warm += 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sdr._native.launch_count()
    t0, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for i in range(args.steps):
        step(i)
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm)
    t1e.record()
    barrier()
    launches = sdr._native.launch_count() - launches0
    total_ms = t0.elapsed_time(t1e)
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in k_events)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(peaks_path))["bf16_tflops"]) if os.path.exists(peaks_path) else 1590.0
        M = -(-T1 // D1)
        pfb = ch.variant.startswith("pfb<")
        samples_this_gpu = x_mine.numel() // 2
        # the per-channel chain costs ~38 flop per input sample per channel; the filter bank does the same job
        # for all channels at once, so the per-channel figure stays the ALGORITHMIC work the line is normalised by
        flops = samples_this_gpu * len(mine) * (12.0 + 4.0 * T1 / D1 + 10.0 / D1 + 2.0 * T2 / (D1 * D2))
        hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
        if pfb:
            # bytes the two launches must move: int8 IQ in, demodulated samples out and back in, audio out
            alg_bytes = 2.0 * samples_this_gpu + len(mine) * (8.0 * n_demod + 4.0 * n_audio)
            roofline = {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                        "kernel": "pfbKernel (polyphase filter bank + fp64 FFT + demod, all channels) + batched audio FIR (windowKernel)",
                        "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                        "per_channel_equivalent_tflops": flops / (k_ms * 1e-3) / 1e12,
                        "note": "the filter bank replaces ~38 flop per sample PER CHANNEL by ~60 flop per sample for all channels, so the "
                                "binding roof is HBM (2 B per sample in + the demodulated streams); today the kernel is bound by FP64 "
                                "latency at 8 warps per SM (DESIGN.md)"}
        else:
            executed_ops = 2.0 * (n_demod + M) * (2 * D1) * (16 if M > 4 else 8) * 3 * len(mine)      # int8 MMA ops incl. digits and padding
            roofline = {"bound": "tensor", "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": flops / (k_ms * 1e-3) / 1e12 / peak, "traffic": None,
                        "kernel": "channelKernel (int8 GEMM RF stage + demod) + batched audio FIR", "kernel_ms": k_ms,
                        "algorithmic_flops_per_launch": flops, "executed_int8_tops": executed_ops / (k_ms * 1e-3) / 1e12,
                        "legacy_imma_peak_tops_measured": 1143.0,
                        "note": "algorithmic flops (SURVEY 8(d): ~38 per sample per channel) against the measured bf16 GEMM peak; the kernel "
                                "executes the contraction as 3 int8 digit MMAs on the legacy IMMA path (tools/imma_bench.cu: 1143 TOP/s on this part)"}
        line = {
            "metric": "input Msps through the 256-channel wideband channelizer (int8 -> mix -> FIR -> AM/FM demod -> audio FIR per channel)",
            "value": n * args.steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 filter bank + FFT, f32 demod / audio FIR" if pfb else "s8 x s8 -> s32 (24-bit fixed-point taps), f32 epilogue",
            "data": "synthetic",
            "config": {"workload": f"C5 wideband channelizer: 2^{log2n} int8 IQ samples per step at 153.6 Msps-class rate, {total} channels on a 600 kHz "
                                   f"raster alternating AM/FM, per channel mix -> {T1}-tap FIR /{D1} -> demod -> {T2}-tap audio FIR /{D2}",
                       "channels_total": total, "channels_this_gpu": len(mine), "samples_per_step": n, "samples_this_gpu": samples_this_gpu,
                       "parallelism": ("overlapped time segments of the wideband stream, every GPU produces all channels for its share of the audio outputs"
                                       if by_time else "channels interleaved over the GPUs (c mod G), input replicated") +
                                      "; NCCL send/recv gather of the audio to rank 0 on a side stream",
                       "l2": f"input block {2 * n >> 20} MiB exceeds the 126 MB L2", "kernel_variant": ch.variant},
            "roofline": roofline,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "channelizer":
        return run_channelizer(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
