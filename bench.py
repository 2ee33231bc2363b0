#!/usr/bin/env python3
"""Headline benchmark: input Msps through the int8 -> mix -> FIR -> demod -> audio-FIR chain.

  python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the fp64 oracle port on host cores
  python bench.py --workload wbfm|channelizer|firsweep     # BASELINE configs C3 / C5 / C4 as their own lines

Default workload = BASELINE.json configs[1] (C2, AM broadcast chain): per GPU and per step one block of 2^28 synthetic
int8 IQ samples (512 MiB, larger than L2, so no L2 flush is needed between iterations) -> mix by -1.234 MHz -> 101-tap
low-pass, decimate by 40 -> |.| -> 129-tap low-pass, decimate by 10 -> 48 kHz-class audio.  At N > 1 every rank processes
its own time segment of one long stream (no collective on the filter path) and the decimated audio is gathered to rank 0
inside the timed region by the library's b200sdr_gather (NCCL send/recv on a side stream, include/b200sdr/b200sdr.h);
per-GPU work is fixed, so scaling is "weak".  Every line also carries a `channelizer` sub-record: BASELINE configs[4]
(C5, 256 channels) on the same N GPUs, strong scaling, with the same workload timed on ONE GPU in the same run.

torch / torch.distributed are plumbing here (device memory, rendezvous of the NCCL id, barriers); the timed work is this
repo's kernels behind the C-ABI.  One JSON line is printed by rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import shutil
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
                    help="also keep running untimed warm-up steps until this much wall time has passed (a step is ~0.1 ms; "
                         "the SM clock needs far longer than 3 steps to leave its idle state); reported as warmup_extra_steps")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--log2-block", type=int, default=LOG2_BLOCK, help="log2 of input samples per GPU per step")
    ap.add_argument("--workload", choices=["am", "wbfm", "channelizer", "firsweep"], default="am")
    ap.add_argument("--channels", type=int, default=256, help="channelizer workload: total channels")
    ap.add_argument("--shard", choices=["auto", "channels", "time"], default="auto",
                    help="channelizer workload, N > 1: shard by channel or by time segment (auto: time on the filter-bank route)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-channelizer", action="store_true", help="leave the C5 sub-record out of an am/wbfm line")
    ap.add_argument("--no-balance", action="store_true",
                    help="N > 1: equal time segments on every GPU instead of segments in proportion to each GPU's measured rate")
    ap.add_argument("--skip-ncu", action="store_true", help="do not measure roofline.traffic with ncu (use the committed capture)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def load_taps_module():
    """cuda_sdr_b200/taps.py by path: pure numpy, and importing it this way does not load libb200sdr.so (the reference arm
    must not map the product library)."""
    spec = importlib.util.spec_from_file_location("_b200sdr_taps", os.path.join(ROOT, "cuda_sdr_b200", "taps.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fm_gain_f32(input_sample_rate: float, deviation: float) -> float:
    """QuadDemodFactory.h:108-110 in float32 (same expression as cuda_sdr_b200.fm_gain, restated so that the reference arm
    needs no product import)."""
    import math
    import numpy as np
    f = np.float32
    return float(f(input_sample_rate) / (f(2.0) * f(math.pi) * f(deviation) * f(5)))


def workload(name: str):
    taps = load_taps_module()
    if name == "am":
        return dict(name="C2 AM broadcast chain", fs=FS, f=F_SHIFT, t1=taps.lowpass(T1, 0.45 * FS / D1, FS), d1=D1, mod=0, gain=1.0,
                    t2=taps.lowpass(T2, 0.45 * 48e3, FS / D1), d2=D2)
    return dict(name="C3 WBFM chain", fs=FS, f=2.5e6, t1=taps.lowpass(545, 100e3, FS), d1=80, mod=1, gain=fm_gain_f32(FS / 80, 75e3),
                t2=taps.lowpass(273, 0.45 * 48e3, FS / 80), d2=5)


GATHER_TRANSPORT = {
    0: "NCCL grouped send/recv on a side stream",
    1: "peer memory: the kernels store into rank 0's slabs over NVLink (CUDA IPC), stream-ordered 32-bit flags",
    2: "peer memory: a copy engine moves each rank's local slab into rank 0's slabs over NVLink (CUDA IPC) on a side stream, "
       "stream-ordered 32-bit flags",
}


def gather_slab_schedule(nsteps, longest):
    """Steps per slab over a run of nsteps steps: `longest` in the steady state, halving towards the end so that one step of audio
    is what remains to be moved when the last kernel ends."""
    sizes, left = [], nsteps
    while left > 0:
        sizes.append(min(longest, max(1, left // 2)))
        left -= sizes[-1]
    return sizes


def config_dict(wl, log2_samples, extra=None):
    """`workload` names the chain (identical in both arms); the block one step processes is a separate key, because the CPU arm
    runs a bounded sample of the GPU arm's block (the metric is a rate)."""
    n = 1 << log2_samples
    cfg = {
        "workload": f"{wl['name']}: synthetic int8 IQ at a {wl['fs'] / 1e6:.1f} Msps-class rate "
                    f"-> mix -> {len(wl['t1'])}-tap FIR decimate-by-{wl['d1']} -> {'AM' if wl['mod'] == 0 else 'FM'} demod -> "
                    f"{len(wl['t2'])}-tap audio FIR decimate-by-{wl['d2']}",
        "samples_per_gpu_per_step": n,
        "rf_taps": len(wl["t1"]), "rf_decimation": wl["d1"], "audio_taps": len(wl["t2"]), "audio_decimation": wl["d2"],
        "modulation": "am" if wl["mod"] == 0 else "fm",
        "phase_mode": "exact (64-bit fixed-point turns)",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region), one sampler per rank
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
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
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


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (fp64, OpenMP) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_chain_rate(wl, target_seconds: float, max_log2: int = 26):
    """Returns (Msps, cores, sample description) for the fp64 oracle on host cores."""
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
    return n / dt / 1e6, orc.num_threads(), f"2^{log2} samples of the same workload, one pass, fp64 oracle (oracle/oracle.c), OpenMP"


def reference_pipeline_rate(wl, log2_samples: int = 24):
    """Baseline B1: the reference's own host framework compiled in place + a plain restated gsdr, driven in
    <= 1 MiB steps through its Filter API with pinned host buffers (oracle/ref/ref_chain.cpp -> oracle/_ref/ref_chain_naive).
    A reported baseline, like cpu_baseline; {} when the binary is not there (it needs /root/reference at build time)."""
    import numpy as np

    exe = os.path.join(ROOT, "oracle", "_ref", "ref_chain_naive")
    if not os.path.exists(exe):
        return {}
    n = 1 << log2_samples
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        rng = np.random.default_rng(0x5D120001)
        rng.integers(-100, 101, size=2 * n, dtype=np.int8).tofile(os.path.join(tmp, "in.i8"))
        np.asarray(wl["t1"], dtype=np.float32).tofile(os.path.join(tmp, "t1.f32"))
        np.asarray(wl["t2"], dtype=np.float32).tofile(os.path.join(tmp, "t2.f32"))
        cmd = [exe, "--fs", repr(wl["fs"]), "--freq", repr(wl["f"]), "--mod", "am" if wl["mod"] == 0 else "fm", "--d1", str(wl["d1"]),
               "--d2", str(wl["d2"]), "--taps1", os.path.join(tmp, "t1.f32"), "--taps2", os.path.join(tmp, "t2.f32"),
               "--in", os.path.join(tmp, "in.i8"), "--repeat", "4"]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            info = json.loads(res.stdout.strip().splitlines()[-1])
            return {"reference_cuda_pipeline": {
                "value": info["msps"], "unit": UNIT, "samples": info["samples"],
                "what": "reference host framework compiled in place + restated one-thread-per-output gsdr kernels (B1), <= 1 MiB steps, "
                        "pinned host in / host out"}}
        except Exception as e:  # a baseline, never fatal for the bench line
            return {"reference_cuda_pipeline": {"value": None, "error": str(e)[:200]}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = workload("wbfm" if args.workload == "wbfm" else "am")
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
    sample = f"each step = 2^{log2} samples of the workload (a bounded sample of the 2^{args.log2_block}-sample block the GPU arm runs), " \
             "fp64 oracle port on host cores, OpenMP"
    line = {
        "impl": "reference", "metric": METRIC, "value": msps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(wl, log2, extra={
            "bounded_sample_of_samples_per_gpu_per_step": 1 << args.log2_block,
            "l2": "n/a (host arm)", "parallelism": "host cores, OpenMP",
            "kernel_variant": "fp64 CPU restatement (oracle/oracle.c): the reference has no CPU DSP path and its kernels (gsdr) are absent"}),
        "cpu_baseline": {"value": msps, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": msps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# shared plumbing of the GPU arms
# ---------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one rank: device, torch.distributed (rendezvous + barriers only), the NCCL id for b200sdr_gather."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner on stdout when the communicator comes up; stdout must carry ONE JSON line, so fd 1
            # points at stderr until the communicators exist
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.all_reduce(torch.zeros(1, device=self.dev))
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)

    def make_gather(self, floats_per_rank, slabs=3):
        """b200sdr_gather over all ranks.  Peer memory first (rank 0's slabs mapped into every rank over NVLink; a copy engine
        moves each rank's local slab into them, or with BENCH_GATHER=store the kernels store their audio straight into them; the
        IPC handles travel through torch.distributed), NCCL send/recv if that is unavailable or BENCH_GATHER=nccl (the 128-byte
        NCCL id then travels from rank 0 the same way)."""
        from cuda_sdr_b200 import sharding
        torch, dist = self.torch, self.dist
        want = os.environ.get("BENCH_GATHER", "copy")  # copy: local slab + copy engine (default); store (= peer): kernels store remotely
        if self.world > 1 and want in ("copy", "store", "peer"):
            g, ok = None, 1
            try:
                g = sharding.Gather(self.rank, self.world, floats_per_rank, slabs=slabs, device=self.local_rank,
                                    mode=sharding.Gather.PEER_COPY if want == "copy" else sharding.Gather.PEER)
                blob = g.export_blob()
            except Exception as e:
                ok, blob = 0, b""
                print(f"rank {self.rank}: peer-mode gather unavailable ({e})", file=sys.stderr)
            blobs = self.gather_objects(blob)
            if ok and all(len(b) == len(blob) and len(b) > 0 for b in blobs):
                try:
                    g.import_blobs(b"".join(blobs))
                except Exception as e:
                    ok = 0
                    print(f"rank {self.rank}: peer-mode import failed ({e})", file=sys.stderr)
            else:
                ok = 0
            if min(self.gather_objects(ok)) == 1:
                return g
            if g is not None:
                g.close()
        uid = None
        if self.world > 1:
            t = torch.zeros(128, dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                t.copy_(torch.frombuffer(bytearray(sharding.Gather.unique_id()), dtype=torch.uint8))
            dist.broadcast(t, 0)
            uid = bytes(t.cpu().numpy().tobytes())
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            return sharding.Gather(self.rank, self.world, floats_per_rank, slabs=slabs, device=self.local_rank, unique_id=uid)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure_traffic_with_ncu(wl, log2_block, variant):
    """DRAM bytes of ONE launch of the chain kernel (dram__bytes_read.sum + dram__bytes_write.sum) from a live ncu pass over a
    one-step child process.  None when ncu is missing or fails; the caller then falls back to the committed capture."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    child = (
        "import sys, torch; sys.path.insert(0, %r); import bench, cuda_sdr_b200 as sdr\n"
        "wl = bench.workload(%r); n = 1 << %d; dev = torch.device('cuda', 0)\n"
        "c = sdr.Chain(wl['fs'], wl['f'], wl['t1'], wl['d1'], wl['mod'], fm_gain=wl['gain'], audio_taps=wl['t2'], audio_decim=wl['d2'])\n"
        "x = sdr.synth.device_int8_iq(n, dev); c.process_device(x); c.process_device(x); torch.cuda.synchronize()\n"
    ) % (ROOT, "am" if wl["mod"] == 0 else "wbfm", log2_block)
    kernel = "toepKernel" if variant.startswith("toeplitz<") else "chainKernel"
    try:
        res = subprocess.run([ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{kernel}",
                              "-s", "1", "-c", "1", "--csv", sys.executable, "-c", child], capture_output=True, text=True, timeout=240)
        total, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        found = 0
        for line in res.stdout.splitlines():
            cells = [c.strip('"') for c in line.split('","')]
            if len(cells) >= 3 and ("dram__bytes_read.sum" in cells or "dram__bytes_write.sum" in cells):
                value, unit = float(cells[-1].replace(",", "")), cells[-2]
                total += value * unit_scale.get(unit, 1.0)
                found += 1
        return total if found == 2 else None
    except Exception:
        return None


def e2e_filter_api(ctx, wl):
    """The headline end-to-end number: build/bin/filter_api_bench (tools/filter_api_bench.cpp) drives host -> CudaMemcpyFilter ->
    ONE fused Filter -> CudaMemcpyFilter -> host through the reference-facing IFactories / Filter API, one process per GPU."""
    import numpy as np

    exe = os.path.join(ROOT, "build", "bin", "filter_api_bench")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        np.asarray(wl["t1"], dtype=np.float32).tofile(os.path.join(tmp, "t1.f32"))
        np.asarray(wl["t2"], dtype=np.float32).tofile(os.path.join(tmp, "t2.f32"))
        threads = max(1, len(os.sched_getaffinity(0)) // max(1, ctx.world))
        if ctx.world > 1:
            threads = max(threads, 4)  # torchrun pins nothing; leave the producer a few threads per rank
        cmd = [exe, "--fs", repr(wl["fs"]), "--freq", repr(wl["f"]), "--mod", "am" if wl["mod"] == 0 else "fm", "--d1", str(wl["d1"]),
               "--d2", str(wl["d2"]), "--taps1", os.path.join(tmp, "t1.f32"), "--taps2", os.path.join(tmp, "t2.f32"), "--device", str(ctx.local_rank),
               "--samples-per-pass", str(1 << 28), "--passes", "6", "--step", str(64 << 20), "--warmup-steps", "4", "--pipeline", "1",
               "--threads", str(threads)]
        def run(extra):
            ctx.barrier()
            try:
                res = subprocess.run(cmd + extra, capture_output=True, text=True, timeout=300)
                return json.loads(res.stdout.strip().splitlines()[-1])
            except Exception as e:
                return {"error": str(e)[:200]}

        # Headline: the samples are IN the pinned block requestBuffer() hands out when the step starts (a capture device that DMAs
        # there; what the C-ABI leg does with its pinned input) -- host -> device copy, kernel, device -> host copy and the host's
        # read of the audio all inside the timed region.  Beside it: the same loop with a pool of host threads copying every block
        # from a synthetic capture first (bound by that copy: ~50 GB/s of host memcpy next to the DMA; at N > 1 the producers of
        # all ranks share the host's cores and memory bandwidth).
        info = run(["--producer", "resident"])
        copied = run(["--producer", "copy"])
    infos = ctx.gather_objects(info)
    copies = ctx.gather_objects(copied)
    if ctx.rank != 0:
        return None
    if any("error" in i for i in infos):
        return {"value": None, "unit": UNIT, "error": next(i["error"] for i in infos if "error" in i)}

    def total(rs):
        return sum(r["timed_samples"] for r in rs) / max(r["seconds"] for r in rs) / 1e6

    steps = infos[0]["timed_steps"]
    out = {"value": total(infos), "unit": UNIT, "h2d_bytes_per_step": infos[0]["h2d_bytes"] // max(1, steps),
           "d2h_bytes_per_step": infos[0]["d2h_bytes"] // max(1, steps), "steps": steps,
           "api": "IFactories fused Filter: pinned block from requestBuffer() -> CudaMemcpyFilter (H2D) -> gsCreateFusedChain Filter -> "
                  "CudaMemcpyFilter (D2H) -> pinned host buffers behind IEventPipeline, the host reads the audio; 64 MiB steps "
                  "(tools/filter_api_bench.cpp), one process per GPU",
           "input": "synthetic int8 IQ resident in the pinned blocks when a step starts (capture-device model)",
           "per_rank_msps": [i["msps"] for i in infos], "host_ms_per_step": infos[0].get("host_ms_per_step")}
    if all("error" not in c for c in copies):
        out["with_producer_copy"] = {"value": total(copies), "unit": UNIT, "producer_threads_per_gpu": copies[0]["threads"],
                                     "per_rank_msps": [c["msps"] for c in copies], "host_ms_per_step": copies[0].get("host_ms_per_step"),
                                     "what": "the same loop, a pool of host threads first copies each 64 MiB block of a synthetic capture "
                                             "into the pinned block (producer-bound)"}
    return out


# ---------------------------------------------------------------------------------------------------
# C2 / C3: the chain, weak scaling over time segments
# ---------------------------------------------------------------------------------------------------
def run_chain(args, ctx):
    import cuda_sdr_b200 as sdr

    torch = ctx.torch
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    wl = workload(args.workload)
    n = 1 << args.log2_block
    chain = sdr.Chain(wl["fs"], wl["f"], wl["t1"], wl["d1"], wl["mod"], fm_gain=wl["gain"], audio_taps=wl["t2"], audio_decim=wl["d2"],
                      device=ctx.local_rank)
    n_rf, n_demod, n_audio = chain.counts(n)
    balance = world > 1 and not args.no_balance
    n_alloc = n + (n >> 3) if balance else n  # balanced segments: the faster GPUs take up to 12 % more than the average
    x_all = sdr.synth.device_int8_iq(n_alloc, dev, seed=0x5D120001 + rank)  # this rank's time segment of the stream
    x = x_all[: 2 * n]
    demod = torch.empty(n_demod + (n_demod >> 3), dtype=torch.float32, device=dev)
    first_index = rank * n  # absolute sample index of the segment (mixer phase)
    # Audio of up to `ge` consecutive steps is collected in one of three slabs; a submitted slab travels to rank 0 on the library's
    # side stream while the next slabs fill.  Slabs are long in the steady state (every submit puts an event between two launches
    # and so costs that pair its programmatic overlap) and shrink towards the end of a run (slab_schedule), so that only ONE step of
    # audio is still to be moved when the last kernel ends -- a streaming application that knows its end flushes the same way.
    ge = int(os.environ.get("BENCH_GATHER_EVERY", "0")) or max(1, min(32, args.steps // 4))
    slabs = 3
    step_no = [0]
    fused = chain.fused
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # the clock sampler polls every 50 ms and the timed region of a 20-step run lasts ~2 ms: it runs from the warm-up (the same
    # kernel back to back, the same load) through the timed region, and reports the samples taken under load
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    # warm-up, part 1: the local kernel alone until the clocks are up (time-based, so NO communication in it -- ranks would run
    # different counts)
    scratch_out = torch.empty(n_audio + (n_audio >> 3), dtype=torch.float32, device=dev)
    extra = 0
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < args.warmup_seconds or extra < 16:
        chain.process_device(x, first_index, out=scratch_out, scratch=demod)
        extra += 1
        if extra % 16 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    # Partition of the stream over the ranks.  The total work of a step is world * 2^log2_block samples whatever the split; with
    # --no-balance every rank takes exactly its 2^log2_block, otherwise b200sdr_chain_segment_weighted cuts the world-wide run of
    # audio outputs in proportion to each GPU's measured rate (power-capped GPUs of one box differ by a few per cent, rank 0 also
    # absorbs the other ranks' audio, and a step ends with the slowest), and every rank reads the input segment its outputs need.
    part = {"audio": n_audio, "first": first_index, "in": n, "counts": [n_audio] * world, "x": x, "views": None}
    cap_audio = n_audio + (n_audio >> 3)
    gather = ctx.make_gather([ge * cap_audio] * world, slabs) if world > 1 else None

    def set_partition(segs):
        if segs is not None:
            part["counts"] = [sg[1] for sg in segs]
            _, part["audio"], part["first"], part["in"] = segs[rank]
            assert part["in"] <= n_alloc and part["audio"] <= cap_audio
            part["x"] = x_all[: 2 * part["in"]]
        a = part["audio"]
        if gather is not None:
            part["views"] = [gather.slab(s)[: ge * a].view(ge, a) for s in range(slabs)]
        else:
            part["views"] = [torch.empty(ge, a, dtype=torch.float32, device=dev) for _ in range(slabs)]

    set_partition(None)

    def slab_schedule(nsteps):
        return gather_slab_schedule(nsteps, ge)

    def run_steps(nsteps, events=None, timed=False):
        """nsteps steps of the chain, slab by slab; events: one (start, end) pair per step around the kernel(s)."""
        i = 0
        for size in slab_schedule(nsteps):
            slab = step_no[0] % slabs
            step_no[0] += 1
            if gather is not None:
                gather.acquire(slab)  # this slab's previous gather has drained
            for slot in range(size):
                ev = events[i] if events is not None else (k_events[i] if timed and not fused else None)
                if ev is not None:
                    ev[0].record()
                # fused: ONE kernel; else K1 + K2.  Exactly part["audio"] outputs of the segment that starts at sample part["first"].
                got = chain.run(part["x"], part["audio"], part["first"], out=part["views"][slab][slot], scratch=demod, n_in=part["in"])
                if ev is not None:
                    ev[1].record()
                i += 1
            assert got.numel() == part["audio"]
            if gather is not None:
                gather.submit(slab, [size * c for c in part["counts"]])

    def quiesce():
        if gather is not None:
            gather.finish()
        ctx.barrier()

    # warm-up, part 1b (N > 1, balanced): the rate of every GPU UNDER THE CONDITIONS OF THE TIMED REGION -- rehearsals of exactly
    # what is timed below (all ranks start together after a barrier, the same number of steps, the gather active, rank 0
    # absorbing the others' audio), each rank timing its own kernels with one event pair around the whole run (median of five).  The GPUs of a box
    # differ by a few per cent and power capping moves them further apart under load; a rate measured on each GPU alone, or over
    # a few kernels, misjudges exactly the ranks that matter.  Two rounds: equal segments first, then a damped correction.
    shares = rates = None
    if balance:
        total_audio = chain.counts(world * n)[2]
        weights = [1.0] * world
        cal_a, cal_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rnd in range(2):
            times = []
            for _ in range(5):
                quiesce()
                cal_a.record()
                run_steps(args.steps)
                cal_b.record()
                quiesce()
                times.append(cal_a.elapsed_time(cal_b))
                extra += args.steps
            mine = statistics.median(times)  # ms for this rank's share of args.steps steps (a GPU's rate moves between windows)
            all_ms = [t if t > 0 else 1.0 for t in ctx.gather_objects(mine)]
            mean_ms = sum(all_ms) / world
            damp = 1.0 if rnd == 0 else 0.7
            weights = [w * (mean_ms / t) ** damp for w, t in zip(weights, all_ms)]
            norm = sum(weights) / world
            weights = [min(max(w / norm, 0.90), 1.10) for w in weights]  # stay inside the allocation whatever a noisy measurement says
            segs = [chain.segment_weighted(total_audio, weights, r) for r in range(world)]
            set_partition(segs)
            rates = [args.steps / t for t in all_ms]  # steps per ms of the last rehearsal round
        shares = [sg[3] / n for sg in segs]
    my_in, my_first_in, my_audio = part["in"], part["first"], part["audio"]
    # warm-up, part 2: max(W, 3) complete steps including the gather, in lockstep on every rank
    n_warm = max(args.warmup, 3)
    run_steps(n_warm)
    quiesce()
    launches0 = sdr._native.launch_count()
    t_start, t_kernels, t_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    quiesce()
    t_start.record()
    run_steps(args.steps, timed=True)
    t_kernels.record()  # the last kernel of this rank
    if gather is not None:
        gather.finish()  # the gathers are part of the timed region
    t_end.record()
    ctx.barrier()
    launches = sdr._native.launch_count() - launches0
    my_ms, my_kernel_ms = t_start.elapsed_time(t_end), t_start.elapsed_time(t_kernels)
    clocks = sampler.stop()
    # fused route: ONE kernel per step and nothing else on the stream, so the kernel's average duration is the step time (launch
    # gaps included: an upper bound); events between the launches would serialise what programmatic dependent launch overlaps.
    # Two-kernel route: events around K1 + K2 of every step.
    k_ms = my_kernel_ms / args.steps if fused else statistics.mean(a.elapsed_time(b) for a, b in k_events)
    total_ms = ctx.max_over_ranks(my_ms)
    per_rank = ctx.gather_objects({"rank": rank, "ms_per_step": my_ms / args.steps, "kernel_ms_per_step": my_kernel_ms / args.steps,
                                   "comm_exposed_ms": my_ms - my_kernel_ms, "clocks": clocks})
    gstats = gather.stats() if gather is not None else None
    gather_mode = gather.mode if gather is not None else 0

    # ---- end to end ------------------------------------------------------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        # (1) the C-ABI host call: pinned host buffers, H2D / kernel / D2H double-buffered inside b200sdr_chain_process_host
        xh = torch.empty(2 * n, dtype=torch.int8).pin_memory()
        xh.copy_(x_all[: 2 * n])
        outh = torch.empty(n_audio, dtype=torch.float32).pin_memory()
        e2e_steps = max(2, min(args.steps, 5))
        chain.process_host(xh, first_index, out=outh)  # warm-up: allocates the staging slots
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = chain.process_host(xh, first_index, out=outh)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        assert got.numel() == n_audio
        c_abi = {"value": world * n * e2e_steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n_audio,
                 "steps": e2e_steps, "api": "b200sdr_chain_process_host (C-ABI; pinned host buffers; double-buffered H2D / kernel / D2H)"}
        del xh
        # (2) the reference-facing plugin API (the headline): one fused Filter between two CudaMemcpyFilters
        api = e2e_filter_api(ctx, wl)
        if rank == 0:
            if api is not None and api.get("value"):
                e2e = dict(api)
                e2e["c_abi"] = c_abi
            else:
                e2e = dict(c_abi)
                e2e["filter_api"] = api

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        if fused:  # one kernel: 2 B/sample of int8 IQ in, 4 B per audio sample out, nothing in between
            alg_bytes = 2.0 * my_in + 4.0 * my_audio  # this rank's segment (rank 0's share when the partition is balanced)
            kernel_name = ("toepKernel" if chain.variant.startswith("toeplitz<") else "chainKernel") + \
                " (convert+mix+FIR+decimate+demod+audio FIR, persistent, one launch per step)"
        else:      # K1 + K2: the demodulated stream makes one round trip through HBM
            alg_bytes = 2.0 * my_in + 8.0 * ((my_audio - 1) * wl["d2"] + len(wl["t2"])) + 4.0 * my_audio
            kernel_name = "rowsKernel + directKernel (two launches per step)"
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        if fused and world == 1 and not args.skip_ncu:
            traffic = measure_traffic_with_ncu(wl, args.log2_block, chain.variant)
            traffic_src = "measured by this run: ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch (child process)" if traffic else None
        if traffic is None:
            try:  # the committed ncu --set full capture of this kernel variant (profiles/)
                t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                for entry in t if isinstance(t, list) else [t]:
                    if chain.variant.startswith(entry["variant_prefix"]) and entry["samples_per_launch"] == n:
                        traffic, traffic_src = entry["dram_bytes_per_launch"], "committed capture: " + entry["source"]
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "warmup_extra_steps": extra, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, args.log2_block, extra={
                "kernel_variant": chain.variant,
                "l2": f"input block {2 * n >> 20} MiB per step exceeds the 126 MB L2; no flush between iterations",
                "parallelism": ("overlapped time segments, one per GPU" + (", lengths balanced by measured GPU rate" if balance else "") +
                                f", no collective on the filter path; b200sdr_gather (libb200sdr.so): up to {ge} steps of audio per slab "
                                f"to rank 0, 3 slabs; {GATHER_TRANSPORT[gather_mode]}") if world > 1 else
                               "single GPU"}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
            "gpu_launches": int(launches), "clocks": clocks,
            "per_rank": per_rank, "comm_exposed_ms": max(p["comm_exposed_ms"] for p in per_rank),
        }
        if gstats:
            line["gather"] = {"transport": GATHER_TRANSPORT[gather_mode],
                              "steps_per_slab": slab_schedule(args.steps), "slabs": slabs, "calls": gstats["gathers"], "nccl_version": gstats["nccl_version"],
                              "bytes_into_rank0_per_step": 4 * sum(part["counts"][1:])}
        if shares is not None:
            line["balance"] = {"what": "time segments in proportion to each GPU's rate measured in warm-up rehearsals of the timed region (b200sdr_chain_segment_weighted); "
                                       "the step's total stays n_gpus * samples_per_gpu_per_step",
                               "segment_samples_over_average": shares, "rehearsal_steps_per_ms": rates}
        if e2e:
            line["e2e"] = e2e
    part.clear()
    del x, x_all, demod, scratch_out
    if gather is not None:
        gather.close()
    torch.cuda.empty_cache()
    return line, wl


# ---------------------------------------------------------------------------------------------------
# C5: wideband channelizer (strong scaling: one fixed block, split over the GPUs)
# ---------------------------------------------------------------------------------------------------
def measure_channelizer(args, ctx, steps, warmup, log2n, only_rank0_unsharded=False):
    """One measurement of the C5 workload.  Sharded over ctx.world GPUs, or (only_rank0_unsharded) the whole block on rank 0 alone
    with the other ranks idle -- the N = 1 figure a speed-up is quoted against, taken in the same run."""
    import cuda_sdr_b200 as sdr
    from cuda_sdr_b200 import sharding

    torch = ctx.torch
    world = 1 if only_rank0_unsharded else ctx.world
    rank = 0 if only_rank0_unsharded else ctx.rank
    dev = ctx.dev
    taps = load_taps_module()
    fs, T1c, D1c, T2c, D2c = 153.6e6, 4097, 640, 273, 5            # 153.6 Msps = 48 kHz x 640 x 5
    total = args.channels
    n = 1 << log2n
    freqs = [(c - total / 2) * 600e3 + 100e3 for c in range(total)]  # 600 kHz raster
    mods = [c & 1 for c in range(total)]                             # alternating AM / FM
    t1 = taps.lowpass(T1c, 100e3, fs)
    t2 = taps.lowpass(T2c, 0.45 * 48e3, fs / D1c)
    gain = sdr.fm_gain(fs / D1c, 75e3)
    x = sdr.synth.device_int8_iq(n, dev, seed=0x5D120005)           # identical on every rank: stands in for a broadcast feed
    # Two decompositions, no collective on the filter path either way:
    #   filter-bank route (all channels on one raster: ONE pass yields every channel) -> overlapped TIME segments, every GPU
    #     produces all channels for its share of the audio outputs (b200sdr_channelizer_segment);
    #   per-channel route -> channel c on rank c mod G, input replicated.
    probe = sdr.Channelizer(fs, freqs, mods, t1, D1c, t2, D2c, fm_gains=[gain] * total, device=ctx.local_rank)
    by_time = probe.variant.startswith("pfb") and args.shard != "channels"
    if by_time:
        mine = list(range(total))
        ch = probe
        _, n_audio_total = ch.counts(n)
        segs = [ch.segment(n_audio_total, world, r) for r in range(world)]
        first_out, n_audio, first_in, in_count = segs[rank]
        x_mine = x[2 * first_in: 2 * (first_in + in_count)]
        floats = [total * s[1] for s in segs]
    else:
        mine = sharding.channels_of_rank(total, world, rank)
        ch = sdr.Channelizer(fs, [freqs[c] for c in mine], [mods[c] for c in mine], t1, D1c, t2, D2c, fm_gains=[gain] * len(mine),
                             device=ctx.local_rank)
        del probe
        x_mine = x
        _, n_audio = ch.counts(n)
        floats = [len(sharding.channels_of_rank(total, world, r)) * n_audio for r in range(world)]
    n_demod = (n_audio - 1) * D2c + T2c
    scratch = torch.empty(len(mine), (n_demod + 3) // 4 * 4, dtype=torch.float32, device=dev)[:, :n_demod]  # rows 16-byte aligned
    slabs = 3
    gather = ctx.make_gather(floats, slabs) if world > 1 else None
    if gather is not None:
        outs = [gather.slab(s)[: len(mine) * n_audio].view(len(mine), n_audio) for s in range(slabs)]
    else:
        outs = [torch.empty(len(mine), n_audio, dtype=torch.float32, device=dev) for _ in range(2)]
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    step_no = [0]

    def step(i=None):
        buf = step_no[0] % len(outs)
        step_no[0] += 1
        if gather is not None:
            gather.acquire(buf)
        if i is not None:
            k_events[i][0].record()
        ch.run(x_mine, n_audio, out=outs[buf], scratch=scratch)
        if i is not None:
            k_events[i][1].record()
        if gather is not None:  # this step's audio of every rank to rank 0 on the side stream
            gather.submit(buf)

    def quiesce():
        if gather is not None:
            gather.finish()
        if only_rank0_unsharded:
            torch.cuda.synchronize()
        else:
            ctx.barrier()

    # the clock sampler polls every 50 ms and the timed region lasts a few milliseconds: it runs from the warm-up (the same kernels
    # back to back) through the timed region
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    t_warm = time.perf_counter()
    extra = 0
    while time.perf_counter() - t_warm < min(args.warmup_seconds, 0.3):
        ch.run(x_mine, n_audio, out=outs[0], scratch=scratch)
        extra += 1
        torch.cuda.synchronize()
    n_warm = max(warmup, 3)
    for _ in range(n_warm):
        step()
    quiesce()
    launches0 = sdr._native.launch_count()
    t0, tk, t1e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    quiesce()
    t0.record()
    for i in range(steps):
        step(i)
    tk.record()
    if gather is not None:
        gather.finish()
    t1e.record()
    quiesce()
    launches = sdr._native.launch_count() - launches0
    my_ms, my_kernel_ms = t0.elapsed_time(t1e), t0.elapsed_time(tk)
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in k_events)
    clocks = sampler.stop()
    if only_rank0_unsharded:
        total_ms, per_rank = my_ms, None
    else:
        total_ms = ctx.max_over_ranks(my_ms)
        per_rank = ctx.gather_objects({"rank": rank, "ms_per_step": my_ms / steps, "kernel_ms_per_step": my_kernel_ms / steps,
                                       "comm_exposed_ms": my_ms - my_kernel_ms, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")})
    rec = None
    if rank == 0:
        hbm, _ = hbm_peak()
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_tc = float(json.load(open(peaks_path))["bf16_tflops"]) if os.path.exists(peaks_path) else 1590.0
        M = -(-T1c // D1c)
        pfb = ch.variant.startswith("pfb")
        fused_audio = "audio FIR fused" in ch.variant
        samples_this_gpu = x_mine.numel() // 2
        # the per-channel chain costs ~38 flop per input sample per channel; the filter bank does the same job
        # for all channels at once, so the per-channel figure stays the ALGORITHMIC work the line is normalised by
        flops = samples_this_gpu * len(mine) * (12.0 + 4.0 * T1c / D1c + 10.0 / D1c + 2.0 * T2c / (D1c * D2c))
        if pfb:
            # bytes the launches must move: int8 IQ in, (unless fused) demodulated samples out and back in, audio out
            alg_bytes = 2.0 * samples_this_gpu + len(mine) * ((0.0 if fused_audio else 8.0 * n_demod) + 4.0 * n_audio)
            roofline = {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm, "traffic": None,
                        "kernel": "polyphase filter bank + fp64 FFT + demod, all channels (pfb kernel) + batched audio FIR",
                        "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                        "per_channel_equivalent_tflops": flops / (k_ms * 1e-3) / 1e12,
                        "fp64_dfma_per_step": samples_this_gpu / D1c * (4.0 * 17 * 256 + 5500.0),
                        "note": "the filter bank replaces ~38 flop per sample PER CHANNEL by ~60 flop per sample for all channels; the binding "
                                "unit of the kernel is the FP64 pipe (filter bank + FFT in fp64, DESIGN.md), HBM is the contract's roof"}
        else:
            executed_ops = 2.0 * (n_demod + M) * (2 * D1c) * (16 if M > 4 else 8) * 3 * len(mine)      # int8 MMA ops incl. digits and padding
            roofline = {"bound": "tensor", "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": peak_tc, "unit": "TFLOP/s",
                        "frac": flops / (k_ms * 1e-3) / 1e12 / peak_tc, "traffic": None,
                        "kernel": "channelKernel (int8 GEMM RF stage + demod) + batched audio FIR", "kernel_ms": k_ms,
                        "algorithmic_flops_per_launch": flops, "executed_int8_tops": executed_ops / (k_ms * 1e-3) / 1e12,
                        "legacy_imma_peak_tops_measured": 1143.0}
        rec = {
            "metric": "input Msps through the 256-channel wideband channelizer (int8 -> mix -> FIR -> AM/FM demod -> audio FIR per channel)",
            "value": n * steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": n_warm, "warmup_extra_steps": extra,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 filter bank + FFT, f32 demod / audio FIR" if pfb else "s8 x s8 -> s32 (24-bit fixed-point taps), f32 epilogue",
            "data": "synthetic",
            "config": {"workload": f"C5 wideband channelizer: 2^{log2n} int8 IQ samples per step at 153.6 Msps-class rate, {total} channels on a 600 kHz "
                                   f"raster alternating AM/FM, per channel mix -> {T1c}-tap FIR /{D1c} -> demod -> {T2c}-tap audio FIR /{D2c}",
                       "channels_total": total, "channels_this_gpu": len(mine), "samples_per_step": n, "samples_this_gpu": samples_this_gpu,
                       "route": ch.variant,
                       "parallelism": (("overlapped time segments of the wideband stream, every GPU produces all channels for its share of the "
                                        "audio outputs" if by_time else "channels interleaved over the GPUs (c mod G), input replicated") +
                                       "; b200sdr_gather: one slab of audio per step to rank 0, 3 slabs; " +
                                       GATHER_TRANSPORT[gather.mode if gather is not None else 0])
                       if world > 1 else "single GPU",
                       "l2": f"input block {2 * n >> 20} MiB exceeds the 126 MB L2"},
            "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
        }
        if per_rank is not None and world > 1:
            rec["per_rank"] = per_rank
            rec["comm_exposed_ms"] = max(p["comm_exposed_ms"] for p in per_rank)
            rec["gather"] = {"bytes_into_rank0_per_step": 4 * sum(floats[1:]), "slabs": slabs, "calls": gather.stats()["gathers"],
                             "transport": GATHER_TRANSPORT[gather.mode]}
    del x, x_mine, scratch, outs
    if gather is not None:
        gather.close()
    ch.close()
    torch.cuda.empty_cache()
    return rec


def channelizer_record(args, ctx, steps, warmup, log2n):
    """C5 on ctx.world GPUs plus, for N > 1, the same block on ONE GPU in the same run (rank 0 alone, the others wait)."""
    rec = measure_channelizer(args, ctx, steps, warmup, log2n)
    if ctx.world > 1:
        single = measure_channelizer(args, ctx, max(3, steps // 2), warmup, log2n, only_rank0_unsharded=True) if ctx.rank == 0 else None
        ctx.barrier()
        if ctx.rank == 0:
            rec["single_gpu_same_run"] = {"value": single["value"], "ms_per_step": single["ms_per_step"], "unit": UNIT, "steps": single["steps"]}
            rec["speedup_vs_single_gpu_same_run"] = rec["value"] / single["value"]
    return rec


# ---------------------------------------------------------------------------------------------------
# C4: FIR roofline sweep through the gsdr C-ABI (gsdrFirFC), one GPU
# ---------------------------------------------------------------------------------------------------
def run_firsweep(args, ctx):
    import cuda_sdr_b200 as sdr
    from cuda_sdr_b200 import ops

    torch = ctx.torch
    if ctx.rank != 0:
        return None
    taps = load_taps_module()
    hbm, peak_src = hbm_peak()
    ffma = 71.0  # TFLOP/s, tools/microbench.cu (ffma_rrr) on this pool's B200 at 1.965 GHz
    log2 = args.log2_block if args.log2_block != LOG2_BLOCK else 26
    n = 1 << log2
    x = torch.view_as_complex(torch.randn(n, 2, device=ctx.dev, dtype=torch.float32))
    cells, launches0 = [], sdr._native.launch_count()
    reps = max(1, min(args.steps, 5))
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    t_all0 = time.perf_counter()
    for T in [32, 64, 128, 256, 512, 1024, 2048, 4096]:
        h = torch.from_numpy(taps.lowpass(T, 0.2, 1.0)).to(ctx.dev)
        for D in [1, 2, 4, 8, 16, 32, 64]:
            n_out = ops.fir_num_outputs(n, T, D)
            for _ in range(max(1, min(args.warmup, 2))):
                ops.fir("fc", h, x, D, n_out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            times, spent = [], 0.0
            while len(times) < reps and (spent < 400.0 or not times):
                e0.record()
                ops.fir("fc", h, x, D, n_out)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                spent += times[-1]
            ms = statistics.median(times)
            gbs = n * (8.0 + 8.0 / D) / (ms * 1e-3) / 1e9
            tflops = n * 4.0 * T / D / (ms * 1e-3) / 1e12
            t_roof = max(n * (8.0 + 8.0 / D) / (hbm * 1e9), n * 4.0 * T / D / (ffma * 1e12)) * 1e3
            cells.append({"taps": T, "decimation": D, "ms": ms, "msps": n / (ms * 1e-3) / 1e6, "gbs": gbs, "tflops": tflops,
                          "bound": "hbm" if gbs / hbm >= tflops / ffma else "ffma", "frac": t_roof / ms})
    wall = time.perf_counter() - t_all0
    clocks = sampler.stop()
    fracs = [c["frac"] for c in cells]
    total_samples = n * len(cells)
    total_ms = sum(c["ms"] for c in cells)
    worst = min(cells, key=lambda c: c["frac"])
    return {
        "metric": "input Msps of the decimating FIR (gsdrFirFC: real taps x complex float), mean over the C4 sweep cells", "value": total_samples / (total_ms * 1e-3) / 1e6,
        "unit": UNIT, "n_gpus": 1, "steps": reps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": total_ms / len(cells), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C4 FIR roofline sweep: T in 32..4096 x D in 1..64 on 2^{log2} complex-float samples (512 MiB > L2), one B200, "
                               "through the gsdr C-ABI", "cells": len(cells), "wall_seconds": wall},
        "roofline": {"bound": worst["bound"], "achieved": worst["gbs"] if worst["bound"] == "hbm" else worst["tflops"],
                     "peak": hbm if worst["bound"] == "hbm" else ffma, "unit": "GB/s" if worst["bound"] == "hbm" else "TFLOP/s",
                     "frac": worst["frac"], "traffic": None, "kernel": f"worst cell: T={worst['taps']} D={worst['decimation']}",
                     "frac_min": min(fracs), "frac_median": statistics.median(fracs), "frac_max": max(fracs),
                     "cells_at_or_above_0.70": sum(f >= 0.70 for f in fracs), "peak_source": peak_src + "; FFMA 71 TFLOP/s measured (tools/microbench.cu)"},
        "cells": cells, "gpu_launches": int(sdr._native.launch_count() - launches0), "clocks": clocks,
    }


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    ctx = Ctx(args)
    line = None
    if args.workload == "firsweep":
        line = run_firsweep(args, ctx)
    elif args.workload == "channelizer":
        log2n = args.log2_block if args.log2_block != LOG2_BLOCK else 29  # 3.5 s of signal per step: long segments amortise per-launch costs at N = 8
        line = channelizer_record(args, ctx, args.steps, args.warmup, log2n)
    else:
        line, wl = run_chain(args, ctx)
        if not args.skip_channelizer:
            # the north-star scaling workload rides on every line: C5, 256 channels, the same N GPUs, strong scaling
            sub = channelizer_record(args, ctx, max(5, min(args.steps, 20)), 3, 29)
            if ctx.rank == 0:
                keep = ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "single_gpu_same_run", "speedup_vs_single_gpu_same_run",
                        "comm_exposed_ms", "gather", "per_rank", "gpu_launches")
                line["channelizer"] = {k: sub[k] for k in keep if k in sub}
                line["channelizer"]["route"] = sub["config"]["route"]
                line["channelizer"]["workload"] = sub["config"]["workload"]
                line["channelizer"]["roofline_frac"] = sub["roofline"]["frac"]
        if ctx.rank == 0 and not args.skip_cpu and ctx.world == 1:  # the CPU baseline is reported at N = 1 only
            v, cores, sample = cpu_chain_rate(wl, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        if ctx.rank == 0 and ctx.world == 1 and not args.skip_e2e:
            line.update(reference_pipeline_rate(wl))
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
